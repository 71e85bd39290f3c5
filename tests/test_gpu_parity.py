"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C-ABI library; the oracle
is only the checker.  Bit-exact for every format (stricter than the <= 1 ulp the north star allows for fp16/fp32:
the kernels evaluate the same IEEE operations in the same order)."""
import hashlib
import json
import os

import numpy as np
import pytest

import floor_b200
from floor_b200.image_types import IMAGE_TYPE as T, MEMORY_FLAG as MF
from floor_b200 import image_types as it

pytestmark = pytest.mark.gpu
M = T.FLAG_MIPMAPPED | T.READ_WRITE
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
FLOAT_ULP_TOLERANCE = 0  # north star allows 1 ulp for fp16 / fp32; we require 0


def gpu_chain(gpu_ctx, l0, dim, t, **kw):
    ctx, dev, q = gpu_ctx
    img = ctx.create_image(q, dim, t, flags=MF.READ_WRITE | MF.HOST_READ_WRITE, **kw)
    img.upload_levels(q, l0, 0, 0)
    img.generate_mip_map_chain(q)
    out = img.download_levels(q)
    plan = img.plan()
    img.destroy()
    return out, plan


def assert_same(a, b, t, dim, what=""):
    assert a.size == b.size, (hex(t), dim, a.size, b.size)
    if not np.array_equal(a, b):
        bad = np.nonzero(a != b)[0]
        raise AssertionError(f"{what} type={t:#x} dim={dim}: {bad.size} differing bytes, first at {int(bad[0])} "
                             f"(gpu={int(a[bad[0]])} oracle={int(b[bad[0]])})")


ALL_FORMATS = [T.R8, T.RG8, T.RGBA8, T.R16, T.RG16, T.RGBA16, T.R8I_NORM, T.RG8I_NORM, T.RGBA8I_NORM, T.R16I_NORM, T.RG16I_NORM,
               T.RGBA16I_NORM, T.R8UI, T.RG8UI, T.RGBA8UI, T.R8I, T.RG8I, T.RGBA8I, T.R16UI, T.RG16UI, T.RGBA16UI, T.R16I, T.RG16I,
               T.RGBA16I, T.R32UI, T.RG32UI, T.RGBA32UI, T.R32I, T.RG32I, T.RGBA32I, T.R16F, T.RG16F, T.RGBA16F, T.R32F, T.RG32F,
               T.RGBA32F]


@pytest.mark.parametrize("fmt", ALL_FORMATS)
def test_single_pass_2d_all_formats(gpu_ctx, oracle_mod, fmt):
    """power-of-two 2D images large enough for the TMA single-pass kernel (several tiles, groups and the layer stage)"""
    bpp = it.bytes_per_pixel(fmt)
    w = max(2 * 512 // bpp, 256)
    for dim in [(w, 256), (w * 2, 128), (512 // bpp, 64)]:
        t = T.IMAGE_2D | fmt | M
        l0 = oracle_mod.fill_synthetic(dim, t, 100 + (fmt & 0xFFFF))
        want = oracle_mod.generate_mip_map_chain(l0, dim, t, threads=8)
        for units in (False, True):  # single tiles, and 2x2-tile units where the tile grid allows it
            got, plan = gpu_chain(gpu_ctx, l0, dim, t, units=units)
            assert plan["single_pass"], (hex(t), dim, plan)
            assert_same(got, want, t, dim, f"single-pass 2D units={units}")


@pytest.mark.parametrize("fmt", ALL_FORMATS)
def test_single_pass_3d_all_formats(gpu_ctx, oracle_mod, fmt):
    bpp = it.bytes_per_pixel(fmt)
    for dim in [(2 * 128 // bpp if bpp < 16 else 32, 32, 32), (128 // bpp if bpp < 16 else 8, 16, 64)]:
        t = T.IMAGE_3D | fmt | M
        l0 = oracle_mod.fill_synthetic(dim, t, 200 + (fmt & 0xFFFF))
        want = oracle_mod.generate_mip_map_chain(l0, dim, t, threads=8)
        for units in (False, True):
            got, plan = gpu_chain(gpu_ctx, l0, dim, t, units=units)
            assert plan["single_pass"], (hex(t), dim, plan)
            assert_same(got, want, t, dim, f"single-pass 3D units={units}")


@pytest.mark.parametrize("fmt", ALL_FORMATS)
def test_general_path_all_formats(gpu_ctx, oracle_mod, fmt):
    """NPOT / small images take the multi-level tile kernel, 1D images the literal per-level kernel; the literal kernel
    (force_generic) must agree with both"""
    for base, dim in [(T.IMAGE_2D, (37, 21)), (T.IMAGE_2D, (100, 60)), (T.IMAGE_2D_ARRAY, (20, 12, 3)), (T.IMAGE_3D, (12, 10, 6)),
                      (T.IMAGE_1D, (33,)), (T.IMAGE_1D_ARRAY, (64, 2)), (T.IMAGE_2D, (16, 16))]:
        t = base | fmt | M
        l0 = oracle_mod.fill_synthetic(dim, t, 300 + (fmt & 0xFFFF))
        got, plan = gpu_chain(gpu_ctx, l0, dim, t)
        assert not plan["single_pass"]
        want = oracle_mod.generate_mip_map_chain(l0, dim, t, threads=4)
        assert_same(got, want, t, dim, "general")
        lit, plan = gpu_chain(gpu_ctx, l0, dim, t, force_generic=True)
        assert plan["launches"] == sum(1 for l in range(1, oracle_mod.mip_level_count(dim, t)) if oracle_mod.level_size(dim, t, l))
        assert_same(lit, want, t, dim, "literal general kernel")


@pytest.mark.parametrize("fmt", ALL_FORMATS)
def test_tile_kernel_all_formats(gpu_ctx, oracle_mod, fmt):
    """multi-level tile kernel (flmip_tile2d / 3d): NPOT sizes with partial border tiles, odd levels, several launches per chain
    (more than 6 / 4 levels), arrays, cubes, and power-of-two images forced onto it"""
    cases = [(T.IMAGE_2D, (333, 129), {}), (T.IMAGE_2D, (1000, 70), {}), (T.IMAGE_2D_ARRAY, (130, 67, 3), {}), (T.IMAGE_CUBE, (96, 96), {}),
             (T.IMAGE_3D, (70, 33, 18), {}), (T.IMAGE_3D, (65, 40, 100), {}), (T.IMAGE_2D, (256, 128), {"force_tiled": True}),
             (T.IMAGE_3D, (64, 32, 32), {"force_tiled": True}), (T.IMAGE_2D, (64, 4), {}), (T.IMAGE_2D, (5, 3), {}),
             # texel-2 quirk sizes (tests/test_npot_weights.py) at level 0 and, via 2624 = 41 << 6 / 188 = 47 << 2, deep inside a tile
             (T.IMAGE_2D, (41, 47), {}), (T.IMAGE_2D, (2624, 188), {}), (T.IMAGE_3D, (55, 61, 41), {}), (T.IMAGE_3D, (328, 94, 82), {}),
             # ... and where the host has to cut a launch short: level 5 is 41 wide (a tile holds 2 texels of it), depth 41 at level 3
             (T.IMAGE_2D, (1312, 200), {}), (T.IMAGE_2D, (96, 1504), {}), (T.IMAGE_3D, (40, 48, 328), {})]
    for base, dim, kw in cases:
        t = base | fmt | M
        l0 = oracle_mod.fill_synthetic(dim, t, 600 + (fmt & 0xFFFF))
        got, plan = gpu_chain(gpu_ctx, l0, dim, t, **kw)
        assert not plan["single_pass"], (dim, plan)
        assert_same(got, oracle_mod.generate_mip_map_chain(l0, dim, t, threads=4), t, dim, "tile kernel")
        if it.bits_per_channel(t) == 16 and (t & T.FLAG_NORMALIZED):
            got, _ = gpu_chain(gpu_ctx, l0, dim, t, no_double=True, **kw)
            assert_same(got, oracle_mod.generate_mip_map_chain(l0, dim, t, no_double=True, threads=4), t, dim, "tile kernel, no_double")


def test_tile_kernel_launch_plan_and_large_npot(gpu_ctx, oracle_mod):
    """3840 x 2160 RGBA8 (12 levels) = 2 launches; 64 layers of 1920 x 1080 RGBA16F; a 300 x 200 x 100 volume"""
    for base, fmt, dim, launches in [(T.IMAGE_2D, T.RGBA8, (3840, 2160), 2), (T.IMAGE_2D_ARRAY, T.RGBA16F, (1920, 1080, 6), 2),
                                     (T.IMAGE_3D, T.R32F, (300, 200, 100), 2), (T.IMAGE_2D, T.R8, (4097, 33), 1), (T.IMAGE_2D, T.RG8, (5000, 3000), 2),
                                     (T.IMAGE_2D, T.RGBA8, (1312, 1312), 2)]:
        t = base | fmt | M
        l0 = oracle_mod.fill_synthetic(dim, t, 77)
        want = oracle_mod.generate_mip_map_chain(l0, dim, t, threads=os.cpu_count() or 8)
        # the LDG tile kernel: ceil((levels - 1) / 6) launches (4 levels per launch for volumes) ...
        got, plan = gpu_chain(gpu_ctx, l0, dim, t, no_tma_tiles=True)
        assert plan["launches"] == launches and plan["tma_tile_launches"] == 0 and not plan["single_pass"], (dim, plan)
        assert_same(got, want, t, dim, "LDG tile kernel, large")
        # ... and whatever the planner picks (2D images with 16-byte row multiples: the persistent TMA tile kernel, fewer launches)
        got, plan = gpu_chain(gpu_ctx, l0, dim, t)
        assert plan["launches"] <= launches and not plan["single_pass"], (dim, plan)
        assert_same(got, want, t, dim, "tile kernel, large")


def test_regenerate_from_dirty_level(gpu_ctx, oracle_mod):
    """flmip_mip_chain_generate_from: overwrite level 2, regenerate only the levels below it (tile kernel from level 2)"""
    ctx, dev, q = gpu_ctx
    for dim, bt in [((512, 256), T.IMAGE_2D | T.RGBA8), ((300, 200), T.IMAGE_2D | T.RGBA16F), ((64, 64, 64), T.IMAGE_3D | T.R32F)]:
        t = bt | M
        l0 = oracle_mod.fill_synthetic(dim, t, 88)
        img = ctx.create_image(q, dim, t)
        img.upload_levels(q, l0, 0, 0)
        img.generate_mip_map_chain(q)
        ldim = oracle_mod.level_dim(dim, t, 2)
        sub = tuple(d for d in ldim[: len(dim)])
        new2 = oracle_mod.fill_synthetic(sub, t, 89)
        img.upload_levels(q, new2, 2, 2)
        img.enqueue_mip_map_chain(q, first_level=2)
        q.finish()
        got = img.download_levels(q)
        img.destroy()
        want_top = oracle_mod.generate_mip_map_chain(l0, dim, t, threads=4)
        o2 = oracle_mod.level_offset(dim, t, 2)
        want_tail = oracle_mod.generate_mip_map_chain(new2, sub, t, threads=4)
        assert np.array_equal(got[:o2], want_top[:o2])
        assert_same(got[o2:], want_tail[: got.size - o2], t, dim, "regenerated tail")


def test_general_equals_single_pass(gpu_ctx, oracle_mod):
    """the POT fast form is a derived identity of the general sampler: both GPU paths must agree"""
    for base, fmt, dim in [(T.IMAGE_2D, T.RGBA16F, (512, 512)), (T.IMAGE_2D_ARRAY, T.RGBA8, (256, 128, 5)), (T.IMAGE_3D, T.R32F, (64, 64, 64)),
                           (T.IMAGE_2D, T.RGBA32UI, (128, 128)), (T.IMAGE_2D, T.RGBA16, (256, 256))]:
        t = base | fmt | M
        l0 = oracle_mod.fill_synthetic(dim, t, 7)
        a, pa = gpu_chain(gpu_ctx, l0, dim, t)
        b, pb = gpu_chain(gpu_ctx, l0, dim, t, force_generic=True)
        assert pa["single_pass"] and not pb["single_pass"]
        assert_same(a, b, t, dim, "single-pass vs general")


def test_layered_and_cube(gpu_ctx, oracle_mod):
    cases = [(T.IMAGE_2D_ARRAY | T.RGBA8, (256, 256, 7)), (T.IMAGE_2D_ARRAY | T.RGBA8, (1024, 1024, 3)), (T.IMAGE_CUBE | T.RGBA32F, (128, 128)),
             (T.IMAGE_CUBE_ARRAY | T.RGBA32F, (64, 64, 3)), (T.IMAGE_CUBE_ARRAY | T.RGBA16F, (256, 256, 2)), (T.IMAGE_DEPTH_ARRAY | T.FORMAT_32 | T.FLOAT, (256, 128, 4))]
    for bt, dim in cases:
        t = bt | M
        l0 = oracle_mod.fill_synthetic(dim, t, 11)
        got, plan = gpu_chain(gpu_ctx, l0, dim, t)
        assert plan["single_pass"] and plan["launches"] == 1, plan
        assert_same(got, oracle_mod.generate_mip_map_chain(l0, dim, t, threads=8), t, dim, "layered")


def test_layered_with_units(gpu_ctx, oracle_mod):
    """2x2-tile units on layered / cube images with several groups per layer"""
    for bt, dim in [(T.IMAGE_2D_ARRAY | T.RGBA8, (1024, 1024, 5)), (T.IMAGE_CUBE_ARRAY | T.RGBA32F, (512, 512, 2)), (T.IMAGE_2D_ARRAY | T.R16F, (2048, 512, 3)),
                    (T.IMAGE_3D | T.RGBA8, (256, 256, 256)), (T.IMAGE_2D | T.RG16, (4096, 2048))]:
        t = bt | M
        l0 = oracle_mod.fill_synthetic(dim, t, 17)
        got, plan = gpu_chain(gpu_ctx, l0, dim, t, units=True)
        assert plan["single_pass"] and plan["launches"] == 1, plan
        assert_same(got, oracle_mod.generate_mip_map_chain(l0, dim, t, threads=16), t, dim, "layered units")


def test_baseline_configs_full_size(gpu_ctx, oracle_mod):
    """BASELINE.json configs at full size where one image fits the oracle's budget: C1, C2, C5 against the oracle on every level;
    C3 / C4 on a contiguous shard of layers / one cube (layers are independent; the oracle runs per layer anyway)"""
    ctx, dev, q = gpu_ctx
    cases = [(1, (1024, 1024), T.IMAGE_2D | T.RGBA8), (2, (8192, 8192), T.IMAGE_2D | T.RGBA16F), (5, (512, 512, 512), T.IMAGE_3D | T.R32F),
             (3, (1024, 1024, 32), T.IMAGE_2D_ARRAY | T.RGBA8), (4, (4096, 4096, 1), T.IMAGE_CUBE_ARRAY | T.RGBA32F)]
    for cid, dim, bt in cases:
        t = bt | M
        img = ctx.create_image(q, dim, t)
        img.fill_synthetic(q, cid, layer_id0=7)
        img.generate_mip_map_chain(q)
        got = img.download_levels(q)
        plan = img.plan()
        img.destroy()
        assert plan["single_pass"] and plan["launches"] == 1, (cid, plan)
        l0 = got[: oracle_mod.level_size(dim, t, 0)]
        assert np.array_equal(l0, oracle_mod.fill_synthetic(dim, t, cid, layer_id0=7)), cid  # device fill == oracle fill
        want = oracle_mod.generate_mip_map_chain(l0, dim, t, threads=os.cpu_count() or 8)
        assert hashlib.sha256(got.tobytes()).digest() == hashlib.sha256(want.tobytes()).digest(), f"config {cid}: GPU chain differs from the oracle"
        del got, want, l0


def test_top_level_is_the_mean_of_means_property(gpu_ctx, oracle_mod):
    """size-independent property at full size: for a constant image every level is that constant (fp16 / fp32 exactly; unorm8 exactly,
    cf. the constant-image KATs of the oracle tests)"""
    ctx, dev, q = gpu_ctx
    for dim, bt, dt, val in [((8192, 8192), T.IMAGE_2D | T.RGBA16F, np.float16, 0.3330078125), ((512, 512, 512), T.IMAGE_3D | T.R32F, np.float32, 1.2345678),
                             ((1024, 1024, 16), T.IMAGE_2D_ARRAY | T.RGBA8, np.uint8, 77)]:
        t = bt | M
        img = ctx.create_image(q, dim, t)
        n0 = img.levels[0]["size"] // np.dtype(dt).itemsize
        host = np.full(n0, val, dtype=dt)
        img.upload_levels(q, host, 0, 0)
        img.generate_mip_map_chain(q)
        got = img.download_levels(q).view(dt)
        img.destroy()
        assert np.all(got == np.array(val, dtype=dt)), (dim, hex(t))
        del host, got


def test_non_square_and_level_limit(gpu_ctx, oracle_mod):
    for dim in [(2048, 64), (64, 2048), (4096, 256), (128, 1024)]:
        t = T.IMAGE_2D | T.RGBA16F | M
        l0 = oracle_mod.fill_synthetic(dim, t, 13)
        got, plan = gpu_chain(gpu_ctx, l0, dim, t)
        assert plan["single_pass"]
        assert_same(got, oracle_mod.generate_mip_map_chain(l0, dim, t, threads=8), t, dim, "non-square")
    for limit in [2, 3, 5, 8]:
        dim, t = (1024, 1024), T.IMAGE_2D | T.RGBA8 | M
        l0 = oracle_mod.fill_synthetic(dim, t, 14)
        got, _ = gpu_chain(gpu_ctx, l0, dim, t, mip_level_limit=limit)
        assert_same(got, oracle_mod.generate_mip_map_chain(l0, dim, t, mip_level_limit=limit, threads=8), t, dim, f"limit {limit}")
    for dim in [(128, 16, 16), (32, 16, 256), (64, 128, 32)]:
        t = T.IMAGE_3D | T.R32F | M
        l0 = oracle_mod.fill_synthetic(dim, t, 15)
        got, plan = gpu_chain(gpu_ctx, l0, dim, t)
        assert plan["single_pass"]
        assert_same(got, oracle_mod.generate_mip_map_chain(l0, dim, t, threads=8), t, dim, "non-cubic")


def test_no_double_variant(gpu_ctx, oracle_mod):
    for fmt in [T.RGBA16, T.RG16I_NORM]:
        dim, t = (256, 256), T.IMAGE_2D | fmt | M
        l0 = oracle_mod.fill_synthetic(dim, t, 16)
        a, _ = gpu_chain(gpu_ctx, l0, dim, t, no_double=True)
        assert_same(a, oracle_mod.generate_mip_map_chain(l0, dim, t, no_double=True, threads=8), t, dim, "no-double")
        b, _ = gpu_chain(gpu_ctx, l0, dim, t)
        assert not np.array_equal(a, b)  # the two encoder modes do differ


def test_adversarial_floats(gpu_ctx, oracle_mod):
    """+-0, subnormals, +-65504 and 1-ulp neighbours (fp16); mixed exponents / cancellation and subnormals (fp32)"""
    rng = np.random.default_rng(3)
    dim = (256, 256)
    h = rng.integers(0, 0x7C00, size=dim[0] * dim[1] * 4, dtype=np.uint16) | (rng.integers(0, 2, size=dim[0] * dim[1] * 4, dtype=np.uint16) << 15)
    special = np.array([0x0000, 0x8000, 0x0001, 0x8001, 0x03FF, 0x0400, 0x7BFF, 0xFBFF, 0x7BFE, 0x3C00, 0x3C01, 0xBC00], np.uint16)
    h[:: 5] = special[rng.integers(0, special.size, size=h[::5].size)]
    t = T.IMAGE_2D | T.RGBA16F | M
    got, _ = gpu_chain(gpu_ctx, h, dim, t)
    assert_same(got, oracle_mod.generate_mip_map_chain(h, dim, t, threads=8), t, dim, "adversarial fp16")
    e = rng.integers(1, 254, size=dim[0] * dim[1] * 4, dtype=np.uint32)
    f = (rng.integers(0, 2, size=e.size, dtype=np.uint32) << 31) | (e << 23) | rng.integers(0, 1 << 23, size=e.size, dtype=np.uint32)
    f[::7] = rng.integers(0, 1 << 23, size=f[::7].size, dtype=np.uint32)  # subnormals
    f[::11] = np.uint32(0x80000000)
    t = T.IMAGE_2D | T.RGBA32F | M
    got, _ = gpu_chain(gpu_ctx, f, dim, t)
    assert_same(got, oracle_mod.generate_mip_map_chain(f, dim, t, threads=8), t, dim, "adversarial fp32")


def test_subnormal_ties_round_like_the_reference(gpu_ctx, oracle_mod):
    """fp32 texels that are tiny multiples of 2^-149: (b - a) * 0.5 is then an inexact subnormal and a fused multiply-add would
    round differently from the reference's separate product and sum (a = 2^-149, b = 2^-148 -> 2^-149, fused: 2^-148).  ptxas
    contracts mul.rn.f32x2 + add.rn.f32x2 on its own, so the packed kernels multiply through mul2_sep().  All three kernels."""
    rng = np.random.default_rng(12)
    small = np.array([0, 1, 2, 3, 4, 5, 6, 7, 9, 11, 0x80000001, 0x80000002, 0x80000003, 0x80000005, 0x00800000, 0x00800001, 0x007FFFFF], np.uint32)
    for base, fmt, dim in [(T.IMAGE_2D, T.RGBA32F, (128, 128)), (T.IMAGE_2D, T.R32F, (256, 128)), (T.IMAGE_2D, T.RG32F, (128, 128)),
                           (T.IMAGE_3D, T.R32F, (64, 32, 32)), (T.IMAGE_3D, T.RGBA32F, (16, 16, 32)), (T.IMAGE_2D, T.RGBA32F, (100, 60)),
                           (T.IMAGE_3D, T.RG32F, (20, 12, 10))]:
        t = base | fmt | M
        n = int(np.prod(dim)) * it.channel_count(t)
        f = small[rng.integers(0, small.size, size=n)]
        want = oracle_mod.generate_mip_map_chain(f, dim, t, threads=4)
        for kw in ({}, {"force_tiled": True}, {"force_generic": True}):
            got, _ = gpu_chain(gpu_ctx, f, dim, t, **kw)
            assert_same(got, want, t, dim, f"subnormal ties {kw}")


def test_golden_fixtures_on_gpu(gpu_ctx, oracle_mod):
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        cases = json.load(f)
    small = np.load(os.path.join(GOLDEN, "golden_small.npz"))
    for c in cases:
        dim, t = tuple(c["dim"]), int(c["type"], 16)
        if it.channel_count(t) == 3:
            continue
        l0 = oracle_mod.fill_synthetic(dim, t, c["config_id"])
        got, _ = gpu_chain(gpu_ctx, l0, dim, t, no_double=c["no_double"])
        assert hashlib.sha256(got.tobytes()).hexdigest() == c["chain_sha256"], c["name"]
        if c["name"] in small.files:
            assert np.array_equal(got, small[c["name"]])


def test_device_fill_matches_oracle_fill(gpu_ctx, oracle_mod):
    ctx, dev, q = gpu_ctx
    for bt, dim in [(T.IMAGE_2D | T.RGBA16F, (128, 64)), (T.IMAGE_2D_ARRAY | T.RGBA8, (64, 64, 5)), (T.IMAGE_3D | T.R32F, (32, 16, 8)),
                    (T.IMAGE_2D | T.RG32I, (64, 64)), (T.IMAGE_2D | T.R16, (64, 64))]:
        t = bt | M
        img = ctx.create_image(q, dim, t)
        img.fill_synthetic(q, 42, layer_id0=3)
        got = img.download_levels(q, 0, 0)
        want = oracle_mod.fill_synthetic(dim, t, 42, layer_id0=3)
        assert np.array_equal(got, want), hex(t)
        img.destroy()


def test_relaunch_is_idempotent_and_counters_reset(gpu_ctx, oracle_mod):
    ctx, dev, q = gpu_ctx
    dim, t = (2048, 2048), T.IMAGE_2D | T.RGBA16F | M
    img = ctx.create_image(q, dim, t)
    img.fill_synthetic(q, 2)
    outs = []
    for _ in range(3):
        img.generate_mip_map_chain(q)
        outs.append(hashlib.sha256(img.download_levels(q).tobytes()).hexdigest())
    assert outs[0] == outs[1] == outs[2]
    l0 = oracle_mod.fill_synthetic(dim, t, 2)
    assert hashlib.sha256(oracle_mod.generate_mip_map_chain(l0, dim, t, threads=8).tobytes()).hexdigest() == outs[0]
    img.destroy()


def test_layout_and_srgb_flags_do_not_change_the_arithmetic(gpu_ctx, oracle_mod):
    """BGRA / ABGR / ARGB layouts and FLAG_SRGB are properties of how a sampler interprets channels; the minify path treats the
    stored channels alike, so the chains must be byte-identical to plain RGBA"""
    for dim, base in [((256, 128), T.IMAGE_2D | T.RGBA8), ((100, 60), T.IMAGE_2D | T.RGBA8), ((64, 64, 2), T.IMAGE_2D_ARRAY | T.RGBA16F)]:
        l0 = oracle_mod.fill_synthetic(dim, base | M, 91)
        plain, _ = gpu_chain(gpu_ctx, l0, dim, base | M)
        assert_same(plain, oracle_mod.generate_mip_map_chain(l0, dim, base | M, threads=4), base | M, dim, "rgba")
        for extra in (T.LAYOUT_BGRA, T.LAYOUT_ABGR, T.LAYOUT_ARGB, T.FLAG_SRGB, T.LAYOUT_BGRA | T.FLAG_SRGB):
            t = base | extra | M
            got, _ = gpu_chain(gpu_ctx, l0, dim, t)
            assert_same(got, plain, t, dim, "layout / srgb variant")
            assert_same(oracle_mod.generate_mip_map_chain(l0, dim, t, threads=4), plain, t, dim, "oracle, layout / srgb variant")


def test_unsupported_types_are_rejected(gpu_ctx):
    ctx, dev, q = gpu_ctx
    for t, dim in [(T.IMAGE_2D | T.CHANNELS_1 | T.FORMAT_4 | T.UINT | T.FLAG_NORMALIZED | M, (64, 64)),  # half-byte texels (see test_reference_pin.py)
                   (T.IMAGE_2D | T.CHANNELS_4 | T.FORMAT_4 | T.UINT | M, (64, 64)),  # 4-bit formats only exist normalized
                   (T.IMAGE_2D | T.FORMAT_64 | T.FLOAT | T.CHANNELS_1 | M, (64, 64)),
                   (T.IMAGE_2D | T.FLAG_MSAA | T.RGBA8 | M, (64, 64)), (T.IMAGE_CUBE | T.RGBA8 | M, (64, 32))]:
        with pytest.raises(floor_b200.FlmipError):
            ctx.create_image(q, dim, t)


def test_generate_mip_maps_life_cycle(gpu_ctx, oracle_mod):
    """ctor upload -> chain, write() -> chain, map()/unmap() -> chain (cuda_image.cpp:533-536, 667-670, 803-806)"""
    ctx, dev, q = gpu_ctx
    dim, t = (512, 256), T.IMAGE_2D | T.RGBA8 | T.FLAG_MIPMAPPED | T.READ
    l0 = oracle_mod.fill_synthetic(dim, t, 21)
    img = ctx.create_image(q, dim, t, data=l0, flags=MF.READ | MF.HOST_READ_WRITE | MF.GENERATE_MIP_MAPS)
    assert img.get_generate_mip_maps() and img.get_image_data_size() == l0.size  # level 0 only (device_image.hpp:486)
    want = oracle_mod.generate_mip_map_chain(l0, dim, t | T.WRITE, threads=4)
    assert np.array_equal(img.download_levels(q), want)
    l0b = oracle_mod.fill_synthetic(dim, t, 22)
    assert img.write(q, l0b, (0, 0, 0), (dim[0], dim[1], 1), (0, 0), (0, 0))
    assert np.array_equal(img.download_levels(q), oracle_mod.generate_mip_map_chain(l0b, dim, t | T.WRITE, threads=4))
    m = img.map(q)
    assert m.size == l0.size and np.array_equal(m, l0b)
    m[:] = l0
    assert img.unmap(q, m)
    assert np.array_equal(img.download_levels(q), want)
    assert not img.write(q, l0, (0, 0, 0), (dim[0] + 1, dim[1], 1), (0, 0), (0, 0))  # write_check failure -> False
    img.destroy()


def test_partial_write_then_regenerate(gpu_ctx, oracle_mod):
    ctx, dev, q = gpu_ctx
    dim, t = (256, 256, 3), T.IMAGE_2D_ARRAY | T.RGBA8 | M
    l0 = oracle_mod.fill_synthetic(dim, t, 31).copy()
    img = ctx.create_image(q, dim, t, data=np.concatenate([l0, np.zeros(oracle_mod.image_data_size(dim, t) - l0.size, np.uint8)]))
    patch = np.arange(64 * 32 * 4, dtype=np.uint32).astype(np.uint8)
    assert img.write(q, patch, (16, 8, 0), (64, 32, 1), (0, 0), (1, 1))
    img.generate_mip_map_chain(q)
    ref0 = l0.reshape(3, 256, 256, 4).copy()
    ref0[1, 8:40, 16:80, :] = patch.reshape(32, 64, 4)
    assert np.array_equal(img.download_levels(q), oracle_mod.generate_mip_map_chain(ref0, dim, t, threads=4))
    img.destroy()


def test_reference_golden_on_gpu(gpu_ctx, oracle_mod):
    """tests/golden/golden_ref.json: sha256 of chains computed by the REFERENCE's own Host-Compute kernels (oracle/_ref, made by
    tests/golden/make_golden_ref.py where /root/reference exists), including BASELINE configs C2 and C5 at full size, a C3 shard
    and one C4 cube.  The CUDA path must reproduce every one bit for bit; level 0 comes from the device-side fill."""
    ctx, dev, q = gpu_ctx
    with open(os.path.join(GOLDEN, "golden_ref.json")) as f:
        cases = json.load(f)
    assert len(cases) >= 30 and sum(c["heavy"] for c in cases) >= 4
    for c in cases:
        dim, t = tuple(c["dim"]), int(c["type"], 16)
        img = ctx.create_image(q, dim, t, mip_level_limit=c["mip_level_limit"], no_double=c["no_double"])
        img.fill_synthetic(q, c["config_id"])
        img.generate_mip_map_chain(q)
        got = img.download_levels(q)
        img.destroy()
        assert got.size == c["bytes"], c["name"]
        n0 = oracle_mod.level_size(dim, t, 0)
        assert hashlib.sha256(got[:n0].tobytes()).hexdigest() == c["level0_sha256"], c["name"]
        assert hashlib.sha256(got.tobytes()).hexdigest() == c["chain_sha256"], f"{c['name']}: CUDA chain differs from the reference's"
        del got


def test_cuda_path_against_reference_library(gpu_ctx, oracle_mod):
    """direct comparison with oracle/_ref (the reference's kernels compiled by oracle/build_ref.py; the .so travels with the
    repo snapshot) on seeded inputs that are in no fixture: both kernels (single-pass, general), POT and NPOT"""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref was not built (needs /root/reference)")
    rng = np.random.default_rng(99)
    fmts = [T.RGBA8, T.RGBA16F, T.R32F, T.RG16, T.RGBA8I_NORM, T.RGBA32UI, T.RG16I, T.R16F, T.RGBA32F, T.R8]
    for i, fmt in enumerate(fmts):
        for base, dim in [(T.IMAGE_2D, (512, 256)), (T.IMAGE_2D, (int(rng.integers(65, 400)), int(rng.integers(65, 400)))),
                          (T.IMAGE_2D_ARRAY, (128, 128, 5)), (T.IMAGE_3D, (64, 64, 32)),
                          (T.IMAGE_3D, tuple(int(x) for x in rng.integers(9, 70, 3)))]:
            t = base | fmt | M
            l0 = oracle_mod.fill_synthetic(dim, t, 500 + i)
            want = ref.generate_mip_map_chain(l0, dim, t, threads=8)
            got, _ = gpu_chain(gpu_ctx, l0, dim, t)
            assert_same(got, want, t, dim, "vs reference library")


def test_blit_and_clone(gpu_ctx, oracle_mod):
    """device_image::blit / clone (device_image.cpp:446-501) on linear images: every level both images have is copied on the
    device; mismatching dims / formats are refused like blit_check does"""
    ctx, dev, q = gpu_ctx
    dim, t = (256, 128, 3), T.IMAGE_2D_ARRAY | T.RGBA16F | M
    l0 = oracle_mod.fill_synthetic(dim, t, 41)
    want = oracle_mod.generate_mip_map_chain(l0, dim, t, threads=4)
    src = ctx.create_image(q, dim, t)
    src.upload_levels(q, l0, 0, 0)
    src.generate_mip_map_chain(q)
    c = src.clone(q, copy_contents=True)
    assert c.get_image_dim() == src.get_image_dim() and c.get_mip_level_count() == src.get_mip_level_count()
    assert c.device_ptr() != src.device_ptr()
    assert_same(c.download_levels(q), want, t, dim, "clone(copy_contents)")
    e = src.clone(q, copy_contents=False)
    e.zero(q)
    assert not e.download_levels(q).any()
    assert e.blit(q, src)
    assert_same(e.download_levels(q), want, t, dim, "blit")
    # a blitted level 0 regenerates to the same chain
    e.zero(q)
    other = ctx.create_image(q, (256, 128, 2), t)
    assert not other.blit(q, src)  # layer count mismatch
    other.destroy()
    other = ctx.create_image(q, dim, T.IMAGE_2D_ARRAY | T.RGBA16 | M)
    assert other.blit(q, src)  # same FORMAT_16 x 4 channels: blit_check only compares the format bits
    other.destroy()
    other = ctx.create_image(q, dim, T.IMAGE_2D_ARRAY | T.RGBA8 | M)
    assert not other.blit(q, src)
    for i in (src, c, e, other):
        i.destroy()


@pytest.mark.parametrize("base,dim,fmt", [(T.IMAGE_2D, (512, 256), T.RGBA8), (T.IMAGE_2D, (300, 200), T.RGBA16F), (T.IMAGE_2D_ARRAY, (128, 128, 5), T.RG16),
                                          (T.IMAGE_CUBE, (64, 64), T.RGBA32F), (T.IMAGE_CUBE_ARRAY, (32, 32, 2), T.R32F), (T.IMAGE_3D, (64, 32, 16), T.R32F),
                                          (T.IMAGE_3D, (20, 12, 10), T.RGBA8UI), (T.IMAGE_2D, (1024, 64), T.R16F)])
def test_tiled_interop_round_trip(gpu_ctx, oracle_mod, base, dim, fmt):
    """linear image -> CUmipmappedArray with the reference's descriptor (cuda_image.cpp:158-539) -> host, and back: the chain the
    new kernel generated arrives in floor's tiled image bit for bit, level by level, and a tiled image's level 0 can be
    pulled into a linear image to have its chain generated"""
    ctx, dev, q = gpu_ctx
    t = base | fmt | M
    l0 = oracle_mod.fill_synthetic(dim, t, 43)
    want = oracle_mod.generate_mip_map_chain(l0, dim, t, threads=4)
    img = ctx.create_image(q, dim, t)
    img.upload_levels(q, l0, 0, 0)
    arr = img.create_tiled_twin()
    img.copy_to_tiled(q, arr, 0, 0)          # the "original cuda_image" holds level 0 only
    img.zero(q)
    img.copy_from_tiled(q, arr, 0, 0)        # pull level 0 in, generate, push the generated levels back
    img.generate_mip_map_chain(q)
    if img.get_mip_level_count() > 1:
        img.copy_to_tiled(q, arr, 1, img.get_mip_level_count() - 1)
    got = img.tiled_download(q, arr)
    assert_same(got, want, t, dim, "tiled array contents")
    img.destroy_tiled(arr)
    img.destroy()


def test_concurrent_host_threads_and_all_devices(built_lib, oracle_mod):
    """the entry points are callable from any thread (each call makes the device's context current on the calling thread, like
    cuda_function.cpp:83-87), and one context serves every GPU of the box (cuda_context.cpp:80-128): 2 threads per visible
    device, each with its own queue and images, all running chains at the same time"""
    import threading
    ctx = floor_b200.device_context()
    devs = ctx.get_devices()
    cases = [((512, 256), T.IMAGE_2D | T.RGBA8 | M), ((300, 200), T.IMAGE_2D | T.RGBA16F | M), ((64, 64, 32), T.IMAGE_3D | T.R32F | M),
             ((128, 128, 4), T.IMAGE_2D_ARRAY | T.RGBA8 | M)]
    want = {}
    for i, (dim, t) in enumerate(cases):
        l0 = oracle_mod.fill_synthetic(dim, t, 400 + i)
        want[i] = (l0, oracle_mod.generate_mip_map_chain(l0, dim, t, threads=4))
    errors = []

    def worker(dev, k):
        try:
            q = ctx.create_queue(dev)
            for rep in range(6):
                i = (k + rep) % len(cases)
                dim, t = cases[i]
                img = ctx.create_image(q, dim, t)
                img.upload_levels(q, want[i][0], 0, 0)
                img.generate_mip_map_chain(q)
                got = img.download_levels(q)
                img.destroy()
                if not np.array_equal(got, want[i][1]):
                    errors.append((dev.index, k, i))
            q.destroy()
        except Exception as e:  # noqa: BLE001
            errors.append((dev.index, k, repr(e)))

    threads = [threading.Thread(target=worker, args=(d, k)) for d in devs for k in range(2)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


def test_back_to_back_chains_on_one_queue(gpu_ctx, oracle_mod):
    """stream-ordered chains without host waits in between: the single-pass and tile kernels are launched as programmatic dependents
    (their prologue may overlap the previous kernel's tail, griddepcontrol.wait guards every global access), the literal kernel
    and the fill kernel are not.  Mixed sequence on one queue, same image relaunched in between, results checked at the end."""
    ctx, dev, q = gpu_ctx
    specs = [((512, 512), T.IMAGE_2D | T.RGBA8 | M, {}), ((300, 200), T.IMAGE_2D | T.RGBA16F | M, {}), ((64, 64, 64), T.IMAGE_3D | T.R32F | M, {}),
             ((33,), T.IMAGE_1D | T.RGBA8 | M, {}), ((128, 128, 3), T.IMAGE_2D_ARRAY | T.RG16 | M, {"force_generic": True}),
             ((2048, 1024), T.IMAGE_2D | T.RGBA16F | M, {})]
    imgs, want = [], []
    for i, (dim, t, kw) in enumerate(specs):
        l0 = oracle_mod.fill_synthetic(dim, t, 700 + i)
        im = ctx.create_image(q, dim, t, **kw)
        im.upload_levels(q, l0, 0, 0)
        imgs.append(im)
        want.append(oracle_mod.generate_mip_map_chain(l0, dim, t, threads=4))
    order = [0, 1, 2, 3, 4, 5, 5, 5, 0, 0, 2, 1, 1, 4, 3, 5, 0, 2] * 6
    for i in order:
        imgs[i].enqueue_mip_map_chain(q)
    imgs[2].fill_synthetic(q, 702)      # a plain kernel in between, rewriting level 0 with the same pattern
    imgs[2].enqueue_mip_map_chain(q)
    q.finish()
    for im, w, (dim, t, kw) in zip(imgs, want, specs):
        assert_same(im.download_levels(q), w, t, dim, "back-to-back chains")
        im.destroy()


def test_batch_of_independent_textures(gpu_ctx, oracle_mod):
    """flmip_batch_*: one CUDA graph launch generates the chains of many independent textures (different sizes, formats and kernels:
    single-pass, tile with two launches, literal 1D); relaunching the batch after rewriting level 0 regenerates them"""
    ctx, dev, q = gpu_ctx
    specs = [((1024, 1024), T.IMAGE_2D | T.RGBA8 | M), ((512, 256), T.IMAGE_2D | T.RGBA16F | M), ((300, 200), T.IMAGE_2D | T.RGBA8 | M),
             ((3840, 2160), T.IMAGE_2D | T.RGBA8 | M), ((64, 64, 64), T.IMAGE_3D | T.R32F | M), ((70, 33, 18), T.IMAGE_3D | T.RG16 | M),
             ((33,), T.IMAGE_1D | T.RGBA8 | M), ((256, 256, 3), T.IMAGE_2D_ARRAY | T.RGBA32F | M), ((1, 1), T.IMAGE_2D | T.RGBA8 | M)]
    specs += [((1024, 1024), T.IMAGE_2D | T.RGBA8 | M)] * 8
    imgs = [ctx.create_image(q, dim, t) for dim, t in specs]
    batch = ctx.create_mip_chain_batch(imgs)
    assert batch.kernel_count >= len(specs) - 1  # the 1 x 1 image has nothing to generate, 1D / NPOT images need several kernels
    for cid in (800, 900):
        want = []
        for i, ((dim, t), im) in enumerate(zip(specs, imgs)):
            l0 = oracle_mod.fill_synthetic(dim, t, cid + i)
            im.upload_levels(q, l0, 0, 0, sync=False)
            want.append(oracle_mod.generate_mip_map_chain(l0, dim, t, threads=4))
        before = floor_b200.lib().flmip_launch_count()
        batch.generate(q)
        assert floor_b200.lib().flmip_launch_count() - before == batch.kernel_count
        for (dim, t), im, w in zip(specs, imgs, want):
            assert_same(im.download_levels(q), w, t, dim, "batch")
    with pytest.raises(floor_b200.FlmipError):
        ctx.create_mip_chain_batch([imgs[0], imgs[0]])
    batch.destroy()
    for im in imgs:
        im.destroy()
