"""Pins the oracle (oracle/minify_oracle.c, the C restatement) against the REFERENCE's own code: oracle/_ref holds the
reference's Host-Compute `libfloor_mip_map_minify_*` kernels, its software sampler and its IMAGE_TYPE size helpers,
compiled from /root/reference by oracle/build_ref.py.  Every comparison is bit-exact.

The known-answer vectors that SURVEY.md 8c could only *derive* from the source are checked here on the reference itself.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from floor_b200.image_types import IMAGE_TYPE as T
from floor_b200 import image_types as it

M = T.FLAG_MIPMAPPED | T.READ_WRITE
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ref_mod():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    return ref


def rchain(ref_mod, arr, dim, t, **kw):
    return ref_mod.generate_mip_map_chain(arr, dim, t | M, **kw)


# ---------------------------------------------------------------------------------------------------------------------
# known answers, on the reference itself


def test_reference_unorm8_2x2_kats(ref_mod):
    for blk, exp in [((0, 255, 255, 0), 127), ((1, 2, 3, 4), 2), ((10, 20, 30, 41), 25), ((255, 255, 255, 254), 254),
                     ((0, 0, 0, 1), 0), ((3, 3, 3, 2), 2)]:
        out = rchain(ref_mod, np.array(blk, np.uint8), (2, 2), T.IMAGE_2D | T.R8)
        assert out[4] == exp, (blk, out[4], exp)


def test_reference_constant_images(ref_mod):
    for v in range(256):
        assert rchain(ref_mod, np.full(4, v, np.uint8), (2, 2), T.IMAGE_2D | T.R8)[4] == v
    v = np.arange(65536, dtype=np.uint16)
    img = np.repeat(v, 4).reshape(-1, 2, 2).transpose(1, 0, 2).reshape(2, -1)
    out = rchain(ref_mod, img, (2 * 65536, 2), T.IMAGE_2D | T.R16, mip_level_limit=2).view(np.uint16)[4 * 65536:]
    # truncating encoder with a double scale: 33 407 of the 65 536 constant unorm16 values lose one count per level
    assert int((out != v).sum()) == 33407 and np.all((out == v) | (out == v - 1)) and out[1] == 0
    out = rchain(ref_mod, img, (2 * 65536, 2), T.IMAGE_2D | T.R16, mip_level_limit=2, no_double=True).view(np.uint16)[4 * 65536:]
    assert int((out != v).sum()) == 512 and int(np.nonzero(out != v)[0][0]) == 257
    s = np.arange(-127, 128, dtype=np.int8)
    img = np.repeat(s, 4).reshape(-1, 2, 2).transpose(1, 0, 2).reshape(2, -1)
    out = rchain(ref_mod, img, (2 * 255, 2), T.IMAGE_2D | T.R8I_NORM, mip_level_limit=2).view(np.int8)[4 * 255:]
    assert sorted(int(x) for x in s[out != s]) == [-104, -72, -52, -36, -26, -18, -13, -9, 9, 13, 18, 26, 36, 52, 72, 104]


def test_reference_integer_quirks(ref_mod):
    o = lambda a, t, dt: rchain(ref_mod, np.array(a, dt), (2, 2), T.IMAGE_2D | t).view(dt)[4]
    assert o([1, 2, 3, 4], T.R32I, np.int32) == 2
    assert o([4, 3, 2, 1], T.R32I, np.int32) == 3
    assert o([4, 3, 2, 1], T.R32UI, np.uint32) == 4  # unsigned (b - a) wraps around
    assert o([4, 3, 2, 1], T.R8UI, np.uint8) == 4
    assert o([-4, -3, -2, -1], T.R8I, np.int8) == -3


# ---------------------------------------------------------------------------------------------------------------------
# image geometry: image_types.hpp helpers of the reference vs the restatement


GEOMETRY = [(T.IMAGE_2D | T.RGBA8, (1024, 1024)), (T.IMAGE_2D | T.RGBA16F, (8192, 8192)), (T.IMAGE_2D_ARRAY | T.RGBA8, (1024, 1024, 2048)),
            (T.IMAGE_CUBE_ARRAY | T.RGBA32F, (4096, 4096, 64)), (T.IMAGE_3D | T.R32F, (512, 512, 512)), (T.IMAGE_2D | T.R8, (8, 2)),
            (T.IMAGE_2D | T.RG16F, (100, 37)), (T.IMAGE_2D | T.RGB8, (33, 65)), (T.IMAGE_3D | T.RGBA16, (12, 10, 6)),
            (T.IMAGE_3D | T.R8UI, (5, 64, 3)), (T.IMAGE_1D | T.R32F, (33,)), (T.IMAGE_1D_ARRAY | T.RG8, (16, 2)),
            (T.IMAGE_CUBE | T.RGBA8, (8, 8)), (T.IMAGE_2D_ARRAY | T.R16UI, (7, 300, 5)), (T.IMAGE_2D | T.R8, (1, 1)),
            (T.IMAGE_2D | T.R8, (65535, 3)), (T.IMAGE_3D | T.R32F, (2048, 2, 2))]


@pytest.mark.parametrize("t,dim", GEOMETRY)
def test_geometry_matches_reference(ref_mod, oracle_mod, t, dim):
    import ctypes
    R, t = ref_mod.lib(), t | M
    d = (ctypes.c_uint32 * 4)(*(list(dim) + [0] * (4 - len(dim))))
    assert R.flr_bytes_per_pixel(t) == oracle_mod.bytes_per_pixel(t) == it.bytes_per_pixel(t)
    assert R.flr_layer_count(d, t) == oracle_mod.layer_count(dim, t)
    n = R.flr_mip_level_count(d, t)
    assert n == oracle_mod.mip_level_count(dim, t)
    for limit in (0, 1, 3, n, n + 5):
        assert R.flr_effective_level_count(d, t, limit) == oracle_mod.mip_level_count(dim, t, limit)
        assert R.flr_image_data_size(d, t, limit) == oracle_mod.image_data_size(dim, t, limit)
    for level in range(n):
        assert R.flr_level_offset(d, t, level) == oracle_mod.level_offset(dim, t, level), level
        assert R.flr_level_data_size(d, t, level) == oracle_mod.level_size(dim, t, level), level


def test_kernel_selection_key(ref_mod):
    """minify_image_base_type (mip_map_minify.hpp:53-69): normalized formats sample as FLOAT, the key keeps dim + array"""
    R = ref_mod.lib()
    assert R.flr_minify_base_type(T.IMAGE_2D | T.RGBA8 | M) == (T.IMAGE_2D | T.FLOAT)
    assert R.flr_minify_base_type(T.IMAGE_2D_ARRAY | T.RG16UI | M) == (T.IMAGE_2D_ARRAY | T.UINT)
    assert R.flr_minify_base_type(T.IMAGE_3D | T.R32I | M) == (T.IMAGE_3D | T.INT)
    assert R.flr_minify_base_type(T.IMAGE_3D | T.RGBA8I_NORM | M) == (T.IMAGE_3D | T.FLOAT)


# ---------------------------------------------------------------------------------------------------------------------
# chains: restatement == reference, bit for bit


FORMATS = [T.R8, T.RG8, T.RGB8, T.RGBA8, T.R16, T.RG16, T.RGBA16, T.R8I_NORM, T.RG8I_NORM, T.RGBA8I_NORM, T.R16I_NORM,
           T.RG16I_NORM, T.RGBA16I_NORM, T.R8UI, T.RG8UI, T.RGBA8UI, T.R8I, T.RG8I, T.RGBA8I, T.R16UI, T.RG16UI, T.RGBA16UI,
           T.R16I, T.RG16I, T.RGBA16I, T.R32UI, T.RG32UI, T.RGBA32UI, T.R32I, T.RG32I, T.RGBA32I,
           T.R16F, T.RG16F, T.RGB16F, T.RGBA16F, T.R32F, T.RG32F, T.RGB32F, T.RGBA32F,
           # the remaining 3-channel formats and the FORMAT_2 / FORMAT_4 normalized formats whose texel is a whole number of bytes
           T.RGB16, T.RGB8I_NORM, T.RGB16I_NORM, T.RGB8UI, T.RGB8I, T.RGB16UI, T.RGB16I, T.RGB32UI, T.RGB32I,
           T.RGBA2, T.RGBA2I_NORM, T.RG4, T.RG4I_NORM, T.RGBA4, T.RGBA4I_NORM]
SHAPES = [(T.IMAGE_2D, (64, 64)), (T.IMAGE_2D, (20, 12)), (T.IMAGE_2D, (37, 5)), (T.IMAGE_2D, (129, 67)), (T.IMAGE_2D, (64, 4)),
          (T.IMAGE_2D_ARRAY, (16, 8, 3)), (T.IMAGE_2D_ARRAY, (11, 23, 2)), (T.IMAGE_3D, (16, 16, 16)), (T.IMAGE_3D, (12, 10, 6)),
          (T.IMAGE_3D, (7, 33, 5)), (T.IMAGE_CUBE, (8, 8)), (T.IMAGE_CUBE_ARRAY, (6, 6, 2)),
          # sizes where the reference's linear fetch for destination texel 0 reads texels 0 and 2 (tests/test_npot_weights.py)
          (T.IMAGE_2D, (41, 47)), (T.IMAGE_2D, (83, 164)), (T.IMAGE_3D, (55, 61, 41)), (T.IMAGE_2D_ARRAY, (97, 94, 2)),
          (T.IMAGE_1D, (33,)), (T.IMAGE_1D, (256,)), (T.IMAGE_1D, (41,)), (T.IMAGE_1D_ARRAY, (64, 2)), (T.IMAGE_1D_ARRAY, (97, 3))]


@pytest.mark.parametrize("fmt", FORMATS)
def test_oracle_matches_reference(ref_mod, oracle_mod, fmt):
    for base, dim in SHAPES:
        t = base | fmt | M
        l0 = oracle_mod.fill_synthetic(dim, t, (fmt & 0xFFFF) + len(dim))
        a = oracle_mod.generate_mip_map_chain(l0, dim, t)
        b = ref_mod.generate_mip_map_chain(l0, dim, t)
        assert a.size == b.size, (hex(t), dim)
        assert np.array_equal(a, b), (hex(t), dim, int(np.nonzero(a != b)[0][0]))
        if it.bits_per_channel(t) == 16 and (t & T.FLAG_NORMALIZED):
            a = oracle_mod.generate_mip_map_chain(l0, dim, t, no_double=True)
            b = ref_mod.generate_mip_map_chain(l0, dim, t, no_double=True)
            assert np.array_equal(a, b), (hex(t), dim, "no_double")


def test_layout_and_srgb_flags_in_the_reference(ref_mod, oracle_mod):
    """channel layouts (BGRA, ABGR, ARGB) and FLAG_SRGB leave the reference's minify arithmetic untouched"""
    dim, base = (100, 60), T.IMAGE_2D | T.RGBA8
    l0 = oracle_mod.fill_synthetic(dim, base | M, 91)
    plain = ref_mod.generate_mip_map_chain(l0, dim, base | M)
    for extra in (T.LAYOUT_BGRA, T.LAYOUT_ABGR, T.LAYOUT_ARGB, T.FLAG_SRGB):
        t = base | extra | M
        assert np.array_equal(ref_mod.generate_mip_map_chain(l0, dim, t), plain), hex(t)
        assert np.array_equal(oracle_mod.generate_mip_map_chain(l0, dim, t), plain), hex(t)


def test_depth_images_match_reference(ref_mod, oracle_mod):
    """libfloor_mip_map_minify_IMAGE_DEPTH[_ARRAY]_FLOAT (mip_map_minify.hpp:22-30): D32F, single channel"""
    for dim, t in [((256, 256), T.D32F | M), ((100, 37), T.D32F | M), ((41, 47), T.D32F | M),
                   ((64, 32, 3), T.IMAGE_DEPTH_ARRAY | T.FORMAT_32 | T.FLOAT | M)]:
        l0 = oracle_mod.fill_synthetic(dim, t, 27)
        assert np.array_equal(oracle_mod.generate_mip_map_chain(l0, dim, t), ref_mod.generate_mip_map_chain(l0, dim, t)), (hex(t), dim)


def test_random_geometry_fuzz(ref_mod, oracle_mod):
    rng = np.random.default_rng(20261017)
    for i in range(160):
        fmt = FORMATS[int(rng.integers(len(FORMATS)))]
        kind = int(rng.integers(5))
        if kind == 3:
            base, dim = T.IMAGE_1D, (int(rng.integers(1, 600)),)
        elif kind == 4:
            base, dim = T.IMAGE_1D_ARRAY, (int(rng.integers(1, 200)), int(rng.integers(1, 4)))
        elif kind == 0:
            base, dim = T.IMAGE_2D, tuple(int(x) for x in rng.integers(1, 90, 2))
        elif kind == 1:
            base, dim = T.IMAGE_2D_ARRAY, tuple(int(x) for x in rng.integers(1, 50, 2)) + (int(rng.integers(1, 4)),)
        else:
            base, dim = T.IMAGE_3D, tuple(int(x) for x in rng.integers(1, 24, 3))
        t = base | fmt | M
        limit = int(rng.integers(0, 5))
        l0 = oracle_mod.fill_synthetic(dim, t, 1000 + i)
        a = oracle_mod.generate_mip_map_chain(l0, dim, t, mip_level_limit=limit)
        b = ref_mod.generate_mip_map_chain(l0, dim, t, mip_level_limit=limit)
        assert a.size == b.size and np.array_equal(a, b), (hex(t), dim, limit)


def test_adversarial_half_and_float(ref_mod, oracle_mod):
    """every finite fp16 bit pattern next to shuffled neighbours (+-0, subnormals, +-65504, 1-ulp pairs); fp32 bit patterns
    with mixed exponents (cancellation in b - a) and subnormals.  No NaN / Inf (undefined under the reference's fast-math)."""
    rng = np.random.default_rng(7)
    h = np.arange(65536, dtype=np.uint16)
    h = h[np.isfinite(h.view(np.float16))]
    cols = [h, np.roll(h, 1), rng.permutation(h), h ^ np.uint16(1)]
    cols[3] = np.where(np.isfinite(cols[3].view(np.float16)), cols[3], h)
    n = (h.size // 64) * 64
    img = np.stack([c[:n] for c in cols], axis=1).reshape(-1)  # RGBA16F texels
    dim = (64, n // 64)
    t = T.IMAGE_2D | T.RGBA16F | M
    assert np.array_equal(oracle_mod.generate_mip_map_chain(img, dim, t), ref_mod.generate_mip_map_chain(img, dim, t))
    bits = rng.integers(0, 1 << 32, 128 * 128 * 2, dtype=np.uint64).astype(np.uint32)
    f = bits.view(np.float32)
    bits = np.where(np.isfinite(f), bits, bits & np.uint32(0x3FFFFFFF))
    bits[::17] &= np.uint32(0x807FFFFF)  # subnormals
    bits[::19] = bits[1::19][: bits[::19].size] ^ np.uint32(1)  # 1-ulp neighbours
    for t, dim in [(T.IMAGE_2D | T.RG32F | M, (128, 128)), (T.IMAGE_3D | T.R32F | M, (32, 32, 32))]:
        a = oracle_mod.generate_mip_map_chain(bits.view(np.uint8), dim, t)
        b = ref_mod.generate_mip_map_chain(bits.view(np.uint8), dim, t)
        assert np.array_equal(a, b), hex(t)


def test_subnormal_ties_in_the_reference(ref_mod, oracle_mod):
    """(b - a) * 0.5 as an inexact subnormal: product and sum are rounded separately (a = 2^-149, b = 2^-148 -> 2^-149)"""
    out = rchain(ref_mod, np.array([1, 2, 1, 2], np.uint32), (2, 2), T.IMAGE_2D | T.R32F).view(np.uint32)[4]
    assert out == 1
    rng = np.random.default_rng(12)
    small = np.array([0, 1, 2, 3, 4, 5, 6, 7, 9, 11, 0x80000001, 0x80000002, 0x80000003, 0x80000005, 0x00800000, 0x00800001, 0x007FFFFF], np.uint32)
    for base, fmt, dim in [(T.IMAGE_2D, T.RGBA32F, (64, 64)), (T.IMAGE_3D, T.R32F, (16, 16, 16)), (T.IMAGE_2D, T.RG32F, (50, 30))]:
        t = base | fmt | M
        f = small[rng.integers(0, small.size, size=int(np.prod(dim)) * it.channel_count(t))]
        assert np.array_equal(oracle_mod.generate_mip_map_chain(f, dim, t), ref_mod.generate_mip_map_chain(f, dim, t)), hex(t)


def test_timing_build_computes_the_same_bytes(ref_mod, oracle_mod):
    """bench.py's reference arm times the -O3 -march=corei7-avx -mf16c build of the reference (hardware half conversions);
    it must produce what the strict build produces"""
    for dim, t in [((256, 256), T.IMAGE_2D | T.RGBA8 | M), ((256, 256), T.IMAGE_2D | T.RGBA16F | M), ((300, 200), T.IMAGE_2D | T.RGBA16 | M),
                   ((32, 32, 32), T.IMAGE_3D | T.R32F | M), ((64, 64, 3), T.IMAGE_2D_ARRAY | T.RGBA32UI | M), ((128, 128), T.IMAGE_2D | T.RG8I_NORM | M)]:
        l0 = oracle_mod.fill_synthetic(dim, t, 77)
        want = ref_mod.generate_mip_map_chain(l0, dim, t)
        buf = np.zeros(want.size + 64, np.uint8)
        buf[: l0.size] = l0
        ref_mod.generate_in_place(buf, dim, t, threads=2, fast=True)
        assert np.array_equal(buf[: want.size], want), (hex(t), dim)


def test_baseline_config_c1_matches_reference(ref_mod, oracle_mod):
    """BASELINE configs[0] at full size: 1024x1024 RGBA8 UNORM, all 11 levels"""
    dim, t = (1024, 1024), T.IMAGE_2D | T.RGBA8 | M
    l0 = oracle_mod.fill_synthetic(dim, t, 1)
    assert np.array_equal(oracle_mod.generate_mip_map_chain(l0, dim, t, threads=4), ref_mod.generate_mip_map_chain(l0, dim, t, threads=4))


def test_scaled_baseline_configs_match_reference(ref_mod, oracle_mod):
    """C2 / C3 / C4 / C5 shapes at a size the CPU finishes in seconds (full sizes run on the GPU box, tests/test_gpu_parity.py)"""
    for cid, dim, t in [(2, (2048, 2048), T.IMAGE_2D | T.RGBA16F | M), (3, (1024, 1024, 4), T.IMAGE_2D_ARRAY | T.RGBA8 | M),
                        (4, (256, 256, 2), T.IMAGE_CUBE_ARRAY | T.RGBA32F | M), (5, (128, 128, 128), T.IMAGE_3D | T.R32F | M)]:
        l0 = oracle_mod.fill_synthetic(dim, t, cid)
        a = oracle_mod.generate_mip_map_chain(l0, dim, t, threads=4)
        b = ref_mod.generate_mip_map_chain(l0, dim, t, threads=4)
        assert np.array_equal(a, b), (cid, dim)


def test_reference_golden_fixtures(ref_mod, oracle_mod):
    """tests/golden/golden_ref.json holds sha256 of chains computed BY THE REFERENCE (tests/golden/make_golden_ref.py);
    the GPU parity tests check the CUDA path against the same file, so they stay pinned where /root/reference is absent"""
    with open(os.path.join(GOLDEN, "golden_ref.json")) as f:
        cases = json.load(f)
    assert len(cases) >= 30
    for c in cases:
        if c["heavy"] and not os.environ.get("FLMIP_HEAVY"):
            continue  # full-size BASELINE configs: verified when the file is made, and on the GPU box against the CUDA path
        dim, t = tuple(c["dim"]), int(c["type"], 16)
        l0 = oracle_mod.fill_synthetic(dim, t, c["config_id"])
        assert hashlib.sha256(l0.tobytes()).hexdigest() == c["level0_sha256"]
        kw = dict(mip_level_limit=c["mip_level_limit"], no_double=c["no_double"])
        r = ref_mod.generate_mip_map_chain(l0, dim, t, threads=4, **kw)
        assert hashlib.sha256(r.tobytes()).hexdigest() == c["chain_sha256"], c["name"]
        o = oracle_mod.generate_mip_map_chain(l0, dim, t, threads=4, **kw)
        assert hashlib.sha256(o.tobytes()).hexdigest() == c["chain_sha256"], c["name"]


def test_sub_byte_texels_are_inconsistent_in_the_reference(ref_mod, oracle_mod):
    """FORMAT_2 / FORMAT_4 formats whose texel is NOT a whole number of bytes (R2, RG2, RGB2, R4, RGB4): the reference sizes the
    image by bits (image_types.hpp:675-691) but its kernels address texels by bytes_per_pixel = ceil(bits / 8)
    (host_image.hpp:235-271), so the image is smaller than what the kernels touch.  This path rejects them (flmip.cpp decode_type,
    oracle image_init); the whole-byte ones (RGBA2, RG4, RGBA4) are supported and pinned above."""
    import ctypes
    L = ref_mod.lib()
    d = (ctypes.c_uint32 * 4)(64, 64, 0, 0)
    for ch, fmt in [(T.CHANNELS_1, T.FORMAT_2), (T.CHANNELS_2, T.FORMAT_2), (T.CHANNELS_3, T.FORMAT_2), (T.CHANNELS_1, T.FORMAT_4), (T.CHANNELS_3, T.FORMAT_4)]:
        t = T.IMAGE_2D | ch | fmt | T.UINT | T.FLAG_NORMALIZED | M
        assert L.flr_level_data_size(d, t, 0) < 64 * 64 * L.flr_bytes_per_pixel(t), hex(t)
        with pytest.raises(RuntimeError):
            oracle_mod.generate_mip_map_chain(np.zeros(64 * 64 * 2, np.uint8), (64, 64), t)
    for fmt in (T.RGBA2, T.RG4, T.RGBA4):
        t = T.IMAGE_2D | fmt | M
        assert L.flr_level_data_size(d, t, 0) == 64 * 64 * L.flr_bytes_per_pixel(t) == oracle_mod.level_size((64, 64), t, 0), hex(t)
