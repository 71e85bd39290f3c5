"""GPU tests added in round 2 (pytest -m gpu): parity at the benchmarked layered sizes (>= 16 sampled layers / faces
per GPU), the one-chain-per-image rule, write() bounds on odd dims, externally owned image memory, and a stress of the
consumer -> finisher slot hand-off of the single-pass kernel.  Everything goes through the C-ABI; the oracle only checks."""
import ctypes
import hashlib
import threading

import numpy as np
import pytest

import floor_b200
from floor_b200.image_types import IMAGE_TYPE as T, MEMORY_FLAG as MF
from floor_b200 import image_types as it

pytestmark = pytest.mark.gpu
M = T.FLAG_MIPMAPPED | T.READ_WRITE


def sampled_layers(n_layers: int, count: int = 16, seed: int = 0x5EED):
    """first, last and seeded picks"""
    picks = {0, n_layers - 1}
    rng = np.random.default_rng(seed)
    while len(picks) < min(count, n_layers):
        picks.add(int(rng.integers(0, n_layers)))
    return sorted(picks)


def check_sampled_layers(img, q, oracle_mod, dim2d, fmt, cid, layer_id0, layers):
    """every level of the sampled layers against the oracle run per layer (device_image.cpp:304-327: layers are independent chains)"""
    t1 = T.IMAGE_2D | fmt | M
    bad = []
    for layer in layers:
        got = img.download_layers(q, layer, 1)
        l0 = oracle_mod.fill_synthetic(dim2d, t1, cid, layer_id0=layer_id0 + layer, layer_num=1)
        if not np.array_equal(got[: l0.size], l0):
            bad.append((layer, "level 0 (device fill)"))
            continue
        want = oracle_mod.generate_mip_map_chain(l0, dim2d, t1, threads=16)
        if not np.array_equal(got, want):
            first = int(np.nonzero(got != want)[0][0])
            bad.append((layer, f"first differing byte {first}"))
    return bad


def test_c3_full_size_sampled_layers(gpu_ctx, oracle_mod):
    """BASELINE config 3 at the benchmarked size: ONE launch over 2048 layers x 1024^2 RGBA8 (11.45 GB, 524 288 tiles, 2048 concurrent
    layer tails), 16 layers compared with the oracle on every level"""
    ctx, dev, q = gpu_ctx
    dim = (1024, 1024, 2048)
    t = T.IMAGE_2D_ARRAY | T.RGBA8 | M
    img = ctx.create_image(q, dim, t)
    assert img.image_data_size_mip_maps == 11453243392
    img.fill_synthetic(q, 3, layer_id0=0)
    img.generate_mip_map_chain(q)
    plan = img.plan()
    assert plan["single_pass"] and plan["launches"] == 1, plan
    layers = sampled_layers(2048)
    assert len(layers) == 16
    bad = check_sampled_layers(img, q, oracle_mod, (1024, 1024), T.RGBA8, 3, 0, layers)
    # relaunch on the same image: counters were left at zero by 2048 layer tails
    img.generate_mip_map_chain(q)
    bad += check_sampled_layers(img, q, oracle_mod, (1024, 1024), T.RGBA8, 3, 0, layers[:4])
    img.destroy()
    assert not bad, bad


def test_c4_many_cubes_sampled_faces(gpu_ctx, oracle_mod):
    """BASELINE config 4: the largest cube array this GPU holds (all 64 cubes = 384 faces = 137 GB if memory allows, at least the
    8 cubes = 48 faces one GPU owns at N = 8): every byte offset beyond the first 12 faces is above 2^32; 16 faces compared"""
    ctx, dev, q = gpu_ctx
    t = T.IMAGE_CUBE_ARRAY | T.RGBA32F | M
    img = None
    for cubes in (64, 32, 16, 8):
        try:
            img = ctx.create_image(q, (4096, 4096, cubes), t)
            break
        except floor_b200.FlmipError as e:
            assert e.code == floor_b200.ERR_OUT_OF_MEMORY, e
    assert img is not None and cubes >= 8
    faces = cubes * 6
    assert img.layer_count == faces and img.image_data_size_mip_maps == 357913936 * faces
    img.fill_synthetic(q, 4, layer_id0=0)
    img.generate_mip_map_chain(q)
    plan = img.plan()
    assert plan["single_pass"] and plan["launches"] == 1, plan
    bad = check_sampled_layers(img, q, oracle_mod, (4096, 4096), T.RGBA32F, 4, 0, sampled_layers(faces))
    img.destroy()
    assert not bad, (cubes, bad)
    print(f"C4 checked with {cubes} cubes")


def test_write_region_that_leaves_a_level_is_rejected(gpu_ctx, oracle_mod):
    """offset >> level + max(extent >> level, 1) can exceed dim >> level for odd dims; on linear memory the z axis of
    a volume has no pitch to catch it.  The call must fail and write nothing."""
    ctx, dev, q = gpu_ctx
    for base, dim, off, ext in [(T.IMAGE_3D, (5, 5, 5), (0, 0, 4), (5, 5, 1)), (T.IMAGE_2D, (5, 7), (4, 0), (1, 7)), (T.IMAGE_3D, (9, 6, 7), (0, 0, 6), (9, 6, 1))]:
        t = base | T.RGBA8 | M
        img = ctx.create_image(q, dim, t)
        before = np.arange(img.image_data_size_mip_maps, dtype=np.uint32).astype(np.uint8)
        img.upload_levels(q, before, 0, img.mip_level_count - 1)
        src = np.full(1 << 16, 0xAB, dtype=np.uint8)
        # level 0 alone is fine ...
        assert img.write(q, src, off, ext, (0, 0), (0, 0))
        # ... levels 0..1 are not: at level 1 the region starts at dim >> 1
        lvl0 = img.download_levels(q)
        assert not img.write(q, src, off, ext, (0, 1), (0, 0))
        assert np.array_equal(img.download_levels(q), lvl0), "a rejected write must not touch the image"
        img.destroy()


def test_generate_from_out_of_range_level(gpu_ctx):
    ctx, dev, q = gpu_ctx
    img = ctx.create_image(q, (64, 64), T.IMAGE_2D | T.RGBA8 | M)
    img.enqueue_mip_map_chain(q, img.mip_level_count - 1)  # last level: nothing to do
    with pytest.raises(floor_b200.FlmipError) as e:
        img.enqueue_mip_map_chain(q, img.mip_level_count)
    assert e.value.code == floor_b200.ERR_INVALID
    q.finish()
    img.destroy()


@pytest.mark.parametrize("base,dim,fmt", [(T.IMAGE_2D, (4096, 4096), T.RGBA8), (T.IMAGE_3D, (256, 256, 128), T.R32F), (T.IMAGE_2D_ARRAY, (1000, 600, 4), T.RGBA16F)])
def test_chains_on_one_image_from_two_streams_serialise(gpu_ctx, oracle_mod, base, dim, fmt):
    """two streams on one image shared the scheduler / group / layer counters.  Now a chain enqueued on another
    stream than the previous one waits for it (event hand-over): alternating streams without any host sync must still give the
    oracle's bytes, also from two host threads, also when the old stream is destroyed in between."""
    ctx, dev, q = gpu_ctx
    t = base | fmt | M
    img = ctx.create_image(q, dim, t)
    l0 = oracle_mod.fill_synthetic(dim, t, 21)
    want = oracle_mod.generate_mip_map_chain(l0, dim, t, threads=16)
    img.upload_levels(q, l0, 0, 0)
    qa, qb = ctx.create_queue(dev), ctx.create_queue(dev)
    for i in range(24):
        img.enqueue_mip_map_chain(qa if i % 2 == 0 else qb)
    qa.finish(); qb.finish()
    assert np.array_equal(img.download_levels(q), want)

    errors = []

    def worker(Q):
        try:
            for _ in range(16):
                img.enqueue_mip_map_chain(Q)
            Q.finish()
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(Q,)) for Q in (qa, qb)]
    [th.start() for th in threads]
    [th.join() for th in threads]
    assert not errors, errors
    assert np.array_equal(img.download_levels(q), want)

    # the stream of the last chain goes away while that chain may still run: the next chain on another stream waits for it
    qc = ctx.create_queue(dev)
    for _ in range(4):
        img.enqueue_mip_map_chain(qc)
    qc.destroy()
    img.enqueue_mip_map_chain(qa)
    qa.finish()
    assert np.array_equal(img.download_levels(q), want)
    img.destroy()


def test_external_image_memory_and_context_attach(gpu_ctx, oracle_mod):
    """the opt-in for other contexts: an image over caller-owned linear memory, on the caller's CUcontext"""
    ctx, dev, q = gpu_ctx
    L = floor_b200.lib()
    cu = ctypes.c_void_p()
    floor_b200.check(L.flmip_device_cu_context(dev.index, ctypes.byref(cu)))
    assert cu.value
    floor_b200.check(L.flmip_device_attach_context(dev.index, cu))       # the context in use: fine
    assert L.flmip_device_attach_context(dev.index, ctypes.c_void_p(cu.value + 64)) == floor_b200.ERR_INVALID  # another one: too late

    dim, t = (2048, 1024, 3), T.IMAGE_2D_ARRAY | T.RGBA16F | M
    owner = ctx.create_image(q, dim, t)  # stands in for the other context's allocation
    l0 = oracle_mod.fill_synthetic(dim, t, 33)
    owner.upload_levels(q, l0, 0, 0)
    h = ctypes.c_void_p()
    d4 = (ctypes.c_uint32 * 4)(*dim, 0)
    floor_b200.check(L.flmip_image_create_external(dev.index, t, d4, 0, 0, owner.device_ptr(), owner.image_data_size_mip_maps, ctypes.byref(h)))
    floor_b200.check(L.flmip_mip_chain_generate(h, q._stream))
    q.finish()
    floor_b200.check(L.flmip_image_destroy(h))  # releases the handle, not the memory
    assert np.array_equal(owner.download_levels(q), oracle_mod.generate_mip_map_chain(l0, dim, t, threads=16))
    # too small / misaligned memory is refused
    assert L.flmip_image_create_external(dev.index, t, d4, 0, 0, owner.device_ptr(), owner.image_data_size_mip_maps - 1, ctypes.byref(h)) == floor_b200.ERR_INVALID
    assert L.flmip_image_create_external(dev.index, t, d4, 0, 0, owner.device_ptr() + 4, owner.image_data_size_mip_maps, ctypes.byref(h)) == floor_b200.ERR_INVALID
    owner.destroy()


@pytest.mark.parametrize("base,dim,fmt", [(T.IMAGE_2D_ARRAY, (1024, 1024, 24), T.RGBA8), (T.IMAGE_2D, (4096, 2048), T.RGBA16F), (T.IMAGE_3D, (256, 128, 128), T.R32F),
                                          (T.IMAGE_2D_ARRAY, (512, 512, 40), T.R8), (T.IMAGE_3D, (128, 128, 128), T.RGBA32F)])
def test_slot_hand_off_stress(gpu_ctx, oracle_mod, base, dim, fmt):
    """racecheck flags the consumer -> finisher hand-off (slot_tile written by thread 0 before its arrive on slot_full); the mbarrier's
    release / acquire orders it, which the tool does not model.  Evidence instead of argument: hundreds of chains with thousands of
    tiles per CTA each (every tile passes through a cascade slot), fresh inputs every round, every result hashed against the oracle."""
    ctx, dev, q = gpu_ctx
    t = base | fmt | M
    imgs = [ctx.create_image(q, dim, t, units=u) for u in (False, True)]
    rounds = 6
    for r in range(rounds):
        l0 = oracle_mod.fill_synthetic(dim, t, 400 + r)
        want = hashlib.sha256(oracle_mod.generate_mip_map_chain(l0, dim, t, threads=16).tobytes()).digest()
        for im in imgs:
            im.upload_levels(q, l0, 0, 0)
            for _ in range(20):  # back to back: the PDL prologue of chain k + 1 overlaps the tail of chain k
                im.enqueue_mip_map_chain(q)
            assert hashlib.sha256(im.download_levels(q).tobytes()).digest() == want, (r, hex(t), dim)
    [im.destroy() for im in imgs]


def test_fastest_gpu_and_device_clock(gpu_ctx):
    ctx, dev, q = gpu_ctx
    assert dev.clock > 500 and dev.mem_bus_width >= 1024 and dev.l2_cache_size > (32 << 20)
    best = ctx.get_device("FASTEST_GPU")
    assert all(d.units * d.clock <= best.units * best.clock for d in ctx.get_devices())


ALL_FORMATS = [T.R8, T.RG8, T.RGBA8, T.R16, T.RG16, T.RGBA16, T.R8I_NORM, T.RG8I_NORM, T.RGBA8I_NORM, T.R16I_NORM, T.RG16I_NORM,
               T.RGBA16I_NORM, T.R8UI, T.RG8UI, T.RGBA8UI, T.R8I, T.RG8I, T.RGBA8I, T.R16UI, T.RG16UI, T.RGBA16UI, T.R16I, T.RG16I,
               T.RGBA16I, T.R32UI, T.RG32UI, T.RGBA32UI, T.R32I, T.RG32I, T.RGBA32I, T.R16F, T.RG16F, T.RGBA16F, T.R32F, T.RG32F,
               T.RGBA32F]


def texel2_size(n: int) -> bool:
    """sizes N with fl(fl(1/N) * N) == pred(1.0f): destination texel 0 reads source texels 0 and 2 (tests/test_npot_weights.py)"""
    f = np.float32
    return n >= 3 and f(f(1.0) / f(n)) * f(n) == np.nextafter(f(1.0), f(0.0))


def chain(gpu_ctx, l0, dim, t, **kw):
    ctx, dev, q = gpu_ctx
    img = ctx.create_image(q, dim, t, **kw)
    img.upload_levels(q, l0, 0, 0)
    img.generate_mip_map_chain(q)
    out = img.download_levels(q)
    plan = img.plan()
    img.destroy()
    return out, plan


@pytest.mark.parametrize("fmt", ALL_FORMATS)
def test_persistent_tma_tile_kernel_all_formats(gpu_ctx, oracle_mod, fmt):
    """flmip_ptile2d_*: NPOT 2D images whose rows are 16-byte multiples -- partial border tiles in x and y, odd level sizes, role
    swaps and weights != 0.5 (1080, 1000, 600 ...), layers, and the texel-2 fetch of the reference deep in the chain (2624 = 41 << 6,
    1312 = 41 << 5: inside the tile stage for narrow texels, cut short for wide ones).  Both forms: two levels per launch (the
    default for large images; the planner's size thresholds are lifted here) and the one-launch form, where the finisher pool
    reduces every tile further and the last tile of a layer finishes the chain."""
    bpp = it.bytes_per_pixel(fmt)
    w16 = 16 // bpp if bpp < 16 else 1  # texels per 16 bytes
    def wd(w):  # a width near w whose row is a 16-byte multiple and at least one tile row
        return max(-(-w // w16) * w16, 512 // bpp)
    cases = [(T.IMAGE_2D, (wd(1080), 600)), (T.IMAGE_2D, (wd(1000), 70)), (T.IMAGE_2D_ARRAY, (wd(600), 333, 3)), (T.IMAGE_2D, (wd(2624), 188)),
             (T.IMAGE_2D, (wd(1312), 200)), (T.IMAGE_CUBE, (wd(720), wd(720))), (T.IMAGE_2D, (wd(517), 64)), (T.IMAGE_2D, (wd(96), 1504))]
    for base, dim in cases:
        t = base | fmt | M
        l0 = oracle_mod.fill_synthetic(dim, t, 900 + (fmt & 0xFFFF))
        want = oracle_mod.generate_mip_map_chain(l0, dim, t, threads=8)
        for mode in ("always+nosplit", "always"):
            got, plan = chain(gpu_ctx, l0, dim, t, tma_tiles=mode)
            # (a level whose first or second destination level takes the reference's texel-2 fetch -- 2624 = 41 << 6, 1312 = 41 << 5,
            #  166 = 83 << 1 ... -- is never a TMA source: levels 1 and 2 are produced in registers; the LDG kernel serves it)
            assert not plan["single_pass"] and (plan["tma_tile_launches"] >= 1 or any(texel2_size(d >> k) for d in dim[:2] for k in (0, 1))), (dim, mode, plan)
            if not np.array_equal(got, want):
                bad = np.nonzero(got != want)[0]
                raise AssertionError(f"ptile {mode} {hex(t)} {dim} plan {plan}: {bad.size} differing bytes, first at {int(bad[0])}")
        got2, plan2 = chain(gpu_ctx, l0, dim, t, no_tma_tiles=True)
        assert plan2["tma_tile_launches"] == 0 and np.array_equal(got2, want), (dim, plan2)
        if it.bits_per_channel(t) == 16 and (t & T.FLAG_NORMALIZED):
            got, _ = chain(gpu_ctx, l0, dim, t, no_double=True, tma_tiles="always+nosplit")
            assert np.array_equal(got, oracle_mod.generate_mip_map_chain(l0, dim, t, no_double=True, threads=8)), (dim, "no_double")


def test_persistent_tma_tile_kernel_plans(gpu_ctx, oracle_mod):
    """which launches the planner picks.  Default: the TMA kernel streams two levels of a level with a long queue of tiles per resident
    CTA (threshold by texel size, never for 16-byte texels), everything else is the LDG kernel's; one-launch form on request: the
    whole chain where the post-tile level of a layer fits the patch (N1, N2 of bench.py), two launches for a larger image; never
    for rows that are no 16-byte multiple / texel-2 sizes at levels 1-2 / images smaller than a tile / volumes"""
    ctx, dev, q = gpu_ctx
    big = 12 * 2 * dev.units  # tiles that make an RGBA8 level qualify by default
    layers = -(-big // (8 * 10))  # 1000 x 600 RGBA8: 8 x 10 tiles per layer
    for base, fmt, dim, kw, launches, tma in [
            (T.IMAGE_2D, T.RGBA8, (3840, 2160), {}, 2, 0), (T.IMAGE_2D_ARRAY, T.RGBA16F, (1920, 1080, 4), {}, 2, 0),
            (T.IMAGE_2D_ARRAY, T.RGBA8, (1000, 600, layers), {}, 3, 1), (T.IMAGE_2D, T.R8, (7680, 4320), {}, None, 1),
            (T.IMAGE_2D, T.RGBA32F, (7680, 4320), {}, 2, 0),
            (T.IMAGE_2D, T.RGBA8, (3840, 2160), {"tma_tiles": "always+nosplit"}, 1, 1), (T.IMAGE_2D_ARRAY, T.RGBA16F, (1920, 1080, 4), {"tma_tiles": "always+nosplit"}, 1, 1),
            (T.IMAGE_2D, T.RGBA8, (8000, 6000), {"tma_tiles": "always+nosplit"}, 2, 1), (T.IMAGE_2D, T.RGBA8, (1002, 600), {"tma_tiles": "always"}, 2, 0),
            (T.IMAGE_2D, T.RGBA8, (164, 1000), {"tma_tiles": "always"}, 2, 0), (T.IMAGE_2D, T.RGBA8, (100, 60), {"tma_tiles": "always"}, 1, 0),
            (T.IMAGE_3D, T.R32F, (300, 200, 100), {"tma_tiles": "always"}, 2, 0), (T.IMAGE_2D, T.RGBA32F, (1001, 999), {"tma_tiles": "always+nosplit"}, 1, 1)]:
        t = base | fmt | M
        l0 = oracle_mod.fill_synthetic(dim, t, 78)
        got, plan = chain(gpu_ctx, l0, dim, t, **kw)
        assert plan["tma_tile_launches"] == tma and (launches is None or plan["launches"] == launches), (dim, kw, plan)
        assert np.array_equal(got, oracle_mod.generate_mip_map_chain(l0, dim, t, threads=16)), (dim, kw)
    # regeneration from a dirty level goes through the same planner
    dim, t = (3840, 2160), T.IMAGE_2D | T.RGBA8 | M
    img = ctx.create_image(q, dim, t, tma_tiles="always+nosplit")
    l0 = oracle_mod.fill_synthetic(dim, t, 79)
    img.upload_levels(q, l0, 0, 0)
    img.generate_mip_map_chain(q)
    full = img.download_levels(q)
    lvl2 = oracle_mod.fill_synthetic((960, 540), t, 80)
    img.upload_levels(q, lvl2, 2, 2)
    img.enqueue_mip_map_chain(q, 2)
    got = img.download_levels(q)
    o2 = img.levels[2]["offset"]
    want_tail = oracle_mod.generate_mip_map_chain(lvl2, (960, 540), t, threads=16)
    assert np.array_equal(got[:o2], full[:o2]) and np.array_equal(got[o2:], want_tail)
    img.destroy()


EXTRA_FORMATS = [T.RGB8, T.RGB16, T.RGB8I_NORM, T.RGB16I_NORM, T.RGB8UI, T.RGB8I, T.RGB16UI, T.RGB16I, T.RGB32UI, T.RGB32I, T.RGB16F, T.RGB32F,
                 T.RGBA2, T.RGBA2I_NORM, T.RG4, T.RG4I_NORM, T.RGBA4, T.RGBA4I_NORM]


@pytest.mark.parametrize("fmt", EXTRA_FORMATS)
def test_three_channel_and_packed_formats(gpu_ctx, oracle_mod, fmt):
    """3-channel images (the reference's CUDA backend cannot store them in a CUarray, its Host-Compute backend
    minifies them: host_image.hpp:1167-1210) and the FORMAT_2 / FORMAT_4 normalized formats (host_image.hpp:341-353, 419-446),
    including the signed variants with the reference's sign fix-up as written.  Linear memory lifts the CUarray restriction:
    3-channel images take the LDG tile kernel (3 / 6 / 12-byte texels) or, 1D, the literal kernel; the packed formats the literal kernel.  Every dimensionality, POT and NPOT, texel-2 sizes, level limit, both encoder modes."""
    ctx, dev, q = gpu_ctx
    for base, dim in [(T.IMAGE_2D, (256, 128)), (T.IMAGE_2D, (100, 37)), (T.IMAGE_2D_ARRAY, (33, 65, 3)), (T.IMAGE_3D, (32, 16, 16)), (T.IMAGE_3D, (12, 10, 6)),
                      (T.IMAGE_CUBE, (24, 24)), (T.IMAGE_1D, (97,)), (T.IMAGE_1D_ARRAY, (64, 2)), (T.IMAGE_2D, (41, 47)), (T.IMAGE_2D, (64, 4))]:
        t = base | fmt | M
        l0 = oracle_mod.fill_synthetic(dim, t, 1200 + (fmt & 0xFFFF))
        for kw in ({}, {"mip_level_limit": 3}):
            want = oracle_mod.generate_mip_map_chain(l0, dim, t, threads=8, **kw)
            got, plan = chain(gpu_ctx, l0, dim, t, **kw)
            assert not plan["single_pass"] and plan["tma_tile_launches"] == 0, plan
            assert np.array_equal(got, want), (hex(t), dim, kw, int(np.nonzero(got != want)[0][0]))
            if it.channel_count(t) == 3:  # 3-channel images take the LDG tile kernel (2D / 3D); the literal kernel must agree
                got, plan = chain(gpu_ctx, l0, dim, t, force_generic=True, **kw)
                assert np.array_equal(got, want), (hex(t), dim, kw, "literal kernel")
        if it.bits_per_channel(t) == 16 and (t & T.FLAG_NORMALIZED):
            got, _ = chain(gpu_ctx, l0, dim, t, no_double=True)
            assert np.array_equal(got, oracle_mod.generate_mip_map_chain(l0, dim, t, no_double=True, threads=8)), (hex(t), dim, "no_double")
    # the device-side synthetic fill defines the same bytes as the oracle's for these formats
    dim, t = (128, 64, 2), T.IMAGE_2D_ARRAY | fmt | M
    img = ctx.create_image(q, dim, t)
    img.fill_synthetic(q, 77, layer_id0=5)
    assert np.array_equal(img.download_levels(q, 0, 0), oracle_mod.fill_synthetic(dim, t, 77, layer_id0=5))
    img.destroy()
