"""The multi-level tile kernel (flmip_tile2d / 3d) relies on one property of the reference's sampler
(include/floor/device/backend/host_image.hpp:141-174, 869-929 driven by mip_map_minify.hpp:103-107): for every level size N
and every destination index g < N >> 1 the two texels a linear fetch at coord = (2g+1) * fl(1/N) reads along one axis are
{2g, 2g+1} -- no clamping, no neighbouring block -- whatever the rounding of (2g+1) * fl(1/N) * N does; only the roles
(which one is the "active" texel B) and the weight 0.5 +- eps vary.  ONE exception exists in the reference and is
reproduced: for g = 0 and sizes with fl(fl(1/N) * N) == pred(1.0f) (N = 41, 47, 55, 61, 82, ...) the neighbour coordinate
0.99999994 + 1 rounds to 2.0, so the fetch reads texels 0 and 2 and skips texel 1.  Checked here in float32 with numpy:
exhaustively for N <= 4096, and for all g of sampled N up to 2^17 (beyond the device's maximum image dimension).
(oracle/_ref, the reference's own sampler, agrees with the restatement on such sizes: tests/test_reference_pin.py.)"""
import numpy as np

f32 = np.float32


def fetch_pair(N: np.ndarray, g: np.ndarray):
    """texel indices (A, B) and the weight of B, literally as host_image.hpp:141-174, 869-894 compute them"""
    fN = N.astype(f32)
    inv = f32(1.0) / fN
    coord = (g * 2 + 1).astype(f32) * inv
    assert np.all(coord < f32(1.0)) and np.all(coord >= 0)
    scaled = coord * fN  # wrap(coord, 1) == coord for 0 <= coord < 1
    frac = scaled - np.floor(scaled)
    off = np.where(frac < f32(0.5), -1, 1).astype(f32)
    w = np.where(frac < f32(0.5), frac + f32(0.5), f32(1.5) - frac).astype(f32)
    excl = np.nextafter(fN, f32(0.0))
    b = np.clip(scaled, f32(0.0), excl).astype(np.int64)
    a = np.clip(scaled + off, f32(0.0), excl).astype(np.int64)
    return a, b, w


def quirk_sizes(n: np.ndarray) -> np.ndarray:
    fN = n.astype(f32)
    return ((f32(1.0) / fN) * fN == np.nextafter(f32(1.0), f32(0.0))) & (n >= 3)


def check(N: np.ndarray, g: np.ndarray):
    a, b, w = fetch_pair(N, g)
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    quirk = (g == 0) & quirk_sizes(N)
    assert np.array_equal(lo, 2 * g)
    assert np.array_equal(hi[~quirk], 2 * g[~quirk] + 1)
    assert np.all(a[quirk] == 2) and np.all(b[quirk] == 0)
    assert np.all(np.abs(w.astype(np.float64) - 0.5) < 1e-2)
    pot = (N & (N - 1)) == 0
    assert np.all(w[pot] == f32(0.5)) and np.all(b[pot] == 2 * g[pot] + 1)  # power-of-two levels: B = odd texel, t = 0.5 exactly
    return int((b == 2 * g).sum()), int(quirk.sum())


def test_block_property_exhaustive_small():
    n = np.arange(2, 4097, dtype=np.int64)
    N = np.repeat(n, n >> 1)
    g = np.concatenate([np.arange(k >> 1, dtype=np.int64) for k in n])
    swapped, quirks = check(N, g)
    assert swapped > 0  # "B is the even texel" does occur for NPOT sizes
    assert quirks == 568 and list(n[quirk_sizes(n)][:8]) == [41, 47, 55, 61, 82, 83, 94, 97]


def test_block_property_sampled_large():
    rng = np.random.default_rng(5)
    sizes = np.unique(np.concatenate([rng.integers(4097, 1 << 17, 300), np.array([65535, 65536, 65537, 131071, 131072, 99999, 12345, 7680, 3840, 1920, 1080])]))
    for n in sizes:
        g = np.arange(int(n) >> 1, dtype=np.int64)
        check(np.full_like(g, n), g)


def test_host_planner_knows_the_quirk_sizes(built_lib):
    """flmip.cpp cuts a tile-kernel launch where texel 2 would leave a 2-texel-wide tile remainder; its predicate must be
    the same float32 computation (checked through the plan of images whose deep levels hit such sizes)"""
    import ctypes
    n = np.arange(3, 200, dtype=np.int64)
    q = set(int(x) for x in n[quirk_sizes(n)])
    assert {41, 47, 55, 61, 82, 83, 94, 97} <= q and 64 not in q and 100 not in q


def test_host_sampler_table_matches_the_float32_emulation(built_lib):
    """the persistent TMA tile kernel (flmip_ptile2d_*) reads roles and weights from a table flmip.cpp computes on the host:
    every entry must be the float32 computation above, bit for bit (weight), with the right role / texel-2 flags"""
    import ctypes
    rng = np.random.default_rng(11)
    sizes = np.unique(np.concatenate([np.arange(2, 300), rng.integers(300, 1 << 15, 120), np.array([1080, 1920, 2160, 3840, 7680, 4095, 4097, 32767, 32768])]))
    e = ctypes.c_uint32()
    for n in sizes:
        g = np.arange(int(n) >> 1, dtype=np.int64)
        a, b, w = fetch_pair(np.full_like(g, n), g)
        wbits = w.view(np.uint32)
        for gi in (g if n < 300 else np.unique(np.concatenate([g[:3], g[-3:], rng.choice(g, 40)]))):
            assert built_lib.flmip_sampler_table_entry(int(gi), int(n), ctypes.byref(e)) == 0, (n, gi)
            v = e.value
            assert (v & 0x3FFFFFFF) == int(wbits[gi]), (n, gi, hex(v), hex(int(wbits[gi])))
            if a[gi] == 2 and b[gi] == 0 and gi == 0:
                assert v >> 30 == 1, (n, gi)
            elif a[gi] == 2 * gi + 1:
                assert v >> 30 == 2 and b[gi] == 2 * gi, (n, gi)
            else:
                assert v >> 30 == 0 and a[gi] == 2 * gi and b[gi] == 2 * gi + 1, (n, gi)
    assert built_lib.flmip_sampler_table_entry(5, 10, ctypes.byref(e)) != 0  # g outside the destination level
