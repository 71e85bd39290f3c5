"""CPU tests of the oracle: survey KATs, cross-check against the independent numpy emulation, golden fixtures."""
import hashlib
import json
import os

import numpy as np
import pytest

from floor_b200.image_types import IMAGE_TYPE as T
from floor_b200 import image_types as it
import np_emulation as npe

M = T.FLAG_MIPMAPPED | T.READ_WRITE
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def chain(oracle, arr, dim, t, **kw):
    return oracle.generate_mip_map_chain(arr, dim, t | M, **kw)


def test_unorm8_2x2_kats(oracle_mod):
    # derived in SURVEY.md section 8c with an independent float32 emulation
    for blk, exp in [((0, 255, 255, 0), 127), ((1, 2, 3, 4), 2), ((10, 20, 30, 41), 25), ((255, 255, 255, 254), 254),
                     ((0, 0, 0, 1), 0), ((3, 3, 3, 2), 2)]:
        out = chain(oracle_mod, np.array(blk, np.uint8), (2, 2), T.IMAGE_2D | T.R8)
        assert out[4] == exp, (blk, out[4], exp)


def test_constant_images(oracle_mod):
    for v in range(256):
        assert chain(oracle_mod, np.full(4, v, np.uint8), (2, 2), T.IMAGE_2D | T.R8)[4] == v
    v = np.arange(65536, dtype=np.uint16)
    img = np.repeat(v, 4).reshape(-1, 2, 2).transpose(1, 0, 2).reshape(2, -1)  # 65536 constant 2x2 blocks side by side
    out = chain(oracle_mod, img, (2 * 65536, 2), T.IMAGE_2D | T.R16, mip_level_limit=2).view(np.uint16)[4 * 65536:]
    assert int((out != v).sum()) == 33407 and np.all((out == v) | (out == v - 1)) and out[1] == 0
    out = chain(oracle_mod, img, (2 * 65536, 2), T.IMAGE_2D | T.R16, mip_level_limit=2, no_double=True).view(np.uint16)[4 * 65536:]
    assert int((out != v).sum()) == 512 and int(np.nonzero(out != v)[0][0]) == 257
    s = np.arange(-127, 128, dtype=np.int8)
    img = np.repeat(s, 4).reshape(-1, 2, 2).transpose(1, 0, 2).reshape(2, -1)
    out = chain(oracle_mod, img, (2 * 255, 2), T.IMAGE_2D | T.R8I_NORM, mip_level_limit=2).view(np.int8)[4 * 255:]
    assert sorted(int(x) for x in s[out != s]) == [-104, -72, -52, -36, -26, -18, -13, -9, 9, 13, 18, 26, 36, 52, 72, 104]


def test_integer_quirks(oracle_mod):
    o = lambda a, t, dt: chain(oracle_mod, np.array(a, dt), (2, 2), T.IMAGE_2D | t).view(dt)[4]
    assert o([1, 2, 3, 4], T.R32I, np.int32) == 2
    assert o([4, 3, 2, 1], T.R32I, np.int32) == 3      # truncation toward zero is order dependent
    assert o([4, 3, 2, 1], T.R32UI, np.uint32) == 4    # unsigned (b - a) wraps
    assert o([4, 3, 2, 1], T.R8UI, np.uint8) == 4
    assert o([-4, -3, -2, -1], T.R8I, np.int8) == -3


def test_half_conversion_matches_numpy(oracle_mod):
    L = oracle_mod.lib()
    h = np.arange(65536, dtype=np.uint16)
    f = h.view(np.float16).astype(np.float32)
    finite = np.isfinite(f)
    got = np.array([L.flo_half_to_float(int(x)) for x in h[finite][::7]], dtype=np.float32)
    assert np.array_equal(got.view(np.uint32), f[finite][::7].view(np.uint32))
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.standard_normal(4000).astype(np.float32) * np.float32(10.0) ** rng.integers(-9, 6, 4000).astype(np.float32),
                        np.array([0.0, -0.0, 65504.0, 65519.9, 65520.0, 6e-8, 2.98e-8, 2.9802322e-8, 5.96e-8, 6.1e-5, 1e-30], np.float32)])
    with np.errstate(over="ignore"):
        want = x.astype(np.float16).view(np.uint16)
    got = np.array([L.flo_float_to_half(float(v)) for v in x], dtype=np.uint16)
    assert np.array_equal(got, want)


FORMATS = [T.R8, T.RG8, T.RGBA8, T.R16, T.RGBA16, T.R8I_NORM, T.RGBA8I_NORM, T.RG16I_NORM, T.RGBA16I_NORM,
           T.R8UI, T.RGBA8UI, T.RG8I, T.R16UI, T.RGBA16I, T.R32UI, T.RG32I, T.RGBA32UI,
           T.R16F, T.RG16F, T.RGBA16F, T.R32F, T.RG32F, T.RGBA32F, T.RGB8, T.RGB16F, T.RGB32F]
SHAPES = [(T.IMAGE_2D, (16, 16)), (T.IMAGE_2D, (20, 12)), (T.IMAGE_2D, (37, 5)), (T.IMAGE_2D_ARRAY, (8, 8, 3)),
          (T.IMAGE_3D, (8, 8, 8)), (T.IMAGE_3D, (12, 10, 6)), (T.IMAGE_1D, (33,)), (T.IMAGE_1D_ARRAY, (16, 2)),
          (T.IMAGE_CUBE, (8, 8)), (T.IMAGE_CUBE_ARRAY, (4, 4, 2)), (T.IMAGE_2D, (64, 4))]


@pytest.mark.parametrize("fmt", FORMATS)
def test_oracle_matches_numpy_emulation(oracle_mod, fmt):
    for base, dim in SHAPES:
        t = base | fmt | M
        cid = (fmt & 0xFFFF) + len(dim)
        l0 = oracle_mod.fill_synthetic(dim, t, cid)
        a = oracle_mod.generate_mip_map_chain(l0, dim, t)
        b = npe.generate_chain(l0, dim, t)
        assert a.size == b.size, (hex(t), dim)
        assert np.array_equal(a, b), (hex(t), dim, int(np.nonzero(a != b)[0][0]))
        if it.bits_per_channel(t) == 16 and (t & T.FLAG_NORMALIZED):
            a = oracle_mod.generate_mip_map_chain(l0, dim, t, no_double=True)
            b = npe.generate_chain(l0, dim, t, no_double=True)
            assert np.array_equal(a, b)


def test_threads_and_level_limit(oracle_mod):
    t = T.IMAGE_2D_ARRAY | T.RGBA8 | M
    dim = (128, 64, 3)
    l0 = oracle_mod.fill_synthetic(dim, t, 5)
    a = oracle_mod.generate_mip_map_chain(l0, dim, t, threads=1)
    b = oracle_mod.generate_mip_map_chain(l0, dim, t, threads=4)
    assert np.array_equal(a, b)
    c = oracle_mod.generate_mip_map_chain(l0, dim, t, mip_level_limit=3)
    assert oracle_mod.mip_level_count(dim, t, 3) == 3 and np.array_equal(c, a[: c.size])


def test_layout_matches_reference_numbers(oracle_mod):
    # SURVEY.md section 8d per-config byte counts
    c1 = ((1024, 1024), T.IMAGE_2D | T.RGBA8 | M)
    assert oracle_mod.mip_level_count(*c1) == 11 and oracle_mod.image_data_size(*c1) == 5592404
    c2 = ((8192, 8192), T.IMAGE_2D | T.RGBA16F | M)
    assert oracle_mod.mip_level_count(*c2) == 14 and oracle_mod.image_data_size(*c2) == 715827880
    assert oracle_mod.level_offset(*c2, 1) == 512 << 20
    c5 = ((512, 512, 512), T.IMAGE_3D | T.R32F | M)
    assert oracle_mod.mip_level_count(*c5) == 10 and oracle_mod.image_data_size(*c5) == 613566756
    c4 = ((4096, 4096, 64), T.IMAGE_CUBE_ARRAY | T.RGBA32F | M)
    assert oracle_mod.layer_count(*c4) == 384 and oracle_mod.image_data_size(*c4) == 137438951424
    # zero-dim quirk: 8x2 has 4 levels, the last two are empty (image_types.hpp:751-766)
    q = ((8, 2), T.IMAGE_2D | T.R8 | M)
    assert oracle_mod.mip_level_count(*q) == 4
    assert [oracle_mod.level_size(*q, l) for l in range(4)] == [16, 4, 0, 0]


def test_golden_fixtures(oracle_mod):
    """frozen oracle outputs (tests/golden/make_golden.py): any change of the restated arithmetic shows up here"""
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        cases = json.load(f)
    assert len(cases) >= 20
    for c in cases:
        dim, t = tuple(c["dim"]), int(c["type"], 16)
        l0 = oracle_mod.fill_synthetic(dim, t, c["config_id"])
        assert hashlib.sha256(l0.tobytes()).hexdigest() == c["level0_sha256"]
        out = oracle_mod.generate_mip_map_chain(l0, dim, t, no_double=c["no_double"])
        assert hashlib.sha256(out.tobytes()).hexdigest() == c["chain_sha256"], c["name"]
