"""Independent numpy-float32 emulation of the Host-Compute minify arithmetic.

Written from the specification, not from oracle/minify_oracle.c, so that the two restatements can be
cross-checked.  Vectorised over all destination texels of a level; supports NPOT levels by replaying the
coordinate arithmetic (reference: include/floor/device/backend/host_image.hpp:141-174, 842-929).
"""
from __future__ import annotations

import numpy as np

from floor_b200.image_types import IMAGE_TYPE as T
from floor_b200 import image_types as it

f32 = np.float32


def _storage_dtype(t: int):
    bpc = it.bits_per_channel(t)
    dt = t & T.DATA_TYPE_MASK
    if dt == T.FLOAT:
        return {16: np.float16, 32: np.float32}[bpc]
    if dt == T.INT:
        return {8: np.int8, 16: np.int16, 32: np.int32}[bpc]
    return {8: np.uint8, 16: np.uint16, 32: np.uint32}[bpc]


def level_dims(dim, t: int, level: int):
    dc = it.dim_count(t)
    return tuple((dim[d] >> level) if d < dc else 0 for d in range(3))


def layer_count(dim, t: int) -> int:
    dc = it.dim_count(t)
    n = 1 if not (t & T.FLAG_ARRAY) else (dim[1] if dc == 1 else dim[2] if dc == 2 else dim[3])
    return n * 6 if t & T.FLAG_CUBE else n


def level_count(dim, t: int, limit: int = 0) -> int:
    if not (t & T.FLAG_MIPMAPPED):
        return 1
    dc = it.dim_count(t)
    m = max(dim[:dc])
    n = 1 if m <= 1 else int(m).bit_length()
    return min(n, limit) if limit else n


def _lerp_f(a, b, w):
    with np.errstate(over="ignore", invalid="ignore"):
        return ((b - a).astype(f32) * w).astype(f32) + a


def _lerp_i(a, b, w, signed: bool):
    # T(float(b - a) * t) + a, with (b - a) evaluated in T (uint32 wraps)
    if signed:
        d = (b.astype(np.int64) - a.astype(np.int64))
        d = ((d + 2**31) % 2**32 - 2**31).astype(np.int32)
        s = (d.astype(f32) * w).astype(f32)
        back = np.trunc(s).astype(np.int64)
        return ((back + a.astype(np.int64) + 2**31) % 2**32 - 2**31).astype(np.int32)
    d = (b.astype(np.int64) - a.astype(np.int64)) % 2**32
    s = (d.astype(np.uint32).astype(f32) * w).astype(f32)
    back = np.trunc(s.astype(np.float64)).astype(np.int64) % 2**32
    return ((back + a.astype(np.int64)) % 2**32).astype(np.uint32)


def minify_level(src: np.ndarray, t: int, no_double: bool = False) -> np.ndarray:
    """src: [layers][z][y][x][c] (missing dims have size 1) in storage dtype -> next level, same layout."""
    dc = it.dim_count(t)
    bpc = it.bits_per_channel(t)
    dt = t & T.DATA_TYPE_MASK
    norm = bool(t & T.FLAG_NORMALIZED)
    L, Dz, Hy, Wx, C = src.shape
    sdim = (Wx, Hy, Dz)
    ddim = tuple((sdim[d] >> 1) if d < dc else 1 for d in range(3))
    if any(ddim[d] == 0 for d in range(dc)):
        return np.zeros((L, 0, 0, 0, C), dtype=src.dtype)
    # decode
    if dt == T.FLOAT:
        val = src.astype(f32)
        kind = "f"
    elif norm:
        scale = f32(1.0 / float((1 << bpc) - 1)) if dt == T.UINT else f32(1.0 / float(((1 << bpc) - 1) >> 1))
        val = src.astype(f32) * scale
        kind = "f"
    else:
        val = src.astype(np.int32 if dt == T.INT else np.uint32)
        kind = "i" if dt == T.INT else "u"
    # per-axis sampling positions and weights
    idx_a, idx_b, wts = [], [], []
    for d in range(3):
        if d >= dc:
            idx_a.append(np.zeros(1, np.int64)); idx_b.append(np.zeros(1, np.int64)); wts.append(None)
            continue
        n_src = sdim[d]
        g = np.arange(ddim[d], dtype=np.uint32)
        inv = f32(1.0) / f32(n_src)
        coord = (g * np.uint32(2) + np.uint32(1)).astype(f32) * inv
        fdim = f32(n_src)
        scaled = np.fmod(coord, f32(1.0)).astype(f32) * fdim
        frac = (scaled - np.floor(scaled)).astype(f32)
        lo = frac < f32(0.5)
        off = np.where(lo, -1, 1).astype(f32)
        w = np.where(lo, frac + f32(0.5), f32(1.5) - frac).astype(f32)
        excl = np.nextafter(fdim, f32(0.0))
        act = np.clip((coord * fdim).astype(f32), f32(0.0), excl).astype(np.int64)
        outside = np.clip(((coord * fdim).astype(f32) + off).astype(f32), f32(0.0), excl).astype(np.int64)
        idx_a.append(outside); idx_b.append(act); wts.append(w)
    def fetch(ix, iy, iz):
        return val[:, iz[:, None, None], iy[None, :, None], ix[None, None, :], :]
    def lerp(a, b, w):
        if kind == "f":
            return _lerp_f(a, b, w)
        return _lerp_i(a, b, w, kind == "i")
    wx = wts[0][None, None, None, :, None]
    res = []
    for iz in ((idx_a[2], idx_b[2]) if dc >= 3 else (idx_a[2],)):
        rows = []
        for iy in ((idx_a[1], idx_b[1]) if dc >= 2 else (idx_a[1],)):
            rows.append(lerp(fetch(idx_a[0], iy, iz), fetch(idx_b[0], iy, iz), wx))
        r = rows[0] if dc < 2 else lerp(rows[0], rows[1], wts[1][None, None, :, None, None])
        res.append(r)
    out = res[0] if dc < 3 else lerp(res[0], res[1], wts[2][None, :, None, None, None])
    # encode
    if dt == T.FLOAT:
        with np.errstate(over="ignore"):
            return out.astype(src.dtype)
    if norm:
        s = ((1 << bpc) - 1) if dt == T.UINT else (((1 << bpc) - 1) >> 1)
        if bpc <= 8 or no_double:
            q = np.trunc((out * f32(s)).astype(f32).astype(np.float64))
        else:
            q = np.trunc(out.astype(np.float64) * np.float64(s))
        return (q.astype(np.int64) % (1 << bpc)).astype(np.dtype(f"uint{bpc}")).view(src.dtype)
    return (out.astype(np.int64) % (1 << bpc)).astype(np.dtype(f"uint{bpc}")).view(src.dtype)


def generate_chain(level0_bytes: np.ndarray, dim, t: int, limit: int = 0, no_double: bool = False) -> np.ndarray:
    """Full level-major image buffer (uint8) like oracle.generate_mip_map_chain."""
    dc = it.dim_count(t)
    C = it.channel_count(t)
    sd = _storage_dtype(t)
    L = layer_count(dim, t)
    n_levels = level_count(dim, t, limit)
    d0 = level_dims(dim, t, 0)
    shape = (L, d0[2] if dc >= 3 else 1, d0[1] if dc >= 2 else 1, d0[0], C)
    cur = np.frombuffer(np.ascontiguousarray(level0_bytes).tobytes(), dtype=sd)[: int(np.prod(shape))].reshape(shape)
    parts = [cur]
    for _ in range(1, n_levels):
        cur = minify_level(cur, t, no_double) if cur.size else cur
        parts.append(cur)
    return np.concatenate([p.reshape(-1).view(np.uint8) for p in parts]) if parts else np.zeros(0, np.uint8)
