"""Freezes oracle outputs as golden fixtures (input = counter-based synthetic data, so only hashes are stored).

    python tests/golden/make_golden.py        # rewrites tests/golden/golden.json and golden_small.npz

The reference ships no golden vectors and cannot run here ("Oracle"); these fixtures pin the CPU
restatement itself -- parity stays "unpinned" with respect to a reference binary.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from floor_b200.image_types import IMAGE_TYPE as T  # noqa: E402

M = T.FLAG_MIPMAPPED | T.READ_WRITE
CASES = [
    ("c1_1024_rgba8", (1024, 1024), T.IMAGE_2D | T.RGBA8, 1, False),
    ("rgba16f_512", (512, 512), T.IMAGE_2D | T.RGBA16F, 2, False),
    ("array_4x256_rgba8", (256, 256, 4), T.IMAGE_2D_ARRAY | T.RGBA8, 3, False),
    ("cube_64_rgba32f", (64, 64), T.IMAGE_CUBE | T.RGBA32F, 4, False),
    ("cubearray_2x32_rgba32f", (32, 32, 2), T.IMAGE_CUBE_ARRAY | T.RGBA32F, 4, False),
    ("vol_64_r32f", (64, 64, 64), T.IMAGE_3D | T.R32F, 5, False),
    ("vol_32x16x64_rgba8", (32, 16, 64), T.IMAGE_3D | T.RGBA8, 6, False),
    ("r8_256", (256, 256), T.IMAGE_2D | T.R8, 7, False),
    ("rg8_128x512", (128, 512), T.IMAGE_2D | T.RG8, 8, False),
    ("r16_256_double", (256, 256), T.IMAGE_2D | T.R16, 9, False),
    ("r16_256_nodouble", (256, 256), T.IMAGE_2D | T.R16, 9, True),
    ("rgba16_128", (128, 128), T.IMAGE_2D | T.RGBA16, 10, False),
    ("rgba8snorm_128", (128, 128), T.IMAGE_2D | T.RGBA8I_NORM, 11, False),
    ("rg16snorm_128", (128, 128), T.IMAGE_2D | T.RG16I_NORM, 12, False),
    ("rgba8ui_128", (128, 128), T.IMAGE_2D | T.RGBA8UI, 13, False),
    ("rgba8i_128", (128, 128), T.IMAGE_2D | T.RGBA8I, 14, False),
    ("rg16ui_128", (128, 128), T.IMAGE_2D | T.RG16UI, 15, False),
    ("r16i_128", (128, 128), T.IMAGE_2D | T.R16I, 16, False),
    ("rgba32ui_64", (64, 64), T.IMAGE_2D | T.RGBA32UI, 17, False),
    ("rg32i_64", (64, 64), T.IMAGE_2D | T.RG32I, 18, False),
    ("r16f_256", (256, 256), T.IMAGE_2D | T.R16F, 19, False),
    ("rg32f_128", (128, 128), T.IMAGE_2D | T.RG32F, 20, False),
    ("npot_1920x1080_rgba8", (1920, 1080), T.IMAGE_2D | T.RGBA8, 21, False),
    ("npot_100x37_rgba16f", (100, 37), T.IMAGE_2D | T.RGBA16F, 22, False),
    ("npot_vol_30x20x10_r32f", (30, 20, 10), T.IMAGE_3D | T.R32F, 23, False),
    ("1d_1000_rgba8", (1000,), T.IMAGE_1D | T.RGBA8, 24, False),
    ("1darray_256x3_r32f", (256, 3), T.IMAGE_1D_ARRAY | T.R32F, 25, False),
    ("nonsquare_1024x64_rgba8", (1024, 64), T.IMAGE_2D | T.RGBA8, 26, False),
    ("depth_256_d32f", (256, 256), T.D32F, 27, False),
]


def main():
    out = []
    small = {}
    for name, dim, t, cid, nd in CASES:
        t |= M
        l0 = oracle.fill_synthetic(dim, t, cid)
        chain = oracle.generate_mip_map_chain(l0, dim, t, no_double=nd, threads=8)
        out.append({"name": name, "dim": list(dim), "type": hex(t), "config_id": cid, "no_double": nd,
                    "levels": oracle.mip_level_count(dim, t), "bytes": int(chain.size),
                    "level0_sha256": hashlib.sha256(l0.tobytes()).hexdigest(),
                    "chain_sha256": hashlib.sha256(chain.tobytes()).hexdigest()})
        if chain.size <= 64 * 1024:
            small[name] = chain
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), **small)
    print(f"wrote {len(out)} cases, {len(small)} with full bytes")


if __name__ == "__main__":
    main()
