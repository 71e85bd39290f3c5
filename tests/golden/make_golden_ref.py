"""Freezes outputs of the REFERENCE's own Host-Compute minify kernels (oracle/_ref, built by oracle/build_ref.py from
/root/reference) as golden fixtures.  Inputs are the counter-based synthetic data of SURVEY.md 8d, so only hashes are stored.

    python tests/golden/make_golden_ref.py        # rewrites tests/golden/golden_ref.json (needs /root/reference)

Every case is also run through the restatement (oracle/minify_oracle.c) and must agree before the file is written.
"heavy" cases are BASELINE.json configs at full size (C3 / C4: the largest shard the reference can address with its
32-bit level offsets, host_image.hpp:44); the CPU suite re-checks only the light ones, the GPU suite checks the CUDA
path against all of them.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from oracle import ref  # noqa: E402
from floor_b200.image_types import IMAGE_TYPE as T  # noqa: E402

M = T.FLAG_MIPMAPPED | T.READ_WRITE
# name, dim, type, config_id, no_double, mip_level_limit, heavy
CASES = [
    ("c1_1024_rgba8", (1024, 1024), T.IMAGE_2D | T.RGBA8, 1, False, 0, False),
    ("rgba16f_512", (512, 512), T.IMAGE_2D | T.RGBA16F, 2, False, 0, False),
    ("array_4x256_rgba8", (256, 256, 4), T.IMAGE_2D_ARRAY | T.RGBA8, 3, False, 0, False),
    ("cube_64_rgba32f", (64, 64), T.IMAGE_CUBE | T.RGBA32F, 4, False, 0, False),
    ("cubearray_2x32_rgba32f", (32, 32, 2), T.IMAGE_CUBE_ARRAY | T.RGBA32F, 4, False, 0, False),
    ("vol_64_r32f", (64, 64, 64), T.IMAGE_3D | T.R32F, 5, False, 0, False),
    ("vol_32x16x64_rgba8", (32, 16, 64), T.IMAGE_3D | T.RGBA8, 6, False, 0, False),
    ("r8_256", (256, 256), T.IMAGE_2D | T.R8, 7, False, 0, False),
    ("rg8_128x512", (128, 512), T.IMAGE_2D | T.RG8, 8, False, 0, False),
    ("r16_256_double", (256, 256), T.IMAGE_2D | T.R16, 9, False, 0, False),
    ("r16_256_nodouble", (256, 256), T.IMAGE_2D | T.R16, 9, True, 0, False),
    ("rgba16_128", (128, 128), T.IMAGE_2D | T.RGBA16, 10, False, 0, False),
    ("rgba16_128_nodouble", (128, 128), T.IMAGE_2D | T.RGBA16, 10, True, 0, False),
    ("rgba8snorm_128", (128, 128), T.IMAGE_2D | T.RGBA8I_NORM, 11, False, 0, False),
    ("rg16snorm_128", (128, 128), T.IMAGE_2D | T.RG16I_NORM, 12, False, 0, False),
    ("rg16snorm_128_nodouble", (128, 128), T.IMAGE_2D | T.RG16I_NORM, 12, True, 0, False),
    ("rgba8ui_128", (128, 128), T.IMAGE_2D | T.RGBA8UI, 13, False, 0, False),
    ("rgba8i_128", (128, 128), T.IMAGE_2D | T.RGBA8I, 14, False, 0, False),
    ("rg16ui_128", (128, 128), T.IMAGE_2D | T.RG16UI, 15, False, 0, False),
    ("r16i_128", (128, 128), T.IMAGE_2D | T.R16I, 16, False, 0, False),
    ("rgba32ui_64", (64, 64), T.IMAGE_2D | T.RGBA32UI, 17, False, 0, False),
    ("rg32i_64", (64, 64), T.IMAGE_2D | T.RG32I, 18, False, 0, False),
    ("r16f_256", (256, 256), T.IMAGE_2D | T.R16F, 19, False, 0, False),
    ("rg32f_128", (128, 128), T.IMAGE_2D | T.RG32F, 20, False, 0, False),
    ("npot_1920x1080_rgba8", (1920, 1080), T.IMAGE_2D | T.RGBA8, 21, False, 0, False),
    ("npot_100x37_rgba16f", (100, 37), T.IMAGE_2D | T.RGBA16F, 22, False, 0, False),
    ("npot_vol_30x20x10_r32f", (30, 20, 10), T.IMAGE_3D | T.R32F, 23, False, 0, False),
    ("nonsquare_1024x64_rgba8", (1024, 64), T.IMAGE_2D | T.RGBA8, 26, False, 0, False),
    ("limit4_512_rgba16f", (512, 512), T.IMAGE_2D | T.RGBA16F, 28, False, 4, False),
    ("npot_array_3x333x111_rg16", (333, 111, 3), T.IMAGE_2D_ARRAY | T.RG16, 29, False, 0, False),
    ("npot_vol_65x33x17_rgba16f", (65, 33, 17), T.IMAGE_3D | T.RGBA16F, 30, False, 0, False),
    ("vol_128_rgba8ui", (128, 128, 128), T.IMAGE_3D | T.RGBA8UI, 31, False, 0, False),
    ("array_2x512_rgba16i_norm", (512, 512, 2), T.IMAGE_2D_ARRAY | T.RGBA16I_NORM, 32, False, 0, False),
    ("rgb8_200x120", (200, 120), T.IMAGE_2D | T.RGB8, 33, False, 0, False),
    ("1d_1000_rgba8", (1000,), T.IMAGE_1D | T.RGBA8, 24, False, 0, False),
    ("1darray_256x3_r32f", (256, 3), T.IMAGE_1D_ARRAY | T.R32F, 25, False, 0, False),
    ("depth_256_d32f", (256, 256), T.D32F, 27, False, 0, False),
    ("deptharray_100x60x2_d32f", (100, 60, 2), T.IMAGE_DEPTH_ARRAY | T.FORMAT_32 | T.FLOAT, 34, False, 0, False),
    ("quirk_2624x188_rgba8", (2624, 188), T.IMAGE_2D | T.RGBA8, 35, False, 0, False),
    # BASELINE.json configs at full size (config ids as in bench.py)
    ("C2_8192_rgba16f", (8192, 8192), T.IMAGE_2D | T.RGBA16F, 2, False, 0, True),
    ("C3_shard_8x1024_rgba8", (1024, 1024, 8), T.IMAGE_2D_ARRAY | T.RGBA8, 3, False, 0, True),
    ("C4_one_cube_4096_rgba32f", (4096, 4096, 1), T.IMAGE_CUBE_ARRAY | T.RGBA32F, 4, False, 0, True),
    ("C5_512_r32f", (512, 512, 512), T.IMAGE_3D | T.R32F, 5, False, 0, True),
]


def main():
    out = []
    threads = os.cpu_count() or 4
    for name, dim, t, cid, nd, limit, heavy in CASES:
        t |= M
        t0 = time.time()
        l0 = oracle.fill_synthetic(dim, t, cid)
        r = ref.generate_mip_map_chain(l0, dim, t, mip_level_limit=limit, no_double=nd, threads=threads)
        o = oracle.generate_mip_map_chain(l0, dim, t, mip_level_limit=limit, no_double=nd, threads=threads)
        if not np.array_equal(r, o):
            raise SystemExit(f"{name}: restatement differs from the reference at byte {int(np.nonzero(r != o)[0][0])}")
        out.append({"name": name, "dim": list(dim), "type": hex(t), "config_id": cid, "no_double": nd, "mip_level_limit": limit,
                    "heavy": heavy, "levels": oracle.mip_level_count(dim, t, limit), "bytes": int(r.size),
                    "level0_sha256": hashlib.sha256(l0.tobytes()).hexdigest(),
                    "chain_sha256": hashlib.sha256(r.tobytes()).hexdigest(),
                    "producer": "reference (oracle/_ref, include/floor/device/backend/mip_map_minify.hpp + host_image.hpp compiled with g++)"})
        print(f"{name}: {r.size} bytes, {time.time() - t0:.1f} s", flush=True)
    with open(os.path.join(HERE, "golden_ref.json"), "w") as f:
        json.dump(out, f, indent=1)
        f.write("\n")


if __name__ == "__main__":
    main()
