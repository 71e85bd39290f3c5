"""compute-sanitizer target (round 2): small chains through the persistent TMA tile kernel in both forms (two levels per launch /
one-launch with finisher pool, units and last-tile stage), the literal kernel with 3-channel and packed formats, external memory"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import floor_b200, oracle
from floor_b200.image_types import IMAGE_TYPE as T
M = T.FLAG_MIPMAPPED | T.READ_WRITE
ctx = floor_b200.device_context(); q = ctx.create_queue(ctx.get_device(0))
for dim, t, kw in [((1080, 600), T.IMAGE_2D | T.RGBA8 | M, {"tma_tiles": "always+nosplit"}), ((1080, 600), T.IMAGE_2D | T.RGBA8 | M, {"tma_tiles": "always"}),
                   ((600, 333, 40), T.IMAGE_2D_ARRAY | T.RGBA16F | M, {"tma_tiles": "always+nosplit"}), ((1001, 999), T.IMAGE_2D | T.RGBA32F | M, {"tma_tiles": "always+nosplit"}),
                   ((2624, 188), T.IMAGE_2D | T.R8 | M, {"tma_tiles": "always+nosplit"}), ((1008, 70), T.IMAGE_2D | T.R16F | M, {"tma_tiles": "always"}),
                   ((100, 37), T.IMAGE_2D | T.RGB8 | M, {}), ((33, 65, 3), T.IMAGE_2D_ARRAY | T.RGBA4 | M, {}), ((12, 10, 6), T.IMAGE_3D | T.RGBA2I_NORM | M, {})]:
    l0 = oracle.fill_synthetic(dim, t, 5)
    img = ctx.create_image(q, dim, t, **kw)
    img.upload_levels(q, l0, 0, 0)
    for _ in range(2):
        img.enqueue_mip_map_chain(q)
    got = img.download_levels(q)
    assert np.array_equal(got, oracle.generate_mip_map_chain(l0, dim, t, threads=4)), (dim, hex(t))
    img.destroy()
print("race_ptile ok")
