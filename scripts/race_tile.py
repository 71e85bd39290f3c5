"""compute-sanitizer target: a few small chains through the tile kernel (2D incl. a texel-2 quirk size, 3D) and the single-pass kernel"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import floor_b200, oracle
from floor_b200.image_types import IMAGE_TYPE as T
M = T.FLAG_MIPMAPPED | T.READ_WRITE
ctx = floor_b200.device_context(); q = ctx.create_queue(ctx.get_device(0))
for dim, t, kw in [((333, 129), T.IMAGE_2D | T.RGBA8 | M, {}), ((2624, 188), T.IMAGE_2D | T.R8 | M, {}), ((130, 67, 3), T.IMAGE_2D_ARRAY | T.RGBA16F | M, {}),
                   ((70, 33, 18), T.IMAGE_3D | T.R32F | M, {}), ((256, 128), T.IMAGE_2D | T.RGBA32F | M, {"force_tiled": True}),
                   ((256, 256), T.IMAGE_2D | T.RGBA8 | M, {}), ((64, 32, 32), T.IMAGE_3D | T.R32F | M, {})]:
    l0 = oracle.fill_synthetic(dim, t, 5)
    img = ctx.create_image(q, dim, t, **kw)
    img.upload_levels(q, l0, 0, 0)
    img.generate_mip_map_chain(q)
    got = img.download_levels(q)
    assert np.array_equal(got, oracle.generate_mip_map_chain(l0, dim, t, threads=4)), (dim, hex(t))
    img.destroy()
print("race_tile ok")
