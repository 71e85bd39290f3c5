"""compute-sanitizer target (round 2, second pass): chains on independent images on a queue with chain overlap (late-waiting first kernels,
pdl_late_wait at the end), and the asynchronous gather of the group / layer patches (cp.async L2 -> shared memory)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import floor_b200, oracle
from floor_b200.image_types import IMAGE_TYPE as T
M = T.FLAG_MIPMAPPED | T.READ_WRITE
ctx = floor_b200.device_context(); q = ctx.create_queue(ctx.get_device(0))
q.set_mip_chain_overlap(True)
cases = [((1024, 512), T.IMAGE_2D | T.RGBA8 | M), ((2048, 256), T.IMAGE_2D | T.RGBA16F | M), ((256, 256, 6), T.IMAGE_2D_ARRAY | T.RG8 | M),
         ((128, 64, 64), T.IMAGE_3D | T.R32F | M), ((64, 64, 64), T.IMAGE_3D | T.RGBA32F | M), ((333, 200), T.IMAGE_2D | T.RGBA8 | M),
         ((1080, 600), T.IMAGE_2D | T.RGBA16F | M), ((70, 33, 18), T.IMAGE_3D | T.R16F | M)]
imgs, wants = [], []
for i, (dim, t) in enumerate(cases):
    l0 = oracle.fill_synthetic(dim, t, 40 + i)
    img = ctx.create_image(q, dim, t)
    img.upload_levels(q, l0, 0, 0, sync=False)
    imgs.append(img); wants.append(oracle.generate_mip_map_chain(l0, dim, t, threads=8))
for rnd in range(3):
    for img in imgs:
        img.enqueue_mip_map_chain(q)
    imgs[rnd].enqueue_mip_map_chain(q)
for img, want, (dim, t) in zip(imgs, wants, cases):
    assert np.array_equal(img.download_levels(q), want), (dim, hex(t))
for img in imgs:
    img.destroy()
print("race_overlap ok")
