"""A/B target for the tile kernel's occupancy choices: NPOT chains with 16-byte texels and NPOT volumes (not covered by bench.py's n1 / n2)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import floor_b200
from floor_b200.image_types import IMAGE_TYPE as T
M = T.FLAG_MIPMAPPED | T.READ_WRITE
ctx = floor_b200.device_context(); q = ctx.create_queue(ctx.get_device(0))
for name, dim, t in [("rgba32f 1920x1080 x32", (1920, 1080, 32), T.IMAGE_2D_ARRAY | T.RGBA32F | M),
                     ("r32f 500x300x200", (500, 300, 200), T.IMAGE_3D | T.R32F | M),
                     ("rgba16f 500x300x200", (500, 300, 200), T.IMAGE_3D | T.RGBA16F | M),
                     ("rgba8 500x300x200", (500, 300, 200), T.IMAGE_3D | T.RGBA8 | M)]:
    imgs = [ctx.create_image(q, dim, t) for _ in range(2)]
    for i, im in enumerate(imgs):
        im.fill_synthetic(q, 2, i)
    for k in range(5):
        imgs[k & 1].enqueue_mip_map_chain(q)
    q.finish()
    best = 1e9
    for rep in range(3):
        e0 = q.record_event()
        for k in range(20):
            imgs[k & 1].enqueue_mip_map_chain(q)
        e1 = q.record_event()
        best = min(best, q.elapsed_ms(e0, e1) / 20)
    size = imgs[0].image_data_size_mip_maps
    print(f"{name:26s} {size / 1e6:8.1f} MB  {best * 1e3:8.1f} us  {size / best / 1e6:8.1f} GB/s")
    for im in imgs:
        im.destroy()
