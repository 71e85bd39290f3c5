"""A/B: single-pass kernel with 2x2(x2)-tile units vs single-tile units over image sizes (tuning experiment for the planner's rule)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import floor_b200
from floor_b200.image_types import IMAGE_TYPE as T
M = T.FLAG_MIPMAPPED | T.READ_WRITE
ctx = floor_b200.device_context(); q = ctx.create_queue(ctx.get_device(0))
cases = []
for fmt, name in [(T.RGBA8, "rgba8"), (T.RGBA16F, "rgba16f"), (T.RGBA32F, "rgba32f"), (T.R8, "r8")]:
    for dim in [(2048, 2048), (4096, 2048), (4096, 4096), (8192, 4096), (8192, 8192), (16384, 8192)]:
        cases.append((f"{name} {dim[0]}x{dim[1]}", dim, T.IMAGE_2D | fmt | M))
for dim in [(256, 256, 256), (512, 256, 256), (512, 512, 256), (512, 512, 512), (1024, 512, 512)]:
    cases.append((f"r32f {dim}", dim, T.IMAGE_3D | T.R32F | M))
    cases.append((f"rgba8 {dim}", dim, T.IMAGE_3D | T.RGBA8 | M))
for layers in (4, 16, 64):
    cases.append((f"rgba8 1024^2 x{layers}", (1024, 1024, layers), T.IMAGE_2D_ARRAY | T.RGBA8 | M))
for name, dim, t in cases:
    res = {}
    for units in (True, False, None):
        try:
            imgs = [ctx.create_image(q, dim, t, units=units) for _ in range(2)]
        except Exception as e:
            res[units] = None
            continue
        tiles = None
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
        res[units] = (best, imgs[0].image_data_size_mip_maps)
        for im in imgs:
            im.destroy()
    if res[True] and res[False]:
        a, b, c = res[True][0], res[False][0], res[None][0]
        print(f"{name:28s} {res[True][1] / 1e6:9.1f} MB  units {a * 1e3:8.1f} us  single {b * 1e3:8.1f} us  ratio {a / b:5.2f}  planner's choice {c * 1e3:8.1f} us"
              f"{'' if c <= 1.03 * min(a, b) else '   <-- not the better one'}")
