"""tuning: microseconds per texture of one batch graph launch (flmip_batch_*) by FLMIP_BATCH_LANES (read once per process: run per value)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import floor_b200
from floor_b200.image_types import IMAGE_TYPE as T
M = T.FLAG_MIPMAPPED | T.READ_WRITE
ctx = floor_b200.device_context(); q = ctx.create_queue(ctx.get_device(0))
for name, dim, t in [("1024^2 rgba8", (1024, 1024), T.IMAGE_2D | T.RGBA8 | M), ("512^2 rgba8", (512, 512), T.IMAGE_2D | T.RGBA8 | M),
                     ("1920x1080 rgba8", (1920, 1080), T.IMAGE_2D | T.RGBA8 | M), ("2048^2 rgba16f", (2048, 2048), T.IMAGE_2D | T.RGBA16F | M)]:
    out = []
    for n in (16, 64, 512):
        imgs = [ctx.create_image(q, dim, t) for _ in range(n)]
        for i, im in enumerate(imgs):
            im.fill_synthetic(q, 1, i)
        b = ctx.create_mip_chain_batch(imgs)
        for _ in range(3):
            b.enqueue(q)
        q.finish()
        best = 1e9
        for rep in range(4):
            e0 = q.record_event()
            for _ in range(4):
                b.enqueue(q)
            e1 = q.record_event()
            best = min(best, q.elapsed_ms(e0, e1) * 1e3 / (4 * n))
        out.append(f"{n}: {best:.2f}")
        b.destroy()
        for im in imgs:
            im.destroy()
    print(f"lanes={os.environ.get('FLMIP_BATCH_LANES', 'default')} {name}: us per texture by batch size  " + "  ".join(out), flush=True)
