"""tuning: microseconds per chain for back-to-back chains on rotating images, by number of chains in flight (late-wait experiments)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import floor_b200
from floor_b200.image_types import IMAGE_TYPE as T
M = T.FLAG_MIPMAPPED | T.READ_WRITE
ctx = floor_b200.device_context(); q = ctx.create_queue(ctx.get_device(0))
want = sys.argv[1:] or ["c2", "c5", "c3x256", "c1"]
for name, dim, t in [("c2", (8192, 8192), T.IMAGE_2D | T.RGBA16F | M), ("c5", (512, 512, 512), T.IMAGE_3D | T.R32F | M),
                     ("c3x256", (1024, 1024, 256), T.IMAGE_2D_ARRAY | T.RGBA8 | M), ("c1", (1024, 1024), T.IMAGE_2D | T.RGBA8 | M)]:
    if name not in want:
        continue
    for nimg in (2, 3, 4):
        imgs = [ctx.create_image(q, dim, t) for _ in range(nimg)]
        for i, im in enumerate(imgs):
            im.fill_synthetic(q, 2, i)
        for k in range(3 * nimg):
            imgs[k % nimg].enqueue_mip_map_chain(q)
        q.finish()
        out = []
        for chains in (4, 8, 16, 32, 64):
            best = 1e9
            for rep in range(3):
                e0 = q.record_event()
                for k in range(chains):
                    imgs[k % nimg].enqueue_mip_map_chain(q)
                e1 = q.record_event()
                best = min(best, q.elapsed_ms(e0, e1) * 1e3 / chains)
            out.append(f"{chains}: {best:.1f}")
        print(f"{name} images={nimg}  us per chain by chains enqueued  " + "  ".join(out), flush=True)
        for im in imgs:
            im.destroy()
