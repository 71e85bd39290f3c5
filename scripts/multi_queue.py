"""many independent small textures: chains on Q queues at once (one texture per queue at a time) vs one queue"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import floor_b200
from floor_b200.image_types import IMAGE_TYPE as T
M = T.FLAG_MIPMAPPED | T.READ_WRITE
ctx = floor_b200.device_context(); dev = ctx.get_device(0)
for name, dim, t in [("1024^2 rgba8", (1024, 1024), T.IMAGE_2D | T.RGBA8 | M), ("512^2 rgba8", (512, 512), T.IMAGE_2D | T.RGBA8 | M),
                     ("2048^2 rgba8", (2048, 2048), T.IMAGE_2D | T.RGBA8 | M), ("1920x1080 rgba8", (1920, 1080), T.IMAGE_2D | T.RGBA8 | M)]:
    for nq in (1, 2, 4, 8):
        qs = [ctx.create_queue(dev) for _ in range(nq)]
        imgs = [ctx.create_image(qs[0], dim, t) for _ in range(16)]
        for i, im in enumerate(imgs):
            im.fill_synthetic(qs[0], 1, i)
        qs[0].finish()
        reps = 10
        for rep in range(2):
            for q in qs: q.finish()
            t0 = time.perf_counter()
            for r in range(reps):
                for i, im in enumerate(imgs):
                    im.enqueue_mip_map_chain(qs[i % nq])
            for q in qs: q.finish()
            dt = time.perf_counter() - t0
        n = reps * len(imgs)
        print(f"{name:16s} queues={nq}: {dt / n * 1e6:7.2f} us per chain, {imgs[0].image_data_size_mip_maps * n / dt / 1e9:8.1f} GB/s")
        if nq == 1:
            batch = ctx.create_mip_chain_batch(imgs)
            for rep in range(2):
                qs[0].finish()
                t0 = time.perf_counter()
                for r in range(reps):
                    batch.enqueue(qs[0])
                qs[0].finish()
                dt = time.perf_counter() - t0
            print(f"{name:16s} batch (1 graph launch per 16 textures): {dt / n * 1e6:7.2f} us per chain, {imgs[0].image_data_size_mip_maps * n / dt / 1e9:8.1f} GB/s")
            batch.destroy()
        for im in imgs: im.destroy()
        for q in qs: q.destroy()
