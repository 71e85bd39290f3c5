"""tuning: per-CTA time stamps of one single-pass launch (needs the FLMIP_TIMELINE build: FLMIP_LIB=.../libfloor_b200_mip_tl.so)"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import floor_b200
from floor_b200.image_types import IMAGE_TYPE as T
M = T.FLAG_MIPMAPPED | T.READ_WRITE
L = floor_b200.lib()
L.flmip_debug_timeline.restype = ctypes.c_int
L.flmip_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32]
ctx = floor_b200.device_context(); q = ctx.create_queue(ctx.get_device(0))
for name, dim, t in [("c2", (8192, 8192), T.IMAGE_2D | T.RGBA16F | M), ("c5", (512, 512, 512), T.IMAGE_3D | T.R32F | M),
                     ("c3x256", (1024, 1024, 256), T.IMAGE_2D_ARRAY | T.RGBA8 | M), ("c1", (1024, 1024), T.IMAGE_2D | T.RGBA8 | M)]:
    imgs = [ctx.create_image(q, dim, t) for _ in range(2)]
    for i, im in enumerate(imgs):
        im.fill_synthetic(q, 2, i)
    for k in range(6):
        imgs[k & 1].enqueue_mip_map_chain(q)
    q.finish()
    n = 296
    buf = np.zeros((n, 16), np.uint64)
    L.flmip_debug_timeline(imgs[0]._handle, buf.ctypes.data, n)   # clears the stamps
    e0 = q.record_event()
    imgs[0].enqueue_mip_map_chain(q)
    e1 = q.record_event()
    ms = q.elapsed_ms(e0, e1)
    L.flmip_debug_timeline(imgs[0]._handle, buf.ctypes.data, n)
    live = buf[buf[:, 0] != 0]
    t0 = live[:, 0].min()
    rel = (live.astype(np.int64) - np.int64(t0)) / 1e3
    f = lambda a: f"min {a.min():7.1f} med {np.median(a):7.1f} max {a.max():7.1f}"
    print(f"{name}: {len(live)} CTAs, event time {ms * 1e3:.1f} us")
    print(f"   CTA start            {f(rel[:, 0])}")
    print(f"   scheduler ran dry    {f(rel[:, 1])}")
    print(f"   consumers done       {f(rel[:, 2])}")
    print(f"   last finisher done   {f(rel[:, 3])}")
    print(f"   histogram of 'last finisher done' (us): {np.histogram(rel[:, 3], bins=8)[0].tolist()} edges {np.round(np.histogram(rel[:, 3], bins=8)[1], 1).tolist()}")
    # last-arriver stages (only the CTAs that ran one have these stamps): publish returned, patch gathered, patch reduced, layer stage done
    stages = rel[:, 4:8]
    ran = live[:, 4] != 0
    if ran.any():
        order = np.argsort(rel[ran][:, 6])[-4:]
        for row in rel[ran][order]:
            print(f"   group stage of a CTA: consumers done {row[2]:7.1f} | publish returned {row[4]:7.1f} gathered {row[5]:7.1f} reduced {row[6]:7.1f}"
                  + (f" | layer stage done {row[7]:7.1f}" if row[7] > 0 else "")
                  + " | levels of the group patch at " + " ".join(f"{v:.2f}" for v in row[8:16] if v > 0))
    for im in imgs:
        im.destroy()
