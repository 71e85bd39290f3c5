"""tuning: per-CTA time stamps of one persistent-tile-kernel launch (needs the FLMIP_TIMELINE build: FLMIP_LIB=.../lib_tl.so)"""
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
for name, dim, t, limit in [("n1 2 levels only", (3840, 2160), T.IMAGE_2D | T.RGBA8 | M, 3), ("n1 full chain", (3840, 2160), T.IMAGE_2D | T.RGBA8 | M, 0),
                            ("1080p rgba16f x4, 2 levels", (1920, 1080, 4), T.IMAGE_2D_ARRAY | T.RGBA16F | M, 3), ("c1 (single-pass kernel)", (1024, 1024), T.IMAGE_2D | T.RGBA8 | M, 0)]:
    imgs = [ctx.create_image(q, dim, t, mip_level_limit=limit, tma_tiles="always+nosplit") for _ in range(6)]
    for i, im in enumerate(imgs):
        im.fill_synthetic(q, 2, i)
    for k in range(12):
        imgs[k % 6].enqueue_mip_map_chain(q)
    q.finish()
    # back-to-back rate for reference
    e0 = q.record_event()
    for k in range(24):
        imgs[k % 6].enqueue_mip_map_chain(q)
    e1 = q.record_event()
    rate = q.elapsed_ms(e0, e1) / 24
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
    print(f"{name}: plan {imgs[0].plan()}, {len(live)} CTAs, isolated event time {ms * 1e3:.1f} us, back-to-back {rate * 1e3:.1f} us per chain")
    print(f"   CTA start            {f(rel[:, 0])}")
    print(f"   scheduler ran dry    {f(rel[:, 1])}")
    print(f"   consumers done       {f(rel[:, 2])}")
    if (live[:, 3] != 0).any():
        print(f"   last finisher done   {f(rel[live[:, 3] != 0][:, 3])}")
    for im in imgs:
        im.destroy()
