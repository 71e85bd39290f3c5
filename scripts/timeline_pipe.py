"""tuning: per-CTA time stamps of back-to-back single-pass launches on rotating images (FLMIP_TIMELINE build: FLMIP_LIB=.../lib_tl.so)"""
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
want = sys.argv[1:] or ["c2", "c5", "c3x256", "c1"]
nimg = int(os.environ.get("NIMG", "2"))
for name, dim, t in [("c2", (8192, 8192), T.IMAGE_2D | T.RGBA16F | M), ("c5", (512, 512, 512), T.IMAGE_3D | T.R32F | M),
                     ("c3x256", (1024, 1024, 256), T.IMAGE_2D_ARRAY | T.RGBA8 | M), ("c1", (1024, 1024), T.IMAGE_2D | T.RGBA8 | M)]:
    if name not in want:
        continue
    imgs = [ctx.create_image(q, dim, t) for _ in range(nimg)]
    for i, im in enumerate(imgs):
        im.fill_synthetic(q, 2, i)
    for k in range(3 * nimg):
        imgs[k % nimg].enqueue_mip_map_chain(q)
    q.finish()
    n = 296
    buf = np.zeros((n, 16), np.uint64)
    for im in imgs:
        L.flmip_debug_timeline(im._handle, buf.ctypes.data, n)   # clears the stamps
    e0 = q.record_event()
    reps = 3
    for k in range(reps * nimg):
        imgs[k % nimg].enqueue_mip_map_chain(q)
    e1 = q.record_event()
    ms = q.elapsed_ms(e0, e1)
    print(f"{name}: {reps * nimg} chains on {nimg} images, {ms * 1e3 / (reps * nimg):.1f} us per chain; stamps of the last launch on each image")
    t0 = None
    for i, im in enumerate(imgs):
        L.flmip_debug_timeline(im._handle, buf.ctypes.data, n)
        live = buf[buf[:, 0] != 0].astype(np.int64)
        if t0 is None:
            t0 = live[:, 0].min()
        rel = (live - t0) / 1e3
        f = lambda a: f"min {a.min():7.1f} p10 {np.percentile(a, 10):7.1f} med {np.median(a):7.1f} p90 {np.percentile(a, 90):7.1f} max {a.max():7.1f}"
        print(f"  image {i}: CTA start          {f(rel[:, 0])}")
        print(f"           scheduler ran dry  {f(rel[:, 1])}")
        print(f"           consumers done     {f(rel[:, 2])}")
        print(f"           last finisher done {f(rel[:, 3])}")
    for im in imgs:
        im.destroy()
