"""tuning: LDG tile kernel vs persistent TMA tile kernel (single launch / two levels per launch) over NPOT image sizes -- sets the
planner's thresholds (tiles per resident CTA) in flmip.cpp"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import floor_b200
from floor_b200.image_types import IMAGE_TYPE as T
M = T.FLAG_MIPMAPPED | T.READ_WRITE
ctx = floor_b200.device_context(); dev = ctx.get_device(0); q = ctx.create_queue(dev)
cases = []
for fmt, name, bpp in [(T.RGBA8, "rgba8", 4), (T.RGBA16F, "rgba16f", 8), (T.RGBA32F, "rgba32f", 16), (T.R8, "r8", 1)]:
    for dim in [(1920, 1080), (3840, 2160), (5120, 2880), (7680, 4320), (10000, 6000), (15360, 8640)]:
        cases.append((f"{name} {dim[0]}x{dim[1]}", dim, T.IMAGE_2D | fmt | M, bpp))
for layers in (4, 16, 64):
    cases.append((f"rgba16f 1920x1080 x{layers}", (1920, 1080, layers), T.IMAGE_2D_ARRAY | T.RGBA16F | M, 8))
    cases.append((f"rgba8 1000x600 x{layers * 4}", (1000, 600, layers * 4), T.IMAGE_2D_ARRAY | T.RGBA8 | M, 4))
for name, dim, t, bpp in cases:
    tiles = -(-dim[0] * bpp // 512) * -(-dim[1] // 64) * (dim[2] if len(dim) > 2 else 1)
    res = {}
    for mode, kw in [("ldg", {"no_tma_tiles": True}), ("tma", {"tma_tiles": "always+nosplit"}), ("tma-split", {"tma_tiles": "always+split"}), ("planner", {})]:
        n_img = 2 if dim[0] * dim[1] * bpp > (64 << 20) else 6
        imgs = [ctx.create_image(q, dim, t, **kw) for _ in range(n_img)]
        for i, im in enumerate(imgs):
            im.fill_synthetic(q, 2, i)
        for k in range(6):
            imgs[k % n_img].enqueue_mip_map_chain(q)
        q.finish()
        best = 1e9
        for rep in range(3):
            e0 = q.record_event()
            for k in range(20):
                imgs[k % n_img].enqueue_mip_map_chain(q)
            e1 = q.record_event()
            best = min(best, q.elapsed_ms(e0, e1) / 20)
        res[mode] = (best, imgs[0].image_data_size_mip_maps, imgs[0].plan())
        for im in imgs:
            im.destroy()
    sz = res["ldg"][1]
    line = f"{name:28s} {sz / 1e6:9.1f} MB {tiles / (2.0 * dev.units):7.1f} tiles/CTA "
    for mode in ("ldg", "tma", "tma-split", "planner"):
        ms, _, plan = res[mode]
        line += f" {mode} {ms * 1e3:8.1f} us ({sz / ms / 1e6:6.0f} GB/s, {plan['launches']}L)"
    best = min(res[m][0] for m in ("ldg", "tma", "tma-split"))
    line += "" if res["planner"][0] <= 1.04 * best else "   <-- planner not the best"
    print(line, flush=True)
