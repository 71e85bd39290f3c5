"""GPU tests of the chain overlap (pytest -m gpu): on a queue that has opted in (flmip_stream_set_chain_overlap), the first kernel of a
chain on an image without a kernel in the queue's open run starts while the chain in front of it is still finishing, and waits for it
before it ends.  Checked here: results stay bit-exact for every kernel family and any order of chains, completion still follows
stream order (whatever is enqueued behind a chain sees every chain before it complete), everything else enqueued through the library
ends the run, and the overlap is real (chains of independent small textures take less time per chain)."""
import numpy as np
import pytest

import floor_b200
from floor_b200.image_types import IMAGE_TYPE as T

pytestmark = pytest.mark.gpu
M = T.FLAG_MIPMAPPED | T.READ_WRITE


@pytest.fixture()
def overlap_queue(gpu_ctx):
    ctx, dev, _ = gpu_ctx
    q = ctx.create_queue(dev)
    q.set_mip_chain_overlap(True)
    yield q
    q.finish()
    q.destroy()


# one image per kernel family: single-pass 2D / 2D-array / 3D (units and single tiles), LDG tile kernel (one and two launches),
# persistent TMA tile kernel + LDG remainder, literal kernel (1D)
FAMILY = [
    (T.IMAGE_2D | T.RGBA8, (1024, 1024)), (T.IMAGE_2D | T.RGBA16F, (2048, 1024)), (T.IMAGE_2D_ARRAY | T.RGBA8, (512, 512, 12)),
    (T.IMAGE_3D | T.R32F, (128, 128, 128)), (T.IMAGE_2D | T.R8, (4096, 256)), (T.IMAGE_CUBE | T.RG16F, (256, 256)),
    (T.IMAGE_2D | T.RGBA8, (1000, 600)), (T.IMAGE_2D | T.RGBA8, (3840, 2160)), (T.IMAGE_2D_ARRAY | T.RGBA16F, (1920, 1080, 24)),
    (T.IMAGE_3D | T.RGBA8, (100, 60, 40)), (T.IMAGE_1D | T.R32F, (4097,)), (T.IMAGE_2D | T.RGB8, (640, 480)),
]
# per-image creation arguments: the two forms of the persistent TMA tile kernel (two levels per launch / the whole chain in one launch)
FAMILY_KW = {6: {"tma_tiles": "always"}, 8: {"tma_tiles": "always+nosplit"}}


def test_overlapped_chains_are_bit_exact_in_any_order(gpu_ctx, overlap_queue, oracle_mod):
    """240 chains over 12 images of every kernel family in a seeded random order (incl. the same image twice in a row and chains on an
    image whose previous chain is still in the open run), level 0 re-uploaded now and then: every image ends bit-exact"""
    ctx, dev, _ = gpu_ctx
    q = overlap_queue
    imgs, l0s = [], []
    for i, (bt, dim) in enumerate(FAMILY):
        t = bt | M
        im = ctx.create_image(q, dim, t, **FAMILY_KW.get(i, {}))
        if i in FAMILY_KW:
            assert im.plan()["tma_tile_launches"] >= 1, (dim, im.plan())
        l0 = oracle_mod.fill_synthetic(dim, t, 700 + i)
        im.upload_levels(q, l0, 0, 0, sync=False)
        imgs.append(im); l0s.append(l0)
    rng = np.random.default_rng(0xC0FFEE)
    launches0 = floor_b200.lib().flmip_launch_count()
    for step in range(240):
        i = int(rng.integers(0, len(imgs)))
        r = rng.random()
        if r < 0.08:
            # new level 0 (closes the run); the chain that follows must see it
            l0s[i] = oracle_mod.fill_synthetic(FAMILY[i][1], FAMILY[i][0] | M, 1000 + step)
            imgs[i].upload_levels(q, l0s[i], 0, 0, sync=False)
        imgs[i].enqueue_mip_map_chain(q)
        if r > 0.9:
            imgs[i].enqueue_mip_map_chain(q)  # the same image again: must wait for its own previous chain
    assert floor_b200.lib().flmip_launch_count() > launches0
    for i, (bt, dim) in enumerate(FAMILY):
        t = bt | M
        got = imgs[i].download_levels(q)
        want = oracle_mod.generate_mip_map_chain(l0s[i], dim, t, threads=16)
        assert np.array_equal(got, want), (hex(t), dim, int(np.nonzero(got != want)[0][0]))
    for im in imgs:
        im.destroy()


@pytest.mark.parametrize("big,small", [((T.IMAGE_2D | T.RGBA16F, (8192, 4096)), (T.IMAGE_2D | T.RGBA8, (512, 512))),
                                       ((T.IMAGE_3D | T.R32F, (512, 256, 256)), (T.IMAGE_2D | T.RGBA8, (64, 64))),
                                       ((T.IMAGE_2D_ARRAY | T.RGBA16F, (1920, 1080, 16)), (T.IMAGE_2D | T.R8, (1024, 64)))])
def test_completion_follows_stream_order(gpu_ctx, overlap_queue, oracle_mod, big, small):
    """a long chain, then a short chain on another image (starts late-waiting, would finish first), then read-backs of BOTH enqueued right
    behind: the read-back of the long chain's image must find its last levels written -- the short chain may not complete before it"""
    ctx, dev, _ = gpu_ctx
    q = overlap_queue
    L = floor_b200.lib()
    (bt_a, dim_a), (bt_b, dim_b) = big, small
    ta, tb = bt_a | M, bt_b | M
    a, b = ctx.create_image(q, dim_a, ta), ctx.create_image(q, dim_b, tb)
    la, lb = oracle_mod.fill_synthetic(dim_a, ta, 31), oracle_mod.fill_synthetic(dim_b, tb, 32)
    want_a = oracle_mod.generate_mip_map_chain(la, dim_a, ta, threads=16)
    want_b = oracle_mod.generate_mip_map_chain(lb, dim_b, tb, threads=16)
    for rep in range(6):
        assert L.flmip_image_zero(a._handle, q._stream) == 0 and L.flmip_image_zero(b._handle, q._stream) == 0
        a.upload_levels(q, la, 0, 0, sync=False)
        b.upload_levels(q, lb, 0, 0, sync=False)
        a.enqueue_mip_map_chain(q)
        b.enqueue_mip_map_chain(q)
        got_a = a.download_levels(q, sync=False)
        got_b = b.download_levels(q, sync=True)
        assert np.array_equal(got_b, want_b), (rep, "short chain")
        assert np.array_equal(got_a, want_a), (rep, "long chain: read back before it was complete?", int(np.nonzero(got_a != want_a)[0][0]))
    a.destroy(); b.destroy()


def test_overlap_is_real_and_off_by_default(gpu_ctx, oracle_mod):
    """16 independent 1024^2 RGBA8 textures, 8 rounds of chains back to back: microseconds per chain on a default queue and on a queue with
    overlap (measured on a B200: 9.5 vs 3.9 us); a fence before every chain brings the default behaviour back.  The host needs about as
    long to enqueue a chain as the GPU to run it, so the chains are enqueued behind 6 ms of other work (chains on a large image) and the
    events bracket what the GPU then finds queued up: device time, not enqueue time."""
    ctx, dev, _ = gpu_ctx
    t = T.IMAGE_2D | T.RGBA8 | M
    dim = (1024, 1024)
    res = {}
    for mode in ("default", "overlap", "overlap+fence"):
        q = ctx.create_queue(dev)
        if mode != "default":
            q.set_mip_chain_overlap(True)
        blocker = ctx.create_image(q, (8192, 8192), T.IMAGE_2D | T.RGBA16F | M)
        blocker.fill_synthetic(q, 2, 0)
        imgs = [ctx.create_image(q, dim, t) for _ in range(16)]
        for i, im in enumerate(imgs):
            im.fill_synthetic(q, 1, i)
        best = 1e9
        for rep in range(4):
            q.finish()
            for _ in range(56):
                blocker.enqueue_mip_map_chain(q)   # ~6 ms of GPU work: the 128 small chains below are enqueued long before it ends
            e0 = q.record_event()
            for k in range(8 * len(imgs)):
                if mode == "overlap+fence":
                    q.fence()
                imgs[k % len(imgs)].enqueue_mip_map_chain(q)
            e1 = q.record_event()
            best = min(best, q.elapsed_ms(e0, e1) * 1e3 / (8 * len(imgs)))
        res[mode] = best
        l0 = oracle_mod.fill_synthetic(dim, t, 1, layer_id0=5)
        assert np.array_equal(imgs[5].download_levels(q), oracle_mod.generate_mip_map_chain(l0, dim, t, threads=8))
        for im in imgs + [blocker]:
            im.destroy()
        q.destroy()
    print("us per chain:", res)
    assert res["overlap"] < 0.8 * res["default"], res
    assert res["overlap+fence"] > 0.9 * res["default"], res


def test_overlap_across_queues_and_batches(gpu_ctx, overlap_queue, oracle_mod):
    """an image whose chain moves between an overlapping queue and a plain one, and a batch graph on the overlapping queue"""
    ctx, dev, q0 = gpu_ctx
    q = overlap_queue
    t = T.IMAGE_2D | T.RGBA8 | M
    dims = [(1024, 512), (512, 512), (256, 1024), (333, 200)]
    imgs, wants = [], []
    for i, dim in enumerate(dims):
        im = ctx.create_image(q, dim, t)
        l0 = oracle_mod.fill_synthetic(dim, t, 50 + i)
        im.upload_levels(q, l0, 0, 0)
        imgs.append(im); wants.append(oracle_mod.generate_mip_map_chain(l0, dim, t, threads=8))
    batch = ctx.create_mip_chain_batch(imgs[:2])
    for rep in range(10):
        for im in imgs:
            im.enqueue_mip_map_chain(q)
        imgs[0].enqueue_mip_map_chain(q0)   # hand-over to the plain queue ...
        imgs[0].enqueue_mip_map_chain(q)    # ... and back
        batch.enqueue(q)
        imgs[3].enqueue_mip_map_chain(q)
    q0.finish()
    for im, want in zip(imgs, wants):
        assert np.array_equal(im.download_levels(q), want)
    batch.destroy()
    for im in imgs:
        im.destroy()


def test_overlap_with_several_host_threads_on_one_queue(gpu_ctx, overlap_queue, oracle_mod):
    """four host threads enqueue chains on their own images, fences and events on ONE overlapping queue at the same time: the bookkeeping
    sees chains and closers in one order per stream (stream_run::mtx), every image ends bit-exact"""
    import threading
    ctx, dev, _ = gpu_ctx
    q = overlap_queue
    t = T.IMAGE_2D | T.RGBA8 | M
    dims = [(1024, 1024), (512, 1024), (1000, 600), (2048, 512), (256, 256), (1024, 512), (640, 480), (512, 512)]
    imgs, wants = [], []
    for i, dim in enumerate(dims):
        im = ctx.create_image(q, dim, t)
        l0 = oracle_mod.fill_synthetic(dim, t, 80 + i)
        im.upload_levels(q, l0, 0, 0)
        imgs.append(im); wants.append(oracle_mod.generate_mip_map_chain(l0, dim, t, threads=8))
    errors = []

    def worker(k):
        try:
            rng = np.random.default_rng(k)
            mine = imgs[2 * k: 2 * k + 2]
            for step in range(400):
                r = rng.random()
                if r < 0.05:
                    q.fence()
                elif r < 0.08:
                    floor_b200.lib().flmip_event_destroy(dev.index, q.record_event())
                mine[int(rng.integers(0, 2))].enqueue_mip_map_chain(q)
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    [th.start() for th in threads]
    [th.join() for th in threads]
    assert not errors, errors
    for im, want, dim in zip(imgs, wants, dims):
        assert np.array_equal(im.download_levels(q), want), dim
    for im in imgs:
        im.destroy()
