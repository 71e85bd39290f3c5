#!/usr/bin/env python
"""bench.py -- mip-chain throughput of the B200-native path (and of the CPU reference arm).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c3|c4|c5] [--impl ours|reference|incumbent]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full mip chain (device_image::generate_mip_map_chain) over one batch of synthetic input.
Workload at every N: one BASELINE `configs[1]` image (8192x8192 RGBA16F 2D, 14 levels) per GPU -- the path does not
split a single 2D image ("a single 2D image stays on one GPU"), so N GPUs run N independent textures (weak scaling,
no collective, NCCL only carries the barrier and the max-over-ranks of the device time).  The same line also carries
`layered`: BASELINE configs 3 and 4 with their layers / cubes sharded across the N ranks (strong scaling: total work fixed),
timed the same way, and `parity_check`: what the timed launches wrote, compared with the oracle (whole chain for the
single images, >= 16 sampled layers / faces per GPU for the layered ones).  `--workload c3|c4|...` makes another
workload the headline of the line instead (tuning runs).

Metric = algorithmic bytes (level 0 read once + every generated level written once) per second, SURVEY.md 8(d).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from floor_b200.image_types import IMAGE_TYPE as T  # noqa: E402

M = T.FLAG_MIPMAPPED | T.READ_WRITE
WORKLOADS = {
    # name: (description, base image dim (layers = total over all GPUs for sharded ones), type, sharded over layers?, config id)
    "c1": ("C1: 1024x1024 RGBA8 UNORM 2D full mip chain", (1024, 1024), T.IMAGE_2D | T.RGBA8 | M, False, 1),
    "c2": ("C2: 8192x8192 RGBA16F 2D full mip chain", (8192, 8192), T.IMAGE_2D | T.RGBA16F | M, False, 2),
    "c3": ("C3: 2D array 2048 layers x 1024^2 RGBA8 UNORM, layers sharded across GPUs", (1024, 1024, 2048), T.IMAGE_2D_ARRAY | T.RGBA8 | M, True, 3),
    "c4": ("C4: cube array 64 x 6 x 4096^2 RGBA32F, cubes sharded across GPUs", (4096, 4096, 64), T.IMAGE_CUBE_ARRAY | T.RGBA32F | M, True, 4),
    "c5": ("C5: 3D volume 512^3 R32F, 2x2x2 minification", (512, 512, 512), T.IMAGE_3D | T.R32F | M, False, 5),
    # not BASELINE configs: non-power-of-two images take the multi-level tile kernel (flmip_tile2d / 3d)
    "n1": ("N1: 3840x2160 RGBA8 UNORM 2D full mip chain (NPOT)", (3840, 2160), T.IMAGE_2D | T.RGBA8 | M, False, 6),
    "n2": ("N2: 2D array 64 layers x 1920x1080 RGBA16F (NPOT)", (1920, 1080, 64), T.IMAGE_2D_ARRAY | T.RGBA16F | M, False, 7),
}
DTYPE = {"c1": "f32 (unorm8 storage)", "c2": "f32 (f16 storage)", "c3": "f32 (unorm8 storage)", "c4": "f32", "c5": "f32", "n1": "f32 (unorm8 storage)",
         "n2": "f32 (f16 storage)"}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy, read+write)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def dram_traffic_per_launch(workload: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu capture, if any"""
    try:
        with open(os.path.join(ROOT, "profiles", "dram_traffic.json")) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self, settle: float = 0.15):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(settle)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for n, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def shard_layers(total: int, world: int, rank: int, multiple: int = 1):
    """contiguous layer ranges: rank r gets [r*L/N, (r+1)*L/N), kept on `multiple` boundaries (6 faces per cube)"""
    units = total // multiple
    lo, hi = rank * units // world, (rank + 1) * units // world
    return lo * multiple, (hi - lo) * multiple


def shard_round_robin(total: int, world: int, rank: int):
    """batches of independent textures: texture i goes to rank i % world"""
    return list(range(rank, total, world))


L2_BYTES = 126 << 20
BATCH_TEXTURES = 512  # the batch leg: independent 1024^2 RGBA8 textures (C1's shape), round-robin over the ranks


def workload_geometry(workload: str, world: int = 1, rank: int = 0, layers_override: int = 0):
    """(per-rank image dim, global id of its first layer, levels, algorithmic bytes of the per-rank image) -- host arithmetic only
    (floor_b200.image_types mirrors image_types.hpp), so both arms print the same `config` without touching a GPU"""
    from floor_b200 import image_types as it
    desc, dim, t, sharded, cid = WORKLOADS[workload]
    rdim, layer_id0 = list(dim), 0
    if sharded:
        lo, n = shard_layers(layers_override or dim[2], world, rank)
        rdim[2] = n
        layer_id0 = lo * (6 if t & T.FLAG_CUBE else 1)
    d4 = tuple(rdim) + (0,) * (4 - len(rdim))
    return tuple(rdim), layer_id0, it.mip_level_count(d4, t), it.image_data_size(d4, t)


def config_for(workload: str, world: int, layers_override: int = 0):
    """the `config` object of the JSON line: a pure function of (workload, N), identical for `--impl ours` and `--impl reference`"""
    desc, dim, t, sharded, cid = WORKLOADS[workload]
    _, _, levels, alg = workload_geometry(workload, world, 0, layers_override)
    n_rot = 1 if sharded else int(min(64, max(2, -(-2 * L2_BYTES // alg))))
    return {"workload": desc + (f"; one such image per GPU ({world} independent textures)" if not sharded and world > 1 else ""),
            "levels": levels, "algorithmic_bytes_per_gpu_step": alg,
            "l2": ("working set per step (%.0f MB) exceeds the 126 MB L2" % (alg / 1e6) if alg > L2_BYTES else
                   "inputs rotate over %d images (%.0f MB in total, more than twice the 126 MB L2)" % (n_rot, n_rot * alg / 1e6))
                  + ("; steps rotate over %d images" % n_rot if n_rot > 1 and alg > L2_BYTES else ""),
            "parallelism": "independent images per GPU, no collective" if not sharded else "contiguous layer ranges per GPU, no collective"}


def sampled_layers(n_layers: int, count: int = 16, seed: int = 0x5EED):
    """first, last and seeded picks"""
    picks = {0, n_layers - 1}
    rng = np.random.default_rng(seed)
    while len(picks) < min(count, n_layers):
        picks.add(int(rng.integers(0, n_layers)))
    return sorted(picks)


def parity_check(img, q, workload: str, fill_id: int, layered: bool):
    """Looks at what the timed launches wrote (the oracle is only the checker here): the whole chain of a single image, or >= 16
    sampled layers / faces (first, last, seeded picks) of a layered one, every level, bit-exact against the oracle run per layer
    (device_image.cpp:304-327: layers are independent chains).  `fill_id` = config id's layer id the image was filled with."""
    import oracle
    desc, dim, t, sharded, cid = WORKLOADS[workload]
    threads = min(os.cpu_count() or 8, 32)
    mismatches, checked = 0, 0
    if not layered:
        rdim = img.image_dim[: len(dim)]
        got = img.download_levels(q)
        n0 = img.levels[0]["size"]
        l0 = oracle.fill_synthetic(rdim, t, cid, layer_id0=fill_id)
        ok = bool(np.array_equal(got[:n0], l0))
        if ok:
            want = oracle.generate_mip_map_chain(l0, rdim, t, threads=threads)
            ok = hashlib.sha256(got.tobytes()).digest() == hashlib.sha256(want.tobytes()).digest()
        return {"layers_checked": 1, "mismatches": 0 if ok else 1, "scope": "whole image, every level, sha256 vs oracle"}
    fmt_bits = t & ~(T.IMAGE_2D_ARRAY | T.IMAGE_CUBE_ARRAY | T.FLAG_CUBE | T.FLAG_ARRAY)
    t1 = T.IMAGE_2D | fmt_bits
    dim2d = dim[:2]
    picks = sampled_layers(img.layer_count)
    for layer in picks:
        got = img.download_layers(q, layer, 1)
        l0 = oracle.fill_synthetic(dim2d, t1, cid, layer_id0=fill_id + layer, layer_num=1)
        want = oracle.generate_mip_map_chain(l0, dim2d, t1, threads=threads)
        checked += 1
        if not np.array_equal(got, want):
            mismatches += 1
    return {"layers_checked": checked, "mismatches": mismatches,
            "scope": f"layers {picks[0]}, {picks[-1]} and seeded picks of this rank's {img.layer_count}, every level, bit-exact vs oracle"}


class Ranks:
    """torch.distributed (NCCL) carries the barrier and the max / sum over ranks of the timings -- nothing else"""

    def __init__(self, world, local_rank):
        self.world, self.dist, self.torch = world, None, None
        if world > 1:
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl")
            self.dist, self.torch = dist, torch

    def barrier(self, q=None):
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()
        if q is not None:
            q.finish()

    def reduce(self, values, op="max"):
        if self.dist is None:
            return [float(v) for v in values]
        tt = self.torch.tensor([float(v) for v in values], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(tt, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return [float(x) for x in tt]

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def time_resident(ctx, dev, q, ranks, rank, world, workload, steps, warmup, layers_override=0, sampler=None, keep=1, pipelined_leg=False):
    """Device-timed leg, inputs resident in HBM: creates this rank's image(s) of `workload`, fills them on the device, runs `warmup`
    untimed and EXACTLY `steps` timed chains between barriers (CUDA events on the launching stream), checks what the timed launches
    wrote against the oracle.  Returns the numbers and the first `keep` images (the rest is destroyed)."""
    import floor_b200
    lib = floor_b200.lib()
    desc, dim, t, sharded, cid = WORKLOADS[workload]
    rdim, layer_id0, levels, alg_bytes = workload_geometry(workload, world, rank, layers_override)
    # rotate over enough images that a step never finds its input (or a previous output) in the 126 MB L2: at least two, and
    # at least 2 x L2 worth of them for the small workloads (C1: 46 images of 5.6 MB); a layered shard is far larger than L2
    n_rot = 1 if sharded else int(min(64, max(2, -(-2 * L2_BYTES // alg_bytes))))
    fill_ids = [layer_id0 if sharded else rank * n_rot + i for i in range(n_rot)]
    rot = []
    for i in range(n_rot):
        im = ctx.create_image(q, rdim, t)
        im.fill_synthetic(q, cid, fill_ids[i])
        rot.append(im)
    q.finish()
    img = rot[0]
    assert img.image_data_size_mip_maps == alg_bytes and img.mip_level_count == levels
    plan = img.plan()
    for i in range(max(warmup, min(n_rot, 8))):
        rot[i % n_rot].enqueue_mip_map_chain(q)
    ranks.barrier(q)
    if sampler is not None:
        sampler.start()
    launches0 = lib.flmip_launch_count()
    ranks.barrier(q)
    ev0 = q.record_event()
    for i in range(steps):
        rot[(i + 1) % n_rot].enqueue_mip_map_chain(q)
    ev1 = q.record_event()
    ms = q.elapsed_ms(ev0, ev1)
    ranks.barrier(q)
    launches = lib.flmip_launch_count() - launches0
    clocks = None
    if sampler is not None:
        # nvidia-smi samples every 100 ms and needs as long to start, the timed region of most workloads lasts a few ms: keep the
        # same steps running (untimed) until the sampler has seen the GPU under this load for a few samples
        t_s = time.perf_counter()
        k = 0
        while len(sampler.lines) < 4 and time.perf_counter() - t_s < 2.0:
            for _ in range(16):
                rot[k % n_rot].enqueue_mip_map_chain(q)
                k += 1
            q.finish()
        clocks = sampler.stop(settle=0.0)
        clocks["window"] = ("the timed region (%.1f ms) followed by %.0f ms of the same steps, untimed, so that the 100 ms sampler sees this load"
                            % (ms, (time.perf_counter() - t_s) * 1e3))
    ranks.barrier(q)
    # per-step spread ("report median and min"): the same steps once more with an event behind every step.  Not the headline:
    # an event between two chains keeps the second from being launched as a programmatic dependent of the first, so each of these
    # steps pays its launch in full (this rank's numbers).
    evs = [q.record_event()]
    for i in range(steps):
        rot[(i + 1) % n_rot].enqueue_mip_map_chain(q)
        evs.append(q.record_event())
    q.finish()
    per = sorted(q.elapsed_ms(a, b, destroy=False) for a, b in zip(evs, evs[1:]))
    for e in evs:
        lib.flmip_event_destroy(dev.index, e)
    per_step = {"median_ms": round(per[len(per) // 2], 6), "min_ms": round(per[0], 6), "max_ms": round(per[-1], 6),
                "note": "one event behind every step: no programmatic dependent launch between consecutive chains"}
    ranks.barrier(q)
    # the same steps on a queue that lets chains on independent images overlap (flmip_stream_set_chain_overlap): the first kernel of a
    # chain starts while the chain in front of it is still finishing, and waits for it before it ends.  Reported beside `value`, which
    # stays the strictly stream-ordered number.  Rotates over 8 images: a chain waits as before when its image still has a kernel in
    # the queue's open run, i.e. for one step in 8 here.
    pipelined = None
    if pipelined_leg and not sharded and alg_bytes * 8 < 48e9:
        n_pipe = max(n_rot, 8)
        for i in range(n_rot, n_pipe):
            im = ctx.create_image(q, rdim, t)
            im.fill_synthetic(q, cid, rank * n_pipe + i + 1000)
            rot.append(im)
        q.finish()
        qo = ctx.create_queue(dev)
        qo.set_mip_chain_overlap(True)
        for i in range(max(warmup, n_pipe)):
            rot[i % n_pipe].enqueue_mip_map_chain(qo)
        ranks.barrier(qo)
        # A host thread needs 5 - 8 us to enqueue a chain (Python + cuLaunchKernelEx), which is more than the GPU needs for an overlapped
        # chain of the small workloads (c1: ~4 us): their steps are enqueued behind ~3 ms of untimed work on the same queue (chains on a
        # 716 MB image), so that the events bracket what the GPU does with the queued steps and not the enqueue loop.
        blocker = None
        if alg_bytes < 64e6:
            blocker = ctx.create_image(qo, (8192, 8192), T.IMAGE_2D | T.RGBA16F | T.FLAG_MIPMAPPED | T.READ_WRITE)
            blocker.fill_synthetic(qo, 2, 0)
            qo.finish()
            for _ in range(28):
                blocker.enqueue_mip_map_chain(qo)
        p0 = qo.record_event()
        for i in range(steps):
            rot[(i + 1) % n_pipe].enqueue_mip_map_chain(qo)
        p1 = qo.record_event()
        ms_p = qo.elapsed_ms(p0, p1)
        ranks.barrier(qo)
        ms_p_all, = ranks.reduce([ms_p], "max")
        tot, = ranks.reduce([alg_bytes], "sum")
        pipelined = {"value": round(tot / (ms_p_all / steps * 1e-3) / 1e9, 3), "unit": "GB/s", "ms_per_step": round(ms_p_all / steps, 6), "steps": steps,
                     "images_in_rotation": n_pipe,
                     "note": "the same chains on a queue with flmip_stream_set_chain_overlap: chains on independent images overlap (the next chain's "
                             "first kernel streams while the previous chain's tail finishes; completion still follows stream order); "
                             "value / roofline above are the strictly stream-ordered numbers"}
        qo.finish()
        if blocker is not None:
            pipelined["note"] += "; the steps of this small workload are enqueued behind ~3 ms of untimed work so that the events time the device, not the host's enqueue loop"
            blocker.destroy()
        qo.destroy()
    # what did the timed launches write?  image 1 % n_rot was the first one of the timed loop (both legs)
    chk_i = 1 % n_rot
    pc = parity_check(rot[chk_i], q, workload, fill_ids[chk_i], sharded)
    texels_in = img.levels[0]["size"] // img.get_bytes_per_pixel()
    ms_all, = ranks.reduce([ms], "max")
    total_bytes, total_texels, mism, checked = ranks.reduce([alg_bytes, texels_in, pc["mismatches"], pc["layers_checked"]], "sum")
    pc = dict(pc, mismatches=int(mism), layers_checked=int(checked))
    for im in rot[keep:]:
        im.destroy()
    ms_per_step = ms_all / steps
    return {"images": rot[:keep], "fill_ids": fill_ids[:keep], "rdim": rdim, "alg_bytes": alg_bytes, "levels": levels, "plan": plan, "ms_rank": ms, "ms_per_step": ms_per_step,
            "value": total_bytes / (ms_per_step * 1e-3) / 1e9, "achieved": alg_bytes / (ms / steps * 1e-3) / 1e9, "total_bytes": total_bytes,
            "mtexels_in_per_s": total_texels / (ms_per_step * 1e-3) / 1e6,
            "mtexels_out_per_s": (total_bytes / img.get_bytes_per_pixel() - total_texels) / (ms_per_step * 1e-3) / 1e6, "launches": int(launches), "clocks": clocks, "parity_check": pc, "n_rot": n_rot, "pipelined": pipelined, "per_step": per_step,
            "kernel": ("flmip_fast%dd_k*" if plan["single_pass"] else "flmip_tile%dd_k*") % (3 if (t >> 16) & 3 == 3 else 2)}


def time_batch(ctx, dev, q, ranks, rank, world, steps, warmup, peak):
    """The third way the path shards (north star: "array layers, cube faces and batches of independent textures"): BATCH_TEXTURES
    independent 1024^2 RGBA8 textures, texture i on rank i % world, no collective.  A step = the chains of all of this rank's textures,
    timed three ways on the device: stream-ordered chains on a plain queue, the same calls on a queue with chain overlap, and one CUDA
    graph per step (flmip_batch_*).  16 sampled textures per rank are compared with the oracle on every level."""
    import oracle
    desc, dim, t, _, cid = WORKLOADS["c1"]
    mine = shard_round_robin(BATCH_TEXTURES, world, rank)
    imgs = []
    for g in mine:
        im = ctx.create_image(q, dim, t)
        im.fill_synthetic(q, cid, g)
        imgs.append(im)
    q.finish()
    alg = imgs[0].image_data_size_mip_maps * len(imgs)
    blocker = ctx.create_image(q, (8192, 8192), T.IMAGE_2D | T.RGBA16F | M)
    blocker.fill_synthetic(q, 2, 0)
    qo = ctx.create_queue(dev)
    qo.set_mip_chain_overlap(True)
    batch = ctx.create_mip_chain_batch(imgs)

    def timed(Q, step_fn, behind_blocker):
        for _ in range(max(1, warmup)):
            step_fn(Q)
        ranks.barrier(Q)
        if behind_blocker:
            # a host thread needs longer to enqueue a chain than the GPU to run it: queue the steps behind ~3 ms of untimed work so that
            # the events time the device and not the enqueue loop (one graph launch per step needs no such help)
            for _ in range(28):
                blocker.enqueue_mip_map_chain(Q)
        e0 = Q.record_event()
        for _ in range(steps):
            step_fn(Q)
        e1 = Q.record_event()
        ms = Q.elapsed_ms(e0, e1) / steps
        ranks.barrier(Q)
        ms_all, = ranks.reduce([ms], "max")
        return ms_all

    def chains(Q):
        for im in imgs:
            im.enqueue_mip_map_chain(Q)

    ms_queued = timed(q, chains, True)
    ms_overlapped = timed(qo, chains, True)
    ms_graph = timed(q, lambda Q: batch.enqueue(Q), False)
    total, = ranks.reduce([alg], "sum")
    # what did the timed launches write?
    picks = sampled_layers(len(imgs))
    mism = 0
    for k in picks:
        got = imgs[k].download_levels(q)
        l0 = oracle.fill_synthetic(dim, t, cid, layer_id0=mine[k])
        if not (np.array_equal(got[: l0.size], l0) and np.array_equal(got, oracle.generate_mip_map_chain(l0, dim, t, threads=min(os.cpu_count() or 8, 32)))):
            mism += 1
    mism_all, checked = ranks.reduce([mism, len(picks)], "sum")
    batch.destroy()
    for im in imgs + [blocker]:
        im.destroy()
    qo.destroy()
    leg = lambda ms: {"value": round(total / (ms * 1e-3) / 1e9, 2), "ms_per_step": round(ms, 5), "us_per_texture": round(ms * 1e3 / len(imgs), 3),
                      "frac": round(total / world / (ms * 1e-3) / 1e9 / peak, 4)}
    return {"workload": "B1: batch of %d independent 1024x1024 RGBA8 textures, texture i on rank i %% N" % BATCH_TEXTURES, "unit": "GB/s", "scaling": "strong",
            "textures_per_gpu": len(imgs), "algorithmic_bytes_per_gpu_step": alg, "steps": steps,
            "queued": leg(ms_queued), "overlapped": leg(ms_overlapped), "graph": leg(ms_graph),
            "parity_check": {"layers_checked": int(checked), "mismatches": int(mism_all), "scope": "16 sampled textures per rank (first, last, seeded), every level, bit-exact vs oracle"},
            "note": "queued = one chain per texture on a plain queue (stream-ordered); overlapped = the same calls on a queue with flmip_stream_set_chain_overlap; "
                    "graph = one CUDA graph launch per step (flmip_batch_*); queued / overlapped steps are enqueued behind ~3 ms of untimed work so that the events time the device, not the host's enqueue loop"}


def pcie_probe(q, img, pin_in, pin_out, h2d, d2h, last, ranks, reps=6):
    """the ceiling of the end-to-end leg: the same copies through the same C-ABI calls with no kernel between them -- upload alone,
    read-back alone, and both directions at once (two queues), every rank at the same time (max over ranks)"""
    def timed(fn, Qs):
        fn(); [Q.finish() for Q in Qs]
        ranks.barrier(q)
        t0 = Qs[0].record_event()
        for _ in range(reps):
            fn()
        ends = [Q.record_event() for Q in Qs]
        ms = max(Qs[0].elapsed_ms(t0, e, destroy=False) for e in ends) / reps
        [Q.finish() for Q in Qs]
        ranks.barrier(q)
        return ms
    q2 = pin_out["queue"]
    up = lambda: img.upload_levels(q, pin_in.ptr, 0, 0, sync=False, nbytes=h2d)
    down = lambda: img.download_levels(q2, 1, last, out=pin_out["buf"].ptr, sync=False)
    ms_up, ms_down = timed(up, [q]), timed(down, [q2])
    ms_both = timed(lambda: (up(), down()), [q, q2])
    ms_up, ms_down, ms_both = ranks.reduce([ms_up, ms_down, ms_both], "max")
    return {"h2d_gbs": round(h2d / ms_up / 1e6, 2), "d2h_gbs": round(d2h / ms_down / 1e6, 2), "duplex_ms_per_step": round(ms_both, 4),
            "note": "pinned host <-> device copies of one step's bytes through flmip_image_upload / _download, no kernel; per rank, all ranks at once, max over ranks"}


def run_ours(args, rank, world, local_rank):
    import floor_b200
    desc, dim, t, sharded, cid = WORKLOADS[args.workload]
    ctx = floor_b200.device_context()
    dev = ctx.get_device(local_rank)
    q = ctx.create_queue(dev)
    ranks = Ranks(world, local_rank)
    peak, peak_src = measured_peak_gbs()

    # ---- resident (HBM -> HBM) timing of the headline workload ----
    n_images = 2 if not sharded else 1  # images of the end-to-end leg (one queue each)
    R = time_resident(ctx, dev, q, ranks, rank, world, args.workload, args.steps, args.warmup, args.layers, ClockSampler(local_rank) if rank == 0 else None,
                      keep=n_images, pipelined_leg=True)
    images, img = R["images"], R["images"][0]
    alg_bytes, level0 = R["alg_bytes"], R["images"][0].levels[0]["size"]

    # ---- end to end through the public API with host buffers (pinned), H2D + chain + D2H every step ----
    e2e_steps = max(4, min(args.steps, 20))
    if level0 > (4 << 30) or args.no_e2e:
        e2e_steps = 0  # layered multi-GB shards: no host staging buffer of that size; e2e is reported for the default workload
    h2d = level0
    d2h = alg_bytes - level0
    e2e_ms = e2e_blocking_ms = float("nan")
    pcie = None
    if e2e_steps:
        pin_in = floor_b200.pinned_buffer(h2d, local_rank, write_combined=not args.no_write_combined)  # upload only: the CPU never reads it
        pin_out = [floor_b200.pinned_buffer(max(d2h, 1), local_rank) for _ in range(n_images)]
        images[0].download_levels(q, 0, 0, out=pin_in.ptr)  # synthetic level 0 back to the host staging buffer
        q.finish()
        last = img.mip_level_count - 1
        qs = [q] + [ctx.create_queue(dev) for _ in range(n_images - 1)]

        def e2e_enqueue(k):
            im, Q = images[k % n_images], qs[k % n_images]
            im.upload_levels(Q, pin_in.ptr, 0, 0, sync=False, nbytes=h2d)
            im.enqueue_mip_map_chain(Q)
            im.download_levels(Q, 1, last, out=pin_out[k % n_images].ptr, sync=False)

        # (a) blocking, the reference's semantics: write -> chain -> read back, one image at a time
        e2e_enqueue(0); q.finish()
        ranks.barrier(q)
        t0 = q.record_event()
        for i in range(e2e_steps):
            e2e_enqueue(0)
            q.finish()
        t1 = q.record_event()
        e2e_blocking_ms = q.elapsed_ms(t0, t1) / e2e_steps
        ranks.barrier(q)
        # (b) the same steps through the non-blocking calls on one queue per image: the read-back of step k overlaps
        #     the upload of step k + 1 (PCIe is full duplex); every step still moves all of its bytes both ways
        t0 = q.record_event()
        for i in range(e2e_steps):
            e2e_enqueue(i)
            if i >= 1:
                qs[(i - 1) % n_images].finish()  # the result of step i - 1 is on the host now
        ends = [Q.record_event() for Q in qs]
        e2e_ms = max(Q.elapsed_ms(t0, e, destroy=False) for Q, e in zip(qs, ends)) / e2e_steps
        for Q in qs:
            Q.finish()
        ranks.barrier(q)
        # the result that arrived on the host in the last step is the chain of the uploaded level 0: compare with the device copy
        e2e_ok = bool(np.array_equal(pin_out[(e2e_steps - 1) % n_images].array[:d2h], images[(e2e_steps - 1) % n_images].download_levels(q, 1, last)))
        # (c) the ceiling: the same copies without the kernel
        pcie = pcie_probe(q, images[0], pin_in, {"queue": qs[-1] if n_images > 1 else ctx.create_queue(dev), "buf": pin_out[-1]}, h2d, d2h, last, ranks)
        e2e_all, = ranks.reduce([e2e_ms], "max")
    for im in images:
        im.destroy()

    # ---- BASELINE configs 3 and 4: layers / cubes sharded across the ranks (strong scaling), same timing rules ----
    layered = None
    if args.workload == "c2" and not args.no_layered:
        layered = {}
        for wl in ("c3", "c4"):
            try:
                L = time_resident(ctx, dev, q, ranks, rank, world, wl, max(3, min(args.steps, 10)), 3, keep=0)
            except floor_b200.FlmipError as e:
                layered[wl] = {"unavailable": str(e)[:200]}
                continue
            layered[wl] = {"workload": WORKLOADS[wl][0], "value": round(L["value"], 2), "unit": "GB/s", "ms_per_step": round(L["ms_per_step"], 5),
                           "steps": max(3, min(args.steps, 10)), "scaling": "strong", "layers_per_gpu": L["rdim"][2] * (6 if wl == "c4" else 1),
                           "algorithmic_bytes_per_gpu_step": L["alg_bytes"], "per_gpu_achieved": round(L["achieved"], 2), "frac": round(L["achieved"] / peak, 4),
                           "launches_per_step": L["plan"]["launches"], "kernel": L["kernel"], "parity_check": L["parity_check"],
                           "limiter": "fixed per-launch cost (launch + ring ramp-up + last-CTA tail of the group / layer stages), paid once per step whatever the shard size"}

        try:
            layered["b1"] = time_batch(ctx, dev, q, ranks, rank, world, max(3, min(args.steps, 10)), 2, peak)
        except floor_b200.FlmipError as e:
            layered["b1"] = {"unavailable": str(e)[:200]}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = run_cpu(args.workload, steps=None, warmup=0)

    # the reference's own GPU kernels on the same box (GPU over GPU): the headline workload, plus a C3 shard and C5 for the default line
    gpu_incumbent = None
    if rank == 0 and world == 1 and not args.no_incumbent:
        gpu_incumbent = {}
        for wl in ([args.workload] + (["c3", "c5"] if args.workload == "c2" else [])):
            inc = run_incumbent(wl, steps=max(3, min(args.steps, 10)), warmup=2, device=local_rank)
            if "value" in inc:
                ours = R["achieved"] if wl == args.workload else None
                if ours is None and layered and wl in layered and "per_gpu_achieved" in layered[wl]:
                    ours = layered[wl]["per_gpu_achieved"]
                if ours is not None:
                    inc["speedup_vs_blocking"] = round(ours / inc["value"], 2)
                    inc["speedup_vs_enqueued"] = round(ours / inc["enqueued_value"], 2)
            gpu_incumbent[wl] = inc

    if rank == 0:
        plan = R["plan"]
        e2e = None
        if e2e_steps:
            e2e_value = R["total_bytes"] / (e2e_all * 1e-3) / 1e9
            # ceiling of the pipelined leg: the same copies of one step, both directions at once, with no kernel between them
            pcie_peak = R["total_bytes"] / (pcie["duplex_ms_per_step"] * 1e-3) / 1e9
            e2e = {"value": round(e2e_value, 3), "unit": "GB/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "ms_per_step": round(e2e_all, 4), "steps": e2e_steps, "blocking_value": round(alg_bytes / (e2e_blocking_ms * 1e-3) / 1e9, 3),
                   "blocking_ms_per_step": round(e2e_blocking_ms, 4), "result_on_host_matches_device": e2e_ok,
                   "pcie_peak_gbs": round(pcie_peak, 3), "frac": round(e2e_value / pcie_peak, 4), "pcie": pcie,
                   "upload_buffer": "pinned" + ("" if args.no_write_combined else ", write-combined"),
                   "note": "per step: pinned host level 0 -> H2D -> chain -> D2H of all generated levels; value = non-blocking calls, one queue per image, so the read-back of step k overlaps the upload of step k+1; blocking_value = the reference's blocking semantics on one queue (this rank); pcie_peak_gbs = the same metric if a step cost only its host <-> device copies (pcie.duplex_ms_per_step: upload and read-back of one step's bytes at the same time, no kernel, all ranks copying at once): the host fabric is the ceiling, not a kernel"}
        out = {
            "metric": "mip_chain_throughput", "value": round(R["value"], 3), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(R["ms_per_step"], 6), "higher_is_better": True, "scaling": "strong" if sharded else "weak",
            "vs_baseline": None, "dtype": DTYPE[args.workload], "data": "synthetic (counter-based splitmix64, SURVEY 8d)",
            "config": config_for(args.workload, world, args.layers),
            "detail": {"mtexels_in_per_s": round(R["mtexels_in_per_s"], 1), "mtexels_out_per_s": round(R["mtexels_out_per_s"], 1), "single_pass": plan["single_pass"], "launches_per_step": plan["launches"],
                       "images_in_rotation": R["n_rot"], "per_step": R["per_step"]},
            "roofline": {"bound": "hbm", "achieved": round(R["achieved"], 2), "peak": peak, "unit": "GB/s", "frac": round(R["achieved"] / peak, 4),
                         "traffic": dram_traffic_per_launch(args.workload), "peak_source": peak_src, "kernel": R["kernel"], "algorithmic_bytes_per_launch": alg_bytes},
            "parity_check": R["parity_check"],
            "pipelined": (dict(R["pipelined"], frac=round(R["pipelined"]["value"] / world / peak, 4)) if R.get("pipelined") else None),
            "layered": layered,
            "e2e": e2e,
            "gpu_launches": R["launches"],
            "clocks": R["clocks"],
            "cpu_baseline": cpu_baseline,
            "gpu_incumbent": gpu_incumbent,
            "device": dev.name,
        }
        print(json.dumps(out), flush=True)
    ranks.close()


def run_cpu(workload: str, steps: int, warmup: int):
    """times the reference's CPU implementation of the path on the box's host cores.  kind "reference": oracle/_ref, the
    reference's own Host-Compute minify kernels + software sampler compiled from /root/reference by oracle/build_ref.py (the
    .so travels with the repo); kind "port": the C restatement (oracle/minify_oracle.c) when oracle/_ref is absent.
    One step = the full chain of the workload image (a few layers / one cube for the layered configs: the reference's
    32-bit level offsets cannot address more, and layers are independent)."""
    import oracle
    from oracle import ref
    use_ref = ref.available()
    desc, dim, t, sharded, cid = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    sample = "full workload image"
    sdim = list(dim)
    if sharded:
        sdim[2] = 8 if workload == "c3" else 1
        sample = f"{sdim[2]} of {dim[2]} {'cubes' if t & T.FLAG_CUBE else 'layers'} (layers are independent)"
    elif workload == "n2":
        sdim[2] = 8
        sample = f"8 of {dim[2]} layers (layers are independent)"
    sdim = tuple(sdim)
    l0 = oracle.fill_synthetic(sdim, t, cid)
    total = oracle.image_data_size(sdim, t)
    buf = np.zeros(total + 64, dtype=np.uint8)
    buf[: l0.size] = l0
    if use_ref:
        run = lambda: ref.generate_in_place(buf, sdim, t, threads=cores, fast=True)
        what = ("reference Host-Compute kernels (mip_map_minify.hpp + host_image.hpp compiled by g++ with the reference's release "
                "flags minus -ffast-math: -O3 -funroll-loops -march=corei7-avx -mf16c; one thread pool per launch, without libfloor's "
                "per-texel fibers)")
    else:
        run = lambda: oracle.generate_in_place(buf, sdim, t, threads=cores)
        what = "Host-Compute restatement (optimistic: omits libfloor's per-launch thread spawn and per-texel fibers)"
    for _ in range(warmup):
        run()
    if steps is None:
        # bounded sample: about 10 s of CPU work, at most 64 chains
        t0 = time.perf_counter()
        run()
        steps = int(min(64, max(1, round(10.0 / max(time.perf_counter() - t0, 1e-4)))))
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return {"value": round(total / dt / 1e9, 4), "unit": "GB/s", "cores": cores, "kind": "reference" if use_ref else "port",
            "sample": sample + f"; {steps} step(s) of {dt:.3f} s; {what}",
            "seconds_per_step": round(dt, 4), "bytes_per_step": int(total)}


def run_incumbent(workload: str, steps: int, warmup: int, device: int = 0):
    """The reference's OWN GPU path on this box (bench infrastructure, oracle/incumbent_harness.cpp): its prebuilt sm_50 PTX kernels
    (tex.level + sust, extracted from mmm.fubar into oracle/_ref) JIT-compiled for the B200, on a CUmipmappedArray, behind its own
    loop of one BLOCKING launch per (layer, level) (device_image.cpp:304-327).  `value` = that loop as an application pays it;
    `enqueued_value` = the same launches without the per-launch sync between CUDA events (the kernels alone)."""
    import oracle
    from oracle import incumbent
    desc, dim, t, sharded, cid = WORKLOADS[workload]
    if not incumbent.available():
        return {"unavailable": "oracle/_ref/mmm_incumbent.ptx or libfloor_incumbent.so missing (python oracle/build_incumbent.py needs /root/reference)"}
    if t & T.FLAG_CUBE:
        return {"unavailable": "the reference has no minify kernel for cube images (mip_map_minify.hpp:95-97; lookup fails device_image.cpp:278-283)"}
    sdim = list(dim)
    sample = "full workload image"
    if sharded or workload == "n2":
        sdim[2] = min(dim[2], 64)
        sample = f"{sdim[2]} of {dim[2]} layers (the reference loops layers outermost: cost is linear in the layer count)"
    sdim = tuple(sdim)
    l0 = oracle.fill_synthetic(sdim, t, cid)
    total = oracle.image_data_size(sdim, t)
    out = np.zeros(total, dtype=np.uint8)
    try:
        blocking_ms, enq_ms, launches = incumbent.run(sdim, t, l0, warmup, steps, device, out)
    except RuntimeError as e:
        return {"unavailable": str(e)[:300]}
    want = oracle.generate_mip_map_chain(l0, sdim, t, threads=min(os.cpu_count() or 8, 32))
    n0 = l0.size
    differing = float(np.count_nonzero(out[n0:] != want[n0:])) / max(total - n0, 1)
    return {"value": round(total / blocking_ms / 1e6, 2), "unit": "GB/s", "ms_per_step": round(blocking_ms, 4), "launches": int(launches),
            "enqueued_value": round(total / enq_ms / 1e6, 2), "enqueued_ms_per_step": round(enq_ms, 4), "bytes_per_step": int(total), "sample": sample,
            "steps": steps, "kind": "reference's prebuilt sm_50 PTX (mmm.fubar binary #7), JIT for this GPU with MAX_REGISTERS=32 / O4, CUmipmappedArray + tex.level.* + sust.b.*, "
                                    "one blocking launch per (layer, level)",
            "bytes_differing_from_host_compute": round(differing, 6),
            "note": "not the parity target: the texture unit filters with 9-bit fixed-point weights (SURVEY 8a row 14)"}


def run_incumbent_arm(args, rank, world):
    if rank != 0:
        return
    inc = run_incumbent(args.workload, args.steps, args.warmup, int(os.environ.get("LOCAL_RANK", "0")))
    out = {"impl": "incumbent", "metric": "mip_chain_throughput", "value": inc.get("value"), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": inc.get("ms_per_step"), "higher_is_better": True, "scaling": "strong" if WORKLOADS[args.workload][3] else "weak",
           "vs_baseline": None, "dtype": DTYPE[args.workload], "data": "synthetic (counter-based splitmix64, SURVEY 8d)",
           "config": config_for(args.workload, world, args.layers), "gpu_incumbent": inc}
    print(json.dumps(out), flush=True)


def run_reference(args, rank, world):
    if rank != 0:
        return
    base = run_cpu(args.workload, steps=args.steps, warmup=args.warmup)
    out = {"impl": "reference", "metric": "mip_chain_throughput", "value": base["value"], "unit": "GB/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": round(base["seconds_per_step"] * 1e3, 3), "higher_is_better": True,
           "scaling": "strong" if WORKLOADS[args.workload][3] else "weak", "vs_baseline": None, "dtype": DTYPE[args.workload],
           "data": "synthetic (counter-based splitmix64, SURVEY 8d)",
           "config": config_for(args.workload, world, args.layers),
           "detail": {"arm": "CPU arm on the host cores: " + ("the reference's own Host-Compute kernels (oracle/_ref)" if base["kind"] == "reference" else "Host-Compute restatement (oracle/_ref absent)"),
                      "sample": base["sample"]},
           "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
           "e2e": {"value": base["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "incumbent"])
    ap.add_argument("--no-incumbent", action="store_true", help="tuning runs only: skip the reference's own GPU kernels (gpu_incumbent)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layers", type=int, default=0, help="tuning runs only: override the layer / cube count of c3 / c4")
    ap.add_argument("--no-e2e", action="store_true", help="tuning runs only: skip the end-to-end leg")
    ap.add_argument("--no-layered", action="store_true", help="tuning runs only: skip the sharded C3 / C4 legs of the default line")
    ap.add_argument("--no-write-combined", action="store_true", help="A/B: plain pinned upload buffer instead of a write-combined one")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.impl == "incumbent":
        run_incumbent_arm(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
