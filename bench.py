#!/usr/bin/env python
"""bench.py -- mip-chain throughput of the B200-native path (and of the CPU reference arm).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c3|c4|c5] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full mip chain (device_image::generate_mip_map_chain) over one batch of synthetic input.
Workload at every N: one BASELINE `configs[1]` image (8192x8192 RGBA16F 2D, 14 levels) per GPU -- the path does not
split a single 2D image ("a single 2D image stays on one GPU"), so N GPUs run N independent textures (weak scaling,
no collective, NCCL only carries the barrier and the max-over-ranks of the device time).  `--workload c3|c4` run the
layered configs with layers / cubes sharded across ranks instead.

Metric = algorithmic bytes (level 0 read once + every generated level written once) per second, SURVEY.md 8(d).
"""
from __future__ import annotations

import argparse
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


def algorithmic_bytes(oracle_free_sizes):
    return int(sum(oracle_free_sizes))


def run_ours(args, rank, world, local_rank):
    import floor_b200
    desc, dim, t, sharded, cid = WORKLOADS[args.workload]
    ctx = floor_b200.device_context()
    dev = ctx.get_device(local_rank)
    q = ctx.create_queue(dev)
    lib = floor_b200.lib()

    # per-rank image: the whole workload image, or this rank's contiguous layer range of it
    layer_id0 = 0
    rdim = list(dim)
    if sharded:
        is_cube = bool(t & T.FLAG_CUBE)
        total_layers = args.layers or dim[2]
        lo, n = shard_layers(total_layers, world, rank)
        rdim[2] = n
        layer_id0 = lo * (6 if is_cube else 1)
    n_images = 2 if not sharded else 1  # images of the end-to-end leg (one queue each)
    images = [ctx.create_image(q, tuple(rdim), t) for _ in range(n_images)]
    for i, im in enumerate(images):
        im.fill_synthetic(q, cid, layer_id0 if sharded else rank * n_images + i)
    q.finish()
    img = images[0]
    # resident leg: rotate over enough images that a step never finds its input (or a previous output) in the 126 MB L2:
    # at least two, and at least 2 x L2 worth of them for the small workloads (C1: 46 images of 5.6 MB)
    L2_BYTES = 126 << 20
    n_rot = n_images if sharded else int(min(64, max(2, -(-2 * L2_BYTES // img.image_data_size_mip_maps))))
    rot = list(images)
    for i in range(len(rot), n_rot):
        im = ctx.create_image(q, tuple(rdim), t)
        im.fill_synthetic(q, cid, rank * n_rot + i)
        rot.append(im)
    q.finish()
    level0 = img.levels[0]["size"]
    alg_bytes = img.image_data_size_mip_maps  # level 0 read once + levels >= 1 written once
    texels_in = level0 // img.get_bytes_per_pixel()
    plan = img.plan()

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_
        torch.cuda.set_device(local_rank)
        dist_.init_process_group("nccl")
        dist = dist_

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()
        q.finish()

    # ---- resident (HBM -> HBM) timing ----
    for i in range(max(args.warmup, min(n_rot, 8))):
        rot[i % n_rot].enqueue_mip_map_chain(q)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.flmip_launch_count()
    barrier()
    ev0 = q.record_event()
    for i in range(args.steps):
        rot[(i + 1) % n_rot].enqueue_mip_map_chain(q)
    ev1 = q.record_event()
    ms = q.elapsed_ms(ev0, ev1)
    barrier()
    launches = lib.flmip_launch_count() - launches0
    clocks = None
    if rank == 0:
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
    barrier()

    # ---- end to end through the public API with host buffers (pinned), H2D + chain + D2H every step ----
    e2e_steps = max(4, min(args.steps, 20))
    if level0 > (4 << 30) or args.no_e2e:
        e2e_steps = 0  # layered multi-GB shards: no host staging buffer of that size; e2e is reported for the default workload
    h2d = level0
    d2h = alg_bytes - level0
    e2e_ms = e2e_blocking_ms = float("nan")
    if e2e_steps:
        pin_in = floor_b200.pinned_buffer(h2d, local_rank)
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
        barrier()
        t0 = q.record_event()
        for i in range(e2e_steps):
            e2e_enqueue(0)
            q.finish()
        t1 = q.record_event()
        e2e_blocking_ms = q.elapsed_ms(t0, t1) / e2e_steps
        barrier()
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
        barrier()

    # max over ranks of the device time
    ms_all, e2e_all = ms, e2e_ms
    if dist is not None:
        import torch
        tt = torch.tensor([ms, e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_all, e2e_all = float(tt[0]), float(tt[1])
        tb = torch.tensor([float(alg_bytes), float(texels_in)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        total_bytes, total_texels = float(tb[0]), float(tb[1])
    else:
        total_bytes, total_texels = float(alg_bytes), float(texels_in)

    ms_per_step = ms_all / args.steps
    value = total_bytes / (ms_per_step * 1e-3) / 1e9
    peak, peak_src = measured_peak_gbs()
    achieved = alg_bytes / (ms / args.steps * 1e-3) / 1e9  # this rank's kernel: algorithmic bytes per launch / avg launch duration

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = run_cpu(args.workload, steps=None, warmup=0)

    if rank == 0:
        out = {
            "metric": "mip_chain_throughput", "value": round(value, 3), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 6), "higher_is_better": True, "scaling": "strong" if sharded else "weak",
            "vs_baseline": None, "dtype": DTYPE[args.workload], "data": "synthetic (counter-based splitmix64, SURVEY 8d)",
            "config": {"workload": desc + (f"; one such image per GPU ({world} independent textures)" if not sharded and world > 1 else ""),
                       "levels": img.mip_level_count, "algorithmic_bytes_per_gpu_step": alg_bytes, "mtexels_in_per_s": round(total_texels / (ms_per_step * 1e-3) / 1e6, 1),
                       "single_pass": plan["single_pass"], "launches_per_step": plan["launches"],
                       "l2": ("working set per step (%.0f MB) exceeds the 126 MB L2" % (alg_bytes / 1e6) if alg_bytes > L2_BYTES else
                              "inputs rotate over %d images (%.0f MB in total, more than twice the 126 MB L2)" % (n_rot, n_rot * alg_bytes / 1e6))
                             + ("; steps rotate over %d images" % n_rot if n_rot > 1 and alg_bytes > L2_BYTES else ""),
                       "parallelism": "independent images per GPU, no collective" if not sharded else "contiguous layer ranges per GPU, no collective"},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": dram_traffic_per_launch(args.workload), "peak_source": peak_src,
                         "kernel": ("flmip_fast%dd_*" if plan["single_pass"] else "flmip_tile%dd_*") % (3 if (t >> 16) & 3 == 3 else 2), "algorithmic_bytes_per_launch": alg_bytes},
            "e2e": None if not e2e_steps else {"value": round(total_bytes / (e2e_all * 1e-3) / 1e9, 3), "unit": "GB/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": round(e2e_all, 4), "steps": e2e_steps, "blocking_value": round(alg_bytes / (e2e_blocking_ms * 1e-3) / 1e9, 3), "blocking_ms_per_step": round(e2e_blocking_ms, 4),
                    "note": "per step: pinned host level 0 -> H2D -> chain -> D2H of all generated levels; value = non-blocking calls, one queue per image, so the read-back of step k overlaps the upload of step k+1; blocking_value = the reference's blocking semantics on one queue (this rank)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "cpu_baseline": cpu_baseline,
            "device": dev.name,
        }
        print(json.dumps(out), flush=True)
    for im in rot:
        im.destroy()
    if dist is not None:
        dist.destroy_process_group()


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


def run_reference(args, rank, world):
    if rank != 0:
        return
    base = run_cpu(args.workload, steps=args.steps, warmup=min(args.warmup, 1))
    desc = WORKLOADS[args.workload][0]
    out = {"impl": "reference", "metric": "mip_chain_throughput", "value": base["value"], "unit": "GB/s", "n_gpus": world, "steps": args.steps,
           "warmup": min(args.warmup, 1), "ms_per_step": round(base["seconds_per_step"] * 1e3, 3), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": DTYPE[args.workload], "data": "synthetic (counter-based splitmix64, SURVEY 8d)",
           "config": {"workload": desc, "note": "CPU arm on the host cores: " + ("the reference's own Host-Compute kernels (oracle/_ref)" if base["kind"] == "reference" else "Host-Compute restatement (oracle/_ref absent)")},
           "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
           "e2e": {"value": base["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layers", type=int, default=0, help="tuning runs only: override the layer / cube count of c3 / c4")
    ap.add_argument("--no-e2e", action="store_true", help="tuning runs only: skip the end-to-end leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
