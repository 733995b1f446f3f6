#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json): Mrays/s (primary + secondary).

Workload at N=1: BASELINE config C2 — procedural ~1M-triangle displaced mesh in an emitter-lit room, 1920x1080,
1 spp per pass, depth 8 (bounces = 6), GGX + diffuse.  A step = one progressive pass over the whole frame.
N>1 (torchrun, one rank per GPU): the scene and its BVH are replicated, rank r renders sample s = step*N + r of every
pixel (weak scaling: one full-frame pass per GPU per step), and the accumulation buffers are combined by one NCCL
reduce per pass.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--width W --height H --tris T]

One JSON line on rank 0.  `value` = rays/s with everything resident in HBM (device-timed, max over ranks);
`e2e` = the same metric through the C ABI with host buffers (instance + camera upload, render, RGBA8 read-back per step);
`roofline` = achieved algorithmic GB/s of the closest-hit traversal kernel against the measured HBM peak;
`cpu_baseline` = the CPU oracle (oracle/, a port of the reference's shaders) on a bounded pixel subsample.
`--impl reference` times that CPU port on all host threads (the reference itself is a Windows/DX12 app and cannot run).
"""
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

METRIC = "Mrays/s (primary+secondary)"
UNIT = "Mrays/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--side", type=int, default=296, help="cube-sphere side: 12*side^2 triangles (296 -> 1,051,392)")
    ap.add_argument("--bounces", type=int, default=6)
    ap.add_argument("--cpu-step", type=int, default=2, help="pixel subsampling of the CPU baseline (every n-th pixel in x and y)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/traffic.json, written by
    tools/ncu_traffic.py); None when no capture is committed."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            t = json.load(f)
        return float(t["kernels"][kernel]["dram_bytes_per_launch"]), t.get("source", "")
    except Exception:
        return None, ""


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).  The sampler is started before the
    warm-up (nvidia-smi takes a while to come up, longer with 8 GPUs); mark_begin()/mark_end() bracket the timed region and only
    the samples whose timestamps fall inside it are reported (all samples under load if none did)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        time.sleep(0.1)
        self.proc.terminate()
        try:
            text, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            return out
        rows = []
        for line in text.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), f[5:9]))
            except ValueError:
                continue
        inside = [r for r in rows if self.t0 is not None and self.t1 is not None and self.t0 - 0.02 <= r[0] <= self.t1 + 0.02]
        used, scope = (inside, "timed region") if inside else (rows, "whole run (no sample fell inside the timed region)")
        reasons = set()
        for r in used:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if used:
            out = {"sm_mhz": float(np.median([r[1] for r in used])), "sm_max_mhz": float(max(r[2] for r in used)), "reasons": sorted(reasons),
                   "samples": len(used), "scope": scope}
        return out


def build_scene(args):
    import rtdx
    sc = rtdx.scenes.mesh_room(n=args.side, seed=1)
    return rtdx, sc


def workload_name(args, sc):
    return "C2 %s %dx%d 1spp/pass depth %d (bounces=%d) GGX+diffuse" % (sc.name, args.width, args.height, args.bounces + 2, args.bounces)


def cpu_oracle_sample(rtdx, sc, args, n_threads, cam, props, lights, first_sample=0, osc=None):
    """Times the CPU port on every cpu_step-th pixel of the same workload.  Returns (Mrays/s, rays, seconds, oracle scene)."""
    from oracle import orc
    if osc is None:
        osc = orc.OracleScene(sc, props, lights)       # BVH2 build is not timed (the GPU's BLAS build is not either)
    t0 = time.perf_counter()
    if n_threads <= 1:
        _, ctr = osc.render(cam, args.width, args.height, first_sample, 1, bounces=args.bounces, step=args.cpu_step)
    else:
        _, ctr = osc.render_threads(cam, args.width, args.height, first_sample, 1, n_threads, bounces=args.bounces, step=args.cpu_step)
    dt = time.perf_counter() - t0
    rays = ctr["closest_rays"] + ctr["shadow_rays"]
    return rays / dt / 1e6, rays, dt, osc


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (the reference is Windows/DX12-only),
    on all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    rtdx, sc = build_scene(args)
    props, descs = rtdx.instance_properties([i[1] for i in sc.instances], [i[0] for i in sc.instances])
    lights = rtdx.collect_emissive_triangles(sc)
    cam = rtdx.camera_params(sc.eye, sc.center, sc.up, args.width / args.height)
    cores = os.cpu_count() or 1
    osc = None
    for w in range(args.warmup):
        _, _, _, osc = cpu_oracle_sample(rtdx, sc, args, cores, cam, props, lights, first_sample=w, osc=osc)
    rays_tot, t_tot = 0, 0.0
    for k in range(args.steps):
        _, rays, dt, osc = cpu_oracle_sample(rtdx, sc, args, cores, cam, props, lights, first_sample=args.warmup + k, osc=osc)
        rays_tot += rays; t_tot += dt
    val = rays_tot / t_tot / 1e6
    sample = "every %d-th pixel in x and y of %dx%d, 1 spp, per step (%d rays/step)" % (args.cpu_step, args.width, args.height, rays_tot // max(args.steps, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * t_tot / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": workload_name(args, sc), "triangles": sc.n_triangles()},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    clocks = ClockSampler(local); clocks.start()          # long before the timed region: nvidia-smi needs time to come up
    rtdx, sc = build_scene(args)
    W, H = args.width, args.height
    stream = torch.cuda.Stream()            # a real (non-default) stream: the engine, the events and NCCL all use it
    torch.cuda.set_stream(stream)
    ctx = rtdx.Context(W, H, bounces=args.bounces, samples_per_pass=1, device=local, stream=stream.cuda_stream)
    up = ctx.upload_scene(sc)
    blas = [ctx.blas_info(i) for i in up["model_ids"]]

    class _Wrap:
        pass
    w = _Wrap()
    w.__cuda_array_interface__ = {"shape": (H, W, 4), "typestr": "<f4", "data": (ctx.accum_device_ptr(), False), "version": 2}
    accum = torch.as_tensor(w, device="cuda")
    total = torch.empty_like(accum) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(k):
        ctx.render_pass(k * world + rank, 1)
        if world > 1:                                 # one NCCL reduce of gPermanentData per progressive pass
            total.copy_(accum)
            dist.reduce(total, dst=0, op=dist.ReduceOp.SUM)

    # ---- per-ray traversal statistics for the roofline (instrumented kernel variant, untimed, same rays as a timed pass)
    ctx.render_pass(rank, 1); ctx.synchronize()
    ctx.reset_counters(); ctx.set_option(rtdx.OPT_TRACE_STATS, 1)
    ctx.reset_accum(); ctx.render_pass(rank, 1); ctx.synchronize()
    st = ctx.counters(); ctx.set_option(rtdx.OPT_TRACE_STATS, 0)
    n_closest = max(st["closest_rays"], 1)
    n_node, n_tri, n_inst = st["nodes_visited"] / n_closest, st["tris_tested"] / n_closest, st["instances_entered"] / n_closest
    b_ray = 32 + 20 + 80 * n_node + 48 * n_tri + 64 * n_inst

    # ---- resident timing (value)
    ctx.reset_accum()
    for k in range(args.warmup):
        step_resident(k)
    barrier(); ctx.reset_counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks.mark_begin()
    e0.record()
    for k in range(args.steps):
        step_resident(args.warmup + k)
    e1.record()
    barrier()
    clocks.mark_end()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    cnt = ctx.counters()
    # ---- the same K steps again with CUDA events around every traversal launch (per-launch durations for the roofline)
    ctx.set_option(rtdx.OPT_STAGE_TIMING, 1)
    ctx.reset_counters()
    trace_ms, pass_ms = 0.0, 0.0
    for k in range(args.steps):
        ctx.render_pass((args.warmup + k) * world + rank, 1)
        tr, tot = ctx.last_pass_ms()
        trace_ms += tr; pass_ms += tot
    cnt_t = ctx.counters()
    ctx.set_option(rtdx.OPT_STAGE_TIMING, 0)
    rays = cnt["closest_rays"] + cnt["shadow_rays"]
    t_all = torch.tensor([ms, float(rays), float(cnt["closest_rays"]), float(cnt["kernel_launches"])], device="cuda", dtype=torch.float64)
    if world > 1:
        tmax = t_all.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t_all.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms = float(tmax[0]); rays_all = float(tsum[1]); launches = int(tsum[3])
    else:
        rays_all = float(rays); launches = int(cnt["kernel_launches"])
    value = rays_all / (ms * 1e-3) / 1e6

    # ---- end to end through the C ABI with host buffers (TLAS refit + camera upload + render + RGBA8 read-back per step)
    _pins = []

    def pinned_copy(a):                     # the step's inputs live in pinned host memory
        t = torch.empty(a.nbytes, dtype=torch.uint8, pin_memory=True)
        _pins.append(t)
        n = t.numpy()
        n[:] = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
        return n.view(a.dtype).reshape(a.shape)
    props, descs, cam = pinned_copy(up["props"]), pinned_copy(up["descs"]), pinned_copy(up["camera"])
    h2d = int(props.nbytes + descs.nbytes + cam.nbytes); d2h = W * H * 4
    ctx.reset_accum()
    if world > 1 and rank == 0:
        ctx.set_resolve_source(total.data_ptr())     # rank 0 displays the reduced image, not its private partial sum

    # the frame's read-back targets (pinned host memory): one frame is kept in flight, so the RGBA8 copy of frame k overlaps the
    # rendering of frame k+1 (rtx_read_output_async / rtx_wait_output); every frame's image is complete on the host before the timed
    # region ends, and frame k's image is waited for before frame k+1's read-back is queued
    out_pinned = [torch.empty((H, W, 4), dtype=torch.uint8, pin_memory=True).numpy() for _ in range(2)]

    def step_e2e(k):
        ctx.set_instances(descs, props)
        ctx.set_camera(cam)
        ctx.render_pass(k * world + rank, 1)
        if world > 1:
            total.copy_(accum)
            dist.reduce(total, dst=0, op=dist.ReduceOp.SUM)
        ctx.wait_output()                       # frame k-1's image is on the host (the consumer may use it now)
        ctx.read_output_async(out_pinned[k & 1])

    for k in range(min(args.warmup, 2)):
        step_e2e(k)
    ctx.wait_output()
    barrier(); ctx.reset_counters()
    e0.record()
    t0 = time.perf_counter()
    for k in range(args.steps):
        step_e2e(args.warmup + k)
    ctx.wait_output()
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    c2 = ctx.counters()
    t2 = torch.tensor([e2e_ms, float(c2["closest_rays"] + c2["shadow_rays"])], device="cuda", dtype=torch.float64)
    if world > 1:
        a = t2.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
        b = t2.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
        e2e_ms = float(a[0]); e2e_rays = float(b[1])
    else:
        e2e_rays = float(t2[1])
    e2e_value = e2e_rays / (e2e_ms * 1e-3) / 1e6

    if rank == 0:
        peak, peak_src = peaks()
        n_launch = (args.bounces + 3) * args.steps      # closest-hit traversal launches in the instrumented K steps
        achieved = (cnt_t["closest_rays"] * b_ray) / (trace_ms * 1e-3) / 1e9 if trace_ms > 0 else 0.0
        traffic, traffic_src = measured_traffic("trace_kernel<closest>")
        roofline = {"bound": "hbm", "kernel": "trace_kernel<closest>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_source": "ncu --set full, dram read+write bytes per launch (profiles/traffic.json <- %s)" % traffic_src,
                    "algorithmic_bytes_per_launch": cnt_t["closest_rays"] * b_ray / n_launch, "peak_source": peak_src,
                    "bytes_per_ray": b_ray, "nodes_per_ray": n_node, "tris_per_ray": n_tri, "instances_per_ray": n_inst,
                    "launches": n_launch, "avg_launch_ms": trace_ms / n_launch, "rays_per_launch": cnt_t["closest_rays"] / n_launch, "trace_share_of_step": trace_ms / pass_ms if pass_ms else None,
                    "note": "BVH (%.1f MB) is L2-resident at this scene size; HBM peak is the conservative denominator (SURVEY.md §8d); per-launch durations are from the same K steps re-run with CUDA events around every traversal launch (RTX_OPT_STAGE_TIMING: one path range per pass, full-size launches), the timed steps of `value` run as 2 concurrent path ranges" % (sum(b["bytes"] for b in blas) / 1e6)}
        cpu = None
        if not args.no_cpu_baseline:
            v, r, dt, _ = cpu_oracle_sample(rtdx, sc, args, 1, cam, props, up["lights"])
            cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": "every %d-th pixel in x and y of %dx%d, 1 spp (%d rays in %.1f s, single thread, host has %d cores)" % (args.cpu_step, W, H, r, dt, os.cpu_count() or 0)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args, sc), "triangles": sc.n_triangles(), "parallelism": "replicated scene, samples mod %d" % world,
                       "l2": "per-step path state + queues (%.0f MB) exceed the 126 MB L2; no explicit flush" % (W * H * 440 / 1e6),
                       "rays_per_path": rays / max(cnt["paths"], 1)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / max(args.steps, 1)},
            "gpu_launches": launches, "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
            "blas": blas,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
