"""BASELINE config C4: a 4K progressive render of S samples per pixel split across the GPUs of one box, one NCCL reduce of the
accumulation buffer (gPermanentData, 133 MB at 4K) per progressive pass.

  python tools/c4_progressive.py --spp 64                                                        (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/c4_progressive.py --gpus 8 --spp 1024

Rank r renders the samples s = r (mod G) (seeds depend on the global sample index only), keeps its private partial sum and reduces a copy
to rank 0 after every pass.  Prints one JSON line: wall/device time, Mrays/s over all ranks, the reduce's share, and — with --check K —
the largest relative difference between the G-GPU image of the first K samples and the same K samples rendered by rank 0 alone
(they differ only by the fp32 summation order)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtdx  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--spp", type=int, default=1024)
ap.add_argument("--scene", default="mesh", help="mesh (C2 scene, 1 M triangles) | inst (C3 scene, 10 M instanced triangles)")
ap.add_argument("--width", type=int, default=3840)
ap.add_argument("--height", type=int, default=2160)
ap.add_argument("--bounces", type=int, default=6)
ap.add_argument("--check", type=int, default=0)
a = ap.parse_args()
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H = a.width, a.height
sc = rtdx.scenes.mesh_room(n=296) if a.scene == "mesh" else rtdx.scenes.instanced_blobs()
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = rtdx.Context(W, H, bounces=a.bounces, device=local, stream=stream.cuda_stream)
ctx.upload_scene(sc)
import importlib  # noqa: E402
accum = importlib.import_module("royaltracer-dx_b200.dist").wrap_device_buffer(ctx.accum_device_ptr(), (H, W, 4))
total = torch.empty_like(accum)


def sync():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(n_samples, timed):
    ctx.reset_accum(); ctx.reset_counters()
    passes = (n_samples + world - 1) // world
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    red_ms = 0.0
    sync(); t0 = time.perf_counter(); ev[0].record()
    for p in range(passes):
        s = p * world + rank
        if s < n_samples:
            ctx.render_pass(s, 1)
        total.copy_(accum)
        if world > 1:
            dist.reduce(total, dst=0, op=dist.ReduceOp.SUM)
    ev[1].record(); sync()
    wall = time.perf_counter() - t0
    ms = ev[0].elapsed_time(ev[1])
    c = ctx.counters()
    t = torch.tensor([ms, float(c["closest_rays"] + c["shadow_rays"])], device="cuda", dtype=torch.float64)
    if world > 1:
        mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, rays = float(mx[0]), float(sm[1])
    else:
        rays = float(t[1])
    return ms, wall, rays, passes


for _ in range(2):                                   # warm-up passes
    ctx.render_pass(rank, 1)
sync()
ms, wall, rays, passes = run(a.spp, True)
img = total.clone()
# the reduce alone (same buffers, nothing else on the stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
sync(); e0.record()
for _ in range(10):
    total.copy_(accum)
    if world > 1:
        dist.reduce(total, dst=0, op=dist.ReduceOp.SUM)
e1.record(); sync()
reduce_ms = e0.elapsed_time(e1) / 10
out = {"config": "C4", "scene": sc.name, "width": W, "height": H, "spp": a.spp, "n_gpus": world, "passes_per_gpu": passes, "bounces": a.bounces,
       "device_ms": ms, "wall_s": wall, "Mrays_per_s": rays / ms / 1e3, "rays": rays, "ms_per_pass": ms / passes,
       "reduce_ms_per_pass": reduce_ms, "reduce_bytes": int(accum.numel() * 4)}
if a.check:
    run(a.check, False)
    multi = total.clone()
    if rank == 0:
        ctx.reset_accum()
        for s in range(a.check):
            ctx.render_pass(s, 1)
        torch.cuda.synchronize()
        single = accum.clone()
        rgb_m, rgb_s = multi[..., :3], single[..., :3]
        denom = rgb_s.abs().clamp_min(1e-3)
        d = (rgb_m - rgb_s).abs()
        # samples can be negative (an emitter hit from behind), so a pixel's sum may cancel: the absolute difference
        # against the image scale is the meaningful number, the per-pixel relative one is reported for completeness
        out["check"] = {"samples": a.check, "counts_equal": bool(torch.equal(multi[..., 3], single[..., 3])),
                        "max_abs_diff_rgb": float(d.max()), "mean_abs_diff_rgb": float(d.mean()), "max_abs_sum_rgb": float(rgb_s.abs().max()),
                        "pixels_differing": int((d.amax(dim=2) > 0).sum()), "max_rel_diff_rgb": float((d / denom).max()),
                        "mean_image": float(rgb_s.mean() / a.check)}
    sync()
if rank == 0:
    print(json.dumps(out), flush=True)
ctx.close()
if world > 1:
    dist.destroy_process_group()
