"""Does the traversal kernel pay for a BVH that the shading stages flushed out of L2?  Traces the same 2^21 incoherent rays on the 1 M-triangle
C5 scene (a) back to back (BVH warm in L2) and (b) after writing 1 GB (L2 flushed), CUDA events around each launch."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import rtdx  # noqa: E402
from sweep import bounce_rays  # noqa: E402

stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
sc = rtdx.scenes.sphere_in_box(int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000)
ctx = rtdx.Context(64, 64, stream=stream.cuda_stream)
ctx.upload_scene(sc); torch.cuda.synchronize()
cam = rtdx.camera_params(sc.eye, sc.center, sc.up, 1.0)
prim = rtdx.scenes.camera_rays(cam, 2048, 1024)
rays = torch.from_numpy(prim.view(np.float32).reshape(-1, 8)).cuda()
n = rays.shape[0]
hits = torch.empty((n, 5), dtype=torch.float32, device="cuda")
ctx.trace_device(rays.data_ptr(), n, hits.data_ptr()); torch.cuda.synchronize()
gen = torch.Generator(device="cuda"); gen.manual_seed(7)
inc = bounce_rays(rays, hits, gen)
flush = torch.empty(1 << 28, dtype=torch.float32, device="cuda")


def run(r, cold, reps=8):
    m = r.shape[0]
    h = torch.empty((m, 5), dtype=torch.float32, device="cuda")
    tot = 0.0
    for i in range(reps + 2):
        if cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.trace_device(r.data_ptr(), m, h.data_ptr()); e1.record(); torch.cuda.synchronize()
        if i >= 2:
            tot += e0.elapsed_time(e1)
    return tot / reps, m


for name, r in (("coherent", rays), ("incoherent", inc)):
    for cold in (False, True):
        ms, m = run(r, cold)
        print("%s %-10s %-4s  %.3f ms  %.0f Mrays/s" % (sc.name, name, "cold" if cold else "warm", ms, m / ms / 1e3), flush=True)
