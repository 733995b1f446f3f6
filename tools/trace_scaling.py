"""Fixed cost of one traversal launch: the same shuffled incoherent rays (1 M-triangle C5 scene), traced in batches of 2^17 .. 2^22;
time = a + b n separates the per-launch cost (start-up + tail) from the per-ray cost."""
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
sc = rtdx.scenes.sphere_in_box(1_000_000)
ctx = rtdx.Context(64, 64, stream=stream.cuda_stream)
ctx.upload_scene(sc); torch.cuda.synchronize()
cam = rtdx.camera_params(sc.eye, sc.center, sc.up, 1.0)
prim = rtdx.scenes.camera_rays(cam, 4096, 2048)
rays = torch.from_numpy(prim.view(np.float32).reshape(-1, 8)).cuda()
n = rays.shape[0]
hits = torch.empty((n, 5), dtype=torch.float32, device="cuda")
ctx.trace_device(rays.data_ptr(), n, hits.data_ptr()); torch.cuda.synchronize()
gen = torch.Generator(device="cuda"); gen.manual_seed(7)
inc = bounce_rays(rays, hits, gen)
for name, r in (("ordered", inc), ("shuffled", inc[torch.randperm(inc.shape[0], device="cuda", generator=gen)].contiguous())):
    for lg in (17, 18, 19, 20, 21, 22):
        m = min(1 << lg, r.shape[0])
        sub = r[:m].contiguous()
        h = torch.empty((m, 5), dtype=torch.float32, device="cuda")
        for _ in range(3):
            ctx.trace_device(sub.data_ptr(), m, h.data_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ctx.trace_device(sub.data_ptr(), m, h.data_ptr())
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("%-9s n=2^%d  %.4f ms  %.0f Mrays/s" % (name, lg, ms, m / ms / 1e3), flush=True)
