"""The per-frame loop of bench.py's `e2e` (instances + camera + pass + asynchronous read-back, one frame in flight), host wall-clock per
frame.  usage: python tools/e2e_time.py [--frames N] [--opt NAME=VALUE ...]"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtdx  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=60)
ap.add_argument("--side", type=int, default=296)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--bounces", type=int, default=6)
ap.add_argument("--tag", default="")
ap.add_argument("--opt", action="append", default=[])
a = ap.parse_args()
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
sc = rtdx.scenes.mesh_room(n=a.side)
ctx = rtdx.Context(a.width, a.height, bounces=a.bounces, stream=stream.cuda_stream)
up = ctx.upload_scene(sc)
for o in a.opt:
    k, v = o.split("=")
    ctx.set_option(getattr(rtdx, "OPT_" + k), int(v, 0))
pinned = [torch.empty((a.height, a.width, 4), dtype=torch.uint8).pin_memory().numpy() for _ in range(2)]


def frames(n, first):
    for f in range(n):
        ctx.set_instances(up["descs"], up["props"])
        ctx.set_camera(up["camera"])
        ctx.render_pass(first + f, 1)
        if f > 0:
            ctx.wait_output()
        ctx.read_output_async(pinned[f & 1])
    ctx.wait_output()


frames(6, 0)
ctx.synchronize()
t0 = time.perf_counter()
frames(a.frames, 6)
ctx.synchronize()
ms = (time.perf_counter() - t0) * 1e3 / a.frames
print("%-34s %.3f ms per frame" % (a.tag or "e2e", ms), flush=True)
ctx.close()
