"""Profiling driver: uploads a bench scene and renders N passes (no stats variant, no host copies). For ncu."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtdx  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="mesh")
ap.add_argument("--side", type=int, default=296)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--bounces", type=int, default=6)
ap.add_argument("--passes", type=int, default=2)
ap.add_argument("--flags", type=int, default=0)
ap.add_argument("--opt", action="append", default=[], help="engine option NAME=VALUE, e.g. PASS_PARTS=1")
a = ap.parse_args()
if a.scene == "mesh":
    sc = rtdx.scenes.mesh_room(n=a.side)
elif a.scene == "inst":
    sc = rtdx.scenes.instanced_blobs()
else:
    sc = rtdx.scenes.cornell()
ctx = rtdx.Context(a.width, a.height, bounces=a.bounces, flags=a.flags)
ctx.upload_scene(sc)
for o in a.opt:
    k, v = o.split("=")
    ctx.set_option(getattr(rtdx, "OPT_" + k), int(v, 0))
for p in range(a.passes):
    ctx.render_pass(p, 1)
ctx.synchronize()
print(ctx.counters())
