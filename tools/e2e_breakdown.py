"""Host-side wall-clock of each call of the reference's per-frame loop through the C ABI (bench.py's e2e leg)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtdx
sc = rtdx.scenes.mesh_room(n=296)
ctx = rtdx.Context(1920, 1080, bounces=6)
up = ctx.upload_scene(sc)
acc = {}
def t(name, fn):
    t0 = time.perf_counter(); r = fn(); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0; return r
for k in range(23):
    if k == 3: acc.clear()
    t("set_instances", lambda: ctx.set_instances(up["descs"], up["props"]))
    t("set_camera", lambda: ctx.set_camera(up["camera"]))
    t("render_pass(launch)", lambda: ctx.render_pass(k, 1))
    t("read_output", lambda: ctx.read_output())
print({k: round(v / 20 * 1e3, 3) for k, v in acc.items()}, "ms per frame (blocking read-back); total", round(sum(acc.values()) / 20 * 1e3, 3))
import numpy as np, torch
outs = [torch.empty((1080, 1920, 4), dtype=torch.uint8, pin_memory=True).numpy() for _ in range(2)]
acc.clear()
for k in range(23):
    if k == 3: acc.clear(); t0 = time.perf_counter()
    t("set_instances", lambda: ctx.set_instances(up["descs"], up["props"]))
    t("set_camera", lambda: ctx.set_camera(up["camera"]))
    t("render_pass(launch)", lambda: ctx.render_pass(k, 1))
    t("wait_output(k-1)", lambda: ctx.wait_output())
    t("read_output_async", lambda: ctx.read_output_async(outs[k & 1]))
ctx.wait_output()
print("one frame in flight: wall %.3f ms per frame" % ((time.perf_counter() - t0) / 20 * 1e3))
print({k: round(v / 20 * 1e3, 3) for k, v in acc.items()}, "ms per frame; total", round(sum(acc.values()) / 20 * 1e3, 3))
