"""Upper bound of what pipelining consecutive passes could buy: two contexts with the same scene render alternate samples on two streams
(no pass boundary common to both), against one context rendering all of them.  usage: python tools/two_context_experiment.py [--parts k]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtdx  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--parts", type=int, default=2)
ap.add_argument("--passes", type=int, default=42)
ap.add_argument("--contexts", type=int, default=2)
a = ap.parse_args()
sc = rtdx.scenes.mesh_room(n=296)
W, H = 1920, 1080
streams = [torch.cuda.Stream() for _ in range(a.contexts)]
ctxs = []
for s in streams:
    c = rtdx.Context(W, H, bounces=6, stream=s.cuda_stream)
    c.upload_scene(sc)
    c.set_option(rtdx.OPT_PASS_PARTS, a.parts)
    c.set_option(rtdx.OPT_PASS_PIPELINE, 0)          # (the engine's own pipelining off: this experiment is what led to it)
    ctxs.append(c)


def run(n_ctx, passes):
    for k in range(6):
        ctxs[k % n_ctx].render_pass(k, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.default_stream())
    for s in streams[:n_ctx]:
        s.wait_event(e0)
    for k in range(passes):
        ctxs[k % n_ctx].render_pass(6 + k, 1)
    for s in streams[:n_ctx]:
        ev = torch.cuda.Event(); ev.record(s); torch.cuda.default_stream().wait_event(ev)
    e1.record(torch.cuda.default_stream())
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / passes


for rep in range(2):
    print("parts=%d  one context %.3f ms/pass   %d contexts, alternate passes %.3f ms/pass" % (a.parts, run(1, a.passes), a.contexts, run(a.contexts, a.passes)), flush=True)
