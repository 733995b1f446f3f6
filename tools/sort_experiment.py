"""Does ordering the rays of an incoherent queue by (direction octant, origin cell) pay?  Traces the same bounce rays of the C2 scene in
emission (pixel) order, shuffled, and sorted by keys of different widths; prints ms per 2^21-ray batch (CUDA events, 10 repeats)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtdx  # noqa: E402

stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
sc = rtdx.scenes.mesh_room(n=296)
ctx = rtdx.Context(64, 64, stream=stream.cuda_stream)
ctx.upload_scene(sc); torch.cuda.synchronize()
cam = rtdx.camera_params(sc.eye, sc.center, sc.up, 1920 / 1080)
prim = rtdx.scenes.camera_rays(cam, 1920, 1080)
rays = torch.from_numpy(prim.view(np.float32).reshape(-1, 8)).cuda()
n = rays.shape[0]
hits = torch.empty((n, 5), dtype=torch.float32, device="cuda")
gen = torch.Generator(device="cuda"); gen.manual_seed(7)


def bounce(r, h):
    o = r[:, 0:3] + r[:, 4:7] * h[:, 0:1]
    d = torch.randn((r.shape[0], 3), device="cuda", generator=gen); d = d / d.norm(dim=1, keepdim=True)
    b = torch.cat([o, torch.full((r.shape[0], 1), 1e-3, device="cuda"), d, torch.full((r.shape[0], 1), 1e4, device="cuda")], dim=1)
    return b[h[:, 4].view(torch.int32) != -1].contiguous()


def timed(r, tag):
    m = r.shape[0]
    h = torch.empty((m, 5), dtype=torch.float32, device="cuda")
    for _ in range(3):
        ctx.trace_device(r.data_ptr(), m, h.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ctx.trace_device(r.data_ptr(), m, h.data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("%-28s n=%d  %.4f ms  %.0f Mrays/s" % (tag, m, ms, m / ms / 1e3), flush=True)
    return h


def keys(r, bits):
    o = r[:, 0:3]
    lo, hi = o.min(0).values, o.max(0).values
    q = ((o - lo) / (hi - lo + 1e-9) * (1 << bits)).clamp(0, (1 << bits) - 1).to(torch.int64)
    k = torch.zeros(r.shape[0], dtype=torch.int64, device="cuda")
    for b in range(bits):
        for a in range(3):
            k |= ((q[:, a] >> b) & 1) << (3 * b + a)
    octant = (r[:, 4] > 0).to(torch.int64) | ((r[:, 5] > 0).to(torch.int64) << 1) | ((r[:, 6] > 0).to(torch.int64) << 2)
    return (octant << (3 * bits)) | k, (k << 3) | octant


def hits_box(r, lo, hi):
    """slab test of rays against one world box (the displaced sphere of the C2 scene)"""
    inv = 1.0 / r[:, 4:7]
    t0 = (torch.tensor(lo, device="cuda") - r[:, 0:3]) * inv; t1 = (torch.tensor(hi, device="cuda") - r[:, 0:3]) * inv
    tn = torch.minimum(t0, t1).max(1).values.clamp(min=1e-3); tf = torch.maximum(t0, t1).min(1).values
    return tn <= tf


ctx.trace_device(rays.data_ptr(), n, hits.data_ptr()); torch.cuda.synchronize()
timed(rays, "primaries")
b1 = bounce(rays, hits)
for gen_i in range(3):
    h1 = timed(b1, "bounce %d emission order" % (gen_i + 1))
    timed(b1[torch.randperm(b1.shape[0], device="cuda", generator=gen)].contiguous(), "bounce %d shuffled" % (gen_i + 1))
    heavy = hits_box(b1, (-1.75, 0.25, -1.75), (1.75, 3.75, 1.75))
    print("   rays that hit the mesh's box: %.1f %%" % (100.0 * heavy.float().mean().item()))
    timed(b1[torch.argsort(heavy.to(torch.int32), descending=True, stable=True)].contiguous(), "bounce %d heavy rays FIRST" % (gen_i + 1))
    timed(b1[torch.argsort(heavy.to(torch.int32), descending=False, stable=True)].contiguous(), "bounce %d heavy rays LAST" % (gen_i + 1))
    # finer cost classes inside the heavy rays: chord length of the segment inside the box, origin inside the box
    lo_t, hi_t = torch.tensor((-1.75, 0.25, -1.75), device="cuda"), torch.tensor((1.75, 3.75, 1.75), device="cuda")
    inv = 1.0 / b1[:, 4:7]
    t0 = (lo_t - b1[:, 0:3]) * inv; t1 = (hi_t - b1[:, 0:3]) * inv
    tn = torch.minimum(t0, t1).max(1).values.clamp(min=1e-3); tf = torch.maximum(t0, t1).min(1).values
    chord = (tf - tn).clamp(min=0) * heavy
    inside = ((b1[:, 0:3] > lo_t) & (b1[:, 0:3] < hi_t)).all(1)
    timed(b1[torch.argsort(chord, descending=True, stable=True)].contiguous(), "bounce %d by chord in box, desc" % (gen_i + 1))
    cls = heavy.to(torch.int32) + (heavy & inside).to(torch.int32)
    timed(b1[torch.argsort(cls, descending=True, stable=True)].contiguous(), "bounce %d 3 classes (inside first)" % (gen_i + 1))
    cls2 = heavy.to(torch.int32) + (heavy & ~inside).to(torch.int32)
    timed(b1[torch.argsort(cls2, descending=True, stable=True)].contiguous(), "bounce %d 3 classes (outside first)" % (gen_i + 1))
    for bits in ():
        ko, kc = keys(b1, bits)
        timed(b1[torch.argsort(ko)].contiguous(), "bounce %d sorted oct|cell%d" % (gen_i + 1, bits))
        timed(b1[torch.argsort(kc)].contiguous(), "bounce %d sorted cell%d|oct" % (gen_i + 1, bits))
    b1 = bounce(b1, h1)
