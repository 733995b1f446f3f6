"""Bring-up diagnostics on the GPU box: prints mismatch statistics between the CUDA path and the oracle."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rtdx  # noqa: E402
from oracle import orc  # noqa: E402
from util import bits, compare_hits, host_inputs, random_rays  # noqa: E402


def trace_check(name, sc, W, H, lo, hi, nrand, mode):
    t0 = time.time()
    ctx = rtdx.Context(W, H)
    up = ctx.upload_scene(sc)
    print(name, "upload+build s", round(time.time() - t0, 3), [ctx.blas_info(i) for i in up["model_ids"]], flush=True)
    osc = orc.OracleScene(sc, up["props"], up["lights"])
    rng = np.random.RandomState(7)
    rays = np.concatenate([rtdx.scenes.camera_rays(up["camera"], W, H), random_rays(rtdx, rng, nrand, lo, hi)])
    t0 = time.time(); g = ctx.trace(rays); t1 = time.time()
    ref = osc.trace(rays, mode=mode); t2 = time.time()
    r = compare_hits(g, ref)
    print(name, "trace", r, "gpu_s", round(t1 - t0, 3), "cpu_s", round(t2 - t1, 3), flush=True)
    bad = np.nonzero((g["inst"] != ref["inst"]) | (g["prim"] != ref["prim"]))[0][:5]
    for b in bad:
        print("   ray", b, rays[b], "gpu", g[b], "ref", ref[b])
    ah = ctx.trace(rays, any_hit=True)
    print(name, "anyhit mismatches", int(((ah["inst"] != rtdx.MISS) != (ref["inst"] != rtdx.MISS)).sum()), flush=True)
    return ctx, up, osc


def render_check(name, sc, W, H, spp, bounces, flags, spp_per_pass=1):
    ctx = rtdx.Context(W, H, bounces=bounces, flags=flags, samples_per_pass=spp_per_pass)
    up = ctx.upload_scene(sc)
    osc = orc.OracleScene(sc, up["props"], up["lights"])
    ctx.reset_counters()
    t0 = time.time(); ctx.render_pass(0, spp); ctx.synchronize(); t1 = time.time()
    gpu = ctx.read_accum(); cnt = ctx.counters()
    ref, octr = osc.render(up["camera"], W, H, 0, spp, bounces=bounces, flags=flags); t2 = time.time()
    mism = (bits(gpu) != bits(ref)).any(axis=-1)
    print(name, "render gpu_s", round(t1 - t0, 3), "cpu_s", round(t2 - t1, 3), "counters", cnt, "oracle", octr)
    print(name, "pixels mismatching:", int(mism.sum()), "of", mism.size, "max abs diff", float(np.nanmax(np.abs(gpu - ref))), flush=True)
    if mism.any() and spp == 1:
        ys, xs = np.nonzero(mism)
        for k in range(min(3, ys.size)):
            x, y = int(xs[k]), int(ys[k])
            g = orc.unpack_debug(ctx.debug_pixel(x, y))
            o = orc.unpack_debug(osc.debug_pixel(up["camera"], W, H, x, y, 0, bounces=bounces, flags=flags))
            print("  pixel", x, y, "gpu accum", gpu[y, x], "ref accum", ref[y, x])
            for key in ["mID", "x1", "n1", "di_x2", "di_w_sum", "di_n2", "di_W", "di_L2", "gi_xn", "gi_w_sum", "gi_W", "gi_E3", "p_hat", "C", "seed"]:
                same = np.array_equal(g[key].view(np.uint32), o[key].view(np.uint32))
                print("     %-9s %s gpu %s ref %s" % (key, "ok " if same else "BAD", g[key], o[key]))
    return ctx


if __name__ == "__main__":
    which = sys.argv[1:] or ["trace", "render"]
    if "trace" in which:
        trace_check("cornell", rtdx.scenes.cornell(), 256, 256, (-1, 0, -1), (1, 2, 1), 100000, 0)
        trace_check("mesh8", rtdx.scenes.mesh_room(n=8), 128, 128, (-5, 0.1, -5), (5, 5.9, 5), 50000, 0)
        trace_check("mesh64", rtdx.scenes.mesh_room(n=64), 128, 128, (-5, 0.1, -5), (5, 5.9, 5), 100000, 1)
        trace_check("inst", rtdx.scenes.instanced_blobs(n_models=3, n_side=6, lattice=4), 128, 128, (-2.5, -2.5, -2.5), (2.5, 2.5, 2.5), 50000, 1)
    if "render" in which:
        F = rtdx.FLAG_JITTER | rtdx.FLAG_LAMBERT_ONLY
        render_check("cornell-1spp", rtdx.scenes.cornell(), 128, 128, 1, 2, F)
        render_check("cornell-ggx-1spp", rtdx.scenes.cornell(), 128, 128, 1, 3, 0)
        render_check("mesh24-1spp", rtdx.scenes.mesh_room(n=24), 96, 64, 1, 6, 0)
        render_check("inst-1spp", rtdx.scenes.instanced_blobs(n_models=3, n_side=6, lattice=4, emissive_fraction=0.1), 96, 64, 1, 3, 0)
        render_check("cornell-16spp", rtdx.scenes.cornell(), 256, 256, 16, 2, F, spp_per_pass=4)
