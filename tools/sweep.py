"""Measurements of the BASELINE.json configurations that are not the bench workload (C1, C3, C5), on one B200.
Prints one JSON line per measurement; run under gpurun and keep the output under profiles/.

  python tools/sweep.py c1            Cornell 256x256, 16 spp, depth 4, Lambert only (+ the CPU port on the full image)
  python tools/sweep.py c3            10 BLAS x ~10k triangles x 1000 instances, 3840x2160, bounces 3, emissive instances + NEE
  python tools/sweep.py c5 [sizes]    rtx_trace throughput: coherent pinhole primaries vs incoherent cosine bounces, 10K..50M triangles

All timings are CUDA events on the engine's stream after warm-up; per-ray node/triangle/instance counts come from the
instrumented traversal variant on the same rays (untimed).  Roofline arithmetic as in bench.py / DESIGN.md section 5."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtdx  # noqa: E402


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def render_config(name, sc, W, H, bounces, flags, spp_per_pass, passes, warm=2):
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    t0 = time.time()
    ctx = rtdx.Context(W, H, bounces=bounces, flags=flags, samples_per_pass=spp_per_pass, stream=stream.cuda_stream)
    up = ctx.upload_scene(sc)
    torch.cuda.synchronize()
    t_up = time.time() - t0
    blas = [ctx.blas_info(i) for i in up["model_ids"]]
    for p in range(warm):
        ctx.render_pass(p * spp_per_pass, spp_per_pass)
    ctx.synchronize()
    ctx.reset_counters(); ctx.set_option(rtdx.OPT_TRACE_STATS, 1)
    ctx.render_pass(0, spp_per_pass); ctx.synchronize()
    s0 = ctx.counters(); ctx.set_option(rtdx.OPT_TRACE_STATS, 0)
    nc = max(s0["closest_rays"], 1)
    n_node, n_tri, n_inst = s0["nodes_visited"] / nc, s0["tris_tested"] / nc, s0["instances_entered"] / nc
    b_ray = 32 + 20 + 80 * n_node + 48 * n_tri + 64 * n_inst
    ctx.set_option(rtdx.OPT_STAGE_TIMING, 1)
    ctx.reset_counters(); ctx.reset_accum()
    acc, tot = {}, 0.0
    for p in range(passes):
        ctx.render_pass(p * spp_per_pass, spp_per_pass)
        k, t = ctx.last_pass_stage_ms()
        tot += t
        for n, v in k.items():
            acc[n] = acc.get(n, 0.0) + v
    c = ctx.counters()
    ctx.set_option(rtdx.OPT_STAGE_TIMING, 0)
    # the same passes without per-launch events: the headline number
    ctx.reset_counters(); ctx.reset_accum()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for p in range(passes):
        ctx.render_pass(p * spp_per_pass, spp_per_pass)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    c2 = ctx.counters()
    rays = c2["closest_rays"] + c2["shadow_rays"]
    peak, src = hbm_peak()
    closest_ms = acc.get("closest", 0.0)
    ach = c["closest_rays"] * b_ray / (closest_ms * 1e-3) / 1e9 if closest_ms else 0.0
    line = {"config": name, "scene": sc.name, "triangles": sc.n_triangles(), "instances": len(sc.instances), "width": W, "height": H,
            "bounces": bounces, "spp_per_pass": spp_per_pass, "passes": passes, "ms_per_pass": ms / passes, "Mrays_per_s": rays / ms / 1e3,
            "rays_per_path": rays / max(c2["paths"], 1), "stage_ms_per_pass": {k: v / passes for k, v in acc.items() if v > 0},
            "nodes_per_ray": n_node, "tris_per_ray": n_tri, "instances_per_ray": n_inst, "bytes_per_ray": b_ray,
            "roofline": {"kernel": "trace_kernel<closest>", "achieved_GBs": ach, "peak_GBs": peak, "peak_source": src, "frac": ach / peak},
            "as_bytes": sum(b["bytes"] for b in blas), "blas_build_ms": sum(b["build_ms"] for b in blas), "upload_s": t_up,
            "n_lights": int(up["lights"].size)}
    return ctx, up, line


def c1():
    from oracle import orc
    sc = rtdx.scenes.cornell()
    flags = rtdx.FLAG_JITTER | rtdx.FLAG_LAMBERT_ONLY
    ctx, up, line = render_config("C1", sc, 256, 256, 2, flags, 16, 8)
    ctx.reset_accum(); ctx.render_pass(0, 16); ctx.synchronize()
    gpu = ctx.read_accum()
    osc = orc.OracleScene(sc, up["props"], up["lights"])
    t0 = time.perf_counter()
    ref, octr = osc.render(up["camera"], 256, 256, 0, 16, bounces=2, flags=flags)
    dt = time.perf_counter() - t0
    rays = octr["closest_rays"] + octr["shadow_rays"]
    line["cpu_port"] = {"Mrays_per_s": rays / dt / 1e6, "seconds": dt, "rays": rays, "threads": 1}
    line["parity"] = {"accum_floats_differing": int((gpu.view(np.uint32) != ref.view(np.uint32)).sum()), "of": int(gpu.size),
                      "rmse": float(np.sqrt(np.mean((gpu[..., :3] / 16 - ref[..., :3] / 16) ** 2)))}
    print(json.dumps(line), flush=True)
    ctx.close()


def c3():
    sc = rtdx.scenes.instanced_blobs()
    ctx, up, line = render_config("C3", sc, 3840, 2160, 3, 0, 1, 6)
    print(json.dumps(line), flush=True)
    ctx.close()


def restir(width=1920, height=1080, side=296, bounces=3, frames=8):
    """The reference's own frame (3 DispatchRays: RayGen, RayGen2, RayGen3) on the C2 scene at the reference's defaults
    (1920x1080, bounces 3): ms/frame and Mrays/s, next to the E0 pass (pass 1 + accumulate only) on the same context."""
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    sc = rtdx.scenes.mesh_room(n=side)
    ctx = rtdx.Context(width, height, bounces=bounces, flags=rtdx.FLAG_RESTIR, stream=stream.cuda_stream)
    ctx.upload_scene(sc)
    out = {"config": "restir_frame", "scene": sc.name, "triangles": sc.n_triangles(), "width": width, "height": height, "bounces": bounces}
    for mode in ("e0_pass", "restir_frame"):
        step = (lambda f: ctx.render_pass(f, 1)) if mode == "e0_pass" else (lambda f: ctx.render_frame(f))
        for f in range(3):
            step(f)
        ctx.synchronize(); ctx.reset_counters()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for f in range(frames):
            step(3 + f)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / frames
        c = ctx.counters()
        rays = (c["closest_rays"] + c["shadow_rays"]) / frames
        out[mode] = {"ms": ms, "rays_per_frame": rays, "Mrays_per_s": rays / ms / 1e3, "closest_rays": c["closest_rays"] / frames,
                     "shadow_rays": c["shadow_rays"] / frames}
    d = ctx.read_restir()
    out["M_di_mean"] = float(d[..., 11].mean()); out["M_gi_mean"] = float(d[..., 23].mean())
    print(json.dumps(out), flush=True)
    ctx.close()


def bounce_rays(rays_t, hits_t, gen):
    """incoherent class: cosine-hemisphere bounces from the primary hits (seed 7), built with torch on the device."""
    o, d = rays_t[:, 0:3], rays_t[:, 4:7]
    t = hits_t[:, 0:1]
    inst = hits_t[:, 4].view(torch.int32)
    hit = inst != -1
    p = o + t * d
    n_sphere = torch.nn.functional.normalize(p, dim=1)
    ax = p.abs().argmax(dim=1)
    n_box = torch.zeros_like(p)
    n_box.scatter_(1, ax[:, None], -torch.sign(p.gather(1, ax[:, None])))
    n = torch.where((inst == 0)[:, None], n_sphere, n_box)
    n = torch.where(((n * d).sum(1, keepdim=True) > 0), -n, n)
    u1 = torch.rand(p.shape[0], device=p.device, generator=gen); u2 = torch.rand(p.shape[0], device=p.device, generator=gen)
    r = u1.sqrt(); th = 2 * np.pi * u2
    x, y, z = r * th.cos(), r * th.sin(), (1 - u1).clamp_min(0).sqrt()
    up = torch.where((n[:, 2:3].abs() < 0.999), torch.tensor([0.0, 0.0, 1.0], device=p.device), torch.tensor([1.0, 0.0, 0.0], device=p.device))
    tx = torch.nn.functional.normalize(torch.cross(up.expand_as(n), n, dim=1), dim=1)
    ty = torch.cross(n, tx, dim=1)
    nd = torch.nn.functional.normalize(x[:, None] * tx + y[:, None] * ty + z[:, None] * n, dim=1)
    out = torch.empty_like(rays_t)
    out[:, 0:3] = p + 1e-3 * n; out[:, 3] = 2e-5; out[:, 4:7] = nd; out[:, 7] = 1e4
    return out[hit].contiguous()


def time_trace(ctx, rays_t, reps=5):
    n = rays_t.shape[0]
    hits = torch.empty((n, 5), dtype=torch.float32, device="cuda")
    for _ in range(2):
        ctx.trace_device(rays_t.data_ptr(), n, hits.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ctx.trace_device(rays_t.data_ptr(), n, hits.data_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ctx.reset_counters()
    ctx.trace_device(rays_t.data_ptr(), n, hits.data_ptr(), stats=True); torch.cuda.synchronize()
    c = ctx.counters()
    return ms, hits, (c["nodes_visited"] / n, c["tris_tested"] / n, c["instances_entered"] / n)


def c5(sizes):
    peak, src = hbm_peak()
    for target in sizes:
        stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
        t0 = time.time()
        sc = rtdx.scenes.sphere_in_box(target)
        t_gen = time.time() - t0
        ctx = rtdx.Context(64, 64, stream=stream.cuda_stream)
        t0 = time.time()
        up = ctx.upload_scene(sc); torch.cuda.synchronize()
        t_up = time.time() - t0
        blas = [ctx.blas_info(i) for i in up["model_ids"]]
        S = 4096
        cam = rtdx.camera_params(sc.eye, sc.center, sc.up, 1.0)
        prim = rtdx.scenes.camera_rays(cam, S, S)                         # 2^24 pinhole primaries
        rays_t = torch.from_numpy(prim.view(np.float32).reshape(-1, 8)).cuda()
        gen = torch.Generator(device="cuda"); gen.manual_seed(7)
        for cls in ("coherent", "incoherent"):
            ms, hits, (nn, nt, ni) = time_trace(ctx, rays_t)
            n = rays_t.shape[0]
            b_ray = 32 + 20 + 80 * nn + 48 * nt + 64 * ni
            ach = n * b_ray / (ms * 1e-3) / 1e9
            print(json.dumps({"config": "C5", "class": cls, "scene": sc.name, "triangles": sc.n_triangles(), "rays": n, "ms": ms,
                              "Mrays_per_s": n / ms / 1e3, "nodes_per_ray": nn, "tris_per_ray": nt, "instances_per_ray": ni,
                              "bytes_per_ray": b_ray, "roofline": {"achieved_GBs": ach, "peak_GBs": peak, "peak_source": src, "frac": ach / peak},
                              "as_bytes": sum(b["bytes"] for b in blas), "blas_build_ms": sum(b["build_ms"] for b in blas),
                              "scene_gen_s": t_gen, "upload_s": t_up}), flush=True)
            if cls == "coherent":
                rays_t = bounce_rays(rays_t, hits, gen)
        ctx.close()
        del rays_t, hits
        torch.cuda.empty_cache()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "c5"
    if which == "c1":
        c1()
    elif which == "c3":
        c3()
    elif which == "restir":
        restir()
    else:
        sizes = [int(float(x)) for x in sys.argv[2:]] or [10_000, 100_000, 1_000_000, 10_000_000, 50_000_000]
        c5(sizes)
