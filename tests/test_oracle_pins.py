"""CPU tests (-m "not gpu"): pin the oracle.

The reference ships no tests, golden vectors or fixtures and cannot run here, so the oracle cannot be
pinned against reference outputs.  What CAN be pinned is every closed form the reference's own source fixes; each test
restates one of them independently (numpy, from the cited file:line) and compares with the C++ oracle:
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from util import host_inputs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_kat.json")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ---- F1 RandomFloat, shaders/Common_v7.hlsl:119-138 -------------------------------------------------------------
def _tea_numpy(sx, sy, n):
    v0, v1 = np.uint32(sx), np.uint32(sy)
    out = []
    with np.errstate(over="ignore"):
        for _ in range(n):
            s = np.uint32(0)
            for _ in range(4):
                s = np.uint32(s + np.uint32(0x9e3779b9))
                v0 = np.uint32(v0 + (np.uint32(np.uint32(v1 << np.uint32(4)) + np.uint32(0xA341316C)) ^ np.uint32(v1 + s) ^ np.uint32(np.uint32(v1 >> np.uint32(5)) + np.uint32(0xC8013EA4))))
                v1 = np.uint32(v1 + (np.uint32(np.uint32(v0 << np.uint32(4)) + np.uint32(0xAD90777D)) ^ np.uint32(v0 + s) ^ np.uint32(np.uint32(v0 >> np.uint32(5)) + np.uint32(0x7E95761E))))
            out.append(np.float32(v0) / np.float32(4294967296.0))
    return np.array(out, dtype=np.float32), int(v0), int(v1)


def _seed_numpy(x, y, p, s):   # shaders/Pass_init_di_v7.hlsl:63-77 with uint(time) := sample index
    M = 0xFFFFFFFF
    return ((y * 73856093) ^ (x * 19349663) ^ (p * 83492791) ^ (s * 293803)) & M, ((x * 37623481) ^ (y * 51964263) ^ (p * 68250729) ^ (s * 423977)) & M


@pytest.mark.parametrize("xy", [(0, 0), (1, 0), (0, 1), (1919, 1079)])
def test_rng_and_seed_match_independent_restatement(orc, xy):
    L = orc.lib()
    seed = (C.c_uint32 * 2)()
    L.orc_kat_seed(xy[0], xy[1], 1, 0, seed)
    assert (seed[0], seed[1]) == _seed_numpy(xy[0], xy[1], 1, 0)
    out = np.zeros(8, dtype=np.float32)
    end = (C.c_uint32 * 2)()
    L.orc_kat_rng(seed[0], seed[1], 8, _p(out), end)
    ref, e0, e1 = _tea_numpy(seed[0], seed[1], 8)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    assert (end[0], end[1]) == (e0, e1)
    assert (out >= 0).all() and (out <= 1.0).all()


def test_rng_can_return_exactly_one(orc):
    """float(v0) rounds to 2^32 for v0 >= 0xFFFFFF80, so u is in the closed interval [0,1]."""
    assert np.float32(0xFFFFFF80) / np.float32(4294967296.0) == np.float32(1.0)


def test_map_pixel_id(orc):   # shaders/Common_v7.hlsl:173-198
    L = orc.lib()
    for (w, h) in [(8, 8), (1920, 1080), (13, 7)]:
        for (x, y) in [(0, 0), (3, 3), (4, 0), (w - 1, h - 1), (5, 2)]:
            tcx = (w + 3) // 4
            ref = ((y // 4) * tcx + (x // 4)) * 16 + (y % 4) * 4 + (x % 4)
            assert L.orc_kat_map_pixel(w, h, x, y) == ref


# ---- numerics contract (oracle/det_math.h) ----------------------------------------------------------------------
def test_half_round_trip_is_ieee_rne(orc):
    rng = np.random.RandomState(1)
    x = np.concatenate([rng.normal(size=20000).astype(np.float32) * np.float32(10.0) ** rng.randint(-9, 6, size=20000).astype(np.float32),
                        np.array([0.0, -0.0, 0.6, 0.73, 65504.0, 65519.9, 65520.0, 1e9, -1e9, 6.0e-8, 2.98e-8, 2.9802322e-8, 5.9604645e-8,
                                  6.1035156e-5, 6.1e-5, np.inf, -np.inf], dtype=np.float32)])
    out = np.zeros_like(x)
    orc.lib().orc_kat_half(_p(x), x.size, _p(out))
    with np.errstate(over="ignore"):
        ref = x.astype(np.float16).astype(np.float32)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    assert out[x == np.float32(0.6)][0] == np.float32(0.60009765625) and out[x == np.float32(0.73)][0] == np.float32(0.72998046875)   # Appendix C.3


def test_sincos_accuracy_and_symmetry(orc):
    x = np.linspace(0, 2 * np.pi * 1.0001, 200001).astype(np.float32)
    s, c = np.zeros_like(x), np.zeros_like(x)
    orc.lib().orc_kat_sincos(_p(x), x.size, _p(s), _p(c))
    xd = x.astype(np.float64)
    assert np.abs(s - np.sin(xd)).max() < 2.5e-7 and np.abs(c - np.cos(xd)).max() < 2.5e-7
    assert np.abs(s * s + c * c - 1).max() < 5e-7
    z = np.zeros(1, dtype=np.float32); s0 = np.zeros(1, dtype=np.float32); c0 = np.zeros(1, dtype=np.float32)
    orc.lib().orc_kat_sincos(_p(z), 1, _p(s0), _p(c0))
    assert s0[0] == 0.0 and c0[0] == 1.0


def test_pow_for_srgb(orc):
    x = np.concatenate([np.linspace(0.0031308, 1.0, 5000), np.linspace(1.0, 50.0, 500)]).astype(np.float32)
    out = np.zeros_like(x)
    orc.lib().orc_kat_pow(_p(x), np.float32(1.0 / 2.4), x.size, _p(out))
    ref = x.astype(np.float64) ** (1.0 / 2.4)
    assert np.abs(out / ref - 1).max() < 4e-7


# ---- camera ----------------------------------------------------------------------------
def test_camera_closed_form(rtdx, orc):
    W, H = 1920, 1080
    eye, center, up = (-1.5, 1.5, 3.5), (0.0, 1.0, 0.0), (0.0, 1.0, 0.0)          # rdn/Renderer.cpp:47-48
    cam = rtdx.camera_params(eye, center, up, W / H)
    cfg = orc.OrcConfig(W, H, 3, 4, 4, 0)
    out = np.zeros(6, dtype=np.float32)
    orc.lib().orc_kat_camera_ray(C.byref(cfg), _p(cam), W // 2, H // 2, 0.0, 0.0, _p(out))
    f = np.array(center) - np.array(eye); f /= np.linalg.norm(f)
    assert np.allclose(out[:3], eye, atol=1e-6) and np.allclose(out[3:], f, atol=1e-6)     # centre pixel -> normalize(center - eye)
    orc.lib().orc_kat_camera_ray(C.byref(cfg), _p(cam), 0, 0, 0.0, 0.0, _p(out))
    s = np.cross(f, up); s /= np.linalg.norm(s); u = np.cross(s, f)
    t = np.tan(np.radians(30.0)); a = W / H
    d = -1.0 * t * a * s + 1.0 * t * u + f                                           # camera space (-t*a, +t, -1): (-1.0264, 0.5774, -1)
    assert abs(t * a - 1.0264) < 1e-4 and abs(t - 0.5774) < 1e-4
    assert np.allclose(out[3:], d / np.linalg.norm(d), atol=2e-6)
    # viewI * view = I in the HLSL reading (M[r][c] = mem[4c + r])
    V = cam["view"][0].reshape(4, 4).T; VI = cam["viewI"][0].reshape(4, 4).T
    P = cam["projection"][0].reshape(4, 4).T; PI = cam["projectionI"][0].reshape(4, 4).T
    assert np.allclose(VI @ V, np.eye(4), atol=1e-5) and np.allclose(PI @ P, np.eye(4), atol=1e-4)
    assert np.allclose(P[0, 0], 1 / (t * a), rtol=1e-6) and np.allclose(P[3, 2], -1.0)      # XMMatrixPerspectiveFovRH(60 deg, aspect, 0.1, 1000)


def _hemi(rng, n, min_cos=0.05):
    while True:
        v = rng.normal(size=3); v /= np.linalg.norm(v)
        if abs(v @ n) >= min_cos:
            return v if v @ n > 0 else -v


# ---- BSDF closed forms (shaders/GGX_v7.hlsl, include/Lambertian_v6.hlsl, shaders/BRDF_v7.hlsl) -----------------------
def _scene_for_bsdf(rtdx, orc):
    sc = rtdx.scenes.mesh_room(n=4)
    props, descs, lights, cam = host_inputs(rtdx, sc, 16, 16)
    return sc, orc.OracleScene(sc, props, lights)


def _bsdf(orc, osc, op, mat, n, i, o, seed=(1, 2)):
    out = np.zeros(4, dtype=np.float32)
    sd = (C.c_uint32 * 2)(*seed)
    a = [np.ascontiguousarray(np.asarray(v, dtype=np.float32)) for v in (n, i, o)]
    orc.lib().orc_kat_bsdf(osc.h, op, mat, _p(a[0]), _p(a[1]), _p(a[2]), sd, _p(out))
    return out, (sd[0], sd[1])


def test_bsdf_closed_forms(rtdx, orc):
    sc, osc = _scene_for_bsdf(rtdx, orc)
    rng = np.random.RandomState(3)
    h16 = lambda v: np.float64(np.float32(v).astype(np.float16))
    PI = np.float64(np.float32(3.1415))
    for mat_id in range(1, 9):
        m = sc.materials[mat_id]
        kd = np.array([h16(v) for v in m["Kd"][:3]]); ks = np.array([h16(v) for v in m["Ks"]])
        rough = h16(m["Pr_Pm_Ps_Pc"][0]); metal = h16(m["Pr_Pm_Ps_Pc"][1])
        for _ in range(20):
            n = rng.normal(size=3); n /= np.linalg.norm(n)
            o = _hemi(rng, n)
            l = _hemi(rng, n)
            inc = -l
            f0, _ = _bsdf(orc, osc, 0, mat_id, n, inc, o)
            assert np.allclose(f0[:3], kd / PI, rtol=2e-6)                                   # Lambertian_v6.hlsl:51-58 (PI = 3.1415f)
            p0, _ = _bsdf(orc, osc, 2, mat_id, n, inc, o)
            assert np.isclose(p0[0], max(n @ l, 1e-6) / PI, rtol=3e-6)                        # :61-64
            # GGX eval / pdf in float64 from GGX_v7.hlsl:26-61,174-224
            V, L = o, l; Hh = (V + L) / np.linalg.norm(V + L)
            NdV, NdL, NdH, VdH = n @ V, n @ L, n @ Hh, V @ Hh
            alpha = h16(np.float32(rough) * np.float32(rough)); a2 = alpha * alpha
            D = (rough ** 2) ** 2 / (PI * (NdH * NdH * ((rough ** 2) ** 2 - 1) + 1) ** 2)
            G2 = 2 * NdL * NdV / (NdV * np.sqrt(a2 + (1 - a2) * NdL * NdL) + NdL * np.sqrt(a2 + (1 - a2) * NdV * NdV))
            F = np.clip(ks + (1 - ks) * abs(1 - VdH) ** 5, 0, 1)
            lut = m["LUT"].astype(np.float64); x = np.clip(NdV, 0, 1) * 15; i0 = int(np.floor(x)); i1 = min(i0 + 1, 15)
            ess = lut[i0] + (x - i0) * (lut[i1] - lut[i0])
            ref = F * D * G2 / (4 * NdV * NdL) * (1 + ks * (1 - ess) / ess)
            f1, _ = _bsdf(orc, osc, 1, mat_id, n, inc, o)
            assert np.allclose(f1[:3], ref, rtol=2e-3, atol=1e-7), (mat_id, f1, ref)
            G1 = 2 * NdV / (np.sqrt(a2 + (1 - a2) * NdV * NdV) + NdV)
            p1, _ = _bsdf(orc, osc, 3, mat_id, n, inc, o)
            assert np.isclose(p1[0], G1 * D / (4 * NdV), rtol=2e-3)
            pr, _ = _bsdf(orc, osc, 4, mat_id, n, inc, o)
            ps = min(1.0, np.clip(ks + (1 - ks) * abs(1 - n @ o) ** 5, 0, 1).mean() + metal)     # BRDF_v7.hlsl:50-70
            assert np.isclose(pr[1], ps, rtol=1e-5, atol=1e-6) and np.isclose(pr[0] + pr[1], 1.0, atol=1e-6)


def test_bsdf_sampling_draw_counts_and_hemisphere(rtdx, orc):
    sc, osc = _scene_for_bsdf(rtdx, orc)
    rng = np.random.RandomState(5)
    for strategy in (0, 1):
        for k in range(50):
            n = rng.normal(size=3); n /= np.linalg.norm(n)
            o = _hemi(rng, n)
            seed = (int(rng.randint(1, 2 ** 31)), int(rng.randint(1, 2 ** 31)))
            d, end = _bsdf(orc, osc, 5 + strategy, 2, n, n, o, seed)
            assert abs(np.linalg.norm(d[:3]) - 1) < 1e-5 and d[:3] @ n >= -1e-6            # unit vector in the normal's hemisphere
            _, e0, e1 = _tea_numpy(seed[0], seed[1], 2)                                      # exactly 2 RandomFloat draws
            assert end == (e0, e1)
    # SelectSamplingStrategy: exactly 1 draw; roughness < 0.04 never picks GGX (BRDF_v7.hlsl:35)
    seed = (123, 456)
    out, end = _bsdf(orc, osc, 7, 5, (0, 0, 1), (0, 0, -1), (0, 0, 1), seed)
    _, e0, e1 = _tea_numpy(seed[0], seed[1], 1)
    assert end == (e0, e1) and out[0] in (0.0, 1.0)


# ---- traversal contract (T3/T4) ---------------------------------------------------------------------------------
def test_trace_known_answers_and_bvh_equals_brute_force(rtdx, orc):
    sc = rtdx.scenes.cornell()
    props, descs, lights, cam = host_inputs(rtdx, sc, 64, 64)
    osc = orc.OracleScene(sc, props, lights)
    rays = np.zeros(4, dtype=rtdx.ray_dt)
    rays["origin"] = [(0, 1, 0.9), (0, 1, 0.9), (0.0, 0.5, 0.9), (0, 1, 5)]
    rays["direction"] = [(0, 1, 0), (0, -1, 0), (1, 0, 0), (0, 0, 1)]
    rays["tmin"] = 1e-4; rays["tmax"] = 1e4
    h = osc.trace(rays, mode=0)
    assert np.isclose(h["t"][0], 1.0) and np.isclose(h["t"][1], 1.0) and np.isclose(h["t"][2], 1.0)   # ceiling y=2, floor y=0, right wall x=1
    assert h["inst"][3] == rtdx.MISS and h["t"][3] == np.float32(1e4)
    assert (h["inst"][:3] == 0).all()
    # barycentrics reconstruct the hit point: p = (1-u-v) v0 + u v1 + v v2
    m = sc.models[0]; V = m["vertices"]["position"]; I = m["indices"].reshape(-1, 3)
    for k in range(3):
        tri = V[I[h["prim"][k]]]
        p = (1 - h["u"][k] - h["v"][k]) * tri[0] + h["u"][k] * tri[1] + h["v"][k] * tri[2]
        assert np.allclose(p, rays["origin"][k] + h["t"][k] * rays["direction"][k], atol=1e-5)
    # strict TMin < t < TMax
    rays["tmax"] = h["t"]
    assert (osc.trace(rays, mode=0)["inst"] == rtdx.MISS).all()
    rng = np.random.RandomState(2)
    r = np.zeros(20000, dtype=rtdx.ray_dt)
    r["origin"] = rng.uniform(-1, 1, size=(20000, 3)) + (0, 1, 0)
    d = rng.normal(size=(20000, 3)); r["direction"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    r["tmin"] = 2e-5; r["tmax"] = 1e4
    a, b = osc.trace(r, mode=0), osc.trace(r, mode=1)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.array_equal(osc.trace(r, any_hit=True, mode=1)["inst"] != rtdx.MISS, a["inst"] != rtdx.MISS)


def test_instance_tree_of_the_ray_caster_equals_brute_force(rtdx, orc):
    """The oracle's ray caster walks a BVH2 over the instances' world boxes when there are more than four (BASELINE config C3 has 1000),
    nearer child first; hits, ties and any-hit answers equal the brute-force definition (mode 0), also after the instances moved and
    with coincident instances."""
    from util import random_rays
    sc = rtdx.scenes.instanced_blobs(n_models=3, n_side=5, lattice=4)                 # 64 instances
    sc.add_instance(sc.instances[0][0], np.asarray(sc.instances[0][1]).reshape(4, 4).T)   # a copy on top of instance 0: ties go to the lower id
    props, descs, lights, cam = host_inputs(rtdx, sc, 64, 64)
    osc = orc.OracleScene(sc, props, lights)
    rng = np.random.RandomState(8)
    rays = np.concatenate([rtdx.scenes.camera_rays(cam, 64, 64), random_rays(rtdx, rng, 8000, (-3, -3, -3), (3, 3, 3))])
    rays["tmax"][::7] = 1.5                                                            # short rays: the far-side culling
    a, b = osc.trace(rays, mode=0), osc.trace(rays, mode=1)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and (a["inst"] != rtdx.MISS).mean() > 0.2
    assert (a["inst"] != len(sc.instances) - 1).all()                                  # the coincident copy never wins a tie
    assert np.array_equal(osc.trace(rays, any_hit=True, mode=1)["inst"] != rtdx.MISS, a["inst"] != rtdx.MISS)
    xf = []
    for k, inst in enumerate(sc.instances):
        m = np.asarray(inst[1], dtype=np.float64).reshape(4, 4).T.copy()
        m[0, 3] += 0.4 * ((k % 3) - 1); m[2, 3] -= 0.2 * (k % 2)
        xf.append(rtdx.xmmatrix_from_colvec(m))
    moved, _ = rtdx.instance_properties(xf, [i[0] for i in sc.instances])
    osc.set_props(moved)
    a, b = osc.trace(rays, mode=0), osc.trace(rays, mode=1)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_closest_hit_tie_break_is_smallest_instance_then_primitive(rtdx, orc):
    """Two coincident copies of the same model: the contract picks the smaller instance id at equal t."""
    sc = rtdx.scenes.cornell()
    sc.add_instance(0)                       # second, identical instance
    props, descs, lights, cam = host_inputs(rtdx, sc, 32, 32)
    osc = orc.OracleScene(sc, props, lights)
    rays = rtdx.scenes.camera_rays(cam, 32, 32)
    for mode in (0, 1):
        h = osc.trace(rays, mode=mode)
        assert (h["inst"][h["inst"] != rtdx.MISS] == 0).all()


# ---- golden fixtures (generated by tests/golden/make_golden.py from this oracle; regression pins, also used by -m gpu) -----
def test_oracle_matches_committed_golden_vectors(rtdx, orc):
    with open(GOLDEN) as f:
        g = json.load(f)
    L = orc.lib()
    for k in g["rng"]:
        out = np.zeros(8, dtype=np.float32); end = (C.c_uint32 * 2)(); seed = (C.c_uint32 * 2)()
        L.orc_kat_seed(k["x"], k["y"], 1, k["sample"], seed)
        L.orc_kat_rng(seed[0], seed[1], 8, _p(out), end)
        assert [int(v) for v in out.view(np.uint32)] == k["bits"] and [seed[0], seed[1]] == k["seed"]
    sc = rtdx.scenes.cornell()
    W = H = g["cornell"]["size"]
    props, descs, lights, cam = host_inputs(rtdx, sc, W, H)
    osc = orc.OracleScene(sc, props, lights)
    acc, ctr = osc.render(cam, W, H, 0, g["cornell"]["spp"], bounces=2, flags=3)
    assert ctr["closest_rays"] == g["cornell"]["closest_rays"] and ctr["shadow_rays"] == g["cornell"]["shadow_rays"]
    assert [int(v) for v in acc.view(np.uint32).reshape(-1)] == g["cornell"]["accum_bits"]
    hits = osc.trace(rtdx.scenes.camera_rays(cam, W, H), mode=0)
    assert [int(v) for v in hits["prim"]] == g["cornell"]["primary_prim"]


def _restir_oracle_run(rtdx, orc, W, H, n_frames, bounces=2):
    sc = rtdx.scenes.cornell()
    props, descs, lights, cam = host_inputs(rtdx, sc, W, H)
    osc = orc.OracleScene(sc, props, lights)
    fr = osc.new_frames(W, H)
    acc = np.zeros((H, W, 4), dtype=np.float32)
    tot = {"closest_rays": 0, "shadow_rays": 0}
    dumps = []
    for f in range(n_frames):
        c = osc.render_frame(cam, W, H, f, fr, acc, bounces=bounces)
        tot["closest_rays"] += c["closest_rays"]; tot["shadow_rays"] += c["shadow_rays"]
        dumps.append(osc.dump_frames(fr, W, H).copy())
    osc.free_frames(fr)
    return osc, cam, acc, tot, dumps


def test_restir_oracle_matches_committed_golden(rtdx, orc):
    """the oracle's 3-pass frame (RayGen, RayGen2, RayGen3) against the committed fixture."""
    import zlib
    with open(GOLDEN) as f:
        g = json.load(f)["restir"]
    _, _, acc, tot, dumps = _restir_oracle_run(rtdx, orc, g["width"], g["height"], g["frames"], g["bounces"])
    assert tot["closest_rays"] == g["closest_rays"] and tot["shadow_rays"] == g["shadow_rays"]
    assert [int(v) for v in acc.view(np.uint32).reshape(-1)] == g["accum_bits"]
    assert int(zlib.crc32(np.ascontiguousarray(dumps[-1]).view(np.uint8).tobytes())) == g["reservoir_crc32"]


def test_restir_oracle_invariants(rtdx, orc):
    """Closed-form properties the reference source fixes for the reuse passes:
    * frame 0 has no history (zero-filled *_last buffers are invalid reservoirs): its temporal pass changes nothing, so M <= 1 + 3
      spatial candidates of M = 1 (Pass_temp_di_v7.hlsl:88-107, Common_v7.hlsl:307-322);
    * the confidence M of a DI reservoir never exceeds spatial cap arithmetic: min(16, M) + min(16, M_last) after RayGen2, and after
      RayGen3 at most 4 * 32 (Pass_temp_di_v7.hlsl:117,135-142; Pass_spat_di_v7.hlsl:95,129);
    * with a static camera every sampled pixel reprojects onto itself (Sampler_v7.hlsl:738-785), so M grows frame over frame;
    * emitter pixels bypass reuse and deliver half3(Ke) (Pass_spat_di_v7.hlsl:458-463); every pixel gets exactly one sample per frame."""
    W, H, N = 40, 32, 4
    osc, cam, acc, tot, dumps = _restir_oracle_run(rtdx, orc, W, H, N)
    kind = dumps[-1][..., 35]
    M = [d[..., 11] for d in dumps]
    sampled = kind == 2
    assert sampled.sum() > 0.3 * W * H
    assert M[0][sampled].max() <= 4 and M[0][sampled].min() >= 1
    for f in range(1, N):
        assert M[f][sampled].mean() > M[f - 1][sampled].mean()          # history is reused
        assert M[f].max() <= 4 * 32
    assert (acc[..., 3] == N).all()
    assert np.isfinite(acc).all()
    # E0 (no reuse) and the ReSTIR estimate agree on the image mean within Monte-Carlo noise
    e0, _ = osc.render(cam, W, H, 0, N, bounces=2)
    m_restir, m_e0 = acc[..., :3][sampled].mean(), e0[..., :3][sampled].mean()
    assert abs(m_restir - m_e0) < 0.25 * m_e0, (m_restir, m_e0)
    # the pairwise-MIS weights of the temporal pass sum to 1 (MIS_v7.hlsl:63-80)
    for cM, nM in [(1, 1), (1, 16), (7, 16), (16, 16)]:
        Ms = np.float32(cM + nM)
        mc = np.float32(cM) / Ms + (np.float32(nM) / Ms) * (np.float32(cM) / (np.float32(cM) + (Ms - np.float32(cM))))
        mt = ((np.float32(nM) / Ms) * (Ms - np.float32(cM))) / ((Ms - np.float32(cM)) + np.float32(cM))
        assert abs(float(mc + mt) - 1.0) < 1e-6


def test_e0_estimator_sanity(rtdx, orc):
    """The E0 image of the Cornell box: emitter pixels carry half3(Ke), walls are lit (finite, non-negative), ray counts
    obey the per-path bound 5 + bounces."""
    sc = rtdx.scenes.cornell()
    props, descs, lights, cam = host_inputs(rtdx, sc, 48, 48)
    osc = orc.OracleScene(sc, props, lights)
    acc, ctr = osc.render(cam, 48, 48, 0, 8, bounces=2, flags=3)
    img = acc[..., :3] / np.maximum(acc[..., 3:4], 1)
    assert np.isfinite(img).all()
    # the reference does not clamp cos_theta at an emitter hit from behind (Sampler_v7.hlsl:466-468), so a few samples are
    # negative (the light quad hangs 0.02 under the ceiling and is two-sided); they stay rare
    assert (img < 0).mean() < 0.01
    assert (acc[..., 3] == 8).all()
    assert img.max() == 15.0                                        # a pixel looking straight at the emitter (Ke = 15, exact in half)
    assert 0.02 < img[24:, 8:40].mean() < 2.0                        # floor region is lit
    assert ctr["closest_rays"] + ctr["shadow_rays"] <= ctr["paths"] * (5 + 2)


def test_oracle_on_the_reference_asset_scene_matches_golden(rtdx, orc):
    """The oracle on the reference's own scene (garage.obj + monke.obj through this repo's ingest, tests/golden/reference_scene.npz):
    E0 render, primary hit ids and 3 ReSTIR frames against tests/golden/reference_scene_golden.json."""
    import zlib
    from util import load_scene_npz
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    with open(os.path.join(here, "reference_scene_golden.json")) as f:
        g = json.load(f)
    sc = load_scene_npz(rtdx, os.path.join(here, "reference_scene.npz"))
    W, H = g["width"], g["height"]
    props, descs, lights, cam = host_inputs(rtdx, sc, W, H)
    assert lights.size == g["n_lights"] and sc.n_triangles() == g["triangles"]
    osc = orc.OracleScene(sc, props, lights)
    acc, ctr = osc.render(cam, W, H, 0, g["e0"]["spp"], bounces=g["bounces"])
    assert (ctr["closest_rays"], ctr["shadow_rays"]) == (g["e0"]["closest_rays"], g["e0"]["shadow_rays"])
    assert int(zlib.crc32(acc.view(np.uint8).tobytes())) == g["e0"]["accum_crc32"]
    hits = osc.trace(rtdx.scenes.camera_rays(cam, W, H), mode=1)
    assert [int(v) for v in hits["inst"]] == g["primary"]["inst"] and [int(v) for v in hits["prim"]] == g["primary"]["prim"]
    fr = osc.new_frames(W, H); acc2 = np.zeros((H, W, 4), dtype=np.float32); tot = [0, 0]
    for f in range(g["restir"]["frames"]):
        c = osc.render_frame(cam, W, H, f, fr, acc2, bounces=g["bounces"])
        tot[0] += c["closest_rays"]; tot[1] += c["shadow_rays"]
    assert tot == [g["restir"]["closest_rays"], g["restir"]["shadow_rays"]]
    assert int(zlib.crc32(acc2.view(np.uint8).tobytes())) == g["restir"]["accum_crc32"]
    assert int(zlib.crc32(np.ascontiguousarray(osc.dump_frames(fr, W, H)).view(np.uint8).tobytes())) == g["restir"]["reservoir_crc32"]


# ---- legacy estimator (include/RayGen.hlsl + include/Hit.hlsl) --------------------------------------------------
def test_legacy_oracle_matches_committed_golden(rtdx, orc):
    with open(GOLDEN) as f:
        g = json.load(f)["legacy"]
    W = H = g["size"]
    sc = rtdx.scenes.cornell()
    props, descs, lights, cam = host_inputs(rtdx, sc, W, H)
    osc = orc.OracleScene(sc, props, lights)
    acc, ctr = osc.render(cam, W, H, 0, g["spp"], bounces=g["bounces"], flags=orc.FLAG_LEGACY_RR)
    assert ctr["closest_rays"] == g["closest_rays"] and ctr["shadow_rays"] == g["shadow_rays"]
    assert [int(v) for v in acc.view(np.uint32).reshape(-1)] == g["accum_bits"]


def test_legacy_oracle_structure_and_cross_check_with_e0(rtdx, orc):
    """Closed-form properties of the legacy path loop, and the purpose SURVEY gives this row: a classic NEE + MIS + Russian-roulette
    estimator as a cross-check of E0's converged image (different estimators of the same transport, so only the means can agree)."""
    W = H = 48
    sc = rtdx.scenes.cornell()
    props, descs, lights, cam = host_inputs(rtdx, sc, W, H)
    osc = orc.OracleScene(sc, props, lights)
    spp = 8
    # (1) a path cap of 1 = primary rays only: one closest ray per path, one shadow ray per non-emitter hit, and a pixel whose samples all
    #     hit the ceiling light head-on carries exactly spp * Ke (Hit.hlsl:129-131: first-bounce emitters are returned unweighted)
    a1, c1 = osc.render(cam, W, H, 0, spp, bounces=1, flags=orc.FLAG_LEGACY_RR)
    assert c1["paths"] == W * H * spp and c1["closest_rays"] == c1["paths"] and c1["shadow_rays"] <= c1["closest_rays"]
    ke = 15.0 * spp
    assert (np.abs(a1[..., :3] - ke).max(axis=2) == 0).any()
    # (2) deterministic, and cutting the cap can only remove light (every term added to payload.emission is an abs())
    a12, c12 = osc.render(cam, W, H, 0, spp, bounces=12, flags=orc.FLAG_LEGACY_RR)
    b12, _ = osc.render(cam, W, H, 0, spp, bounces=12, flags=orc.FLAG_LEGACY_RR)
    assert np.array_equal(a12.view(np.uint32), b12.view(np.uint32))
    assert (a12[..., :3] - a1[..., :3] >= 0).all()
    # (3) Russian roulette: with the cap far away the mean path is short, and raising the cap from 12 to 40 changes (almost) nothing
    a40, c40 = osc.render(cam, W, H, 0, spp, bounces=40, flags=orc.FLAG_LEGACY_RR)
    assert c40["closest_rays"] < 1.02 * c12["closest_rays"] and c40["closest_rays"] < 6 * c40["paths"]
    # (4) cross-check with E0 (bounces 6, jitter): mean radiance of the image within 30 % (Cornell: measured 0.143 vs 0.116 at 64x64x16)
    e0, _ = osc.render(cam, W, H, 0, spp, bounces=6, flags=orc.FLAG_JITTER)
    m_leg = a12[..., :3].sum() / a12[..., 3].sum(); m_e0 = e0[..., :3].sum() / e0[..., 3].sum()
    assert 0.7 < m_leg / m_e0 < 1.45, (m_leg, m_e0)
