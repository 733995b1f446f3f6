"""Shared helpers of the parity tests: build the same inputs for the oracle (CPU) and the engine (GPU)."""
import numpy as np


def host_inputs(rtdx, scene, width, height):
    props, descs = rtdx.instance_properties([i[1] for i in scene.instances], [i[0] for i in scene.instances])
    lights = rtdx.collect_emissive_triangles(scene)
    cam = rtdx.camera_params(scene.eye, scene.center, scene.up, width / height)
    return props, descs, lights, cam


def random_rays(rtdx, rng, n, lo, hi, tmin=1e-4, tmax=1e4):
    rays = np.zeros(n, dtype=rtdx.ray_dt)
    rays["origin"] = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["direction"] = d.astype(np.float32)
    rays["tmin"] = tmin
    rays["tmax"] = tmax
    return rays


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def compare_hits(gpu, ref):
    """returns dict of mismatch counts; t/u/v compared bit-exactly on hits."""
    hit = ref["inst"] != 0xFFFFFFFF
    r = {
        "n": int(ref.size), "hits": int(hit.sum()),
        "inst": int((gpu["inst"] != ref["inst"]).sum()),
        "prim": int((gpu["prim"][hit] != ref["prim"][hit]).sum()),
        "t": int((bits(gpu["t"][hit]) != bits(ref["t"][hit])).sum()),
        "u": int((bits(gpu["u"][hit]) != bits(ref["u"][hit])).sum()),
        "v": int((bits(gpu["v"][hit]) != bits(ref["v"][hit])).sum()),
    }
    return r


def save_scene_npz(path, sc):
    """A scenes.SceneDesc as plain arrays (fixture form of an ingested OBJ scene)."""
    d = {"n_models": np.int32(len(sc.models)), "material_ids": sc.material_ids, "materials": sc.materials,
         "inst_model": np.array([i[0] for i in sc.instances], dtype=np.int32), "inst_xm": np.stack([i[1] for i in sc.instances]),
         "camera": np.array([sc.eye, sc.center, sc.up], dtype=np.float64)}
    for k, m in enumerate(sc.models):
        d["v%d" % k] = m["vertices"]; d["i%d" % k] = m["indices"]; d["o%d" % k] = np.int64(m["material_id_offset"])
    np.savez_compressed(path, **d)


def load_scene_npz(rtdx, path):
    z = np.load(path)
    sc = rtdx.scenes.SceneDesc(); sc.name = "npz"
    for k in range(int(z["n_models"])):
        sc.models.append({"vertices": z["v%d" % k], "indices": z["i%d" % k], "material_id_offset": int(z["o%d" % k])})
    sc.material_ids = z["material_ids"]; sc.materials = z["materials"]
    sc.instances = [(int(m), np.ascontiguousarray(x, dtype=np.float32)) for m, x in zip(z["inst_model"], z["inst_xm"])]
    cam = z["camera"]
    sc.eye, sc.center, sc.up = tuple(cam[0]), tuple(cam[1]), tuple(cam[2])
    return sc


def oracle_trace_threads(osc, rays, any_hit=False, mode=1, n_threads=None):
    """osc.trace over all host threads (ctypes releases the GIL; the oracle scene is read-only while tracing)."""
    import os
    import threading
    n_threads = n_threads or min(os.cpu_count() or 1, 32)
    rays = np.ascontiguousarray(rays)
    bounds = np.linspace(0, rays.size, n_threads + 1).astype(np.int64)
    parts = [None] * n_threads

    def work(i):
        parts[i] = osc.trace(rays[bounds[i]:bounds[i + 1]], any_hit=any_hit, mode=mode)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(n_threads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    return np.concatenate(parts)


def bounce_rays(rtdx, rays, hits, rng, tmin=1e-3, tmax=1e4):
    """Incoherent secondary rays: from the hit points of `rays`, uniformly random directions (seeded)."""
    hit = hits["inst"] != 0xFFFFFFFF
    o = rays["origin"][hit] + rays["direction"][hit] * hits["t"][hit][:, None]
    d = rng.normal(size=o.shape)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    out = np.zeros(o.shape[0], dtype=rtdx.ray_dt)
    out["origin"] = o.astype(np.float32); out["direction"] = d.astype(np.float32)
    out["tmin"] = tmin; out["tmax"] = tmax
    return out


def hostile_material_scene(rtdx):
    """A closed room with mirror-smooth (Pr = 0, Pr < 0.04), over-unity and zero Ks / Kd materials, the loader's default material
    (LUT = 0), huge and tiny emitters, a zero-area light triangle, missing and partly-zero vertex normals (tests/test_gpu_parity.py::
    test_render_fuzz_hostile_materials_and_lights, tests/test_ref_pins.py)."""
    rng = np.random.RandomState(11)
    sc = rtdx.scenes.SceneDesc(); sc.name = "fuzz_mat"
    mk = rtdx.scenes.make_material
    mats = [rtdx.scenes.default_material(), mk((.7, .7, .7)), mk((.2, .4, .9), ks=(.9, .9, .9), roughness=0.0, metallic=1.0),
            mk((.9, .1, .1), ks=(.04, .04, .04), roughness=0.03), mk((0, 0, 0), ks=(2.5, 1.5, 0.0), roughness=0.35, metallic=0.5),
            mk((1, 1, 1), ks=(0, 0, 0), roughness=1.0), mk((0, 0, 0), ke=(1e4, 2e4, 5e3)), mk((0, 0, 0), ke=(3e-4, 0, 0)),
            mk((.5, .5, .5), ke=(4, 4, 4), roughness=0.5)]
    sc.materials = rtdx.scenes._fill_luts(np.concatenate(mats))
    # a closed room of 6 quads (materials 1..5 and the default), two blobs, three emitters (one with a zero-area triangle)
    pos, nrm, idx, tm = rtdx.scenes._quads_to_mesh(rtdx.scenes._box_quads((-2, 0, -2), (2, 3, 2)), [1, 2, 3, 4, 5, 0], flip=True)
    room = sc.add_model(pos, nrm, idx, tm)
    p, vn, i3 = rtdx.scenes._cube_sphere(4, 0.6, 9, 0.15)
    vn[::3] = 0.0                                                           # every third vertex has no normal
    vn[1::7, 1] = 0.0                                                       # and some have one zero component (Hit_v7.hlsl:36-44: treated as missing)
    blob = sc.add_model(p, vn, i3, rng.randint(1, 6, size=i3.shape[0]))
    lq = np.array([[-.5, 2.95, -.5], [.5, 2.95, -.5], [.5, 2.95, .5], [-.5, 2.95, .5], [0, 1.5, 0], [0, 1.5, 0]], dtype=np.float32)
    light = sc.add_model(lq, np.zeros_like(lq), np.array([[0, 1, 2], [0, 2, 3], [4, 4, 5], [0, 3, 1]]), np.array([6, 8, 8, 7]))
    sc.add_instance(room); sc.add_instance(light)
    a = np.eye(4); a[:3, 3] = (-0.8, 0.8, 0.3)
    b = np.diag([1.5, 0.6, 1.0, 1.0]); b[:3, 3] = (0.9, 1.2, -0.4)
    sc.add_instance(blob, a); sc.add_instance(blob, b)
    sc.eye, sc.center, sc.up = (0.0, 1.5, 1.9), (0.0, 1.3, 0.0), (0.0, 1.0, 0.0)
    return sc
