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
