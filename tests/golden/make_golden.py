"""Generates tests/golden/golden_kat.json from the CPU oracle (run in the build container: python tests/golden/make_golden.py).
The reference has no golden vectors of its own; these pin the oracle against regressions and give the
-m gpu tests committed fixtures to compare the CUDA path with (SURVEY.md §8c items i, iii, iv)."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rtdx  # noqa: E402
from oracle import orc  # noqa: E402
from util import host_inputs  # noqa: E402

L = orc.lib()
g = {"rng": []}
for (x, y) in [(0, 0), (1, 0), (0, 1), (1919, 1079)]:
    for sample in (0, 7):
        seed = (C.c_uint32 * 2)(); end = (C.c_uint32 * 2)(); out = np.zeros(8, dtype=np.float32)
        L.orc_kat_seed(x, y, 1, sample, seed)
        L.orc_kat_rng(seed[0], seed[1], 8, out.ctypes.data_as(C.c_void_p), end)
        g["rng"].append({"x": x, "y": y, "sample": sample, "seed": [seed[0], seed[1]], "bits": [int(v) for v in out.view(np.uint32)]})
sc = rtdx.scenes.cornell()
W = H = 24
props, descs, lights, cam = host_inputs(rtdx, sc, W, H)
osc = orc.OracleScene(sc, props, lights)
acc, ctr = osc.render(cam, W, H, 0, 4, bounces=2, flags=3)
hits = osc.trace(rtdx.scenes.camera_rays(cam, W, H), mode=0)
g["cornell"] = {"size": W, "spp": 4, "closest_rays": ctr["closest_rays"], "shadow_rays": ctr["shadow_rays"],
                "accum_bits": [int(v) for v in acc.view(np.uint32).reshape(-1)], "primary_prim": [int(v) for v in hits["prim"]]}
# ReSTIR frames: 3 frames of the 3-pass sequence, static camera, GGX allowed
import zlib  # noqa: E402
W2, H2 = 32, 24
props, descs, lights, cam2 = host_inputs(rtdx, sc, W2, H2)
fr = osc.new_frames(W2, H2)
acc2 = np.zeros((H2, W2, 4), dtype=np.float32)
tot = {"closest_rays": 0, "shadow_rays": 0}
for f_i in range(3):
    c = osc.render_frame(cam2, W2, H2, f_i, fr, acc2, bounces=2)
    tot["closest_rays"] += c["closest_rays"]; tot["shadow_rays"] += c["shadow_rays"]
dump = osc.dump_frames(fr, W2, H2)
g["restir"] = {"width": W2, "height": H2, "frames": 3, "bounces": 2, "closest_rays": tot["closest_rays"], "shadow_rays": tot["shadow_rays"],
               "accum_bits": [int(v) for v in acc2.view(np.uint32).reshape(-1)],
               "reservoir_crc32": int(zlib.crc32(np.ascontiguousarray(dump).view(np.uint8).tobytes())),
               "M_di_sum": int(dump[..., 11].sum()), "M_gi_sum": int(dump[..., 23].sum())}
# legacy estimator: Cornell, 24x24, 4 samples, path cap 12, GGX allowed
acc3, ctr3 = osc.render(cam, W, H, 0, 4, bounces=12, flags=orc.FLAG_LEGACY_RR)
g["legacy"] = {"size": W, "spp": 4, "bounces": 12, "closest_rays": ctr3["closest_rays"], "shadow_rays": ctr3["shadow_rays"],
               "accum_bits": [int(v) for v in acc3.view(np.uint32).reshape(-1)]}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_kat.json"), "w") as f:
    json.dump(g, f)
print("wrote golden_kat.json", len(json.dumps(g)), "bytes")
