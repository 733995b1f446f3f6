"""Builds tests/golden/reference_scene.npz and its golden outputs from the reference's own assets
(/root/reference/Pathtracer/include/{garage,monke}.{obj,mtl}, rdn/Renderer.cpp:363) — run in the build container only:
    python tests/golden/make_reference_scene.py
The .npz is what this repo's OBJ/MTL ingest (royaltracer-dx_b200/host/ObjLoader.cpp, mirroring src/Util/ObjLoader.h:393-495)
produces from those files: vertices de-duplicated by position, indices, one material id per face-vertex, the material blocks
[default, mtl...] with fixed-seed ESS LUTs.  The GPU box has no /root/reference, so the derived arrays travel instead, together
with the oracle's outputs on them (E0 render, ReSTIR frames, primary hits)."""
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rtdx  # noqa: E402
from oracle import orc  # noqa: E402
from util import host_inputs, load_scene_npz, save_scene_npz  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sc = rtdx.scenes.reference_scene("/root/reference/Pathtracer/include")
save_scene_npz(os.path.join(HERE, "reference_scene.npz"), sc)
sc = load_scene_npz(rtdx, os.path.join(HERE, "reference_scene.npz"))
W, H = 64, 36                                      # the reference's 16:9 frame, scaled down
props, descs, lights, cam = host_inputs(rtdx, sc, W, H)
osc = orc.OracleScene(sc, props, lights)
g = {"width": W, "height": H, "bounces": 3, "triangles": sc.n_triangles(), "n_lights": int(lights.size)}
acc, ctr = osc.render(cam, W, H, 0, 2, bounces=3)
g["e0"] = {"spp": 2, "closest_rays": ctr["closest_rays"], "shadow_rays": ctr["shadow_rays"],
           "accum_crc32": int(zlib.crc32(acc.view(np.uint8).tobytes())), "mean": float(acc[..., :3].mean())}
hits = osc.trace(rtdx.scenes.camera_rays(cam, W, H), mode=0)
g["primary"] = {"inst": [int(v) for v in hits["inst"]], "prim": [int(v) for v in hits["prim"]]}
fr = osc.new_frames(W, H); acc2 = np.zeros((H, W, 4), dtype=np.float32); tot = [0, 0]
for f in range(3):
    c = osc.render_frame(cam, W, H, f, fr, acc2, bounces=3)
    tot[0] += c["closest_rays"]; tot[1] += c["shadow_rays"]
g["restir"] = {"frames": 3, "closest_rays": tot[0], "shadow_rays": tot[1], "accum_crc32": int(zlib.crc32(acc2.view(np.uint8).tobytes())),
               "reservoir_crc32": int(zlib.crc32(np.ascontiguousarray(osc.dump_frames(fr, W, H)).view(np.uint8).tobytes()))}
with open(os.path.join(HERE, "reference_scene_golden.json"), "w") as f:
    json.dump(g, f)
print("wrote reference_scene.npz (%d bytes) and reference_scene_golden.json" % os.path.getsize(os.path.join(HERE, "reference_scene.npz")), g["e0"], g["restir"])
