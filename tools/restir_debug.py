"""Bring-up diagnostics for the ReSTIR frame: field-level differences between the engine's and the oracle's *_last buffers."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rtdx
from oracle import orc
from util import bits
NAMES = ["x2.x","x2.y","x2.z","w_sum","n2.x","n2.y","n2.z","W","L2.x","L2.y","L2.z","M","xn.x","xn.y","xn.z","gw_sum","nn.x","nn.y","nn.z","gW","E3.x","E3.y","E3.z","gM",
         "x1.x","x1.y","x1.z","mID","n1.x","n1.y","n1.z","objID","o.x","o.y","o.z","kind","L1.x","L1.y","L1.z","-"]
W, H, frames_n = 96, 80, int(sys.argv[1]) if len(sys.argv) > 1 else 2
sc = rtdx.scenes.cornell()
ctx = rtdx.Context(W, H, bounces=2, flags=rtdx.FLAG_RESTIR)
up = ctx.upload_scene(sc)
osc = orc.OracleScene(sc, up["props"], up["lights"])
fr = osc.new_frames(W, H); acc = np.zeros((H, W, 4), np.float32)
for f in range(frames_n):
    ctx.render_frame(f); ctx.synchronize()
    osc.render_frame(up["camera"], W, H, f, fr, acc, bounces=2)
    g, o = ctx.read_restir(), osc.dump_frames(fr, W, H)
    bad = bits(g) != bits(o)
    print("frame", f, "pixels differing", int(bad.any(-1).sum()), "accum floats differing", int((bits(ctx.read_accum()) != bits(acc)).sum()))
    cnt = bad.reshape(-1, 40).sum(0)
    print("  fields:", {NAMES[k]: int(c) for k, c in enumerate(cnt) if c})
    for (y, x) in np.argwhere(bad.any(-1))[:4]:
        print("  pixel", x, y, "kind", o[y, x, 35])
        for k in np.nonzero(bad[y, x])[0]:
            print("     %-6s gpu %r  ref %r" % (NAMES[k], g[y, x, k], o[y, x, k]))
