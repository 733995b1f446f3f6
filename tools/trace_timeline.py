"""Per-warp timeline of one traversal launch: needs a build with -DRTX_TRACE_TIMELINE (python tools/build_variant.py timeline -DRTX_TRACE_TIMELINE;
RTX_B200_LIB=build/variants/timeline.so python tools/trace_timeline.py > out.txt).  Every 37th warp prints when it found the queue drained and
when it finished; `python tools/trace_timeline.py --analyse out.txt` prints the percentiles (profiles/r01_s4_trace_timeline.txt)."""
import re
import sys
if len(sys.argv) > 2 and sys.argv[1] == "--analyse":
    import numpy as np
    rows = []; sec = 0
    for l in open(sys.argv[2]):
        if l.startswith("==="): sec = 1
        m = re.match(r"TL (\d+) n (\d+) start (\d+) exhausted_at (\d+) end (\d+) steps (\d+) drain_steps (\d+)", l)
        if m: rows.append((sec,) + tuple(int(x) for x in m.groups()))
    a = np.array(rows)
    for sec, name in ((0, "coherent primaries"), (1, "incoherent bounces")):
        b = a[a[:, 0] == sec]
        if len(b) == 0: continue
        ex = b[:, 4] / 1e3; en = b[:, 5] / 1e3
        print("%s: %d warps sampled, %d rays" % (name, len(b), b[0, 2]))
        for nm, v in (("queue found drained at [us]", ex), ("warp finished at [us]", en), ("drain = finished - drained [us]", en - ex), ("steps", b[:, 6]),
                      ("steps after the queue drained", b[:, 7]), ("us per step", en / b[:, 6]), ("us per step while draining", (en - ex) / np.maximum(b[:, 7], 1))):
            print("  %-34s min %7.1f  p10 %7.1f  p50 %7.1f  p90 %7.1f  max %7.1f" % (nm, v.min(), np.percentile(v, 10), np.percentile(v, 50), np.percentile(v, 90), v.max()))
    sys.exit(0)
import os, sys
import numpy as np, torch
ROOT="/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import rtdx
from sweep import bounce_rays
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
sc = rtdx.scenes.sphere_in_box(1_000_000)
ctx = rtdx.Context(64, 64, stream=stream.cuda_stream)
ctx.upload_scene(sc); torch.cuda.synchronize()
cam = rtdx.camera_params(sc.eye, sc.center, sc.up, 1.0)
prim = rtdx.scenes.camera_rays(cam, 2048, 1024)
rays = torch.from_numpy(prim.view(np.float32).reshape(-1, 8)).cuda()
n = rays.shape[0]
hits = torch.empty((n, 5), dtype=torch.float32, device="cuda")
os.environ["X"]="1"
ctx.trace_device(rays.data_ptr(), n, hits.data_ptr()); torch.cuda.synchronize()
print("=== incoherent")
gen = torch.Generator(device="cuda"); gen.manual_seed(7)
inc = bounce_rays(rays, hits, gen)
m = inc.shape[0]
h2 = torch.empty((m, 5), dtype=torch.float32, device="cuda")
ctx.trace_device(inc.data_ptr(), m, h2.data_ptr()); torch.cuda.synchronize()
