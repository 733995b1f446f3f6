"""A/B builds of the CUDA engine: python tools/build_variant.py NAME -DRTX_X=1 ...  ->  build/variants/NAME.so
Run a variant with RTX_B200_LIB=build/variants/NAME.so (royaltracer-dx_b200/__init__.py honours it)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "royaltracer-dx_b200"))
import build as B  # noqa: E402

name, defs = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "build", "variants")
os.makedirs(out, exist_ok=True)
cu = B._sources(B.CSRC, (".cu",))
objdir = os.path.join(out, name + "_obj")
os.makedirs(objdir, exist_ok=True)
jobs = [(c, "", []) for c in cu] + [(os.path.join(B.CSRC, "wavefront.cu"), "_fast", ["-DRTX_FAST_MATH", "-fmad=true", "-use_fast_math"])]
procs = []
for src, suffix, extra in jobs:      # same translation units as build.py (incl. the fast-math copy of the shading stages), plus the defines
    obj = os.path.join(objdir, os.path.basename(src)[:-3] + suffix + ".o")
    cmd = ["nvcc"] + [f for f in B.NVCC_FLAGS if not (extra and f == "-fmad=false")] + extra + defs + ["-c", "-o", obj, src]
    procs.append((subprocess.Popen(cmd), obj))
objs = []
for p, obj in procs:
    if p.wait() != 0:
        raise SystemExit("nvcc failed for " + obj)
    objs.append(obj)
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", os.path.join(out, name + ".so")] + objs
print(" ".join(cmd), flush=True)
subprocess.check_call(cmd)
import shutil
shutil.rmtree(objdir)
