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
cmd = ["nvcc", "-shared"] + B.NVCC_FLAGS + defs + ["-o", os.path.join(out, name + ".so")] + cu
print(" ".join(cmd), flush=True)
subprocess.check_call(cmd)
