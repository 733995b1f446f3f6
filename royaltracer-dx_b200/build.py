"""Build recipe: compiles the CUDA engine for sm_100a into librtx_b200.so (in-tree, next to this file) and the
C++ host-side mirror of the reference's Renderer slices into librtx_host.so.  Run: python build.py [--force]."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIB = os.path.join(HERE, "librtx_b200.so")
HOSTLIB = os.path.join(HERE, "librtx_host.so")
PREPLIB = os.path.join(HERE, "librdx_prep.so")    # host-side data preparation only: no engine, no CUDA dependency

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # reproducible fp32: no implicit FMA contraction (see csrc/dmath.cuh)
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]
OBJ = os.path.join(HERE, "..", "build", "obj")


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(d, exts):
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(exts))


def build(force=False, verbose=False):
    cu = _sources(CSRC, (".cu",))
    deps = _sources(CSRC, (".cu", ".cuh", ".h")) + [os.path.join(HERE, "..", "include", "rtx_b200.h")]
    if force or _newer(LIB, deps):
        # one object per translation unit, compiled concurrently (no relocatable device code: every kernel lives in one file), then linked
        from concurrent.futures import ThreadPoolExecutor
        os.makedirs(OBJ, exist_ok=True)
        hdrs = [d for d in deps if not d.endswith(".cu")]

        def compile_one(job):
            src, suffix, extra = job
            obj = os.path.join(OBJ, os.path.basename(src)[:-3] + suffix + ".o")
            if force or _newer(obj, [src] + hdrs):
                cmd = ["nvcc"] + [f for f in NVCC_FLAGS if not (extra and f == "-fmad=false")] + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
                print(" ".join(cmd), flush=True)
                r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                if verbose or r.returncode:
                    print(r.stdout, flush=True)
                if r.returncode:
                    raise subprocess.CalledProcessError(r.returncode, cmd)
            return obj
        jobs = [(c, "", []) for c in cu]
        # the shading stages a second time with fast math, into namespace rtx::fast (RTX_FLAG_FAST_MATH; wavefront.cu)
        jobs.append((os.path.join(CSRC, "wavefront.cu"), "_fast", ["-DRTX_FAST_MATH", "-fmad=true", "-use_fast_math"]))
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            objs = list(ex.map(compile_one, jobs))
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs
        print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    hs = _sources(HOST, (".cpp",))
    hdeps = _sources(HOST, (".cpp", ".h")) + [os.path.join(HERE, "..", "include", "rtx_b200.h")]
    if hs and (force or _newer(HOSTLIB, hdeps)):
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wall", "-o", HOSTLIB] + hs + [
            "-L" + HERE, "-lrtx_b200", "-Wl,-rpath,$ORIGIN"]
        print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    if hs and (force or _newer(PREPLIB, hdeps)):
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wall", "-DRDX_PREP_ONLY", "-o", PREPLIB] + hs
        print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB, HOSTLIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
