#!/usr/bin/env python
"""ORACLE — TEST INFRASTRUCTURE ONLY.  Recipe that compiles the reference's OWN shader text for the CPU (oracle/_ref/libref.so).

The reference (ML200/RoyalTracer-DX) is a Windows/DX12 program; its host cannot be built here, but the arithmetic of the hot path
lives in plain-function HLSL (Pathtracer/shaders/*_v7.hlsl, which #include Pathtracer/include/*_v6.hlsl).  This recipe

  1. reads those files WHERE THEY LIE under /root/reference (nothing is copied into the repository: the filtered text goes to
     oracle/_ref/gen/, which is git-ignored together with the library),
  2. applies a purely mechanical token filter for the handful of HLSL spellings C++ has no equivalent for (listed in FILTER below;
     no expression, statement, constant or function body is rewritten by hand),
  3. compiles the result with g++ against oracle/ref/hlsl_shim.h (vector types, intrinsics with the numerics contract of
     oracle/det_math.h, resource types, DXR system values) and oracle/ref/ref_harness.cpp (plays the DXR runtime: binds the buffers,
     runs RayGen / RayGen2 / RayGen3 per pixel, routes TraceRay to the oracle's ray caster and then into the reference's own
     ClosestHit / Miss / ShadowClosestHit / ShadowMiss).

tests/test_ref_pins.py then drives the hand-written oracle (oracle/rtx_oracle.cpp) and this library with the same inputs.
/root/reference does not exist on the GPU box: the library is built here and travels with the repository snapshot.

usage: python oracle/ref/make_ref.py [--force] [--bounces N] [--legacy]
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.dirname(HERE)
OUT = os.path.join(ORACLE, "_ref")
GEN = os.path.join(OUT, "gen")
LIB = os.path.join(OUT, "libref.so")
REF = os.environ.get("RTX_REFERENCE_ROOT", "/root/reference")
SHADERS = os.path.join(REF, "Pathtracer", "shaders")
INCLUDE = os.path.join(REF, "Pathtracer", "include")

# translation units = the DXIL libraries the reference's host compiles (rdn/Renderer.cpp:1034-1041); the top-level files are the
# shaders/ copies BASELINE.json names, their #include "..._v6.hlsl" lines resolve in include/ exactly as they do for DXC
UNITS = {
    "rg1": os.path.join(SHADERS, "Pass_init_di_v7.hlsl"),       # RayGen   (== include/RayGen_v6_pass1.hlsl)
    "rg2": os.path.join(SHADERS, "Pass_temp_di_v7.hlsl"),       # RayGen2  (== include/RayGen_v6_pass2.hlsl)
    "rg3": os.path.join(SHADERS, "Pass_spat_di_v7.hlsl"),       # RayGen3  (== include/RayGen_v6_pass3.hlsl)
    "hit": os.path.join(SHADERS, "Hit_v7.hlsl"),                # ClosestHit
    "miss": os.path.join(SHADERS, "Miss_v7.hlsl"),              # Miss
    "shadow": os.path.join(INCLUDE, "ShadowRay.hlsl"),          # ShadowClosestHit / ShadowMiss
}

# the reference's first estimator (RTX_FLAG_LEGACY_RR): include/RayGen.hlsl + include/Hit.hlsl + include/Miss.hlsl
# (-> Common.hlsl, BRDF.hlsl, GGX.hlsl, Lambertian.hlsl), with the same shadow shaders; compiled into its own library because its
# headers define the same names differently (HitInfo, Material accessors, SampleBRDF ...)
LEGACY_UNITS = {
    "leg_rg": os.path.join(INCLUDE, "RayGen.hlsl"),             # RayGen: path loop, Russian roulette, accumulation
    "leg_hit": os.path.join(INCLUDE, "Hit.hlsl"),               # ClosestHit: RIS-10 NEE + shadow ray + BSDF sample + MIS on emitter hits
    "leg_miss": os.path.join(INCLUDE, "Miss.hlsl"),             # Miss
    "shadow": os.path.join(INCLUDE, "ShadowRay.hlsl"),
}
LEGACY_LIB = os.path.join(OUT, "libref_legacy.so")

FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])")

# (pattern, replacement, what it is)
FILTER = [
    (re.compile(r"^\s*#pragma[^\n]*$", re.M), "", "#pragma warning"),
    (re.compile(r"\[\[raypayload\]\]"), "", "payload attribute"),
    (re.compile(r":\s*read\s*\([^)]*\)\s*:\s*write\s*\([^)]*\)"), "", "payload access qualifiers"),
    (re.compile(r":\s*register\s*\(\s*\w+\s*\)"), "", "register bindings"),
    (re.compile(r":\s*SV_RayPayload"), "", "semantic"),
    (re.compile(r"\[shader\(\"[^\"]*\"\)\]"), "", "shader stage attribute"),
    (re.compile(r"\[(?:loop|unroll|branch|flatten)\]"), "", "loop / branch hints"),
    (re.compile(r"\brow_major\s+"), "", "matrix packing qualifier of a parameter"),
    (re.compile(r"cbuffer\s+\w+\s*\{([^{}]*)\}"), r"\1", "cbuffer members are globals"),
    (re.compile(r"\b(?:inout|out)\s+([A-Za-z_]\w*)\s+(?=[A-Za-z_])"), r"\1& ", "inout / out parameters are references"),
    (re.compile(r"\bin\s+([A-Z][A-Za-z_]\w*)\s+(?=[A-Za-z_])"), r"\1 ", "in parameters are values"),
    # HLSL flattens nested initialiser lists ("Row 2: { float3(...), uint16_t(0) }" initialises the two members L2, M)
    (re.compile(r"\{\s*(float3\([^)]*\))\s*,\s*(uint16_t\(0\))\s*\}"), r"\1, \2", "flattened initialiser list"),
    # C++ has no swizzle members: .xyz / .xy become calls; the one swizzled compound assignment becomes a setter
    (re.compile(r"([A-Za-z_][\w\.\[\]\(\)]*)\.xyz\s*([-+*/])=\s*([^;]+);"), r"\1.set_xyz(\1.xyz \2 (\3));", "swizzled compound assignment"),
    (re.compile(r"\.xyz\b"), ".xyz()", "swizzle"),
    (re.compile(r"\.xy\b"), ".xy()", "swizzle"),
    # unsuffixed literals are float in HLSL, double in C++
    (FLOAT_LIT, r"\1f", "float literals"),
]


def expand(path, depth=0):
    """The file with its #include "x.hlsl" lines replaced by the (expanded) files from Pathtracer/include."""
    out = []
    with open(path) as f:
        for line in f.read().splitlines():
            m = re.match(r'\s*#include\s+"([^"]+)"', line)
            if m:
                inc = os.path.join(INCLUDE, m.group(1))
                out.append("// ---- begin %s" % os.path.relpath(inc, REF))
                out.append(expand(inc, depth + 1))
                out.append("// ---- end %s" % os.path.relpath(inc, REF))
            else:
                out.append(line)
    return "\n".join(out)


def filtered(path, bounces=None):
    text = expand(path)
    for pat, rep, _ in FILTER:
        text = pat.sub(rep, text)
    if bounces is not None:      # the reference's compile-time path length (Common_v7.hlsl:11); BASELINE configs vary it
        text = re.sub(r"(#define\s+bounces\s+)\d+", r"\g<1>%d" % bounces, text)      # (ShadowRay.hlsl has no such line)
    return text


def filtered_legacy(path, bounces=None):
    text = filtered(path)
    if bounces is not None:      # RayGen.hlsl:63 `uint bounces = 10000000;` (Russian roulette ends the paths); the engine's legacy mode has a cap
        text = re.sub(r"(uint\s+bounces\s*=\s*)\d+", r"\g<1>%d" % bounces, text)
    return text


def build_legacy(force=False, bounces=None, lib=None, verbose=False):
    lib = lib or (LEGACY_LIB if bounces is None else os.path.join(OUT, "libref_legacy_b%d.so" % bounces))
    if not os.path.isdir(INCLUDE):
        if os.path.exists(lib):
            return lib
        raise RuntimeError("reference sources not found under %s and no prebuilt %s" % (REF, lib))
    srcs = [os.path.join(HERE, f) for f in ("ref_legacy_harness.cpp", "hlsl_shim.h", "make_ref.py")] + [os.path.join(ORACLE, "det_math.h")]
    refs = [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE) if f.endswith(".hlsl")]
    if not force and os.path.exists(lib) and all(os.path.getmtime(s) <= os.path.getmtime(lib) for s in srcs + refs):
        return lib
    gen = os.path.join(OUT, "gen_legacy" + ("" if bounces is None else "_b%d" % bounces))
    os.makedirs(gen, exist_ok=True)
    for name, path in LEGACY_UNITS.items():
        with open(os.path.join(gen, "%s.inc" % name), "w") as f:
            f.write("// GENERATED by oracle/ref/make_ref.py from %s — reference text, do not commit\n" % os.path.relpath(path, REF))
            f.write(filtered_legacy(path, bounces))
            f.write("\n")
    cmd = ["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-Wno-narrowing", "-w",
           "-I", gen, "-o", lib, os.path.join(HERE, "ref_legacy_harness.cpp")]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return lib


def build(force=False, bounces=None, lib=LIB, verbose=False):
    if not os.path.isdir(SHADERS):
        if os.path.exists(lib):
            return lib           # the GPU box: no /root/reference, the library travelled with the snapshot
        raise RuntimeError("reference sources not found under %s and no prebuilt %s" % (REF, lib))
    srcs = [os.path.join(HERE, f) for f in ("ref_harness.cpp", "hlsl_shim.h", "make_ref.py")] + [os.path.join(ORACLE, "det_math.h")]
    refs = list(UNITS.values()) + [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE) if f.endswith(".hlsl")]
    if not force and os.path.exists(lib) and all(os.path.getmtime(s) <= os.path.getmtime(lib) for s in srcs + refs):
        return lib
    gen = GEN if bounces is None else GEN + "_b%d" % bounces
    os.makedirs(gen, exist_ok=True)
    for name, path in UNITS.items():
        with open(os.path.join(gen, "%s.inc" % name), "w") as f:
            f.write("// GENERATED by oracle/ref/make_ref.py from %s — reference text, do not commit\n" % os.path.relpath(path, REF))
            f.write(filtered(path, bounces))
            f.write("\n")
    cmd = ["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-Wno-narrowing", "-w",
           "-I", gen, "-o", lib, os.path.join(HERE, "ref_harness.cpp")]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    b = None
    if "--bounces" in sys.argv:
        b = int(sys.argv[sys.argv.index("--bounces") + 1])
    if "--legacy" in sys.argv:
        print(build_legacy(force="--force" in sys.argv, bounces=b, verbose=True))
        sys.exit(0)
    lib = LIB if b is None else os.path.join(OUT, "libref_b%d.so" % b)
    print(build(force="--force" in sys.argv, bounces=b, lib=lib, verbose=True))
