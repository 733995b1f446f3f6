"""CPU tests (-m "not gpu"): the C++ host mirror of the reference's Renderer slices, the C-ABI library (loads, exports
every declared symbol, refuses to compute without a GPU), and the multi-GPU plumbing on gloo with world_size 2."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from util import host_inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- F21 light list, rdn/Renderer.cpp:2123-2233 ------------------------------------------------------------------
def _light_list_numpy(rtdx, sc):
    rows = []
    for ii, (mi, _) in enumerate(sc.instances):
        m = sc.models[mi]
        V = m["vertices"]["position"]; I = m["indices"].reshape(-1, 3)
        ids = sc.material_ids[m["material_id_offset"]: m["material_id_offset"] + I.size].reshape(-1, 3)
        for t in range(I.shape[0]):
            mat = sc.materials[ids[t, 0]]
            ke = mat["Ke"]
            if np.float32(np.float32(ke[0] + ke[1]) + ke[2]) > 0:
                a, b, c = V[I[t, 0]], V[I[t, 1]], V[I[t, 2]]
                area = np.float32(0.5) * np.float32(np.linalg.norm(np.cross((b - a).astype(np.float32), (c - a).astype(np.float32)).astype(np.float32)))
                w = area * np.float32(np.float32(np.float32(ke[0] + ke[1]) + ke[2]) / np.float32(3.0))
                rows.append((a, b, c, ii, w, ke))
    order = sorted(range(len(rows)), key=lambda k: -rows[k][4])     # stable, descending
    return [rows[k] for k in order]


def test_light_list_matches_numpy_restatement(rtdx):
    sc = rtdx.scenes.instanced_blobs(n_models=2, n_side=3, lattice=3, emissive_fraction=0.2)
    lights = rtdx.collect_emissive_triangles(sc)
    ref = _light_list_numpy(rtdx, sc)
    assert len(lights) == len(ref) > 0
    assert (lights["triCount"] == len(ref)).all()
    tot = np.float32(0)
    for r in ref:
        tot = np.float32(tot + r[4])
    assert np.allclose(lights["total_weight"], tot, rtol=1e-5)
    assert np.all(np.diff(lights["weight"]) <= 1e-9)                 # sorted by weight, descending
    assert lights["cdf"][-1] == np.float32(1.0) and np.all(np.diff(lights["cdf"]) >= 0)
    assert np.allclose(lights["weight"].sum(), 1.0, atol=1e-4)
    for k in (0, len(ref) // 2, len(ref) - 1):
        assert lights["instanceID"][k] == ref[k][3] and np.allclose(lights["x"][k], ref[k][0]) and np.allclose(lights["emission"][k], ref[k][5])
        assert np.isclose(lights["weight"][k] * tot, ref[k][4], rtol=1e-4)


def test_light_list_cornell(rtdx):
    sc = rtdx.scenes.cornell()
    lights = rtdx.collect_emissive_triangles(sc)
    assert len(lights) == 2 and (lights["triCount"] == 2).all()
    assert np.allclose(lights["weight"], 0.5) and np.allclose(lights["total_weight"], 2 * 0.125 * 15.0)   # 2 tris of area 0.125, mean Ke 15
    assert (lights["instanceID"] == 0).all()


# ---- F22 instance properties, rdn/Renderer.cpp:2091-2121 + TopLevelASGenerator.cpp:181-199 ---------------------------
def test_instance_properties(rtdx):
    rng = np.random.RandomState(4)
    ms = []
    for _ in range(5):
        A = rng.normal(size=(3, 3)) + 2 * np.eye(3)
        M = np.eye(4); M[:3, :3] = A; M[:3, 3] = rng.normal(size=3) * 3
        ms.append(M)
    props, descs = rtdx.instance_properties([rtdx.xmmatrix_from_colvec(m) for m in ms], [0, 1, 2, 3, 4])
    for k, M in enumerate(ms):
        O = props["objectToWorld"][k].reshape(4, 4).T            # HLSL reading: M[r][c] = mem[4c + r]
        OI = props["objectToWorldInverse"][k].reshape(4, 4).T
        N = props["objectToWorldNormal"][k].reshape(4, 4).T
        assert np.allclose(O, M, atol=1e-6)
        assert np.allclose(OI @ O, np.eye(4), atol=1e-5)
        assert np.allclose(N[:3, :3], np.linalg.inv(M[:3, :3]).T, atol=1e-5)      # inverse-transpose of the upper 3x3
        assert np.allclose(descs["transform"][k], M[:3, :], atol=1e-6)           # 3x4 row-major objectToWorld
        assert descs["instance_id_mask"][k] == (k | (0xFF << 24)) and descs["hit_group_flags"][k] == 2 * k and descs["blas"][k] == k
        assert np.array_equal(props["prevObjectToWorld"][k], props["objectToWorld"][k])


def test_ess_lut_is_deterministic_and_plausible(rtdx):
    m = rtdx.scenes.make_material((.5, .5, .5), roughness=0.5)
    a = m.copy(); b = m.copy()
    rtdx.generate_ess_lut(a, seed=7); rtdx.generate_ess_lut(b, seed=7)
    assert np.array_equal(a["LUT"], b["LUT"])
    lut = a["LUT"][0]
    assert (lut > 0.2).all() and (lut <= 1.0001).all()            # directional albedo of a white GGX lobe


# ---- C ABI ------------------------------------------------------------------------------------------------------
def _declared_symbols(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rtx_[a-z_0-9]+)\s*\(", text)))


def test_abi_library_exports_every_declared_symbol(rtdx):
    lib = rtdx.load_library()
    declared = _declared_symbols("rtx_b200.h")
    assert len(declared) >= 24
    for name in declared:
        assert hasattr(lib, name), "librtx_b200.so does not export %s" % name
    assert sorted(rtdx.ABI_SYMBOLS) == declared
    host = rtdx.load_host_library()
    for name in rtdx.HOST_SYMBOLS:
        assert hasattr(host, name)


def test_abi_struct_sizes_match_reference_layouts(rtdx):
    assert rtdx.vertex_dt.itemsize == 28 and rtdx.material_dt.itemsize == 128       # S1, S4
    assert rtdx.props_dt.itemsize == 384 and rtdx.light_dt.itemsize == 80           # S6, S7
    assert rtdx.camera_dt.itemsize == 512 and rtdx.desc_dt.itemsize == 64           # S8, D3D12_RAYTRACING_INSTANCE_DESC
    assert C.sizeof(rtdx.RtxConfig) == 48


def test_no_cpu_fallback(rtdx):
    """Without a CUDA device the product must fail loudly, never fall back to the oracle or any CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(rtdx.RtxError):
        rtdx.Context(16, 16)
    # argument validation happens before any device work
    lib = rtdx.load_library()
    cfg = rtdx.RtxConfig(0, 0, 16, 16, 3, 4, 4, 0, 1, None)      # wrong struct_size
    h = C.c_void_p()
    assert lib.rtx_create(C.byref(cfg), C.byref(h)) == 1 and b"struct_size" in lib.rtx_last_error()


def test_product_does_not_reference_the_oracle():
    """The product path may not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "royaltracer-dx_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle/" not in text.replace("oracle/rtx_oracle.cpp tri_test()", "") and "liborc" not in text and "from oracle" not in text, f
    out = subprocess.run(["ldd", os.path.join(pkg, "librtx_b200.so")], capture_output=True, text=True).stdout
    assert "liborc" not in out


# ---- multi-GPU plumbing on gloo, world_size 2 ---------------------------------------------------------------------------
_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch, torch.distributed as dist
import rtdx, importlib
from oracle import orc
from util import host_inputs
rdist = importlib.import_module("royaltracer-dx_b200.dist")
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()
sc = rtdx.scenes.cornell(); W = H = 16; SPP = 4
props, descs, lights, cam = host_inputs(rtdx, sc, W, H)
osc = orc.OracleScene(sc, props, lights)          # the CPU oracle stands in for the GPU render in this plumbing test
local = torch.zeros(H, W, 4); scratch = torch.zeros(H, W, 4)
for step in range(SPP // world):
    s = rdist.sample_for(step, rank, world)
    acc = local.numpy()
    osc.render(cam, W, H, s, 1, bounces=2, flags=3, accum=acc)
    total = rdist.reduce_accum(local, scratch, dst=0)
if rank == 0:
    np.save(%(out)r, total.numpy())
dist.barrier(); dist.destroy_process_group()
"""


def test_two_rank_sample_partition_and_reduce(rtdx, orc, tmp_path):
    import importlib
    rdist = importlib.import_module("royaltracer-dx_b200.dist")
    assert [rdist.sample_for(s, r, 2) for s in range(2) for r in range(2)] == [0, 1, 2, 3]
    assert rdist.samples_of_rank(7, 1, 2) == [1, 3, 5]
    out = str(tmp_path / "total.npy")
    import socket
    with socket.socket() as so:
        so.bind(("127.0.0.1", 0)); port = so.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT, "port": port, "out": out})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    total = np.load(out)
    # reference: one process, samples summed per rank in rank-local order, then rank 0 + rank 1 (fp32 sum order of the reduce)
    sc = rtdx.scenes.cornell(); W = H = 16
    props, descs, lights, cam = host_inputs(rtdx, sc, W, H)
    osc = orc.OracleScene(sc, props, lights)
    parts = []
    for r in range(2):
        acc = np.zeros((H, W, 4), dtype=np.float32)
        for s in rdist.samples_of_rank(4, r, 2):
            osc.render(cam, W, H, s, 1, bounces=2, flags=3, accum=acc)
        parts.append(acc)
    assert np.array_equal(total, parts[0] + parts[1])
    # and it equals the 1-GPU image up to fp32 summation order
    one, _ = osc.render(cam, W, H, 0, 4, bounces=2, flags=3)
    assert (total[..., 3] == 4).all() and np.allclose(total, one, rtol=1e-5, atol=1e-6)
