"""-m gpu: the C++ host (rdx::Renderer, royaltracer-dx_b200/host) executed end to end on the GPU — VERDICT r1 missing 3 / weak 13.
tests/host/host_main.cpp plays rdn/Main.cpp + Renderer::OnInit/OnUpdate/OnRender: CreateVB(path) per OBJ, OnInit, per frame
SetInstanceTransform(1, rotY(1.57)) (rdn/Renderer.cpp:444-449) + OnUpdate + OnRenderFrame.  Its accumulation / reservoir CRCs and ray
counters must equal the Python-driven engine on the same files, which in turn is bit-compared with the oracle."""
import os
import subprocess
import zlib

import numpy as np
import pytest

from util import bits

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "royaltracer-dx_b200")


def _icosphere(sub):
    t = (1 + 5 ** 0.5) / 2
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    v = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(sub):
        cache, nf = {}, []

        def mid(a, b):
            k = (min(a, b), max(a, b))
            if k not in cache:
                m = v[a] + v[b]; v.append(m / np.linalg.norm(m)); cache[k] = len(v) - 1
            return cache[k]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return np.array(v), f


def write_scene(d):
    """A garage-like room (floor, walls, emissive strip; per-vertex normals) and a normal-less metallic blob, as OBJ + MTL files."""
    room = ["mtllib room.mtl"]
    q = []          # quads (4 corner points), material
    lo, hi = (-5.0, 0.0, -4.0), (5.0, 3.5, 4.0)
    q.append(([(lo[0], 0, hi[2]), (hi[0], 0, hi[2]), (hi[0], 0, lo[2]), (lo[0], 0, lo[2])], "floor", (0, 1, 0)))
    q.append(([(lo[0], hi[1], lo[2]), (hi[0], hi[1], lo[2]), (hi[0], hi[1], hi[2]), (lo[0], hi[1], hi[2])], "walls", (0, -1, 0)))
    q.append(([(lo[0], 0, lo[2]), (hi[0], 0, lo[2]), (hi[0], hi[1], lo[2]), (lo[0], hi[1], lo[2])], "walls", (0, 0, 1)))
    q.append(([(hi[0], 0, hi[2]), (lo[0], 0, hi[2]), (lo[0], hi[1], hi[2]), (hi[0], hi[1], hi[2])], "walls", (0, 0, -1)))
    q.append(([(lo[0], 0, hi[2]), (lo[0], 0, lo[2]), (lo[0], hi[1], lo[2]), (lo[0], hi[1], hi[2])], "walls", (1, 0, 0)))
    q.append(([(hi[0], 0, lo[2]), (hi[0], 0, hi[2]), (hi[0], hi[1], hi[2]), (hi[0], hi[1], lo[2])], "walls", (-1, 0, 0)))
    q.append(([(-1.5, 3.45, -0.4), (1.5, 3.45, -0.4), (1.5, 3.45, 0.4), (-1.5, 3.45, 0.4)], "lights", (0, -1, 0)))
    nv = 0
    normals = []
    for pts, mat, n in q:
        for p in pts:
            room.append("v %.6f %.6f %.6f" % p)
        normals.append(n)
    for n in normals:
        room.append("vn %g %g %g" % n)
    for k, (pts, mat, n) in enumerate(q):
        room.append("usemtl " + mat)
        b = 4 * k + 1
        room.append("f %d//%d %d//%d %d//%d %d//%d" % (b, k + 1, b + 1, k + 1, b + 2, k + 1, b + 3, k + 1))
    (d / "room.obj").write_text("\n".join(room) + "\n")
    (d / "room.mtl").write_text("newmtl floor\nKd 0.6 0.6 0.55\nKs 0.04 0.04 0.04\nPr 0.5\nnewmtl walls\nKd 0.15 0.15 0.18\nKs 0.04 0.04 0.04\nPr 1.0\n"
                                "newmtl lights\nKd 0 0 0\nKe 5 5 5\nPr 1.0\n")
    v, f = _icosphere(2)
    rng = np.random.RandomState(4)
    v = v * (0.8 + 0.08 * rng.rand(v.shape[0], 1)) + np.array([0.0, 1.2, 0.0])
    blob = ["mtllib blob.mtl", "usemtl metal"] + ["v %.6f %.6f %.6f" % tuple(p) for p in v] + ["f %d %d %d" % (a + 1, b + 1, c + 1) for a, b, c in f]
    (d / "blob.obj").write_text("\n".join(blob) + "\n")
    (d / "blob.mtl").write_text("newmtl metal\nKd 0.9 0.7 0.3\nKs 0.9 0.7 0.3\nPr 0.3\nPm 1.0\n")
    return [str(d / "room.obj"), str(d / "blob.obj")]


def build_host_main(tmp_path):
    exe = str(tmp_path / "host_main")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "host", "host_main.cpp"),
                           "-I", os.path.join(PKG, "host"), "-L", PKG, "-lrtx_host", "-lrtx_b200", "-lpthread", "-Wl,-rpath," + PKG])
    return exe


def test_cpp_host_renderer_end_to_end(rtdx, orc, tmp_path):
    paths = write_scene(tmp_path)
    W, H, FRAMES = 96, 64, 3
    rot = np.eye(4); a = np.float32(1.57)
    rot[0, 0] = np.cos(a); rot[0, 2] = np.sin(a); rot[2, 0] = -np.sin(a); rot[2, 2] = np.cos(a)
    xms = [rtdx.xmmatrix_from_colvec(np.eye(4)), rtdx.xmmatrix_from_colvec(rot)]           # instance 1: rotY(1.57), rdn/Renderer.cpp:444-449
    np.ascontiguousarray(np.stack(xms), dtype=np.float32).tofile(str(tmp_path / "xforms.bin"))
    exe = build_host_main(tmp_path)
    out = subprocess.run([exe, str(W), str(H), str(FRAMES), str(tmp_path / "xforms.bin")] + paths, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    cpp = [dict(zip(l.split()[2::2], (int(x) for x in l.split()[3::2]))) for l in out.stdout.splitlines() if l.startswith("frame")]
    tail = [l for l in out.stdout.splitlines() if l.startswith("output_crc")][0].split()
    assert len(cpp) == FRAMES and int(tail[3]) == 2                                         # the emissive strip = 2 light triangles

    # the same frames driven from Python through the same C ABI, and by the oracle
    sc = rtdx.scenes.from_obj_files(paths, name="host")
    sc.instances = [(0, xms[0]), (1, xms[1])]
    ctx = rtdx.Context(W, H, bounces=3, flags=rtdx.FLAG_RESTIR)
    up = ctx.upload_scene(sc)
    assert up["lights"].size == 2
    osc = orc.OracleScene(sc, up["props"], up["lights"])
    frames = osc.new_frames(W, H)
    acc = np.zeros((H, W, 4), dtype=np.float32)
    tot = {"closest_rays": 0, "shadow_rays": 0}
    ctx.reset_counters()
    for f in range(FRAMES):
        ctx.set_instances(up["descs"], up["props"]); ctx.set_camera(up["camera"])         # OnUpdate (static transforms: prev == current)
        ctx.render_frame(f); ctx.synchronize()
        o = osc.render_frame(up["camera"], W, H, f, frames, acc, bounces=3)
        for k in tot:
            tot[k] += o[k]
        gpu, cnt = ctx.read_accum(), ctx.counters()
        assert (bits(gpu) != bits(acc)).sum() == 0
        assert (cnt["closest_rays"], cnt["shadow_rays"]) == (tot["closest_rays"], tot["shadow_rays"])
        rs = ctx.read_restir()
        assert (bits(rs) != bits(osc.dump_frames(frames, W, H))).sum() == 0
        # ... and the C++ host produced exactly this frame
        assert cpp[f]["closest"] == cnt["closest_rays"] and cpp[f]["shadow"] == cnt["shadow_rays"], (f, cpp[f], cnt)
        assert cpp[f]["accum_crc"] == zlib.crc32(gpu.tobytes()) and cpp[f]["restir_crc"] == zlib.crc32(np.ascontiguousarray(rs).tobytes())
    assert int(tail[1]) == zlib.crc32(ctx.read_output().tobytes())
    osc.free_frames(frames)
    ctx.close()


def test_cpp_host_two_ranks_with_the_engine_side_reduce(rtdx, tmp_path):
    """Two host_main processes, one per GPU, joined through rtx_comm_init (NCCL id handed over in a file) render the samples
    s = rank (mod 2) and reduce gPermanentData to rank 0 after every pass (rtx_reduce_accum): rank 0's reduced buffer must equal the sum
    of the two ranks' partial sums rendered here in one process (a + b is the same in either order)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    paths = write_scene(tmp_path)
    W, H, FRAMES = 96, 64, 3
    xms = [rtdx.xmmatrix_from_colvec(np.eye(4)), rtdx.xmmatrix_from_colvec(np.eye(4))]
    np.ascontiguousarray(np.stack(xms), dtype=np.float32).tofile(str(tmp_path / "xforms.bin"))
    exe = build_host_main(tmp_path)
    idf = str(tmp_path / "nccl.id")
    procs = []
    for r in range(2):
        env = dict(os.environ, RTX_HOST_NCCL_ID=idf, RTX_HOST_RANK=str(r), RTX_HOST_WORLD="2")
        procs.append(subprocess.Popen([exe, str(W), str(H), str(FRAMES), str(tmp_path / "xforms.bin")] + paths, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True, env=env))
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), [o[1] for o in outs]
    rank0 = [dict(zip(l.split()[2::2], (int(x) for x in l.split()[3::2]))) for l in outs[0][0].splitlines() if l.startswith("frame")]
    sc = rtdx.scenes.from_obj_files(paths, name="host")
    parts = []
    for r in range(2):
        ctx = rtdx.Context(W, H, bounces=3)
        ctx.upload_scene(sc)
        acc = []
        for f in range(FRAMES):
            ctx.render_pass(f * 2 + r, 1); ctx.synchronize()
            acc.append(ctx.read_accum())
        parts.append(acc); ctx.close()
    for f in range(FRAMES):
        total = parts[0][f] + parts[1][f]
        assert rank0[f]["accum_crc"] == zlib.crc32(total.tobytes()), f
