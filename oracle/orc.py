"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes wrapper of oracle/liborc.so (the CPU restatement of the reference's hot
path).  May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liborc.so")

FLAG_JITTER, FLAG_LAMBERT_ONLY, FLAG_LEGACY_RR = 1, 2, 16


class OrcConfig(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("bounces", C.c_uint32), ("nee_samples", C.c_uint32),
                ("nee_samples_di", C.c_uint32), ("flags", C.c_uint32)]


class OrcCounters(C.Structure):
    _fields_ = [("closest_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("paths", C.c_uint64),
                ("bvh_nodes_visited", C.c_uint64), ("tris_tested", C.c_uint64), ("instances_entered", C.c_uint64)]


# layout of orc_debug_pixel's 64 floats: name -> (offset, count, is_uint)
PIXEL_DEBUG_FIELDS = {
    "cam_o": (0, 3, False), "cam_d": (3, 3, False), "hit_inst": (6, 1, True), "hit_prim": (7, 1, True), "hit_t": (8, 1, False),
    "mID": (9, 1, True), "x1": (10, 3, False), "n1": (13, 3, False),
    "di_x2": (16, 3, False), "di_w_sum": (19, 1, False), "di_n2": (20, 3, False), "di_W": (23, 1, False), "di_L2": (24, 3, False),
    "gi_xn": (27, 3, False), "gi_w_sum": (30, 1, False), "gi_nn": (31, 3, False), "gi_W": (34, 1, False), "gi_E3": (35, 3, False),
    "p_hat": (38, 1, False), "C": (39, 3, False), "seed": (42, 2, True), "acc_L": (44, 3, False),
    "closest_rays": (47, 1, True), "shadow_rays": (48, 1, True),
}


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or any(
            os.path.getmtime(os.path.join(HERE, f)) > os.path.getmtime(LIB_PATH) for f in ("rtx_oracle.cpp", "rtx_oracle.h", "det_math.h")):
        subprocess.check_call(["make", "-C", HERE, "-B", "liborc.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        vp, u32 = C.c_void_p, C.c_uint32
        L.orc_scene_create.restype = vp
        L.orc_scene_destroy.argtypes = [vp]
        L.orc_add_model.argtypes = [vp, vp, u32, vp, u32, u32]
        L.orc_set_material_ids.argtypes = [vp, vp, u32]
        L.orc_set_materials.argtypes = [vp, vp, u32]
        L.orc_set_instances.argtypes = [vp, vp, vp, u32]
        L.orc_set_lights.argtypes = [vp, vp, u32]
        L.orc_build.argtypes = [vp]
        L.orc_trace.argtypes = [vp, vp, u32, vp, C.c_int, C.c_int]
        L.orc_render.argtypes = [vp, C.POINTER(OrcConfig), vp, u32, u32, u32, vp, C.POINTER(OrcCounters), C.c_int]
        L.orc_render_rows.argtypes = [vp, C.POINTER(OrcConfig), vp, u32, u32, u32, u32, u32, vp, C.POINTER(OrcCounters), C.c_int]
        L.orc_resolve.argtypes = [vp, u32, vp]
        L.orc_frames_create.argtypes = [u32, u32]
        L.orc_frames_create.restype = vp
        L.orc_frames_destroy.argtypes = [vp]
        L.orc_render_frame.argtypes = [vp, C.POINTER(OrcConfig), vp, u32, vp, vp, C.POINTER(OrcCounters), C.c_int]
        L.orc_frames_dump.argtypes = [vp, vp]
        L.orc_kat_rng.argtypes = [u32, u32, u32, vp, vp]
        L.orc_kat_seed.argtypes = [u32, u32, u32, u32, vp]
        L.orc_kat_sincos.argtypes = [vp, u32, vp, vp]
        L.orc_kat_half.argtypes = [vp, u32, vp]
        L.orc_kat_pow.argtypes = [vp, C.c_float, u32, vp]
        L.orc_kat_map_pixel.argtypes = [u32, u32, u32, u32]
        L.orc_kat_map_pixel.restype = u32
        L.orc_kat_bsdf.argtypes = [vp, C.c_int, u32, vp, vp, vp, vp, vp]
        L.orc_kat_camera_ray.argtypes = [C.POINTER(OrcConfig), vp, u32, u32, C.c_float, C.c_float, vp]
        L.orc_debug_pixel.argtypes = [vp, C.POINTER(OrcConfig), vp, u32, u32, u32, vp, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


class OracleScene:
    """Takes a scenes.SceneDesc plus the host-produced buffers (instance props, light list) — the same inputs the GPU gets."""

    def __init__(self, scene, props, lights):
        L = lib()
        self.L = L
        self.h = C.c_void_p(L.orc_scene_create())
        self._keep = []
        for m in scene.models:
            v, i = np.ascontiguousarray(m["vertices"]), np.ascontiguousarray(m["indices"])
            L.orc_add_model(self.h, _p(v), v.size, _p(i), i.size, int(m["material_id_offset"]))
        ids = np.ascontiguousarray(scene.material_ids, dtype=np.uint32)
        L.orc_set_material_ids(self.h, _p(ids), ids.size)
        mats = np.ascontiguousarray(scene.materials)
        L.orc_set_materials(self.h, _p(mats), mats.size)
        self._model_ids = [i[0] for i in scene.instances]
        mids = np.ascontiguousarray(np.array(self._model_ids, dtype=np.uint32))
        pr = np.ascontiguousarray(props)
        L.orc_set_instances(self.h, _p(mids), _p(pr), mids.size)
        lt = np.ascontiguousarray(lights)
        if lt.size == 0:
            lt = np.zeros(1, dtype=lt.dtype)     # OOB reads of t6 return zeros
        L.orc_set_lights(self.h, _p(lt), lt.size)
        L.orc_build(self.h)

    def close(self):
        if self.h:
            self.L.orc_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def trace(self, rays, any_hit=False, mode=1):
        r = np.ascontiguousarray(rays)
        out = np.zeros(r.size, dtype=np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("prim", "<u4"), ("inst", "<u4")]))
        self.L.orc_trace(self.h, _p(r), r.size, _p(out), 1 if any_hit else 0, mode)
        return out

    def render(self, cam, width, height, first_sample, n_samples, bounces=3, nee_samples=4, nee_samples_di=4, flags=0, step=1,
               accum=None, mode=1):
        cfg = OrcConfig(width, height, bounces, nee_samples, nee_samples_di, flags)
        if accum is None:
            accum = np.zeros((height, width, 4), dtype=np.float32)
        ctr = OrcCounters()
        c = np.ascontiguousarray(cam)
        self.L.orc_render(self.h, C.byref(cfg), _p(c), first_sample, n_samples, step, _p(accum), C.byref(ctr), mode)
        return accum, {k: getattr(ctr, k) for k, _ in OrcCounters._fields_}

    def render_threads(self, cam, width, height, first_sample, n_samples, n_threads, bounces=3, nee_samples=4, nee_samples_di=4,
                       flags=0, step=1, accum=None, mode=1):
        """Same result as render(); rows are split over n_threads host threads (ctypes releases the GIL)."""
        import threading
        cfg = OrcConfig(width, height, bounces, nee_samples, nee_samples_di, flags)
        if accum is None:
            accum = np.zeros((height, width, 4), dtype=np.float32)
        c = np.ascontiguousarray(cam)
        rows = [r for r in range(0, height, step)]
        chunks = [rows[i::n_threads] for i in range(n_threads)]
        ctrs = [OrcCounters() for _ in range(n_threads)]

        def work(i):
            for y in chunks[i]:
                self.L.orc_render_rows(self.h, C.byref(cfg), _p(c), first_sample, n_samples, step, y, y + 1, _p(accum), C.byref(ctrs[i]), mode)

        ts = [threading.Thread(target=work, args=(i,)) for i in range(n_threads)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        tot = {k: sum(getattr(ct, k) for ct in ctrs) for k, _ in OrcCounters._fields_}
        return accum, tot

    def set_props(self, props):
        """OnUpdate: new per-instance matrices (UpdateInstancePropertiesBuffer); rebuilds the oracle's instance boxes."""
        mids = np.ascontiguousarray(np.array(self._model_ids, dtype=np.uint32))
        pr = np.ascontiguousarray(props)
        self.L.orc_set_instances(self.h, _p(mids), _p(pr), mids.size)
        self.L.orc_build(self.h)

    def new_frames(self, width, height):
        return C.c_void_p(self.L.orc_frames_create(width, height))

    def free_frames(self, frames):
        self.L.orc_frames_destroy(frames)

    def render_frame(self, cam, width, height, frame_index, frames, accum, bounces=3, nee_samples=4, nee_samples_di=4, flags=0, mode=1):
        """One frame of the reference's 3-pass sequence (pass 1 + temporal + spatial reuse); accum is updated in place."""
        cfg = OrcConfig(width, height, bounces, nee_samples, nee_samples_di, flags)
        ctr = OrcCounters()
        c = np.ascontiguousarray(cam)
        self.L.orc_render_frame(self.h, C.byref(cfg), _p(c), frame_index, frames, _p(accum), C.byref(ctr), mode)
        return {k: getattr(ctr, k) for k, _ in OrcCounters._fields_}

    def dump_frames(self, frames, width, height):
        out = np.zeros((height, width, 40), dtype=np.float32)
        self.L.orc_frames_dump(frames, _p(out))
        return out

    def debug_pixel(self, cam, width, height, x, y, sample, bounces=3, nee_samples=4, nee_samples_di=4, flags=0, mode=1):
        cfg = OrcConfig(width, height, bounces, nee_samples, nee_samples_di, flags)
        out = np.zeros(64, dtype=np.float32)
        c = np.ascontiguousarray(cam)
        self.L.orc_debug_pixel(self.h, C.byref(cfg), _p(c), x, y, sample, _p(out), mode)
        return out


def unpack_debug(rec):
    d = {}
    for k, (o, n, is_u) in PIXEL_DEBUG_FIELDS.items():
        v = rec[o:o + n]
        d[k] = v.view(np.uint32).copy() if is_u else v.copy()
    return d


def resolve(accum):
    a = np.ascontiguousarray(accum, dtype=np.float32)
    n = a.size // 4
    out = np.zeros((n, 4), dtype=np.uint8)
    lib().orc_resolve(_p(a), n, _p(out))
    return out.reshape(a.shape[:-1] + (4,))
