"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes wrapper of oracle/_ref/libref.so: the reference's own shader text (Pathtracer/shaders/*.hlsl)
compiled for the CPU by oracle/ref/make_ref.py.  Only tests/ may import this (it validates the hand-written oracle; nothing times it)."""
import ctypes as C
import os

import numpy as np

from . import make_ref

TRACE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_int)
_libs = {}


def available(bounces=None):
    lib = make_ref.LIB if bounces is None else os.path.join(make_ref.OUT, "libref_b%d.so" % bounces)
    return os.path.isdir(make_ref.SHADERS) or os.path.exists(lib)


def lib(bounces=None):
    """The library generated with the reference's own `#define bounces 3`, or with that one define set to `bounces`."""
    if bounces not in _libs:
        path = make_ref.LIB if bounces is None else os.path.join(make_ref.OUT, "libref_b%d.so" % bounces)
        make_ref.build(bounces=bounces, lib=path)
        L = C.CDLL(path)
        vp, u32 = C.c_void_p, C.c_uint32
        L.ref_set_tracer.argtypes = [vp, vp, C.c_int]
        L.ref_set_scene.argtypes = [u32, vp, vp, vp, vp, u32, vp, vp, vp, u32, vp, u32, vp, u32]
        L.ref_set_camera.argtypes = [vp]
        L.ref_set_time.argtypes = [C.c_float]
        L.ref_alloc_frame.argtypes = [u32, u32]
        L.ref_raygen.argtypes = [C.c_int, u32, u32]
        L.ref_dispatch.argtypes = [C.c_int]
        L.ref_ray_counts.argtypes = [vp, vp, C.c_int]
        L.ref_frames_dump.argtypes = [C.c_int, vp]
        L.ref_read_permanent.argtypes = [vp]
        L.ref_write_permanent.argtypes = [vp]
        L.ref_read_output.argtypes = [vp]
        L.ref_kat_rng.argtypes = [u32, u32, u32, vp, vp]
        L.ref_kat_map_pixel.argtypes = [u32, u32, u32, u32]
        L.ref_kat_map_pixel.restype = u32
        L.ref_kat_srgb.argtypes = [vp, u32, vp]
        L.ref_kat_bsdf.argtypes = [C.c_int, u32, vp, vp, vp, vp, vp]
        L.ref_kat_update_reservoir.argtypes = [C.c_int, vp, vp, C.c_float, C.c_float, vp]
        L.ref_config.argtypes = [vp]
        L.ref_e0.argtypes = [vp]
        _libs[bounces] = L
    return _libs[bounces]


def legacy_available(bounces=None):
    path = make_ref.LEGACY_LIB if bounces is None else os.path.join(make_ref.OUT, "libref_legacy_b%d.so" % bounces)
    return os.path.isdir(make_ref.INCLUDE) or os.path.exists(path)


def legacy_lib(bounces=None):
    """The reference's FIRST estimator (include/RayGen.hlsl + Hit.hlsl + Miss.hlsl + ShadowRay.hlsl) compiled for the CPU; `bounces`
    replaces RayGen.hlsl:63's `uint bounces = 10000000` (the engine's and the oracle's legacy mode cap the path length)."""
    key = ("legacy", bounces)
    if key not in _libs:
        L = C.CDLL(make_ref.build_legacy(bounces=bounces))
        vp, u32 = C.c_void_p, C.c_uint32
        L.ref_set_tracer.argtypes = [vp, vp, C.c_int]
        L.ref_set_scene.argtypes = [u32, vp, vp, vp, vp, u32, vp, vp, vp, u32, vp, u32, vp, u32]
        L.ref_set_camera.argtypes = [vp]
        L.ref_set_time.argtypes = [C.c_float]
        L.ref_alloc_frame.argtypes = [u32, u32]
        L.ref_raygen.argtypes = [C.c_int, u32, u32]
        L.ref_dispatch.argtypes = [C.c_int]
        L.ref_ray_counts.argtypes = [vp, vp, C.c_int]
        L.ref_read_permanent.argtypes = [vp]
        L.ref_write_permanent.argtypes = [vp]
        L.ref_read_output.argtypes = [vp]
        L.ref_config.argtypes = [vp]
        _libs[key] = L
    return _libs[key]


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


class RefScene:
    """Binds a scenes.SceneDesc + the host-produced buffers (instance props, light list) to the reference shaders' resources and
    routes TraceRay to the oracle's ray caster (orc.OracleScene): the reference has no source for traversal / intersection."""

    def __init__(self, scene, props, lights, oracle_scene, trace_mode=1, bounces=None, legacy=False):
        self.L = L = legacy_lib(bounces) if legacy else lib(bounces)
        self.osc = oracle_scene
        nm = len(scene.models)
        self._keep = [np.ascontiguousarray(m["vertices"]) for m in scene.models] + [np.ascontiguousarray(m["indices"]) for m in scene.models]
        vp = (C.c_void_p * nm)(*[a.ctypes.data for a in self._keep[:nm]])
        ip = (C.c_void_p * nm)(*[a.ctypes.data for a in self._keep[nm:]])
        nv = np.array([a.size for a in self._keep[:nm]], dtype=np.uint32)
        ni = np.array([a.size for a in self._keep[nm:]], dtype=np.uint32)
        im = np.array([i[0] for i in scene.instances], dtype=np.uint32)
        self.props = np.ascontiguousarray(props)
        self.ids = np.ascontiguousarray(scene.material_ids, dtype=np.uint32)
        self.mats = np.ascontiguousarray(scene.materials)
        self.lights = np.ascontiguousarray(lights)
        self._keep += [vp, ip, nv, ni, im]
        L.ref_set_scene(nm, C.cast(vp, C.c_void_p), _p(nv), C.cast(ip, C.c_void_p), _p(ni), im.size, _p(im), _p(self.props), _p(self.ids), self.ids.size,
                        _p(self.mats), self.mats.size, _p(self.lights), self.lights.size)
        from .. import orc
        self._fn = C.cast(orc.lib().orc_trace, C.c_void_p)
        L.ref_set_tracer(self._fn, oracle_scene.h, trace_mode)

    def set_props(self, props):
        self.props[...] = props

    def frame(self, width, height):
        self.w, self.h = width, height
        self.L.ref_alloc_frame(width, height)

    def set_camera(self, cam, sample_index):
        c = np.ascontiguousarray(cam).view(np.float32).reshape(-1).copy()
        self.L.ref_set_camera(_p(c))
        self.L.ref_set_time(float(sample_index))

    def dispatch(self, which_pass):
        self.L.ref_dispatch(which_pass)

    def raygen(self, which_pass, x, y):
        self.L.ref_raygen(which_pass, x, y)

    def dump(self, last=False):
        out = np.zeros((self.h, self.w, 40), dtype=np.float32)
        self.L.ref_frames_dump(1 if last else 0, _p(out))
        return out

    def e0(self):
        """Estimator E0 per pixel from the *_current buffers of pass 1 (float4: C.rgb, 1 sampled / 3 emitter)."""
        out = np.zeros((self.h, self.w, 4), dtype=np.float32)
        self.L.ref_e0(_p(out))
        return out

    def output(self):
        out = np.zeros((self.h, self.w, 4), dtype=np.float32)
        self.L.ref_read_output(_p(out))
        return out

    def permanent(self):
        out = np.zeros((self.h, self.w, 4), dtype=np.float32)
        self.L.ref_read_permanent(_p(out))
        return out

    def ray_counts(self, reset=True):
        a, b = C.c_uint64(), C.c_uint64()
        self.L.ref_ray_counts(C.byref(a), C.byref(b), 1 if reset else 0)
        return a.value, b.value


def kat_rng(sx, sy, n, bounces=None):
    out = np.zeros(n, dtype=np.float32); seed = np.zeros(2, dtype=np.uint32)
    lib(bounces).ref_kat_rng(sx, sy, n, _p(out), _p(seed))
    return out, seed


def kat_bsdf(op, mat_id, n, i, o, seed=(1, 2)):
    out = np.zeros(4, dtype=np.float32); s = np.array(seed, dtype=np.uint32)
    a, b, c = (np.ascontiguousarray(v, dtype=np.float32) for v in (n, i, o))
    lib().ref_kat_bsdf(op, mat_id, _p(a), _p(b), _p(c), _p(s), _p(out))
    return out, s
