"""The hand-written oracle (oracle/rtx_oracle.cpp) PINNED against the reference's own shader text.

oracle/_ref/libref.so is /root/reference/Pathtracer/shaders/*.hlsl (+ the include/*_v6.hlsl files they #include) compiled for the CPU by
oracle/ref/make_ref.py: a mechanical token filter + oracle/ref/hlsl_shim.h, with TraceRay routed to the oracle's ray caster (the
reference has no source for traversal / intersection) and then into the reference's OWN ClosestHit / Miss / Shadow shaders.  These tests
run RayGen, RayGen2, RayGen3 and the leaf functions of the reference and demand bit-identical results from the oracle:
  F1 RandomFloat, F2 seeds, F3 MapPixelID, F5 ClosestHit, F7-F11 BSDF, F12-F18 samplers + SamplePathSimple, F14 reservoirs, F16 reconnection,
  F19 estimator E0, F20 accumulation + sRGB, and the ReSTIR temporal + spatial passes.
Documented deviations of the oracle from reference undefined behaviour are the only differences allowed
and each is asserted to be exactly that deviation.  The library is built where /root/reference exists and travels to the GPU box."""
import ctypes as C

import numpy as np
import pytest

from util import bits, hostile_material_scene, host_inputs, load_scene_npz

ref = pytest.importorskip("oracle.ref.ref")
pytestmark = pytest.mark.skipif(not ref.available(), reason="neither /root/reference nor a prebuilt oracle/_ref/libref.so")

import os
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_library_is_generated_from_the_reference_defaults():
    cfg = np.zeros(4, dtype=np.uint32)
    ref.lib().ref_config(_p(cfg))
    assert list(cfg) == [3, 4, 4, 1]          # bounces, nee_samples, nee_samples_DI, bsdf_samples_DI (Common_v7.hlsl:8-11)
    from oracle.ref import make_ref
    if os.path.isdir(make_ref.SHADERS):       # the filter touches spellings only: every line of the reference survives, in order
        import difflib
        src = make_ref.expand(make_ref.UNITS["rg1"]).split()
        gen = make_ref.filtered(make_ref.UNITS["rg1"]).split()
        sm = difflib.SequenceMatcher(None, src, gen, autojunk=False)
        assert sm.ratio() > 0.9
        # tokens the filter removed or respelled: nothing but qualifiers, attributes, literals' suffixes, swizzle calls, references
        touched = [t for tag, i1, i2, j1, j2 in sm.get_opcodes() if tag != "equal" for t in src[i1:i2]]
        assert len(touched) < 0.12 * len(src)


def test_rng_seeds_and_pixel_map_equal_the_reference(orc):
    L, O = ref.lib(), orc.lib()
    rng = np.random.RandomState(1)
    for sx, sy in [(0, 0), (1, 2), (0xFFFFFFFF, 0x9e3779b9)] + [tuple(int(v) for v in rng.randint(0, 2 ** 32, size=2, dtype=np.uint64)) for _ in range(200)]:
        a, sa = ref.kat_rng(sx, sy, 500)
        b = np.zeros(500, dtype=np.float32); sb = np.zeros(2, dtype=np.uint32)
        O.orc_kat_rng(sx, sy, 500, _p(b), _p(sb))
        assert np.array_equal(bits(a), bits(b)) and np.array_equal(sa, sb)
    for w, h in [(1920, 1080), (37, 23), (4, 4), (5, 9)]:
        for _ in range(3000):
            x, y = int(rng.randint(0, w)), int(rng.randint(0, h))
            assert L.ref_kat_map_pixel(w, h, x, y) == O.orc_kat_map_pixel(w, h, x, y)


def test_bsdf_functions_equal_the_reference(rtdx, orc):
    """EvaluateBRDF / BRDF_PDF (both lobes), CalculateStrategyProbabilities, SampleBRDF (both), SelectSamplingStrategy on every material
    class of the bench scene (fp16 MaterialOptimized, GGX roughness 0.1..1, metallic 0/1, the LUT-less default material)."""
    sc = rtdx.scenes.mesh_room(n=4)
    props, descs, lights, cam = host_inputs(rtdx, sc, 8, 8)
    osc = orc.OracleScene(sc, props, lights)
    rs = ref.RefScene(sc, props, lights, osc)
    L, O = ref.lib(), orc.lib()
    rng = np.random.RandomState(2)
    n_mat = len(sc.materials)

    def unit(v):
        return (v / np.linalg.norm(v)).astype(np.float32)
    bad = {}
    N = 6000
    for k in range(N):
        n, i, o = unit(rng.normal(size=3)), unit(rng.normal(size=3)), unit(rng.normal(size=3))
        if k % 3 == 0:
            o = unit(n + 0.3 * rng.normal(size=3)); i = unit(-n + 0.3 * rng.normal(size=3))   # the common configuration
        if k % 7 == 0:
            o = (o * rng.uniform(0.1, 4.0)).astype(np.float32)                                   # un-normalised outgoing (pdf paths)
        mat = int(rng.randint(0, n_mat + 1))                                                    # n_mat = out of bounds -> zeros
        for op in range(8):
            seed = rng.randint(0, 2 ** 32, size=2, dtype=np.uint64).astype(np.uint32)
            sa, sb = seed.copy(), seed.copy()
            a = np.zeros(4, dtype=np.float32); b = np.zeros(4, dtype=np.float32)
            L.ref_kat_bsdf(op, mat, _p(n), _p(i), _p(o), _p(sa), _p(a))
            O.orc_kat_bsdf(osc.h, op, mat, _p(n), _p(i), _p(o), _p(sb), _p(b))
            if not (np.array_equal(bits(a + 0), bits(b + 0)) and np.array_equal(sa, sb)):      # + 0: -0 == +0
                bad[op] = bad.get(op, 0) + 1
    assert not bad, "ops differing from the reference (of %d inputs each): %s" % (N, bad)


def test_bsdf_functions_equal_the_reference_on_hostile_inputs(rtdx, orc):
    """The same eight BSDF entry points with zero, negative-zero, denormal, huge, infinite and NaN components in the normal, incidence and
    outgoing vectors (a quarter of all components): every result equals the reference text's bit for bit (a NaN equals any NaN)."""
    sc = rtdx.scenes.mesh_room(n=4)
    props, descs, lights, cam = host_inputs(rtdx, sc, 8, 8)
    osc = orc.OracleScene(sc, props, lights)
    rs = ref.RefScene(sc, props, lights, osc)
    L, O = ref.lib(), orc.lib()
    n_mat = len(sc.materials)
    specials = [0.0, -0.0, 1.0, -1.0, 1e-30, -1e-30, 1e30, np.inf, -np.inf, np.nan, 1e-45, 0.5, 3.4e38]
    rng = np.random.RandomState(5)

    def vec():
        v = rng.normal(size=3).astype(np.float32)
        v /= np.linalg.norm(v)
        for c in range(3):
            if rng.rand() < 0.25:
                v[c] = np.float32(specials[rng.randint(len(specials))])
        return v
    bad = {}
    for k in range(4000):
        n, i, o = vec(), vec(), vec()
        mat = int(rng.randint(0, n_mat + 1))
        for op in range(8):
            seed = rng.randint(0, 2 ** 32, size=2, dtype=np.uint64).astype(np.uint32)
            sa, sb = seed.copy(), seed.copy()
            a = np.zeros(4, dtype=np.float32); b = np.zeros(4, dtype=np.float32)
            L.ref_kat_bsdf(op, mat, _p(n), _p(i), _p(o), _p(sa), _p(a))
            O.orc_kat_bsdf(osc.h, op, mat, _p(n), _p(i), _p(o), _p(sb), _p(b))
            same = np.array_equal(sa, sb) and all((np.isnan(x) and np.isnan(y)) or bits(np.float32(x) + 0) == bits(np.float32(y) + 0) for x, y in zip(a, b))
            if not same:
                bad[op] = bad.get(op, 0) + 1
    assert not bad, bad


def test_reservoir_updates_equal_the_reference(orc):
    """UpdateReservoir / UpdateReservoir_GI (Reservoir_v7.hlsl:30-80) through the pass-1 comparison below; here the uint16 M arithmetic."""
    L = ref.lib()
    rng = np.random.RandomState(3)
    for gi in (0, 1):
        for _ in range(2000):
            w = np.array([rng.uniform(0, 5)], dtype=np.float32); M = np.array([float(rng.randint(0, 200))], dtype=np.float32)
            wi, M_in = float(np.float32(rng.uniform(0, 3))), float(rng.randint(0, 130))
            seed = rng.randint(0, 2 ** 32, size=2, dtype=np.uint64).astype(np.uint32)
            u, _ = ref.kat_rng(int(seed[0]), int(seed[1]), 1)
            w0, M0 = w.copy(), M.copy()
            acc = L.ref_kat_update_reservoir(gi, _p(w), _p(M), wi, M_in, _p(seed))
            assert bits(w)[0] == bits(np.float32(w0[0] + np.float32(wi)))[0] and M[0] == float((int(M0[0]) + int(M_in)) & 0xFFFF)
            assert acc == int(u[0] < np.float32(wi) / w[0])


SCENES = {
    "mesh_room": lambda rtdx: rtdx.scenes.mesh_room(n=8),                                  # closed room, GGX + diffuse, two instances
    "garage_monke": lambda rtdx: load_scene_npz(rtdx, os.path.join(GOLDEN, "reference_scene.npz")),   # the reference's own assets
    "cornell": lambda rtdx: rtdx.scenes.cornell(),                                         # open front: primary misses (deviation D1)
    "hostile_materials": hostile_material_scene,                                           # Pr = 0, Ks > 1, LUT = 0, zero-area light, missing normals
    "instanced_emitters": lambda rtdx: rtdx.scenes.instanced_blobs(n_models=2, n_side=5, lattice=3, emissive_fraction=0.15),   # 27 instances, many lights, open
}
OPEN_SCENES = ("cornell", "instanced_emitters")                                            # primary rays can miss (deviation D1)


def _primary_miss_mask(orc, osc, cam, W, H, sample, bounces):
    m = np.zeros((H, W), dtype=bool)
    for y in range(H):
        for x in range(W):
            m[y, x] = orc.unpack_debug(osc.debug_pixel(cam, W, H, x, y, sample, bounces=bounces))["hit_inst"][0] == 0xFFFFFFFF
    return m


@pytest.mark.parametrize("name", sorted(SCENES))
@pytest.mark.parametrize("bounces", [3, 6])
def test_pass1_and_estimator_e0_equal_the_reference(rtdx, orc, name, bounces):
    """RayGen (Pass_init_di_v7.hlsl:48-190) run from the reference's text per pixel, then E0 (F19) from the reference's ReconnectDI /
    GetP_Hat_GI on its reservoirs: the oracle's accumulation of the same sample is bit-identical on every pixel whose primary ray hits.
    bounces = 6 is BASELINE config C2's path length (the reference's `#define bounces 3` set to 6 in the generated copy)."""
    if bounces != 3 and not ref.available(bounces):
        pytest.skip("libref_b%d.so not built" % bounces)
    W, H = 48, 32
    sc = SCENES[name](rtdx)
    props, descs, lights, cam = host_inputs(rtdx, sc, W, H)
    osc = orc.OracleScene(sc, props, lights)
    rs = ref.RefScene(sc, props, lights, osc, bounces=None if bounces == 3 else bounces)
    rs.frame(W, H)
    for sample in (0, 11):
        rs.set_camera(cam, sample)
        rs.dispatch(1)
        e0 = rs.e0()
        rc, rsh = rs.ray_counts()
        acc, ctr = osc.render(cam, W, H, sample, 1, bounces=bounces)
        miss = _primary_miss_mask(orc, osc, cam, W, H, sample, bounces)
        if name not in OPEN_SCENES:
            assert not miss.any()
        differ = (bits(e0[..., :3] + 0) != bits(acc[..., :3])).any(-1)
        assert not (differ & ~miss).any(), "%d pixels differ from the reference" % int((differ & ~miss).sum())
        assert (acc[..., 3] == 1).all()
        if not miss.any():
            # ray counts: the oracle issues the reference's rays minus the visibility rays whose result cannot matter (deviation D6)
            # (with the hostile materials most candidates have no contribution at all: most visibility rays are skipped)
            # and more paths end on emitters without contribution, where the reference traces on for nothing: deviation D4)
            slack, slack_c = (2.0, 1.03) if name == "hostile_materials" else (1.2, 1.01)
            assert ctr["closest_rays"] <= rc <= ctr["closest_rays"] * slack_c + 4 and ctr["shadow_rays"] <= rsh <= ctr["shadow_rays"] * slack + 8


def _frames(rtdx, orc, sc, W, H, script, allow_miss=False):
    """The reference's frame loop (rdn/Renderer.cpp:431-452,611-673) on both sides: per frame OnUpdate (camera with prevView /
    prevProjection, instance matrices with prev*), then RayGen, RayGen2, RayGen3."""
    props, descs, lights, cam = host_inputs(rtdx, sc, W, H)
    osc = orc.OracleScene(sc, props, lights)
    rs = ref.RefScene(sc, props, lights, osc)
    rs.frame(W, H)
    frames = osc.new_frames(W, H)
    acc = np.zeros((H, W, 4), dtype=np.float32)
    model_ids = [i[0] for i in sc.instances]
    prev_props = props
    worst = 0
    for f, (new_cam, xforms) in enumerate(script):
        if xforms is not None:
            props, _ = rtdx.instance_properties(xforms, model_ids)
            if f > 0:
                for k in ("objectToWorld", "objectToWorldInverse", "objectToWorldNormal"):
                    props["prev" + k[0].upper() + k[1:]] = prev_props[k]
            osc.set_props(props); rs.set_props(props)
            prev_props = props
        if new_cam is not None:
            new_cam = new_cam.copy()
            new_cam["prevView"] = cam["view"]; new_cam["prevProjection"] = cam["projection"]
            cam = new_cam
        elif f > 0:
            cam = cam.copy(); cam["prevView"] = cam["view"]; cam["prevProjection"] = cam["projection"]
        rs.set_camera(cam, f)
        for p in (1, 2, 3):
            rs.dispatch(p)
        view_changed = np.abs(cam["view"] - cam["prevView"]).max() > 2e-5
        if view_changed:
            acc[:] = 0                                   # the oracle's host-side form of Pass_spat_di_v7.hlsl:407-423
        osc.render_frame(cam, W, H, f, frames, acc, bounces=3)
        a, b = rs.dump(last=True), osc.dump_frames(frames, W, H)
        kind = b[..., 35].copy(); b[..., 35] = 0         # oracle-internal tag: 0 primary miss, 2 sampled, 3 emitter
        differ = (bits(a) != bits(b)).any(-1)
        if allow_miss:
            assert not (differ & (kind != 0)).any(), "frame %d: %d non-miss pixels differ" % (f, int((differ & (kind != 0)).sum()))
        else:
            assert not differ.any(), "frame %d: %d pixels' reservoirs / samples differ from the reference" % (f, int(differ.sum()))
        # accumulation (F20): identical except where the primary hit is an emitter (deviation D5: the reference bypasses gPermanentData)
        pm = rs.permanent()
        dacc = (bits(pm + 0) != bits(acc + 0)).any(-1)       # + 0: the first frame stores -0 as it is, the oracle's running sum holds +0
        assert not (dacc & (kind == 2)).any(), "frame %d: accumulation differs on %d sampled pixels" % (f, int((dacc & (kind == 2)).sum()))
        assert (pm[kind == 3] == 0).all()
        # output (sRGB -> UNORM8) on the sampled pixels.  On the frame of a view change the reference displays the average that still
        # contains the stale history (averagedColor is computed at :405 BEFORE the reset at :407-423); gPermanentData, compared above, is
        # identical after the frame (DESIGN.md deviation D14), so the images agree again from the next frame on.
        if view_changed:
            continue
        out = np.zeros((H, W, 4), dtype=np.float32); rs.L.ref_read_output(_p(out))
        u8 = (np.clip(out[..., :3], 0, 1) * np.float32(255.0) + np.float32(0.5)).astype(np.int32)
        assert np.array_equal(u8[kind == 2], orc.resolve(acc)[..., :3].astype(np.int32)[kind == 2])
        worst = max(worst, int(differ.sum()))
    osc.free_frames(frames)
    return worst


def test_restir_frames_static_equal_the_reference(rtdx, orc):
    """RayGen + RayGen2 (temporal) + RayGen3 (spatial, shade, accumulate, sRGB) from the reference's text, 4 frames."""
    _frames(rtdx, orc, SCENES["mesh_room"](rtdx), 48, 32, [(None, None)] * 4)
    _frames(rtdx, orc, SCENES["cornell"](rtdx), 40, 32, [(None, None)] * 3, allow_miss=True)
    _frames(rtdx, orc, SCENES["hostile_materials"](rtdx), 40, 30, [(None, None)] * 3)          # reuse across NaN-prone materials and a zero-area light


def test_restir_frames_reference_scene_with_rotating_instance_equal_the_reference(rtdx, orc):
    """garage.obj + monke.obj, instance 1 = rotY(1.57 * frame) as rdn/Renderer.cpp:444-449 animates it, default camera."""
    sc = SCENES["garage_monke"](rtdx)
    base = [np.asarray(i[1], dtype=np.float64).reshape(4, 4).T for i in sc.instances]

    def at(t):
        a = np.float32(1.57) * t
        R = np.eye(4); R[0, 0] = np.cos(a); R[0, 2] = np.sin(a); R[2, 0] = -np.sin(a); R[2, 2] = np.cos(a)
        return [rtdx.xmmatrix_from_colvec(base[0]), rtdx.xmmatrix_from_colvec(R)]
    _frames(rtdx, orc, sc, 48, 32, [(None, at(0.0)), (None, at(0.02)), (None, at(0.04)), (None, None)])


def test_restir_frames_moving_camera_equal_the_reference(rtdx, orc):
    sc = SCENES["mesh_room"](rtdx)
    W, H = 40, 32
    cams = [None, rtdx.camera_params((sc.eye[0] + 0.05, sc.eye[1], sc.eye[2]), sc.center, sc.up, W / H),
            rtdx.camera_params((sc.eye[0] + 0.10, sc.eye[1] + 0.02, sc.eye[2]), sc.center, sc.up, W / H), None]
    _frames(rtdx, orc, sc, W, H, [(c, None) for c in cams])


# ---- the reference's first estimator: include/RayGen.hlsl + include/Hit.hlsl + include/Miss.hlsl -----------------
@pytest.mark.parametrize("name", sorted(SCENES))
@pytest.mark.parametrize("bounces", [2, 12])
def test_legacy_estimator_equals_the_reference_text(rtdx, orc, name, bounces):
    """RayGen (include/RayGen.hlsl:48-187: jittered camera ray, path loop, Russian roulette after depth 3, accumulation) and ClosestHit
    (include/Hit.hlsl:59-357: face-forwarded normals, RIS over 10 light candidates, the nested shadow TraceRay, BSDF sample, MIS on
    emitter hits) run from the reference's own text, one DispatchRays per sample: gPermanentData and the closest / shadow ray counts of
    the oracle's `legacy` estimator (what RTX_FLAG_LEGACY_RR is tested against) are bit-identical on every pixel, for one sample and
    accumulated over three.  The only edit to the text is RayGen.hlsl:63's `uint bounces = 10000000` -> the cap under test (D13);
    the legacy text's older Material / InstanceProperties / LightTriangle layouts are filled field by field (ref_legacy_harness.cpp)."""
    if not ref.legacy_available(bounces):
        pytest.skip("libref_legacy_b%d.so not built" % bounces)
    from oracle.ref import make_ref
    if os.path.isdir(make_ref.INCLUDE) and name == "cornell":      # the filter touches spellings only, here too
        import difflib
        for unit in ("leg_rg", "leg_hit"):
            src = make_ref.expand(make_ref.LEGACY_UNITS[unit]).split()
            gen = make_ref.filtered_legacy(make_ref.LEGACY_UNITS[unit], bounces).split()
            sm = difflib.SequenceMatcher(None, src, gen, autojunk=False)
            touched = [t for tag, i1, i2, j1, j2 in sm.get_opcodes() if tag != "equal" for t in src[i1:i2]]
            assert sm.ratio() > 0.9 and len(touched) < 0.12 * len(src), (unit, sm.ratio(), len(touched), len(src))
    W, H = 40, 24
    sc = SCENES[name](rtdx)
    props, descs, lights, cam = host_inputs(rtdx, sc, W, H)
    osc = orc.OracleScene(sc, props, lights)
    rs = ref.RefScene(sc, props, lights, osc, bounces=bounces, legacy=True)
    cfg = np.zeros(4, dtype=np.uint32)
    rs.L.ref_config(_p(cfg))
    assert cfg[0] == 10                                                   # RIS_M (include/Common.hlsl:8)
    flags = orc.FLAG_LEGACY_RR | orc.FLAG_JITTER                          # RayGen.hlsl:86-87 always jitters
    rs.frame(W, H)
    rs.ray_counts()                                                       # (the library is shared between tests: start the counters at zero)
    zeros = np.zeros((H, W, 4), dtype=np.float32)
    for sample in (0, 5):
        rs.L.ref_write_permanent(_p(zeros))
        rs.set_camera(cam, sample)
        rs.dispatch(1)
        a, rc = rs.permanent(), rs.ray_counts()
        b, ctr = osc.render(cam, W, H, sample, 1, bounces=bounces, flags=flags)
        assert rc == (ctr["closest_rays"], ctr["shadow_rays"]), (sample, rc, ctr)
        assert np.array_equal(bits(a + 0), bits(b + 0)), (sample, int((bits(a + 0) != bits(b + 0)).any(axis=2).sum()))
        assert np.isfinite(a).all() and a[..., :3].sum() > 0 and ((a[..., 3] == 1).all() or name == "hostile_materials")   # (non-finite samples are dropped: RayGen.hlsl:145)
    # three samples accumulated by the reference's own temporal accumulation (RayGen.hlsl:140-156), then its averaged output (:176-183)
    rs.L.ref_write_permanent(_p(zeros))
    for sample in range(3):
        rs.set_camera(cam, sample)
        rs.dispatch(1)
    a = rs.permanent()
    b, _ = osc.render(cam, W, H, 0, 3, bounces=bounces, flags=flags)
    assert np.array_equal(bits(a + 0), bits(b + 0))
    # The legacy RayGen displays sum / max(count BEFORE this frame, 1) (RayGen.hlsl:142,176-177: frameCount is read before the
    # accumulation): its third frame shows the three-sample sum over 2.  The engine resolves every accumulation buffer with the current
    # count (F20, Common_v7 / Pass_spat_di_v7.hlsl:425-444); only gPermanentData is the legacy row's contract.  Asserted as found:
    out = rs.output()
    if name != "hostile_materials":        # (there the count before the frame differs per pixel: dropped samples)
        assert np.array_equal(bits(out[..., :3] + 0), bits((b[..., :3] / np.float32(2.0)) + 0))
