"""-m gpu parity tests on the BASELINE configurations AT THEIR REAL SIZE (VERDICT r1 "missing 1"): every TraceRay site of the reference
(shaders/Pass_init_di_v7.hlsl:91-99, Sampler_v7.hlsl:223-229,428-434, Path_Sampler_v7.hlsl:45-52,271-283) compared with the CPU
oracle on C2 (1 M triangles), C3 (10 M instanced triangles, 1000 instances, 4K, TLAS refit) and C5 (10 M triangles): deep BVHs,
quantisation at large extents, the traversal stack and the 1000-instance refit are what small scenes do not reach.
Hits: (instance, primitive, t, u, v) bit-exact against the oracle's BVH2 (which small-scene tests pin to its brute force).
Radiance: the E0 accumulation of a pixel subsample bit-exact (tolerance 0)."""
import numpy as np
import pytest

from util import bits, bounce_rays, compare_hits, oracle_trace_threads, random_rays

pytestmark = pytest.mark.gpu


def _assert_hits_equal(gpu, ref, what):
    r = compare_hits(gpu, ref)
    assert r["inst"] == 0 and r["prim"] == 0 and r["t"] == 0 and r["u"] == 0 and r["v"] == 0, (what, r)
    return r


def _trace_batches(rtdx, ctx, osc, cam, W, H, step, n_random, lo, hi, seed, what):
    """camera rays + incoherent bounces from their hits + uniformly random rays; closest bit-exact, any-hit == closest-exists."""
    rng = np.random.RandomState(seed)
    prim = rtdx.scenes.camera_rays(cam, W, H, step=step)
    g_prim = ctx.trace(prim)
    inc = bounce_rays(rtdx, prim, g_prim, rng)
    rnd = random_rays(rtdx, rng, n_random, lo, hi)
    rays = np.concatenate([prim, inc, rnd])
    ref = oracle_trace_threads(osc, rays)
    gpu = ctx.trace(rays)
    r = _assert_hits_equal(gpu, ref, what)
    ah = ctx.trace(rays, any_hit=True)
    assert np.array_equal(ah["inst"] != rtdx.MISS, ref["inst"] != rtdx.MISS), what
    return rays, ref, r


def test_c2_full_size_vs_oracle(rtdx, orc):
    """C2: 1 051 406 triangles, 1920x1080, bounces 6.  >= 2^20 rays (coherent primaries, incoherent bounces, random) and the E0
    accumulation of every 8th pixel in x and y."""
    sc = rtdx.scenes.mesh_room(n=296)
    W, H = 1920, 1080
    ctx = rtdx.Context(W, H, bounces=6)
    up = ctx.upload_scene(sc)
    osc = orc.OracleScene(sc, up["props"], up["lights"])
    rays, ref, r = _trace_batches(rtdx, ctx, osc, up["camera"], W, H, 2, 120000, (-5.9, 0.1, -5.9), (5.9, 5.9, 5.9), 21, "C2")
    assert r["n"] >= (1 << 20) and r["hits"] > 0.7 * r["n"]
    ctx.reset_accum(); ctx.render_pass(0, 1); ctx.synchronize()
    gpu = ctx.read_accum()
    step = 8
    ref_img, octr = osc.render_threads(up["camera"], W, H, 0, 1, 16, bounces=6, step=step)
    mism = bits(gpu[::step, ::step]) != bits(ref_img[::step, ::step])
    assert mism.sum() == 0, "C2 full size: %d of %d accumulation floats differ" % (mism.sum(), mism.size)
    assert octr["paths"] == gpu[::step, ::step].shape[0] * gpu[::step, ::step].shape[1]
    ctx.close()


def test_c3_full_size_vs_oracle_and_after_refit(rtdx, orc):
    """C3: 10 M instanced triangles behind a TLAS of 1000 instances, 3840x2160, bounces 3, ~0.5 M light triangles.  E0 accumulation
    of every 16th pixel, 2^18+ rays, and the hits again after every instance moved and the TLAS was REFITTED (rdn/Renderer.cpp:594)."""
    sc = rtdx.scenes.instanced_blobs()
    W, H, bounces = 3840, 2160, 3
    ctx = rtdx.Context(W, H, bounces=bounces)
    up = ctx.upload_scene(sc)
    assert len(sc.instances) == 1000 and sc.n_triangles() > 10_000_000
    osc = orc.OracleScene(sc, up["props"], up["lights"])
    ctx.reset_accum(); ctx.render_pass(0, 1); ctx.synchronize()
    gpu = ctx.read_accum()
    step = 16
    ref_img, _ = osc.render_threads(up["camera"], W, H, 0, 1, 16, bounces=bounces, step=step)
    mism = bits(gpu[::step, ::step]) != bits(ref_img[::step, ::step])
    assert mism.sum() == 0, "C3 full size: %d of %d accumulation floats differ" % (mism.sum(), mism.size)
    assert ref_img[::step, ::step, :3].sum() > 0
    _, _, r = _trace_batches(rtdx, ctx, osc, up["camera"], W, H, 6, 100000, (-6, -6, -6), (6, 6, 6), 22, "C3")
    assert r["n"] >= (1 << 18)
    # OnUpdate: every instance gets a new transform (seeded rotation + translation), TLAS refit on the GPU, instance boxes rebuilt in the oracle
    rng = np.random.RandomState(23)
    xms = []
    for _, xm in sc.instances:
        m = np.asarray(xm, dtype=np.float64).reshape(4, 4).T          # column-vector form
        a = rng.uniform(-0.6, 0.6)
        R = np.eye(4); R[0, 0] = np.cos(a); R[0, 2] = np.sin(a); R[2, 0] = -np.sin(a); R[2, 2] = np.cos(a)
        T = np.eye(4); T[:3, 3] = rng.uniform(-0.4, 0.4, size=3)
        xms.append(rtdx.xmmatrix_from_colvec(T @ m @ R))
    props, descs = rtdx.instance_properties(xms, [up["model_ids"][m] for m, _ in sc.instances])
    ctx.set_instances(descs, props)
    osc.set_props(props)
    _, _, r = _trace_batches(rtdx, ctx, osc, up["camera"], W, H, 8, 60000, (-6, -6, -6), (6, 6, 6), 24, "C3 after refit")
    ctx.set_option(rtdx.OPT_TLAS_REBUILD, 1)                            # and the refitted TLAS answers like a fresh build
    before = ctx.trace(rtdx.scenes.camera_rays(up["camera"], W, H, step=8))
    ctx.set_instances(descs, props)
    after = ctx.trace(rtdx.scenes.camera_rays(up["camera"], W, H, step=8))
    assert np.array_equal(before.view(np.uint8), after.view(np.uint8))
    ctx.close()


def test_c5_ten_million_triangles_vs_oracle(rtdx, orc):
    """C5 at 10 M triangles (one BLAS): a 2^20-ray subset of the sweep's coherent and incoherent batches."""
    sc = rtdx.scenes.sphere_in_box(10_000_000)
    assert sc.n_triangles() > 9_900_000
    ctx = rtdx.Context(64, 64)
    up = ctx.upload_scene(sc)
    osc = orc.OracleScene(sc, up["props"], up["lights"])
    cam = rtdx.camera_params(sc.eye, sc.center, sc.up, 1.0)
    rays, ref, r = _trace_batches(rtdx, ctx, osc, cam, 1024, 1024, 2, 262144, (-3.9, -3.9, -3.9), (3.9, 3.9, 3.9), 25, "C5 10M")
    assert r["n"] >= (1 << 19) + 262144
    ctx.close()
