// wavefront.cu — the logic stages of the wavefront path tracer.
//
// The reference runs one megakernel raygen thread per pixel (shaders/Pass_init_di_v7.hlsl:48-190) that calls TraceRay
// up to 5 + bounces times.  Here the same per-path program is cut at every TraceRay into stages; between stages
// the rays sit in compacted queues that the persistent traversal kernels (trace.cu) drain:
//
//   generate -> [extend] -> shade_primary -> [extend] -> di_finish -> [connect DI] [extend] -> gi_step(0) ->
//   ([extend] -> gi_step(i))* -> [connect GI] -> finalize -> accumulate
//
// Each path draws its random numbers in exactly the order of the reference's sequential program (the RNG state
// travels in the path state), so the result is independent of queue order.  Path state is a structure of float4
// arrays (wavefront.h); queue slots are claimed with one atomic per warp (ballot + popc compaction).
#include "trace.h"
#include "wavefront.h"
#include "wave_dev.cuh"

namespace rtx {
// This file is compiled twice (build.py): as is — every float operation an IEEE binary32 operation in source order (-fmad=false), the
// parity mode whose results are bit-identical to the CPU oracle — and with -DRTX_FAST_MATH -fmad=true -use_fast_math into rtx::fast
// (FMA contraction, approximate division / sqrt / rsqrt / sincos), the opt-in RTX_FLAG_FAST_MATH mode of rtx_render_pass.
#ifdef RTX_FAST_MATH
namespace fast {
#endif

#define CKE(call)                            \
    do {                                     \
        cudaError_t e__ = (call);            \
        if (e__ != cudaSuccess) return e__;  \
    } while (0)

#define WF_BLOCK 128
#ifndef RTX_SP_MINB
#define RTX_SP_MINB 8     // resident CTAs per SM the register budget of k_shade_primary / k_di_finish is set for: 64 registers with ~150 B
                          // of spills measured best (di_finish 0.344 -> 0.305 ms per C2 pass against the unconstrained 108 registers)
#endif
#ifndef RTX_GI_BLOCK
#define RTX_GI_BLOCK 384    // CTA size of k_gi_step: 2 x 384 threads per SM at 85 registers measured best (profiles/)
#endif
#ifndef RTX_GI_MINB
#define RTX_GI_MINB 2       // resident CTAs per SM the register budget of k_gi_step is set for
#endif
#define GI_STAGED 4         // path-state planes k_gi_step (iterations >= 1) stages through shared memory: 4 x 16 B x RTX_GI_BLOCK = 24 KB per CTA
#ifndef RTX_GI_STAGE
#define RTX_GI_STAGE 1      // 1: stage them (cp.async into per-thread shared-memory slots); 0: plain loads.  Bit-identical either way.
#endif                      // Measured on B200 (profiles/r02_gi_step_staging.txt): with the twelve planes a bounce read at the start of round 2
                            // (72 KB per CTA, 144 KB per SM taken from the L1 that holds materials and lights) k_gi_step 3.81 -> 4.11 ms per C2
                            // pass; with today's four planes (48 KB per SM) the stage alone is unchanged and the C2 pass with its concurrent
                            // path ranges gains 1.2 % (7.41 -> 7.33 ms), fast-math 6.01 -> 5.94; C3 13.11 -> 13.17

// camera ray, shaders/Pass_init_di_v7.hlsl:59,80-95
__device__ __forceinline__ void CameraRay(const rtx_camera_params* cam, uint32_t W, uint32_t H, uint32_t x, uint32_t y, float jx, float jy,
                                          f3& o, f3& dir) {
    float dimx = (float)W, dimy = (float)H;
    o = mul43(cam->viewI, 0.0f, 0.0f, 0.0f, 1.0f);
    float dx = (((float)x + jx) / dimx) * 2.0f - 1.0f;
    float dy = (((float)y + jy) / dimy) * 2.0f - 1.0f;
    f3 target = mul43(cam->projectionI, dx, -dy, 1.0f, 1.0f);
    f3 d = mul43(cam->viewI, target.x, target.y, target.z, 0.0f);
    dir = normalize3(d);
}

__global__ void __launch_bounds__(WF_BLOCK)
k_generate(StateView st, SceneData S, PartMap pm, uint32_t np, RayQueue q0, const rtx_camera_params* __restrict__ cam, uint32_t W, uint32_t H,
           const uint32_t* __restrict__ first_sample_ptr, uint32_t flags, float* __restrict__ vis_di, float* __restrict__ vis_gi,
           unsigned long long* ray_counters) {
    const uint32_t first_sample = *first_sample_ptr;     // a device word, so that a captured pass (CUDA graph) can be replayed for any sample
    // paths [p0, p0 + np) of the pass (one part of the frame, see wave_render_pass); queue slot = index within the part
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) { *q0.count = np; atomicAdd(&ray_counters[2], (unsigned long long)np); }
    const bool live = k < np;
    f3 o = mk3(0, 0, 0), dir = mk3(0, 0, 1);
    if (live) {
        const uint32_t p = pm.path(k);
        const uint32_t npx = W * H;
        const uint32_t pixel = p % npx, s = p / npx;
        const uint32_t x = pixel % W, y = pixel / W;
        uint2 seed = init_seed(x, y, 1u, first_sample + s);
        float jx = 0.0f, jy = 0.0f;
        if (flags & RTX_FLAG_JITTER) { jx = RandomFloat(seed); jy = RandomFloat(seed); }
        CameraRay(cam, W, H, x, y, jx, jy, o, dir);
        q0.o_tmin[k] = f4(o, 0.0001f);
        q0.d_tmax[k] = f4(dir, 10000.0f);
        q0.pid[k] = p;
        st.seed[p] = seed;
        st.at(SP_O, p) = f4(-dir, 0.0f);            // (plain stores: k_shade_primary reads them one launch later)
        st.at(SP_RESULT, p) = make_float4(0, 0, 0, 0);
        vis_di[p] = 1.0f; vis_gi[p] = 1.0f;
    }
    if (q0.order) {             // trace order of the primary rays (every lane of the warp reaches this point)
        const unsigned mask = __ballot_sync(0xffffffffu, live);
        if (mask) record_order(q0, mask, live, live && ray_is_heavy(S, o, 0.0001f, dir, 10000.0f), k);
    }
}

// shaders/Reservoir_v7.hlsl:57-80 / :30-53 on unpacked fields
__device__ __forceinline__ bool UpdateReservoir(f3& rx, f3& rn, f3& rL, float& w_sum, float wi, f3 x2, f3 n2, f3 L2, uint2& seed) {
    w_sum += wi;
    if (RandomFloat(seed) < wi / w_sum) { rx = x2; rn = n2; rL = q16v(L2); return true; }
    return false;
}

// SampleLightNEE with useVisibility=false, shaders/Sampler_v7.hlsl:273-396 (call site :677-693)
// `lobe` = make_lobe_ctx(material, normal, on = normalize3(outgoing), CalculateStrategyProbabilities(material, on, normal)):
// the part of :348-374 that is the same for every candidate of one shading point.
__device__ __forceinline__ void SampleLightNEE(const SceneData& S, float& pdf_light, float& pdf_bsdf, float& p_hat, uint2& seed, f3 worldOrigin,
                                               f3 normal, const LobeCtx& lobe, f3& emission, f3& x2, f3& n2) {
    LightSample ls;
    SampleLightPoint(S, worldOrigin, seed, ls);
    x2 = ls.point; n2 = ls.normal_l;
    float cos_theta_x = dot3(normal, ls.L_norm);
    float cos_theta_y = dot3(ls.normal_l, -ls.L_norm);
    float G = fmaxf((cos_theta_y * cos_theta_x) / ls.dist2, RTX_EPS);
    emission = ls.emission;
    f3 brdf_light; float P;
    lobe_FP<true, true>(lobe, -ls.L_norm, true, cos_theta_y, ls.dist2, brdf_light, P);
    p_hat = length3((ls.emission * brdf_light) * G);
    pdf_light = fmaxf(RTX_EPS, ls.pdf_l);
    pdf_bsdf = P;
}

// ---- stage: primary hit -> RIS over NEE candidates -> BSDF candidate ray (Pass_init_di_v7.hlsl:99-159, Sampler_v7.hlsl:653-700)
__global__ void __launch_bounds__(WF_BLOCK, RTX_SP_MINB)
k_shade_primary(StateView st, SceneData S, RayQueue qin, const float4* __restrict__ hit_a, const uint32_t* __restrict__ hit_inst,
                RayQueue qout, unsigned long long* ray_counters) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = *qin.count;
    if (j == 0) atomicAdd(&ray_counters[0], (unsigned long long)n);
    bool emit = false; f3 ro = mk3(0, 0, 0), rd = mk3(0, 0, 0); uint32_t pid = 0;
    if (j < n) {
        pid = ld_stream(qin.pid + j);
        const uint32_t inst = ld_stream(hit_inst + j);
        if (inst != 0xFFFFFFFFu) {                                  // miss: radiance 0 (DESIGN.md deviation D1)
            const float4 ha = ld_stream(hit_a + j);
            const f3 o = xyz(ld_stream(qin.o_tmin + j)), d = xyz(ld_stream(qin.d_tmax + j));
            HitInfo payload;
            ClosestHit(S, o, d, ha.x, ha.y, ha.z, __float_as_uint(ha.w), inst, payload);
            f3 ke_full;
            const MatOpt mat = load_matopt(S, payload.materialID, &ke_full);
            if (length3(ke_full) > 0.0f) {                          // :103-106 ; L1 = half3(Ke) (deviation D5: accumulated)
                st.stt(SP_RESULT, pid, f4(mat.Ke, 3.0f));          // .w: 0 miss, 3 emitter, 1 sampled (2 once finalized)
                st.stt(SP_X1, pid, f4u(mk3(0, 0, 0), payload.materialID));
                st.stt(SP_DI_L2, pid, f4u(mk3(0, 0, 0), inst));
            } else {
                uint2 seed = st.ld_seed(pid);
                const f3 outgoing = -d;
                const uint32_t strategy = SelectSamplingStrategy(S, mat, outgoing, payload.hitNormal, seed);
                f3 rx = mk3(0, 0, 0), rn = mk3(0, 0, 0), rL = mk3(0, 0, 0); float w_sum = 0.0f;
                const float fM1 = (float)S.nee_samples_di, fM2 = 1.0f;
                const f3 on = normalize3(outgoing);
                float p_d, p_s;
                CalculateStrategyProbabilities(S, mat, on, payload.hitNormal, p_d, p_s);
                const LobeCtx lobe = make_lobe_ctx(S, mat, payload.hitNormal, on, p_d, p_s);
                for (uint32_t i = 0; i < S.nee_samples_di; i++) {
                    float pdf_light = 0.0f, pdf_bsdf = 0.0f, p_hat = 0.0f; f3 emission, x2, n2;
                    SampleLightNEE(S, pdf_light, pdf_bsdf, p_hat, seed, payload.hitPosition, payload.hitNormal, lobe, emission, x2, n2);
                    float mi = pdf_light / (fM1 * pdf_light + fM2 * pdf_bsdf);
                    float wi = (mi * p_hat) / pdf_light;
                    if (p_hat > 0.0f) UpdateReservoir(rx, rn, rL, w_sum, wi, x2, n2, emission, seed);
                }
                const f3 sample = SampleBRDF(strategy, mat, outgoing, payload.hitNormal, seed);   // Sampler_v7.hlsl:218-220
                emit = true; ro = payload.hitPosition; rd = sample;
                st.stt(SP_X1, pid, f4u(payload.hitPosition, payload.materialID));
                st.stt(SP_N1, pid, f4(payload.hitNormal, 0.0f));
                st.stt(SP_O, pid, f4(outgoing, 0.0f));
                st.st_seed(pid, seed);
                st.stt(SP_DI_X2, pid, f4(rx, w_sum));
                st.stt(SP_DI_N2, pid, f4(rn, 0.0f));
                st.stt(SP_DI_L2, pid, f4u(rL, inst));              // .w: primary-hit instance (SampleData::objID)
                st.stt(SP_RESULT, pid, make_float4(0, 0, 0, 1.0f));
            }
        }
    }
    push_ray2(qout, emit, emit && ray_is_heavy(S, ro, RTX_S_BIAS, rd, 10000.0f), ro, RTX_S_BIAS, rd, 10000.0f, pid);
}

// ---- stage: BSDF candidate of the DI reservoir, DI visibility ray, first indirect ray
// (Sampler_v7.hlsl:231-270,729-735; Pass_init_di_v7.hlsl:161-167; Path_Sampler_v7.hlsl:13-52)
__global__ void __launch_bounds__(WF_BLOCK, RTX_SP_MINB)
k_di_finish(StateView st, SceneData S, RayQueue qin, const float4* __restrict__ hit_a, const uint32_t* __restrict__ hit_inst,
            RayQueue q_shadow, RayQueue qout, unsigned long long* ray_counters) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = *qin.count;
    if (j == 0) atomicAdd(&ray_counters[0], (unsigned long long)n);
    bool emit = false, emit_sh = false; uint32_t pid = 0;
    f3 ro = mk3(0, 0, 0), rd = mk3(0, 0, 0), so = mk3(0, 0, 0), sd = mk3(0, 0, 0); float stmax = 0.0f;
    if (j < n) {
        pid = ld_stream(qin.pid + j);
        const float4 a0 = st.ld(SP_X1, pid), a1 = st.ld(SP_N1, pid), a2 = st.ld(SP_O, pid);
        const f3 x1 = xyz(a0), hitNormal = xyz(a1), o = xyz(a2);
        const uint32_t mID = __float_as_uint(a0.w);
        uint2 seed = st.ld_seed(pid);
        const MatOpt mat = load_matopt(S, mID, nullptr);
        float4 b0 = st.ld(SP_DI_X2, pid);
        const float4 b2 = st.ld(SP_DI_L2, pid);
        f3 rx = xyz(b0), rn = xyz(st.ld(SP_DI_N2, pid)), rL = xyz(b2); float w_sum = b0.w;
        const f3 sample = xyz(ld_stream(qin.d_tmax + j));
        const uint32_t inst = ld_stream(hit_inst + j);
        if (inst != 0xFFFFFFFFu) {                                  // miss => materials[MISS] reads 0 => p_hat = 0
            const float4 ha = ld_stream(hit_a + j);
            HitInfo sp;
            ClosestHit(S, x1, sample, ha.x, ha.y, ha.z, __float_as_uint(ha.w), inst, sp);
            float4 kd, ks, ke, pr;
            fetch_material_head(S, sp.materialID, kd, ks, ke, pr);
            const float Ke = (ke.x + ke.y) + ke.z;
            if (Ke > RTX_EPS) {
                const f3 emission = mk3(ke.x, ke.y, ke.z);
                f3 L = sp.hitPosition - x1;
                float dist = length3(L);
                float dist2 = dist * dist;
                float cos_theta = dot3(sp.hitNormal, -sample);
                float pdf_light = (Ke / 3.0f) / __ldg(&S.lights[0].total_weight);
                float p_d, p_s;
                f3 on = normalize3(o);
                CalculateStrategyProbabilities(S, mat, on, hitNormal, p_d, p_s);
                // F sees V = normalize3(on), the pdf sees V = normalize3(o) = on (Sampler_v7.hlsl:248-261)
                const bool lo = (S.cfg_flags & RTX_FLAG_LAMBERT_ONLY) != 0u;
                const f3 N = lo ? hitNormal : normalize3(hitNormal);
                const LobeCtx lobeF = make_lobe_ctx_nv(S, mat, hitNormal, N, lo ? on : normalize3(on), p_d, p_s);
                const LobeCtx lobeP = make_lobe_ctx_nv(S, mat, hitNormal, N, on, p_d, p_s);
                f3 brdf, unusedF; float pdf_bsdf, unusedP;
                lobe_FP<true, false>(lobeF, -sample, false, 1.0f, 1.0f, brdf, unusedP);
                lobe_FP<false, true>(lobeP, -sample, true, cos_theta, dist2, unusedF, pdf_bsdf);
                float ndot = dot3(hitNormal, sample);
                float p_hat = length3((((brdf * emission) * ndot) * cos_theta) / dist2);
                float mi = pdf_bsdf / ((float)S.nee_samples_di * pdf_light + 1.0f * pdf_bsdf);
                float wi = (mi * p_hat) / pdf_bsdf;
                if (p_hat > 0.0f) UpdateReservoir(rx, rn, rL, w_sum, wi, sp.hitPosition, sp.hitNormal, emission, seed);
            }
        }
        // GetP_Hat(..., true): Sampler_v7.hlsl:163-171
        const f3 n1 = normalize3(hitNormal);
        const f3 rdi = ReconnectDI(S, x1, n1, rx, rn, rL, o, mat);
        const float f_g = length3(rdi);
        if (f_g > RTX_EPS) {                                        // deviation D6; VisibilityCheck :86-104
            const f3 d21 = rx - x1;
            const float dist = length3(d21);
            emit_sh = true;
            so = x1 + normalize3(n1) * RTX_S_BIAS;
            sd = normalize3(d21);
            stmax = fmaxf(dist - 10.0f * RTX_S_BIAS, 2.0f * RTX_S_BIAS);
        }
        st.stt(SP_DI_X2, pid, f4(rx, w_sum));
        st.stt(SP_DI_N2, pid, f4(rn, f_g));
        st.stt(SP_DI_L2, pid, f4(rL, b2.w));
        st.stt(SP_DI_R, pid, f4(rdi, 0.0f));
        // SamplePathSimple step 1: Path_Sampler_v7.hlsl:24-52
        const f3 outgoing = normalize3(o);
        const uint32_t strategy = SelectSamplingStrategy(S, mat, outgoing, hitNormal, seed);
        const f3 s2 = SampleBRDF(strategy, mat, outgoing, hitNormal, seed);
        emit = true; ro = x1; rd = s2;
        st.st_seed(pid, seed);            // (SP_N1 / SP_O keep what k_shade_primary wrote)
        // the path state of SamplePathSimple's start (origin = x1, normal, outgoing = normalize(o), acc_f = 1, empty GI reservoir) is not
        // stored: k_gi_step<ITER0> derives it from SP_X1 / SP_N1 / SP_O and constants
    }
    push_ray_pair(q_shadow, emit_sh, emit_sh && ray_is_heavy(S, so, 0.0f, sd, stmax), so, 0.0f, sd, stmax,
                  qout, emit, emit && ray_is_heavy(S, ro, RTX_S_BIAS, rd, 10000.0f), ro, RTX_S_BIAS, rd, 10000.0f, pid);
}

// SampleLightNEE_GI with useVisibility=false, Sampler_v7.hlsl:508-647 (call site Path_Sampler_v7.hlsl:133-151)
// `lobe` = make_lobe_ctx(material, normal, on = normalize3(outgoing), CalculateStrategyProbabilities(material, on, normal))
__device__ __forceinline__ f3 SampleLightNEE_GI(const SceneData& S, float& pdf_light, float& pdf_bsdf, f3& x2_pos, uint2& seed, f3 origin, f3 normal,
                                                const LobeCtx& lobe, f3 acc_l, float acc_pdf, f3& throughput, f3& emission) {
    LightSample ls;
    SampleLightPoint(S, origin, seed, ls);
    x2_pos = ls.point;
    float cos_theta_x = fabsf(dot3(normal, ls.L_norm));
    if (cos_theta_x < RTX_EPS) cos_theta_x = 0.0f;
    float cos_theta_y = fabsf(dot3(ls.normal_l, -ls.L_norm));
    if (cos_theta_y < RTX_EPS) cos_theta_y = 0.0f;
    float G = cos_theta_x;
    f3 brdf_light; float P;
    lobe_FP<true, true>(lobe, -ls.L_norm, false, 1.0f, 1.0f, brdf_light, P);
    if (cos_theta_y > 0.0f) pdf_light = (fmaxf(RTX_EPS, ls.pdf_l) * ls.dist2) / cos_theta_y;
    pdf_bsdf = P;
    acc_pdf *= pdf_light;
    acc_l = acc_l * (brdf_light * G);
    throughput = brdf_light * G;
    emission = ls.emission;
    if (acc_pdf > 0.0f) return (ls.emission * acc_l) / acc_pdf;
    return mk3(0, 0, 0);
}

// ---- stage: one step of the indirect path (Path_Sampler_v7.hlsl:54-269 + Sampler_v7.hlsl:436-504).
// iter == 0 consumes the hit of the initial indirect ray; iter >= 1 consumes the BSDF ray of loop iteration iter-1.
// If the path goes on and iter < bounces it runs iteration `iter`'s NEE candidates and emits its BSDF ray.
template <bool ITER0>       // two instantiations: the hit-consuming halves differ, and the kernel is I-cache bound
__global__ void __launch_bounds__(RTX_GI_BLOCK, RTX_GI_MINB)
k_gi_step(StateView st, SceneData S, RayQueue qin, const float4* __restrict__ hit_a, const uint32_t* __restrict__ hit_inst,
          RayQueue q_shadow, RayQueue qout, uint32_t iter, unsigned long long* ray_counters, const uint32_t* __restrict__ perm) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = *qin.count;
    if (i == 0) atomicAdd(&ray_counters[0], (unsigned long long)n);
    // RTX_FLAG_SORT_MATERIAL: queue entries are visited in material-binned order (perm), otherwise in queue order
    const uint32_t j = (perm != nullptr && i < n) ? perm[i] : i;
    bool emit = false, emit_sh = false; uint32_t pid = 0;
    f3 ro = mk3(0, 0, 0), rd = mk3(0, 0, 0), so = mk3(0, 0, 0), sd = mk3(0, 0, 0); float stmax = 0.0f;
    if (j < n) {
        pid = ld_stream(qin.pid + j);
        // ---- path state staged through shared memory (iterations >= 1, build option RTX_GI_STAGE): the four 16-byte planes a bounce reads are requested with
        // cp.async straight into this thread's shared-memory slots — no registers are held while they are in flight (at 80 registers the
        // compiler could keep two or three of the float4 loads outstanding and serialised the rest next to their uses) — and they
        // land while the hit's attributes (instance -> model -> attribute record -> material) are fetched, a chain of dependent
        // loads that does not need the state.  Each thread reads back only its own slots, so no CTA barrier is needed.
        extern __shared__ float4 s_stage[];
        constexpr bool STAGE = RTX_GI_STAGE && !ITER0;
        if (STAGE) {
            const int planes[GI_STAGED] = {SP_ORIGIN, SP_ACC_F, SP_ACC_FR, SP_GI_SC};
#pragma unroll
            for (int k = 0; k < GI_STAGED; k++) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(&s_stage[k * RTX_GI_BLOCK + threadIdx.x]);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(&st.at(planes[k], pid)));
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        const f3 sample = xyz(ld_stream(qin.d_tmax + j));
        const uint32_t inst = ld_stream(hit_inst + j);
        float4 ha = make_float4(0, 0, 0, 0);
        HitInfo sp; MatOpt hm = MatOpt(); f3 ke_full = mk3(0, 0, 0);
        if (inst != 0xFFFFFFFFu) {                                  // hit attributes: independent of the path state (hitPosition is set below)
            ha = ld_stream(hit_a + j);
            ClosestHit(S, mk3(0, 0, 0), sample, ha.x, ha.y, ha.z, __float_as_uint(ha.w), inst, sp);
            hm = load_matopt(S, sp.materialID, &ke_full);
        }
        if (STAGE) asm volatile("cp.async.wait_group 0;" ::: "memory");
#define GI_STATE(K, PLANE) (!STAGE ? st.ld(PLANE, pid) : s_stage[(K) * RTX_GI_BLOCK + threadIdx.x])
        const float4 zero4 = make_float4(0, 0, 0, 0);
        // SP_N1 / SP_O (the primary hit's normal and outgoing direction) are only read by iteration 0
        const float4 a1 = ITER0 ? st.ld(SP_N1, pid) : zero4, a2 = ITER0 ? st.ld(SP_O, pid) : zero4;
        // iteration 0 starts from the state SamplePathSimple begins with (Path_Sampler_v7.hlsl:9-23): the path vertex is the primary hit
        // (SP_X1 / SP_N1 / SP_O), acc_f = acc_f_reconnection = 1, an empty GI reservoir, acc_pdf = 1 — k_di_finish does not write those
        // ten planes and this kernel does not read them (330 MB less written and 300 MB less read per 1080p pass)
        const float4 one3 = make_float4(1, 1, 1, 0);
        // Between bounces a path carries its origin, acc_f, acc_f_reconnection and the reservoir scalars: FOUR planes.  The BSDF value and
        // pdf of the ray in flight (Sampler_v7.hlsl:436-456: evaluated by the reference when the ray comes back) are evaluated by the
        // iteration that SAMPLED the direction, where the vertex's material, normal, outgoing direction and lobe context are live — same
        // expressions on the same values, so bit-identical — and acc_f / acc_pdf / acc_f_reconnection are stored already multiplied
        // (nothing else touches them in between; a ray that misses ends the path).  Before: five planes, and every bounce re-loaded the
        // previous vertex's material and rebuilt its strategy probabilities and lobe context just for that one evaluation.
        const float4 d0 = ITER0 ? st.ld(SP_X1, pid) : GI_STATE(0, SP_ORIGIN);         // origin; ITER0: bits(material id), else: pdf of the ray in flight
        const float4 pa = ITER0 ? one3 : GI_STATE(1, SP_ACC_F), pf = ITER0 ? one3 : GI_STATE(2, SP_ACC_FR);
        f3 origin = xyz(d0), normal = xyz(a1);
        f3 outgoing = ITER0 ? normalize3(xyz(a2)) : mk3(0, 0, 0);
        f3 acc_f = xyz(pa), acc_fr = xyz(pf);
        // The GI reservoir of the path.  Per bounce only its scalars change (w_sum, acc_pdf, "a sample was accepted"): they live in their
        // own plane SP_GI_SC; xn / nn are written once by iteration 0; E3 and the winner's shadow end points are only ever REPLACED, so
        // later iterations neither load them (the end points: only where the path ends) nor store them unless this bounce replaced them.
        // (Before: ten planes read and ten written per path and bounce; now six and three to six + what changed.)
        const float4 sc = ITER0 ? make_float4(0, 1.0f, 0, 0) : GI_STATE(3, SP_GI_SC);
        f3 xn = mk3(0, 0, 0), nn = mk3(0, 0, 0), E3 = mk3(0, 0, 0);
        float w_sum = sc.x, acc_pdf = sc.y;
        f3 x1s = mk3(0, 0, 0), x2s = mk3(0, 0, 0);
        bool e3_set = false, sh_set = false;
#undef GI_STATE
        float gi_has = sc.z;                                        // 1 once UpdateReservoir_GI has accepted a sample (ReSTIR: reservoir.xn/nn)
        uint2 seed = st.ld_seed(pid);
        MatOpt material;                                            // the vertex the path stands on: ITER0 the primary hit, then the hit being consumed
        if (ITER0) material = load_matopt(S, __float_as_uint(d0.w), nullptr);
        else material = hm;
        const float fnee = (float)S.nee_samples;
        bool cont = false;
        float next_pdf = 0.0f;
        if (inst != 0xFFFFFFFFu) {                                  // miss: path ends (deviation D1)
            sp.hitPosition = origin + ha.x * sample;                // Hit_v7.hlsl:15 (the one line of ClosestHit that needs the path state)
            if (ITER0) {
                if (!(length3(ke_full) > 0.0f)) {                   // Path_Sampler_v7.hlsl:55-98
                    const f3 incoming = normalize3(-sample);
                    float p_d, p_s;
                    CalculateStrategyProbabilities(S, material, outgoing, normal, p_d, p_s);
                    const LobeCtx lobe = make_lobe_ctx(S, material, normal, outgoing, p_d, p_s);
                    f3 F; float P;
                    lobe_FP<true, true>(lobe, incoming, false, 1.0f, 1.0f, F, P);
                    const float NdotL = dot3(normal, sample);
                    acc_pdf *= P;
                    acc_f = acc_f * (F * NdotL);
                    outgoing = incoming; material = hm; normal = sp.hitNormal; origin = sp.hitPosition;
                    xn = origin; nn = normalize3(normal);           // :104-106
                    cont = true;
                }
            } else {                                                // Sampler_v7.hlsl:436-504 (brdf, pdf_bsdf, NdotL and the three products: see below)
                const float pdf_bsdf = d0.w;
                const bool emitter = (hm.Ke.x != 0.0f || hm.Ke.y != 0.0f || hm.Ke.z != 0.0f);
                f3 contribution = mk3(0, 0, 0), emission = mk3(0, 0, 0);
                float pdf_light = 1.0f;
                if (emitter) {
                    f3 L = sp.hitPosition - origin;
                    float dist = length3(L);
                    float dist2 = dist * dist;
                    float cos_theta = dot3(sp.hitNormal, -sample);
                    float kesum = hadd(hadd(hm.Ke.x, hm.Ke.y), hm.Ke.z);
                    pdf_light = (((kesum / 3.0f) / __ldg(&S.lights[0].total_weight)) * dist2) / cos_theta;
                    emission = hm.Ke;
                    contribution = (hm.Ke * acc_f) / acc_pdf;
                }
                if (length3(contribution) > 0.0f) {                 // Path_Sampler_v7.hlsl:235-261
                    float mi = pdf_bsdf / (fnee * pdf_light + pdf_bsdf);
                    f3 E_reconnection = (acc_fr * mi) * emission;
                    f3 E_path = mi * contribution;
                    float wi = length3(E_path);
                    if (isnan1(wi) || isinf1(wi)) wi = 0.0f;
                    w_sum += wi;
                    if (RandomFloat(seed) < wi / w_sum) { E3 = q16v(E_reconnection); gi_has = 1.0f; e3_set = true; }   // UpdateReservoir_GI; (xn, nn) are the path's
                } else if (!emitter) {
                    origin = sp.hitPosition; material = hm; outgoing = -sample; normal = sp.hitNormal;
                    cont = true;
                }                                                   // emitter with zero contribution: ends (deviation D4)
            }
        }
        if (cont && iter < S.bounces) {                             // Path_Sampler_v7.hlsl:114-225 (iteration `iter`)
            const float ps_sel = StrategyPs(S, material, outgoing, normal);     // both selections of this vertex see the same p_s
            uint32_t strategy = SelectWithPs(ps_sel, material, seed);
            LobeCtx lobe;
            const f3 Nn = normalize3(normal);
            const f3 on = normalize3(outgoing);
            float p_d, p_s;
            CalculateStrategyProbabilities(S, material, on, normal, p_d, p_s);
            lobe = make_lobe_ctx_nv(S, material, normal, Nn, normalize3(on), p_d, p_s);
            for (uint32_t k = 0; k < S.nee_samples; k++) {
                float pdf_light = 1.0f, pdf_bsdf = 1.0f;
                f3 throughput_NEE = mk3(1, 1, 1), emission_NEE = mk3(0, 0, 0), x2;
                f3 contribution = SampleLightNEE_GI(S, pdf_light, pdf_bsdf, x2, seed, origin, normal, lobe, acc_f, acc_pdf,
                                                    throughput_NEE, emission_NEE);
                float mi = pdf_light / (fnee * pdf_light + pdf_bsdf);
                f3 E_reconnection = ((acc_fr * mi) * emission_NEE) * throughput_NEE;
                f3 E_path = mi * contribution;
                float wi = length3(E_path);
                if (isnan1(wi) || isinf1(wi)) wi = 0.0f;
                w_sum += wi;
                if (RandomFloat(seed) < wi / w_sum) {
                    E3 = q16v(E_reconnection); gi_has = 1.0f; e3_set = true;
                    x1s = origin + RTX_S_BIAS * Nn;
                    x2s = x2; sh_set = true;
                }
            }
            strategy = SelectWithPs(ps_sel, material, seed);
            const f3 s2 = SampleBRDF(strategy, material, outgoing, normal, seed);
            emit = true; ro = origin; rd = s2;
            // What the reference evaluates when this ray comes back (Sampler_v7.hlsl:436-456, Path_Sampler_v7.hlsl:232): F sees
            // V = normalize3(on) — the NEE context above (with RTX_FLAG_LAMBERT_ONLY a context ignores N and V) — the pdf sees V = on.
            {
                const LobeCtx lobeP = make_lobe_ctx_nv(S, material, normal, Nn, on, p_d, p_s);
                f3 brdf, unusedF; float unusedP;
                lobe_FP<true, false>(lobe, -s2, false, 1.0f, 1.0f, brdf, unusedP);
                lobe_FP<false, true>(lobeP, -s2, false, 1.0f, 1.0f, unusedF, next_pdf);
                const float NdotL = dot3(normal, s2);
                const f3 throughput = brdf * NdotL;
                acc_pdf *= next_pdf;
                acc_f = acc_f * throughput;
                acc_fr = acc_fr * throughput;
            }
        } else {
            // the path's sampling is over: one shadow ray for the reservoir winner (Path_Sampler_v7.hlsl:271-283)
            if (!ITER0) { x1s = xyz(st.ld(SP_SH1, pid)); x2s = xyz(st.ld(SP_SH2, pid)); }     // the winner's end points, set by an earlier bounce
            const f3 ds = x2s - x1s;
            const float len = length3(ds);
            if (S.nee_samples > 0u && len > RTX_EPS) {
                emit_sh = true; so = x1s; sd = normalize3(ds);
                stmax = fmaxf(RTX_S_BIAS, len - (RTX_S_BIAS * 5.0f));
            }
        }
        if (emit) {                                                 // the path vertex: read again only if another bounce follows
            st.stt(SP_ORIGIN, pid, f4(origin, next_pdf));
            st.stt(SP_ACC_F, pid, f4(acc_f, 0.0f));
            st.stt(SP_ACC_FR, pid, f4(acc_fr, 0.0f));
        }
        st.stt(SP_GI_SC, pid, make_float4(w_sum, acc_pdf, gi_has, 0.0f));
        if (ITER0) {                                                // written once (Path_Sampler_v7.hlsl:104-106)
            st.stt(SP_GI_XN, pid, f4(xn, 0.0f));
            st.stt(SP_GI_NN, pid, f4(nn, 0.0f));
        }
        if (ITER0 || e3_set) st.stt(SP_GI_E3, pid, f4(E3, 0.0f));
        if (ITER0 || sh_set) { st.stt(SP_SH1, pid, f4(x1s, 0.0f)); st.stt(SP_SH2, pid, f4(x2s, 0.0f)); }
        st.st_seed(pid, seed);
    }
    push_ray_pair(q_shadow, emit_sh, emit_sh && ray_is_heavy(S, so, 0.5f * RTX_S_BIAS, sd, stmax), so, 0.5f * RTX_S_BIAS, sd, stmax,
                  qout, emit, emit && ray_is_heavy(S, ro, RTX_S_BIAS, rd, 10000.0f), ro, RTX_S_BIAS, rd, 10000.0f, pid);
}

// ---- stage: estimator E0 (Pass_init_di_v7.hlsl:166-181 + Pass_spat_di_v7.hlsl:334-372 with no accepted neighbours)
__global__ void __launch_bounds__(WF_BLOCK)
k_finalize(StateView st, PartMap pm, uint32_t np, SceneData S, const float* __restrict__ vis_di, const float* __restrict__ vis_gi,
           const uint32_t* __restrict__ shadow_counts, unsigned long long* ray_counters) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) atomicAdd(&ray_counters[1], (unsigned long long)shadow_counts[0] + (unsigned long long)shadow_counts[1]);
    if (k >= np) return;
    const uint32_t p = pm.path(k);
    const float4 res = st.ld(SP_RESULT, p);
    if (res.w != 1.0f) return;
    const float4 a0 = st.ld(SP_X1, p);
    const f3 x1 = xyz(a0), hitNormal = xyz(st.ld(SP_N1, p)), o = xyz(st.ld(SP_O, p));
    const MatOpt mat = load_matopt(S, __float_as_uint(a0.w), nullptr);
    const f3 n1 = normalize3(hitNormal);
    const f3 rdi = xyz(st.ld(SP_DI_R, p));
    const float f_g = st.ld(SP_DI_N2, p).w, w_sum = st.ld(SP_DI_X2, p).w;
    const float p_hat = f_g * vis_di[p];
    const float W = (p_hat > RTX_EPS) ? w_sum / p_hat : 0.0f;
    const f3 Cdi = rdi * W;
    const float4 c0 = st.ld(SP_GI_XN, p);
    const float w_sum_gi = st.ld(SP_GI_SC, p).x * vis_gi[p];
    const f3 f_gi = ReconnectGI(S, x1, n1, xyz(c0), xyz(st.ld(SP_GI_E3, p)), o, mat);
    const float p_hat_gi = length3(f_gi);
    const float W_GI = (p_hat_gi > RTX_EPS) ? w_sum_gi / p_hat_gi : 0.0f;
    const f3 C = Cdi + f_gi * W_GI;
    st.at(SP_RESULT, p) = f4(C, 2.0f);                 // (plain: k_accumulate reads it next)
    st.stt(SP_GI_XN, p, make_float4(c0.x, c0.y, c0.z, w_sum_gi));
    st.stt(SP_DI_R, p, f4(rdi, W));
    st.at(SP_GI_E3, p).w = W_GI;
    st.at(SP_SH1, p).w = p_hat;
}

// ---- F20 accumulation, Pass_spat_di_v7.hlsl:383-404: drop non-finite samples, sum += C, n += 1 (samples in index order)
__global__ void __launch_bounds__(WF_BLOCK)
k_accumulate(StateView st, uint32_t npx, uint32_t spp, float4* __restrict__ accum) {
    const uint32_t px = blockIdx.x * blockDim.x + threadIdx.x;
    if (px >= npx) return;
    float4 a = accum[px];
    for (uint32_t s = 0; s < spp; s++) {
        const float4 c = st.at(SP_RESULT, s * npx + px);
        if (!any_nan_inf(mk3(c.x, c.y, c.z))) { a.x += c.x; a.y += c.y; a.z += c.z; a.w += 1.0f; }
    }
    accum[px] = a;
}

// sRGBGammaCorrection, shaders/Common_v7.hlsl:353-376; output write Pass_spat_di_v7.hlsl:405,428-441
__device__ __forceinline__ float srgb1(float c) {
    if (c <= 0.0031308f) return 12.92f * c;
    return 1.055f * d_pow(c, 1.0f / 2.4f) - 0.055f;
}
__global__ void k_resolve(const float4* __restrict__ accum, uint32_t npx, uchar4* __restrict__ out) {
    const uint32_t px = blockIdx.x * blockDim.x + threadIdx.x;
    if (px >= npx) return;
    const float4 a = accum[px];
    float c[3] = {a.x / a.w, a.y / a.w, a.z / a.w};
    const bool nan = isnan1(c[0]) || isnan1(c[1]) || isnan1(c[2]);
    const bool inf = isinf1(c[0]) || isinf1(c[1]) || isinf1(c[2]);
    if (nan) { c[0] = 1; c[1] = 0; c[2] = 1; }
    if (inf) { c[0] = 0; c[1] = 1; c[2] = 1; }
    unsigned char o[3];
    for (int k = 0; k < 3; k++) o[k] = (unsigned char)(int)(saturate1(srgb1(c[k])) * 255.0f + 0.5f);
    out[px] = make_uchar4(o[0], o[1], o[2], 255);
}

__global__ void k_debug_pixel(StateView st, uint32_t p, const float* vis_di, const float* vis_gi, float* out) {
    if (threadIdx.x || blockIdx.x) return;
    for (int i = 0; i < 64; i++) out[i] = 0.0f;
    auto put3 = [&](int i, float4 v) { out[i] = v.x; out[i + 1] = v.y; out[i + 2] = v.z; };
    const float4 a0 = st.at(SP_X1, p), a1 = st.at(SP_N1, p), a2 = st.at(SP_O, p);
    out[9] = a0.w; put3(10, a0); put3(13, a1);
    const float4 b0 = st.at(SP_DI_X2, p), b1 = st.at(SP_DI_N2, p), b2 = st.at(SP_DI_L2, p), b3 = st.at(SP_DI_R, p);
    put3(16, b0); out[19] = b0.w; put3(20, b1); out[23] = b3.w; put3(24, b2);
    const float4 c0 = st.at(SP_GI_XN, p), c1 = st.at(SP_GI_NN, p), c2 = st.at(SP_GI_E3, p);
    put3(27, c0); out[30] = c0.w; put3(31, c1); out[34] = c2.w; put3(35, c2);
    out[38] = st.at(SP_SH1, p).w;
    put3(39, st.at(SP_RESULT, p));
    out[42] = __uint_as_float(st.seed[p].x); out[43] = __uint_as_float(st.seed[p].y);
    out[49] = st.at(SP_RESULT, p).w; out[50] = vis_di[p]; out[51] = vis_gi[p]; out[52] = b1.w;
}

// ---- RTX_FLAG_SORT_MATERIAL: bin the entries of a traced queue by the material of their hit (counting sort, 16 bins) so
// that the warps of the next shading kernel see one material class each.  Measured on C2: every
// material runs the same combined-lobe code, so the coherence gained is small and does not pay for the binning passes.
#define WF_NBIN 16
__device__ __forceinline__ uint32_t hit_material_bin(const SceneData& S, uint32_t inst, uint32_t prim) {
    if (inst == 0xFFFFFFFFu) return WF_NBIN - 1;
    const ModelRef M = S.models[S.inst_model[inst]];
    const uint32_t mslot = 3u * prim + M.mat_offset;
    const uint32_t mID = mslot < S.n_material_ids ? __ldg(&S.material_ids[mslot]) : 0u;
    return mID < WF_NBIN - 2 ? mID : WF_NBIN - 2;
}
__global__ void __launch_bounds__(WF_BLOCK)
k_bin_count(SceneData S, const uint32_t* __restrict__ n_ptr, const float4* __restrict__ hit_a, const uint32_t* __restrict__ hit_inst,
            unsigned char* __restrict__ keys, uint32_t* __restrict__ bins) {
    __shared__ uint32_t h[WF_NBIN];
    if (threadIdx.x < WF_NBIN) h[threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < *n_ptr) {
        const uint32_t k = hit_material_bin(S, hit_inst[j], __float_as_uint(hit_a[j].w));
        keys[j] = (unsigned char)k;
        atomicAdd(&h[k], 1u);
    }
    __syncthreads();
    if (threadIdx.x < WF_NBIN && h[threadIdx.x]) atomicAdd(&bins[threadIdx.x], h[threadIdx.x]);
}
__global__ void k_bin_scan(uint32_t* bins) {       // bins[0..15] counts -> bins[16..31] running cursors (exclusive prefix)
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int k = 0; k < WF_NBIN; k++) { bins[WF_NBIN + k] = run; run += bins[k]; bins[k] = 0u; }
    }
}
__global__ void __launch_bounds__(WF_BLOCK)
k_bin_scatter(const uint32_t* __restrict__ n_ptr, const unsigned char* __restrict__ keys, uint32_t* __restrict__ bins, uint32_t* __restrict__ perm) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = j < *n_ptr;
    const uint32_t k = live ? keys[j] : 0xffu;
    const unsigned peers = __match_any_sync(0xffffffffu, k);          // one atomic per distinct bin per warp
    const unsigned lane = threadIdx.x & 31u;
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (live && (int)lane == leader) base = atomicAdd(&bins[WF_NBIN + k], (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (live) perm[base + __popc(peers & ((1u << lane) - 1u))] = j;
}

// ------------------------------------------------------------------------------------------------ host side
static cudaError_t alloc_queue(RayQueue* q, uint32_t n) {
    CKE(cudaMalloc((void**)&q->o_tmin, (size_t)n * 16));
    CKE(cudaMalloc((void**)&q->d_tmax, (size_t)n * 16));
    CKE(cudaMalloc((void**)&q->pid, (size_t)n * 4));
    CKE(cudaMalloc((void**)&q->order, (size_t)n * 4));
    q->count = nullptr; q->n_heavy = q->n_light = nullptr; q->cap = n;
    return cudaSuccess;
}
static void free_queue(RayQueue* q) {
    if (q->o_tmin) cudaFree(q->o_tmin);
    if (q->d_tmax) cudaFree(q->d_tmax);
    if (q->pid) cudaFree(q->pid);
    if (q->order) cudaFree(q->order);
    q->o_tmin = q->d_tmax = nullptr; q->pid = nullptr; q->order = nullptr;
}

static cudaError_t wave_alloc_impl(WaveBuffers* B, const WaveBuffers* shared, uint32_t width, uint32_t height, uint32_t spp) {
    const uint32_t npx = width * height;
    const uint32_t n = npx * spp;
    B->n_paths = n;
    memset(B->q, 0, sizeof B->q); memset(B->sq, 0, sizeof B->sq);
    CKE(cudaMalloc((void**)&B->state, (size_t)NSTATE * n * 16));
    CKE(cudaMalloc((void**)&B->seeds, (size_t)n * 8));
    for (int i = 0; i < 2; i++) { CKE(alloc_queue(&B->q[i], n)); CKE(alloc_queue(&B->sq[i], n)); }
    CKE(cudaMalloc((void**)&B->hit_a, (size_t)n * 16));
    CKE(cudaMalloc((void**)&B->hit_inst, (size_t)n * 4));
    CKE(cudaMalloc((void**)&B->vis_di, (size_t)n * 4));
    CKE(cudaMalloc((void**)&B->vis_gi, (size_t)n * 4));
    CKE(cudaMalloc((void**)&B->counts, WAVE_MAX_PARTS * 384 * 4));     // per part: 128 queue counters + their heavy / light counts
    CKE(cudaMalloc((void**)&B->cursor, WAVE_MAX_PARTS * 16));
    CKE(cudaMemset(B->cursor, 0, WAVE_MAX_PARTS * 16));      // launch_trace keeps the cursor words at zero between launches
    CKE(cudaMalloc((void**)&B->ray_counters, 8 * 8));
    CKE(cudaMemset(B->ray_counters, 0, 64));
    if (shared) { B->accum = shared->accum; B->output = shared->output; }
    else {
        CKE(cudaMalloc((void**)&B->accum, (size_t)npx * 16));
        CKE(cudaMemset(B->accum, 0, (size_t)npx * 16));
        CKE(cudaMalloc((void**)&B->output, (size_t)npx * 4));
    }
    CKE(cudaMalloc((void**)&B->cam, sizeof(rtx_camera_params)));     // (a lane has its own camera block: pipelined frames)
    CKE(cudaMalloc((void**)&B->first_sample, 4));
    CKE(cudaMalloc((void**)&B->debug, 64 * 4));
    CKE(cudaMalloc((void**)&B->perm, (size_t)n * 4));
    CKE(cudaMalloc((void**)&B->bin_keys, (size_t)n));
    CKE(cudaMalloc((void**)&B->bins, 2 * WF_NBIN * 4));
    return cudaSuccess;
}

cudaError_t wave_alloc(WaveBuffers* B, uint32_t width, uint32_t height, uint32_t spp) { return wave_alloc_impl(B, nullptr, width, height, spp); }
cudaError_t wave_alloc_lane(WaveBuffers* L, const WaveBuffers& B, uint32_t width, uint32_t height, uint32_t spp) {
    return wave_alloc_impl(L, &B, width, height, spp);
}
void wave_free_lane(WaveBuffers* L) {
    L->accum = nullptr; L->output = nullptr;       // owned by the context's first set
    wave_free(L);
}

void wave_free(WaveBuffers* B) {
    for (int i = 0; i < WAVE_MAX_PARTS - 1; i++) {
        if (B->aux[i]) cudaStreamDestroy(B->aux[i]);
        if (B->ev_join[i]) cudaEventDestroy(B->ev_join[i]);
    }
    if (B->ev_fork) cudaEventDestroy(B->ev_fork);
    for (int i = 0; i < WAVE_MAX_PARTS; i++) {
        if (B->aux_sh[i]) cudaStreamDestroy(B->aux_sh[i]);
        if (B->ev_sh_fork[i]) cudaEventDestroy(B->ev_sh_fork[i]);
        if (B->ev_sh_join[i]) cudaEventDestroy(B->ev_sh_join[i]);
    }
    if (B->state) cudaFree(B->state);
    for (int i = 0; i < 2; i++) { free_queue(&B->q[i]); free_queue(&B->sq[i]); }
    for (auto& g : B->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    void* ptrs[] = {B->seeds, B->first_sample, B->hit_a, B->hit_inst, B->vis_di, B->vis_gi, B->counts, B->cursor, B->ray_counters, B->accum, B->output, B->cam, B->debug, B->perm, B->bin_keys, B->bins};
    for (void* p : ptrs) if (p) cudaFree(p);
    *B = WaveBuffers();
}

// One DispatchRays-equivalent.  The frame's paths are cut into `parts` contiguous ranges that run the whole stage sequence independently,
// part 0 on the caller's stream and the others on auxiliary streams (forked from and joined back into the caller's stream with events):
// every persistent traversal launch ends in a tail in which a few warps finish the longest rays on an otherwise empty GPU (0.14 ms of a
// 0.43 ms launch at 1080p, profiles/r01_s4_*), and the other parts' kernels fill it.  Paths never interact before the accumulation, so the
// result does not depend on `parts`.
static int pass_parts(const WaveBuffers& B, const SceneData& S, const PassTiming* T, uint32_t n, bool accumulate, int parts_override) {
    int parts = parts_override > 0 ? parts_override : B.parts;
    parts = parts < 1 ? 1 : (parts > WAVE_MAX_PARTS ? WAVE_MAX_PARTS : parts);
    // per-launch events, the counting traversal variant and the material-binned queues run as one part (the ReSTIR frame's first
    // pass runs in parts too: what it hands to the reuse passes is per-path state, complete once the parts have joined)
    (void)accumulate;
    if (T->stage_timing || T->stats || (S.cfg_flags & RTX_FLAG_SORT_MATERIAL) || n < 65536u) parts = 1;
    return parts;
}

__global__ void k_set_word(uint32_t* p, uint32_t v) { *p = v; }

// the launches of one pass on `stream` (+ the auxiliary streams); the sample index comes from B.first_sample, so the same sequence serves
// every pass of a static configuration: wave_render_pass captures it into a CUDA graph
static cudaError_t wave_pass_body(WaveBuffers& B, const SceneData& S, const SceneAS& AS, uint32_t spp, cudaStream_t stream, uint64_t* launches,
                                  PassTiming* T, bool accumulate, int parts_override, bool defer_accumulate) {
    const uint32_t npx = S.width * S.height;
    const uint32_t n = npx * spp;
    StateView st{B.state, n, B.seeds};
    uint64_t L = 0;
    T->n_marks = 0;
    const int parts = pass_parts(B, S, T, n, accumulate, parts_override);
    // traversal launches that run beside each other share the machine's resident CTAs: the parts of this pass, and (defer_accumulate:
    // a pipelined pass, api.cu) the pass of the other lane
    const int share = parts * (defer_accumulate ? 2 : 1);

    struct Part {
        cudaStream_t stream; uint32_t p0, np; PartMap pm; unsigned grid, ggrid;
        uint32_t* counts; unsigned int* cursor; float4* hit_a; uint32_t* hit_inst;
        RayQueue q[2], sdi, sgi, qin, qout;
        int cur;
    } P[WAVE_MAX_PARTS];
    // `ordered` views carry the queue's trace order (wavefront.h RayQueue); the heavy / light counts of queue counter i are counters
    // 128 + i / 256 + i of the part's block
    // Interleaved parts: chunks of a few image rows dealt round-robin, so that the concurrent parts see the same mix of the image and
    // finish together (a part that finishes early leaves the other with its 1/parts share of the traversal grid): C2 7.34 -> 7.24 ms,
    // and three or four parts stop losing (profiles/r02_part_interleave_ab.txt).  B.part_rows: rows per chunk, < 0 = the largest count
    // <= 32 that divides the frame evenly among the parts, 0 (or no such count) = contiguous ranges.
    uint32_t chunk = 0u;
    if (parts > 1 && B.part_rows != 0) {
        uint32_t rows = 0u;
        if (B.part_rows > 0) rows = (S.height % ((uint32_t)B.part_rows * (uint32_t)parts)) == 0u ? (uint32_t)B.part_rows : 0u;
        else for (uint32_t r = 32u; r >= 1u && !rows; r--) if ((S.height % (r * (uint32_t)parts)) == 0u) rows = r;
        chunk = S.width * rows;
    }
    const bool lpt = S.heavy_valid != 0u;
    auto view = [lpt](const RayQueue& q, uint32_t off, uint32_t cap, bool ordered) {
        RayQueue v = q; v.o_tmin += off; v.d_tmax += off; v.pid += off; v.cap = cap;
        v.order = (ordered && lpt) ? q.order + off : nullptr; v.n_heavy = v.n_light = nullptr;
        return v;
    };
    auto counters = [](RayQueue& q, uint32_t* block, int i) { q.count = block + i; q.n_heavy = block + 128 + i; q.n_light = block + 256 + i; };
    for (int h = 0; h < parts; h++) {
        Part& p = P[h];
        p.stream = h == 0 ? stream : B.aux[h - 1];
        p.p0 = (uint32_t)((uint64_t)n * h / parts); p.np = (uint32_t)((uint64_t)n * (h + 1) / parts) - p.p0;
        p.pm = PartMap{p.p0, chunk, (uint32_t)parts, (uint32_t)h};
        p.grid = (p.np + WF_BLOCK - 1) / WF_BLOCK; p.ggrid = (p.np + RTX_GI_BLOCK - 1) / RTX_GI_BLOCK;
        p.counts = B.counts + 384 * h; p.cursor = B.cursor + 4 * h;
        p.hit_a = B.hit_a + p.p0; p.hit_inst = B.hit_inst + p.p0;
        // counter slots: 0 = primary queue, 1 = DI BSDF queue, 2 = DI shadow, 3 = GI shadow, 4.. = indirect queues
        p.q[0] = view(B.q[0], p.p0, p.np, true); p.q[1] = view(B.q[1], p.p0, p.np, true);
        p.sdi = view(B.sq[0], p.p0, p.np, true); p.sgi = view(B.sq[1], p.p0, p.np, true);
        counters(p.sdi, p.counts, 2); counters(p.sgi, p.counts, 3);
        p.cur = 0;
    }
    auto mark = [&](StageKind k) -> cudaError_t {       // per-launch CUDA events (RTX_OPT_STAGE_TIMING only: one part)
        if (T->stage_timing && T->n_marks + 2 < WAVE_MAX_EVENTS) {
            T->kind[T->n_marks] = (unsigned char)k;
            CKE(cudaEventRecord(T->ev[2 + T->n_marks], stream));
            T->n_marks++;
        }
        L++;
        return cudaSuccess;
    };
    auto closest = [&](Part& p, const RayQueue& q) -> cudaError_t {
        CKE(mark(SK_CLOSEST));
        return launch_trace(AS, q.o_tmin, q.d_tmax, q.count, 0, p.cursor, p.hit_a, p.hit_inst, false, T->stats, p.stream, share, q.order, q.n_heavy, q.cap);
    };
    const bool side_shadow = B.shadow_overlap && !T->stage_timing && !T->stats;
    auto shadow = [&](Part& p, const RayQueue& q, float* vis, bool side) -> cudaError_t {
        CKE(mark(SK_ANY));
        // the occlusion result goes straight to the per-path visibility array (the separate scatter kernel of round 1 is gone);
        // a launch on the side stream has its own cursor words (it runs beside the part's closest-hit launches)
        const int h = (int)(&p - P);
        return launch_trace(AS, q.o_tmin, q.d_tmax, q.count, 0, side ? p.cursor + 2 : p.cursor, p.hit_a, p.hit_inst, true, nullptr,
                            side ? B.aux_sh[h] : p.stream, share, q.order, q.n_heavy, q.cap, q.pid, vis);
    };
    // stage s of one part; the parts are issued round-robin stage by stage so that every stream always has work queued
    const int n_stages = 7 + 2 * ((int)S.bounces + 1) + 2;
    auto stage = [&](Part& p, int s) -> cudaError_t {
        RayQueue q0 = p.q[0], q1 = p.q[1], qa = p.q[0];
        counters(q0, p.counts, 0); counters(q1, p.counts, 1); counters(qa, p.counts, 4);
        const int last = 7 + 2 * ((int)S.bounces + 1);
        if (s == 0) {
            CKE(cudaMemsetAsync(p.counts, 0, 384 * 4, p.stream));
            CKE(mark(SK_GENERATE));
            k_generate<<<p.grid, WF_BLOCK, 0, p.stream>>>(st, S, p.pm, p.np, q0, B.cam, S.width, S.height, B.first_sample, S.cfg_flags, B.vis_di, B.vis_gi,
                                                          B.ray_counters);
        } else if (s == 1) {
            CKE(closest(p, q0));
        } else if (s == 2) {
            CKE(mark(SK_SHADE_PRIMARY));
            k_shade_primary<<<p.grid, WF_BLOCK, 0, p.stream>>>(st, S, q0, p.hit_a, p.hit_inst, q1, B.ray_counters);
        } else if (s == 3) {
            CKE(closest(p, q1));
        } else if (s == 4) {
            CKE(mark(SK_DI_FINISH));
            k_di_finish<<<p.grid, WF_BLOCK, 0, p.stream>>>(st, S, q1, p.hit_a, p.hit_inst, p.sdi, qa, B.ray_counters);
        } else if (s == 5) {
            // DI visibility (connect): writes vis_di only, read by k_finalize
            const int h = (int)(&p - P);
            if (side_shadow) {
                CKE(cudaEventRecord(B.ev_sh_fork[h], p.stream));
                CKE(cudaStreamWaitEvent(B.aux_sh[h], B.ev_sh_fork[h], 0));
            }
            CKE(shadow(p, p.sdi, B.vis_di, side_shadow));
            if (side_shadow) CKE(cudaEventRecord(B.ev_sh_join[h], B.aux_sh[h]));
        } else if (s == 6) {
            CKE(closest(p, qa));
            p.qin = qa; p.cur = 0;
        } else if (s < last) {
            const uint32_t iter = (uint32_t)(s - 7) / 2u;
            if (((s - 7) & 1) == 0) {               // k_gi_step(iter)
                p.qout = p.q[p.cur ^ 1]; counters(p.qout, p.counts, 5 + (int)iter);
                const uint32_t* perm = nullptr;
                if (S.cfg_flags & RTX_FLAG_SORT_MATERIAL) {
                    CKE(mark(SK_SORT));
                    CKE(cudaMemsetAsync(B.bins, 0, 2 * WF_NBIN * 4, p.stream));
                    k_bin_count<<<p.grid, WF_BLOCK, 0, p.stream>>>(S, p.qin.count, p.hit_a, p.hit_inst, B.bin_keys, B.bins);
                    k_bin_scan<<<1, 32, 0, p.stream>>>(B.bins);
                    k_bin_scatter<<<p.grid, WF_BLOCK, 0, p.stream>>>(p.qin.count, B.bin_keys, B.bins, B.perm);
                    L += 2;
                    perm = B.perm;
                }
                CKE(mark(SK_GI_STEP));
                if (iter == 0u) k_gi_step<true><<<p.ggrid, RTX_GI_BLOCK, 0, p.stream>>>(st, S, p.qin, p.hit_a, p.hit_inst, p.sgi, p.qout, iter, B.ray_counters, perm);
                else k_gi_step<false><<<p.ggrid, RTX_GI_BLOCK, RTX_GI_STAGE ? GI_STAGED * RTX_GI_BLOCK * sizeof(float4) : 0, p.stream>>>(st, S, p.qin, p.hit_a, p.hit_inst, p.sgi, p.qout, iter, B.ray_counters, perm);
            } else if (iter < S.bounces) {
                CKE(closest(p, p.qout));
                p.qin = p.qout; p.cur ^= 1;
            }
        } else if (s == last) {
            CKE(shadow(p, p.sgi, B.vis_gi, false));
        } else {
            if (side_shadow) CKE(cudaStreamWaitEvent(p.stream, B.ev_sh_join[(int)(&p - P)], 0));
            CKE(mark(SK_FINALIZE));
            k_finalize<<<p.grid, WF_BLOCK, 0, p.stream>>>(st, p.pm, p.np, S, B.vis_di, B.vis_gi, p.counts + 2, B.ray_counters);
        }
        return cudaSuccess;
    };
    if (parts > 1) {
        CKE(cudaEventRecord(B.ev_fork, stream));
        for (int h = 1; h < parts; h++) CKE(cudaStreamWaitEvent(P[h].stream, B.ev_fork, 0));
    }
    for (int s = 0; s < n_stages; s++)
        for (int h = 0; h < parts; h++) CKE(stage(P[h], s));
    for (int h = 1; h < parts; h++) {
        CKE(cudaEventRecord(B.ev_join[h - 1], P[h].stream));
        CKE(cudaStreamWaitEvent(stream, B.ev_join[h - 1], 0));
    }
    if (accumulate && !defer_accumulate) {       // E0: the pass's samples go straight to gPermanentData; the ReSTIR frame accumulates after RayGen3
        if (B.wait_before_accumulate) CKE(cudaStreamWaitEvent(stream, B.wait_before_accumulate, 0));
        CKE(mark(SK_ACCUMULATE));
        k_accumulate<<<(npx + WF_BLOCK - 1) / WF_BLOCK, WF_BLOCK, 0, stream>>>(st, npx, spp, B.accum);
    }
    CKE(cudaGetLastError());
    if (launches) *launches += L;
    return cudaSuccess;
}

// What a captured pass depends on besides the sample index: every kernel argument is in here (device pointers, counts, flags, bounds).
struct GraphKey { SceneData S; SceneAS AS; uint32_t spp; int parts; int variant; int side_shadow; cudaStream_t stream; };
static_assert(sizeof(GraphKey) <= sizeof(WaveBuffers().last_key), "WaveBuffers::last_key / GraphSlot::key too small");

cudaError_t wave_render_pass(WaveBuffers& B, const SceneData& S, const SceneAS& AS, uint32_t first_sample, uint32_t spp, cudaStream_t stream,
                             uint64_t* launches, PassTiming* T, bool accumulate, int parts_override, bool defer_accumulate) {
    const uint32_t n = S.width * S.height * spp;
    if (n > B.n_paths) return cudaErrorInvalidValue;
    const int parts = pass_parts(B, S, T, n, accumulate, parts_override);
    for (int h = 1; h < parts; h++) {       // (never inside a capture)
        if (!B.aux[h - 1]) CKE(cudaStreamCreateWithFlags(&B.aux[h - 1], cudaStreamNonBlocking));
        if (!B.ev_join[h - 1]) CKE(cudaEventCreateWithFlags(&B.ev_join[h - 1], cudaEventDisableTiming));
    }
    if (parts > 1 && !B.ev_fork) CKE(cudaEventCreateWithFlags(&B.ev_fork, cudaEventDisableTiming));
    for (int h = 0; h < parts && B.shadow_overlap; h++) {
        if (!B.aux_sh[h]) CKE(cudaStreamCreateWithFlags(&B.aux_sh[h], cudaStreamNonBlocking));
        if (!B.ev_sh_fork[h]) CKE(cudaEventCreateWithFlags(&B.ev_sh_fork[h], cudaEventDisableTiming));
        if (!B.ev_sh_join[h]) CKE(cudaEventCreateWithFlags(&B.ev_sh_join[h], cudaEventDisableTiming));
    }
    // (the staged planes of k_gi_step take 24 KB of dynamic shared memory per CTA: below the 48 KB a kernel gets without opting in)
    static_assert(GI_STAGED * RTX_GI_BLOCK * sizeof(float4) <= 48 * 1024, "k_gi_step's staging area needs cudaFuncAttributeMaxDynamicSharedMemorySize");
    k_set_word<<<1, 1, 0, stream>>>(B.first_sample, first_sample);
    if (launches) *launches += 1;

    // ---- CUDA graph of the pass.  A pass is ~50 dependent launches per path range plus memsets and the fork / join of the auxiliary
    // streams; for a static configuration (same buffers, counts, flags, bounds as the pass before) the sequence is captured once and
    // replayed: one graph launch per pass instead of ~100 stream operations.  The first pass of a new configuration runs directly, the
    // second one captures (a scene whose instance LIST changes every frame never pays for captures; transforms that move do not
    // change the configuration: the buffers are rewritten in place).  Passes with per-launch events, the
    // counting variant, the ReSTIR frame or a pending multi-GPU reduce (an event from outside the capture) always run directly.
    const bool graph_ok = B.use_graph && accumulate && !T->stage_timing && !T->stats && (defer_accumulate || !B.wait_before_accumulate);
    GraphKey key;
    memset(&key, 0, sizeof key);
    memcpy(&key.S, &S, sizeof S); memcpy(&key.AS, &AS, sizeof AS);
    key.spp = spp; key.parts = parts; key.stream = stream; key.side_shadow = (B.shadow_overlap ? 1 : 0) | (B.part_rows << 8);
#ifdef RTX_FAST_MATH
    key.variant = 1;
#endif
    if (defer_accumulate) key.variant |= 2;
    CKE(cudaEventRecord(T->ev[0], stream));
    bool done = false;
    if (graph_ok) {
        // a small cache of captured passes: a context alternates between a few configurations (the first pass of a sequence and the
        // pipelined ones, the two sets of per-frame state), none of them should evict the others
        WaveBuffers::GraphSlot* slot = nullptr;
        for (auto& g : B.graphs) if (g.exec && memcmp(&key, g.key, sizeof key) == 0) slot = &g;
        // A pipelined pass (defer_accumulate) is captured the first time it is seen: the configurations of the two lanes are stable by
        // construction, and a capture in the middle of a sequence stalls it.  Any other pass is captured when it repeats the one before.
        if (!slot && (defer_accumulate || (B.have_last_key && memcmp(&key, B.last_key, sizeof key) == 0))) {
            WaveBuffers::GraphSlot* victim = &B.graphs[0];
            for (auto& g : B.graphs) if (!g.exec) { victim = &g; break; } else if (g.last_used < victim->last_used) victim = &g;
            if (victim->exec) { cudaGraphExecDestroy(victim->exec); victim->exec = nullptr; }
            cudaGraph_t g = nullptr;
            uint64_t L = 0;
            cudaError_t e = cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal);
            if (e == cudaSuccess) {
                e = wave_pass_body(B, S, AS, spp, stream, &L, T, accumulate, parts_override, defer_accumulate);
                const cudaError_t e2 = cudaStreamEndCapture(stream, &g);
                if (e == cudaSuccess) e = e2;
            }
            if (e == cudaSuccess) e = cudaGraphInstantiate(&victim->exec, g, 0);
            if (g) cudaGraphDestroy(g);
            if (e != cudaSuccess) {            // no graphs for this context then: the direct path below is always valid
                cudaGetLastError();
                victim->exec = nullptr; B.use_graph = false;
            } else {
                memcpy(victim->key, &key, sizeof key); victim->launches = L;
                slot = victim;
            }
        }
        if (slot) {
            CKE(cudaGraphLaunch(slot->exec, stream));
            slot->last_used = ++B.graph_clock;
            if (launches) *launches += slot->launches;
            T->n_marks = 0;
            done = true;
        }
    }
    memcpy(B.last_key, &key, sizeof key); B.have_last_key = true;
    if (!done) CKE(wave_pass_body(B, S, AS, spp, stream, launches, T, accumulate, parts_override, defer_accumulate));
    CKE(cudaEventRecord(T->ev[1], stream));
    return cudaGetLastError();
}

cudaError_t wave_accumulate(WaveBuffers& B, uint32_t npx, uint32_t spp, cudaStream_t stream) {
    StateView st{B.state, npx * spp, B.seeds};
    k_accumulate<<<(npx + WF_BLOCK - 1) / WF_BLOCK, WF_BLOCK, 0, stream>>>(st, npx, spp, B.accum);
    return cudaGetLastError();
}

cudaError_t wave_resolve(WaveBuffers& B, uint32_t n_pixels, cudaStream_t stream, uint64_t* launches) {
    k_resolve<<<(n_pixels + 255) / 256, 256, 0, stream>>>(B.resolve_source ? B.resolve_source : B.accum, n_pixels, (uchar4*)B.output);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

// ---- rtx_selftest_dmath: every binary32 bit pattern through the fast path and through the IEEE operations it restates
__global__ void k_selftest_dmath(unsigned long long* bad) {
    unsigned long long n_rsqrt = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < (1ull << 32); i += stride) {
        const float x = __uint_as_float((uint32_t)i);
        if (__float_as_uint(d_rsqrt(x)) != __float_as_uint(1.0f / sqrtf(x))) n_rsqrt++;
    }
    if (n_rsqrt) atomicAdd(&bad[0], n_rsqrt);
}
cudaError_t wave_selftest_dmath(cudaStream_t stream, unsigned long long* host_out, uint32_t n_out) {
    unsigned long long* d = nullptr;
    CKE(cudaMalloc((void**)&d, 8 * sizeof(unsigned long long)));
    CKE(cudaMemsetAsync(d, 0, 8 * sizeof(unsigned long long), stream));
    k_selftest_dmath<<<148 * 8, 256, 0, stream>>>(d);
    unsigned long long h[8];
    cudaError_t e = cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(d);
    if (e != cudaSuccess) return e;
    for (uint32_t i = 0; i < n_out && i < 8; i++) host_out[i] = h[i];
    return cudaSuccess;
}

cudaError_t wave_debug_pixel(WaveBuffers& B, const SceneData& S, uint32_t x, uint32_t y, cudaStream_t stream, float* host_out64) {
    StateView st{B.state, B.n_paths, B.seeds};
    k_debug_pixel<<<1, 32, 0, stream>>>(st, y * S.width + x, B.vis_di, B.vis_gi, B.debug);
    CKE(cudaMemcpyAsync(host_out64, B.debug, 64 * 4, cudaMemcpyDeviceToHost, stream));
    return cudaStreamSynchronize(stream);
}

#ifdef RTX_FAST_MATH
}  // namespace fast
#endif
}  // namespace rtx
