// legacy.cu — the reference's older single-pass estimator as a wavefront (RTX_FLAG_LEGACY_RR).
//
// Reference: include/RayGen.hlsl:60-137 (path loop, Russian roulette after depth 3), include/Hit.hlsl:58-357 (all shading inside
// ClosestHit: MIS-weighted emitter hits, RIS over RIS_M = 10 light candidates with one shadow ray per bounce, BSDF sample),
// include/Miss.hlsl, include/BRDF.hlsl, include/GGX.hlsl, include/Lambertian.hlsl, include/Common.hlsl (PI 3.1415, s_bias 1e-5,
// EPSILON 1e-4).  Full fp32 materials, face-forwarded normals, primary direction not normalised (RayGen.hlsl:89).
// The megakernel loop is cut at its two TraceRay calls: per bounce  closest trace -> k_legacy_hit -> any-hit trace -> k_legacy_shadow.
// Every path carries its TEA state and draws in the reference's order; additions to payload.emission happen in bounce order
// (the NEE term of bounce y is added by k_legacy_shadow before bounce y+1 is shaded), so the result is bit-identical to a
// sequential CPU evaluation in the same fp32 operation order (tests/test_gpu_parity.py).  Deviations: D2 (seed from the sample index), D13 (path length capped at cfg.bounces).
#include "trace.h"
#include "wave_dev.cuh"

namespace rtx {

#define CKE(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return e__; } while (0)
#define LG_BLOCK 256
#define LG_EPS 0.0001f
#define LG_BIAS 0.00001f
#define LG_RIS_M 10

// path state planes (same storage as the E0 stages, wavefront.h)
enum { LP_SEED = SP_X1,          // bits(seed.x), bits(seed.y), -, -
       LP_COLOR = SP_ACC_F,      // payload.colorAndDistance.xyz, payload.pdf
       LP_RISF = SP_SH1,         // ris_f of the selected candidate, WX
       LP_COLW = SP_SH2,         // payload.colorAndDistance.xyz * weight_light at the time of the NEE, -
       LP_EMISSION = SP_RESULT };// payload.emission, -

struct MatFull { f3 Kd, Ks, Ke; float Pr, Pm, Pc; uint32_t id; };

__device__ __forceinline__ MatFull load_mat_full(const SceneData& S, uint32_t id) {
    float4 kd, ks_ni, ke_pad, pr;
    fetch_material_head(S, id, kd, ks_ni, ke_pad, pr);
    MatFull m;
    m.Kd = mk3(kd.x, kd.y, kd.z); m.Ks = mk3(ks_ni.x, ks_ni.y, ks_ni.z); m.Ke = mk3(ke_pad.x, ke_pad.y, ke_pad.z);
    m.Pr = pr.x; m.Pm = pr.y; m.Pc = pr.w; m.id = id;
    return m;
}
__device__ __forceinline__ f3 abs3(f3 v) { return mk3(fabsf(v.x), fabsf(v.y), fabsf(v.z)); }

// include/GGX.hlsl:4-27
__device__ __forceinline__ float lg_ESS_LUT(const SceneData& S, const MatFull& mat, float NdotV) {
    NdotV = saturate1(NdotV);
    float thetaIdxF = NdotV * 15.0f;
    int i0 = (int)floorf(thetaIdxF);
    int i1 = min(i0 + 1, 15);
    float w = thetaIdxF - (float)i0;
    float v0 = 0.0f, v1 = 0.0f;
    if (mat.id < S.n_materials) { v0 = __ldg(&S.materials[mat.id].LUT[i0]); v1 = __ldg(&S.materials[mat.id].LUT[i1]); }
    return lerp1(v0, v1, w);
}
// :37-46
__device__ __forceinline__ float lg_D_GGX(float NdotH, float roughness) {
    float alpha = roughness * roughness;
    float alpha2 = alpha * alpha;
    float NdotH2 = NdotH * NdotH;
    float denom = (NdotH2 * (alpha2 - 1.0f) + 1.0f);
    denom = fmaxf(denom, 1e-7f);
    return alpha2 / ((RTX_PI_REF * denom) * denom);
}
// :87-142
__device__ __forceinline__ f3 lg_SampleBRDF_GGX(const MatFull& mat, f3 outgoing, f3 normal, uint2& seed) {
    float alpha = mat.Pr * mat.Pr;
    f3 N = normalize3(normal), V = normalize3(outgoing);
    float e0 = RandomFloat(seed), e1 = RandomFloat(seed);
    f3 T1, T2;
    CoordinateSystem(N, T1, T2);
    f3 Vh = normalize3(mk3(dot3(T1, V), dot3(T2, V), dot3(N, V)));
    if (Vh.z < 0.0f) Vh = -Vh;
    f3 Vs = normalize3(mk3(alpha * Vh.x, alpha * Vh.y, Vh.z));
    float lensq = Vs.x * Vs.x + Vs.y * Vs.y;
    f3 T1h, T2h;
    if (lensq > 0.0f) { T1h = mk3(-Vs.y, Vs.x, 0.0f) / sqrtf(lensq); T2h = cross3(Vs, T1h); }
    else { T1h = mk3(1, 0, 0); T2h = mk3(0, 1, 0); }
    float r = sqrtf(e0);
    float phi = (2.0f * RTX_PI_REF) * e1;
    float sn, cs; d_sincos(phi, &sn, &cs);
    float x = r * cs, y = r * sn;
    f3 Nhs = (x * T1h + y * T2h) + sqrtf(fmaxf(0.0f, (1.0f - x * x) - y * y)) * Vs;
    f3 Nh = normalize3(mk3(alpha * Nhs.x, alpha * Nhs.y, Nhs.z));
    f3 H = (Nh.x * T1 + Nh.y * T2) + Nh.z * N;
    f3 sample = reflect3(-V, H);
    if (dot3(sample, N) <= 0.0f) sample = mk3(0, 0, 0);
    return sample;
}
// :145-176
__device__ __forceinline__ f3 lg_EvaluateBRDF_GGX(const SceneData& S, const MatFull& mat, f3 normal, f3 incoming, f3 outgoing) {
    f3 N = normalize3(normal), V = normalize3(outgoing), L = normalize3(-incoming);
    f3 H = normalize3(V + L);
    float NdotV = saturate1(dot3(N, V)), NdotL = saturate1(dot3(N, L)), NdotH = saturate1(dot3(N, H)), VdotH = saturate1(dot3(V, H));
    f3 F = SchlickFresnel(mat.Ks, VdotH);
    float D = lg_D_GGX(NdotH, mat.Pr);
    float G = G2_SmithGGX(NdotV, NdotL, mat.Pr * mat.Pr);
    float denominator = (4.0f * NdotV) * NdotL;
    denominator = fmaxf(denominator, 1e-7f);
    f3 specular = ((F * D) * G) / denominator;
    float Ess = lg_ESS_LUT(S, mat, NdotV);
    float kms = (1.0f - Ess) / Ess;
    return specular * mk3(1.0f + mat.Ks.x * kms, 1.0f + mat.Ks.y * kms, 1.0f + mat.Ks.z * kms);
}
// :179-197
__device__ __forceinline__ float lg_BRDF_PDF_GGX(const MatFull& mat, f3 normal, f3 incoming, f3 outgoing) {
    f3 N = normalize3(normal), V = normalize3(outgoing), L = normalize3(-incoming);
    f3 H = normalize3(V + L);
    float NdotH = saturate1(dot3(N, H)), NdotV = saturate1(dot3(N, V));
    float alpha = mat.Pr * mat.Pr;
    float G1 = G1_SmithGGX(NdotV, alpha);
    float D = lg_D_GGX(NdotH, mat.Pr);
    return (G1 * D) / (NdotV * 4.0f);
}
// include/BRDF.hlsl:10-53
__device__ __forceinline__ uint32_t lg_SelectSamplingStrategy(const SceneData& S, const MatFull& mat, f3 outgoing, f3 normal, uint2& seed) {
    float r = RandomFloat(seed);
    float cosTheta = dot3(normal, outgoing);
    f3 fresnel = SchlickFresnel(mat.Ks, cosTheta);
    float p_s = fminf(1.0f, (((fresnel.x + fresnel.y) + fresnel.z) / 3.0f + mat.Pc) + mat.Pm);
    if (S.cfg_flags & RTX_FLAG_LAMBERT_ONLY) return 0u;
    return r <= p_s ? 1u : 0u;
}
// BRDF.hlsl:72-106, include/Lambertian.hlsl:54-66
__device__ __forceinline__ f3 lg_EvaluateBRDF(const SceneData& S, uint32_t strategy, const MatFull& mat, f3 normal, f3 incidence, f3 outgoing) {
    return strategy == 0u ? mat.Kd / RTX_PI_REF : lg_EvaluateBRDF_GGX(S, mat, normal, incidence, outgoing);
}
__device__ __forceinline__ float lg_BRDF_PDF(uint32_t strategy, const MatFull& mat, f3 normal, f3 incidence, f3 outgoing) {
    return strategy == 0u ? fmaxf(dot3(normal, -incidence), 0.0001f) / RTX_PI_REF : lg_BRDF_PDF_GGX(mat, normal, incidence, outgoing);
}

// RayGen.hlsl:81-98
__global__ void __launch_bounds__(LG_BLOCK)
k_legacy_generate(StateView st, RayQueue q0, const rtx_camera_params* __restrict__ cam, uint32_t W, uint32_t H, uint32_t first_sample,
                  unsigned long long* ray_counters) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p == 0) { *q0.count = st.n; atomicAdd(&ray_counters[2], (unsigned long long)st.n); }
    if (p >= st.n) return;
    const uint32_t npx = W * H;
    const uint32_t pixel = p % npx, s = p / npx;
    const uint32_t x = pixel % W, y = pixel / W;
    uint2 seed = init_seed(x, y, 2u, first_sample + s);                 // uint(samples + 1) = 2; D2
    const float jx = RandomFloat(seed), jy = RandomFloat(seed);
    const float dimx = (float)W, dimy = (float)H;
    const f3 o = mul43(cam->viewI, 0.0f, 0.0f, 0.0f, 1.0f);
    const float dx = (((float)x + jx) / dimx) * 2.0f - 1.0f;
    const float dy = (((float)y + jy) / dimy) * 2.0f - 1.0f;
    const f3 target = mul43(cam->projectionI, dx, -dy, 1.0f, 1.0f);
    const f3 d = mul43(cam->viewI, target.x, target.y, target.z, 0.0f);  // not normalised (:89)
    q0.o_tmin[p] = f4(o, 0.0001f);
    q0.d_tmax[p] = f4(d, 10000.0f);
    q0.pid[p] = p;
    st.at(LP_SEED, p) = make_float4(__uint_as_float(seed.x), __uint_as_float(seed.y), 0.0f, 0.0f);
    st.at(LP_COLOR, p) = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
    st.at(LP_EMISSION, p) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

// Hit.hlsl:58-357 + Miss.hlsl + the loop tail of RayGen.hlsl:110-131 for bounce `yb`
__global__ void __launch_bounds__(LG_BLOCK)
k_legacy_hit(StateView st, SceneData S, RayQueue qin, const float4* __restrict__ hit_a, const uint32_t* __restrict__ hit_inst,
             RayQueue q_shadow, RayQueue qout, uint32_t yb, unsigned long long* ray_counters) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = *qin.count;
    if (i == 0) atomicAdd(&ray_counters[0], (unsigned long long)n);
    bool emit = false, emit_sh = false; uint32_t pid = 0;
    f3 no = mk3(0, 0, 0), nd = mk3(0, 0, 0), so = mk3(0, 0, 0), sd = mk3(0, 0, 0); float stmax = 0.0f;
    if (i < n) {
        pid = qin.pid[i];
        const f3 ro = xyz(qin.o_tmin[i]), rd = xyz(qin.d_tmax[i]);     // payload.origin / payload.direction
        const float4 sv = st.at(LP_SEED, pid), cv = st.at(LP_COLOR, pid);
        uint2 seed = make_uint2(__float_as_uint(sv.x), __float_as_uint(sv.y));
        f3 color = xyz(cv); float pdf = cv.w;
        f3 emission = xyz(st.at(LP_EMISSION, pid));
        const uint32_t inst = hit_inst[i];
        float util_x = 0.0f;
        if (inst == 0xFFFFFFFFu) {                                      // Miss.hlsl:9-11
            emission = emission + mk3(0.0f, 0.0f, 0.0f) * color;
            util_x = 1.0f;
        } else {
            const float4 ha = hit_a[i];
            const uint32_t prim = __float_as_uint(ha.w);
            const ModelRef M = S.models[S.inst_model[inst]];
            const uint32_t vertId = 3u * prim;
            const uint32_t mslot = vertId + M.mat_offset;
            const uint32_t materialID = mslot < S.n_material_ids ? __ldg(&S.material_ids[mslot]) : 0u;
            const MatFull mat = load_mat_full(S, materialID);
            const float bary[3] = {(1.0f - ha.y) - ha.z, ha.y, ha.z};
            const uint32_t vi[3] = {__ldg(&M.idx[vertId]), __ldg(&M.idx[vertId + 1]), __ldg(&M.idx[vertId + 2])};
            f3 pos[3], nrm[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float* v = reinterpret_cast<const float*>(M.verts + (size_t)vi[k] * 28);
                pos[k] = mk3(__ldg(v), __ldg(v + 1), __ldg(v + 2));
                nrm[k] = mk3(__ldg(v + 3), __ldg(v + 4), __ldg(v + 5));
            }
            f3 flatNormal = normalize3(cross3(pos[1] - pos[0], pos[2] - pos[0]));
            f3 smooth = mk3(0, 0, 0);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                if (nrm[k].x != 0.0f && nrm[k].y != 0.0f && nrm[k].z != 0.0f) smooth = smooth + nrm[k] * bary[k];
                else smooth = smooth + flatNormal * bary[k];
            }
            f3 normal;
            if (length3(smooth) > 0.0001f) normal = normalize3(smooth); else normal = flatNormal;
            const rtx_instance_props* ip = S.props + inst;
            normal = normalize3(mul43(ip->objectToWorldNormal, normal.x, normal.y, normal.z, 0.0f));
            flatNormal = normalize3(mul43(ip->objectToWorldNormal, flatNormal.x, flatNormal.y, flatNormal.z, 0.0f));
            if (dot3(normal, -rd) < 0.0f) normal = -normal;             // :108-111
            if (dot3(flatNormal, -rd) < 0.0f) flatNormal = -flatNormal;
            const f3 worldOrigin = ro + ha.x * rd;
            f3 emissive = mk3(0, 0, 0);
            float pdf_sample = 1.0f;
            f3 brdf_sample = mk3(0, 0, 0);
            f3 incoming = -rd;
            if (length3(mat.Ke) > 0.0f) {                               // :127-171
                if (yb == 0u) { emissive = mat.Ke; util_x = 1.0f; }
                else {
                    const f3 L = worldOrigin - ro;
                    const float dist2 = fmaxf(dot3(L, L), LG_EPS);
                    const float dist = fmaxf(sqrtf(dist2), LG_EPS);
                    const f3 Ln = L / dist;
                    const float cos_e = fmaxf(LG_EPS, dot3(normal, -Ln));
                    const f3 x_v = mul43(ip->objectToWorld, pos[0].x, pos[0].y, pos[0].z, 1.0f);
                    const f3 y_v = mul43(ip->objectToWorld, pos[1].x, pos[1].y, pos[1].z, 1.0f);
                    const f3 z_v = mul43(ip->objectToWorld, pos[2].x, pos[2].y, pos[2].z, 1.0f);
                    const f3 cross_l = cross3(y_v - x_v, z_v - x_v);
                    const float area_l = fabsf(length3(cross_l) * 0.5f);
                    const float s_weight = area_l * (((mat.Ke.x + mat.Ke.y) + mat.Ke.z) / 3.0f);
                    const float t_weight = __ldg(&S.lights[0].total_weight);
                    const float weight = s_weight / t_weight;
                    const float pdf_l = fmaxf(LG_EPS, (weight * dist2) / cos_e);
                    const float weight_emissive = pdf / (pdf + pdf_l);
                    emissive = (mat.Ke * color) * weight_emissive;
                    util_x = 1.0f;
                }
                emission = emission + (abs3(mk3(0, 0, 0)) + abs3(emissive));      // :337 (direct = 0)
            } else {                                                    // :173-331
                const f3 outgoing = -rd;
                const uint32_t strategy = lg_SelectSamplingStrategy(S, mat, outgoing, normal, seed);
                float ris_cdf[LG_RIS_M];
                // the selection needs every cdf entry, the selected candidate's record is recomputed from its own seed state afterwards
                // instead of keeping 10 records of 13 floats in local memory
                uint2 seeds[LG_RIS_M];
                float run = 0.0f;
                struct Cand { f3 f, LDir; float lum, dist, cos_y, pdf_brdf, pdf_l; };
                auto candidate = [&](uint2& sd_, Cand& c) -> float {
                    const float randomValue = RandomFloat(sd_);
                    const float4* lt = reinterpret_cast<const float4*>(S.lights + SelectLight(S, randomValue));
                    const float4 l0 = __ldg(lt), l1 = __ldg(lt + 1), l2 = __ldg(lt + 2), l3 = __ldg(lt + 3);
                    const float* MM = S.props[__float_as_uint(l1.w)].objectToWorld;
                    const f3 x_v = mul43(MM, l0.x, l0.y, l0.z, 1.0f);
                    const f3 y_v = mul43(MM, l1.x, l1.y, l1.z, 1.0f);
                    const f3 z_v = mul43(MM, l2.x, l2.y, l2.z, 1.0f);
                    float xi1 = RandomFloat(sd_), xi2 = RandomFloat(sd_);
                    if (xi1 + xi2 > 1.0f) { xi1 = 1.0f - xi1; xi2 = 1.0f - xi2; }
                    const float u = (1.0f - xi1) - xi2, v = xi1, w = xi2;
                    const f3 samplePoint = (u * x_v + v * y_v) + w * z_v;
                    const f3 L = samplePoint - (worldOrigin + LG_BIAS * flatNormal);
                    const float dist2 = fmaxf(dot3(L, L), LG_EPS);
                    const float dist = fmaxf(sqrtf(dist2), LG_EPS);
                    const f3 L_norm = L / dist;
                    const f3 cross_l = cross3(y_v - x_v, z_v - x_v);
                    const f3 normal_l = normalize3(cross_l);
                    const float area_l = fabsf(length3(cross_l) * 0.5f);
                    const float cos_theta_x = fmaxf(LG_EPS, dot3(normal, L_norm));
                    const float cos_theta_y = fmaxf(LG_EPS, dot3(normal_l, -L_norm));
                    const float G = fmaxf((cos_theta_x * cos_theta_y) / dist2, LG_EPS);
                    const float pdf_l = l2.w / fmaxf(area_l, LG_EPS);
                    const f3 emission_l = mk3(l3.x, l3.y, l3.z);
                    const f3 brdf_light = lg_EvaluateBRDF(S, strategy, mat, normal, -L_norm, -rd);
                    const float pdf_brdf_light = fmaxf(lg_BRDF_PDF(strategy, mat, normal, -L_norm, -rd), LG_EPS);
                    c.f = (emission_l * brdf_light) * G;
                    const float lum = ((emission_l.x + emission_l.y) + emission_l.z) / 3.0f;
                    c.lum = (lum * brdf_light.x) * G;                   // float3 -> float keeps .x (:268)
                    c.LDir = L_norm; c.dist = dist; c.cos_y = cos_theta_y; c.pdf_brdf = pdf_brdf_light; c.pdf_l = pdf_l;
                    return (1.0f / 10.0f) * (((lum * brdf_light.x) * G) / pdf_l);
                };
                for (int k = 0; k < LG_RIS_M; k++) {
                    seeds[k] = seed;
                    Cand c;
                    const float wgt = candidate(seed, c);
                    run = (k == 0) ? wgt : run + wgt;
                    ris_cdf[k] = run;
                }
                const float ris_total_weight = ris_cdf[LG_RIS_M - 1];
                const float threshold = RandomFloat(seed) * ris_total_weight;
                int sel = 0;
                for (int k = 0; k < LG_RIS_M; k++) if (threshold < ris_cdf[k]) { sel = k; break; }
                Cand c;
                uint2 ss = seeds[sel];
                candidate(ss, c);
                const float WX = fmaxf(LG_EPS, (1.0f / fmaxf(LG_EPS, c.lum)) * ris_total_weight);
                so = worldOrigin + LG_BIAS * flatNormal; sd = c.LDir; stmax = fabsf(c.dist) - LG_BIAS;
                emit_sh = true;
                const float pdf_l_sa = fmaxf(LG_EPS, ((c.pdf_l * c.dist) * c.dist) / c.cos_y);
                const float weight_light = pdf_l_sa / (pdf_l_sa + c.pdf_brdf);
                st.at(LP_RISF, pid) = f4(c.f, WX);
                st.at(LP_COLW, pid) = f4(color * weight_light, 0.0f);
                const f3 sample = strategy == 0u ? RandomUnitVectorInHemisphere(normal, seed) : lg_SampleBRDF_GGX(mat, outgoing, normal, seed);
                nd = sample;
                no = worldOrigin + LG_BIAS * flatNormal;
                incoming = -nd;
                if (length3(nd) < 0.01f) util_x = 1.1f;                 // :320-323
                else {
                    pdf_sample = fmaxf(lg_BRDF_PDF(strategy, mat, normal, incoming, outgoing), 0.0001f);
                    brdf_sample = lg_EvaluateBRDF(S, strategy, mat, normal, incoming, outgoing);
                }
                // payload.emission += abs(direct) + abs(emissive) is finished by k_legacy_shadow once the visibility is known
            }
            color = ((color * brdf_sample) * dot3(normal, -incoming)) / pdf_sample;     // :342
            pdf = pdf_sample;
        }
        bool go = !(util_x >= 1.0f);
        if (go && yb > 3u) {                                            // RayGen.hlsl:118-130
            const float max_throughput = fmaxf(color.x, fmaxf(color.y, color.z));
            const float q = fminf(fmaxf(max_throughput, 0.05f), 1.0f);
            const float random = RandomFloat(seed);
            if (random > q) go = false;
            else color = color * (1.0f / q);
        }
        emit = go && (yb + 1u < S.bounces);                             // D13
        st.at(LP_SEED, pid) = make_float4(__uint_as_float(seed.x), __uint_as_float(seed.y), 0.0f, 0.0f);
        st.at(LP_COLOR, pid) = f4(color, pdf);
        st.at(LP_EMISSION, pid) = f4(emission, 0.0f);
    }
    push_ray(q_shadow, emit_sh, so, LG_BIAS, sd, stmax, pid);
    push_ray(qout, emit, no, 0.0001f, nd, 10000.0f, pid);
}

// Hit.hlsl:297-337: direct = ris_f * visible * WX; direct *= throughput * weight_light; payload.emission += abs(direct) + abs(emissive = 0)
__global__ void __launch_bounds__(LG_BLOCK)
k_legacy_shadow(StateView st, const uint32_t* __restrict__ n_ptr, const uint32_t* __restrict__ pid_of, const uint32_t* __restrict__ hit_inst,
                unsigned long long* ray_counters) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = *n_ptr;
    if (j == 0) atomicAdd(&ray_counters[1], (unsigned long long)n);
    if (j >= n) return;
    const uint32_t pid = pid_of[j];
    const float visible = hit_inst[j] != 0xFFFFFFFFu ? 0.0f : 1.0f;
    const float4 rf = st.at(LP_RISF, pid);
    f3 direct = (xyz(rf) * visible) * rf.w;
    direct = direct * xyz(st.at(LP_COLW, pid));
    const float4 e = st.at(LP_EMISSION, pid);
    st.at(LP_EMISSION, pid) = f4(xyz(e) + (abs3(direct) + abs3(mk3(0, 0, 0))), 0.0f);
}

cudaError_t wave_render_pass_legacy(WaveBuffers& B, const SceneData& S, const SceneAS& AS, uint32_t first_sample, uint32_t spp, cudaStream_t stream,
                                    uint64_t* launches, PassTiming* T) {
    const uint32_t npx = S.width * S.height;
    const uint32_t n = npx * spp;
    if (n > B.n_paths || S.bounces > 60u) return cudaErrorInvalidValue;
    StateView st{B.state, n};
    const unsigned grid = (n + LG_BLOCK - 1) / LG_BLOCK;
    uint64_t L = 0;
    T->n_marks = 0;
    auto mark = [&](StageKind k) -> cudaError_t {
        if (T->stage_timing && T->n_marks + 2 < WAVE_MAX_EVENTS) {
            T->kind[T->n_marks] = (unsigned char)k;
            CKE(cudaEventRecord(T->ev[2 + T->n_marks], stream));
            T->n_marks++;
        }
        L++;
        return cudaSuccess;
    };
    CKE(cudaMemsetAsync(B.counts, 0, 128 * 4, stream));
    CKE(cudaEventRecord(T->ev[0], stream));
    // counter slots: 0 = primary queue, 1 + y = queue emitted by bounce y, 64 + y = shadow queue of bounce y
    RayQueue qin = B.q[0]; qin.count = B.counts + 0;
    CKE(mark(SK_GENERATE));
    k_legacy_generate<<<grid, LG_BLOCK, 0, stream>>>(st, qin, B.cam, S.width, S.height, first_sample, B.ray_counters);
    int cur = 0;
    for (uint32_t yb = 0; yb < S.bounces; yb++) {
        CKE(mark(SK_CLOSEST));
        CKE(launch_trace(AS, qin.o_tmin, qin.d_tmax, qin.count, 0, B.cursor, B.hit_a, B.hit_inst, false, T->stats, stream));
        RayQueue qout = B.q[cur ^ 1]; qout.count = B.counts + 1 + yb;
        RayQueue qsh = B.sq[0]; qsh.count = B.counts + 64 + yb;
        CKE(mark(SK_GI_STEP));
        k_legacy_hit<<<grid, LG_BLOCK, 0, stream>>>(st, S, qin, B.hit_a, B.hit_inst, qsh, qout, yb, B.ray_counters);
        CKE(mark(SK_ANY));
        CKE(launch_trace(AS, qsh.o_tmin, qsh.d_tmax, qsh.count, 0, B.cursor, B.hit_a, B.hit_inst, true, nullptr, stream));
        CKE(mark(SK_SCATTER));
        k_legacy_shadow<<<grid, LG_BLOCK, 0, stream>>>(st, qsh.count, qsh.pid, B.hit_inst, B.ray_counters);
        qin = qout; cur ^= 1;
    }
    CKE(mark(SK_ACCUMULATE));
    CKE(wave_accumulate(B, npx, spp, stream));
    CKE(cudaEventRecord(T->ev[1], stream));
    CKE(cudaGetLastError());
    if (launches) *launches += L;
    return cudaSuccess;
}

}  // namespace rtx
