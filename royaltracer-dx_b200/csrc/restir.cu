// restir.cu — ReSTIR DI + GI reuse: the reference's RayGen2 (temporal) and RayGen3 (spatial reuse, final shade,
// accumulation) — shaders/Pass_temp_di_v7.hlsl:46-204, shaders/Pass_spat_di_v7.hlsl:46-464, shaders/MIS_v7.hlsl,
// include/MIS_GI_v6.hlsl, shaders/Common_v7.hlsl:203-350, shaders/Sampler_v7.hlsl:163-195,738-785 (paths relative to
// /root/reference/Pathtracer/).  SURVEY.md §8f rank 1.
//
// Wavefront form: every per-pixel program is cut where a visibility TraceRay needs its answer.
//   temporal:  A (reproject, accept, emit <= 2 shadow rays) -> any-hit trace -> B (pairwise MIS, reservoir merge)
//   spatial:   A (neighbour search, emit <= 9 shadow rays)  -> trace -> B (pairwise MIS, merges, emit the winner's ray)
//              -> trace -> C (W, final shade, accumulation F20)
// Visibility answers come back as one bit per (pixel, slot) in a per-pixel mask.  A ray is traced only when the unshadowed
// value it multiplies is > 0 (deviation D6'): the product is 0 or NaN otherwise, whatever V is.
// Arithmetic order follows the reference line by line; the draw order of each pixel's RNG is the sequential one.
#include "restir.h"
#include "trace.h"
#include "wave_dev.cuh"

namespace rtx {

#define CKE(call)                            \
    do {                                     \
        cudaError_t e__ = (call);            \
        if (e__ != cudaSuccess) return e__;  \
    } while (0)

#define RS_BLOCK 128
#ifndef RTX_RS_MINB
#define RTX_RS_MINB 8      // resident CTAs per SM the register budget of the reuse kernels (k_temporal_a/b, k_spatial_a/b) is set for:
                           // 64 registers instead of 72-122 measured best (ReSTIR frame on C2 13.2 -> 12.96 ms)
#endif
#define RS_TEMPORAL_M_CAP 16u        // Common_v7.hlsl:19-20
#define RS_SPATIAL_M_CAP 128u        // :17-18
#define RS_CANDIDATES 3u             // :13
#define RS_MAX_TRIES 9u              // :14
#define RS_RADIUS 20u                // :15
#define RS_WSUM_THRESHOLD 5.0f       // :22
#define RS_J_THRESHOLD 5.0f          // :23

struct Res { f3 x; float w_sum; f3 n; float W; f3 L; uint32_t M; };                 // Reservoir_DI / Reservoir_GI
struct SD { f3 x1; uint32_t mID; f3 n1; uint32_t objID; f3 o; uint32_t kind; f3 L1; };

__device__ __forceinline__ Res load_res(const float4* __restrict__ b, uint32_t n, uint32_t i) {
    const float4 a = b[i], c = b[(size_t)n + i], d = b[(size_t)2 * n + i];
    Res r; r.x = xyz(a); r.w_sum = a.w; r.n = xyz(c); r.W = c.w; r.L = xyz(d); r.M = __float_as_uint(d.w);
    return r;
}
__device__ __forceinline__ void store_res(float4* b, uint32_t n, uint32_t i, const Res& r) {
    b[i] = f4(r.x, r.w_sum); b[(size_t)n + i] = f4(r.n, r.W); b[(size_t)2 * n + i] = f4u(r.L, r.M);
}
__device__ __forceinline__ SD load_sd(const float4* __restrict__ b, uint32_t n, uint32_t i) {
    const float4 a = b[i], c = b[(size_t)n + i], d = b[(size_t)2 * n + i], e = b[(size_t)3 * n + i];
    SD s; s.x1 = xyz(a); s.mID = __float_as_uint(a.w) & 0xFFFFu; s.kind = __float_as_uint(a.w) >> 16;
    s.n1 = xyz(c); s.objID = __float_as_uint(c.w); s.o = xyz(d); s.L1 = xyz(e);
    return s;
}
__device__ __forceinline__ void store_sd(float4* b, uint32_t n, uint32_t i, const SD& s) {
    b[i] = f4u(s.x1, (s.mID & 0xFFFFu) | (s.kind << 16)); b[(size_t)n + i] = f4u(s.n1, s.objID);
    b[(size_t)2 * n + i] = f4(s.o, 0.0f); b[(size_t)3 * n + i] = f4(s.L1, 0.0f);
}
__device__ __forceinline__ Res zero_res() {
    Res r; r.x = r.n = r.L = mk3(0, 0, 0); r.w_sum = r.W = 0.0f; r.M = 0u; return r;
}

__device__ __forceinline__ float lin(f3 v) { return length3(v); }                                  // LinearizeVector, Sampler_v7.hlsl:1-5
__device__ __forceinline__ float GetW(float w_sum, float p_hat) { return p_hat > RTX_EPS ? w_sum / p_hat : 0.0f; }   // :183-195
__device__ __forceinline__ bool nz3(f3 v) { return v.x != 0.0f || v.y != 0.0f || v.z != 0.0f; }    // length(half3) != 0 (D10)
__device__ __forceinline__ uint32_t mcap(uint32_t cap, uint32_t M) { return M < cap ? M : cap; }
// Common_v7.hlsl:297-304
__device__ __forceinline__ bool RejectDistance(f3 x1, f3 x2, f3 camPos, float threshold) {
    const float d1 = length3(x1 - camPos), d2 = length3(x2 - camPos);
    const float rel = fabsf(d1 - d2) / fmaxf(d1, d2);
    return rel > threshold;
}
__device__ __forceinline__ bool RejectJacobian(float J, float threshold) {                         // :268-272
    return J > threshold || J < 1.0f / threshold || isnan1(J) || isinf1(J);
}
__device__ __forceinline__ bool IsValidDI(const Res& r) { return length3(r.n) > 0.0f && nz3(r.L) && r.w_sum > 0.0f && r.M > 0u; }   // :307-314
__device__ __forceinline__ bool IsValidGI(const Res& r) { return r.w_sum > 0.0f && r.M > 0u; }     // :317-322
// Jacobian_Reconnection(sdata_r, sdata_q, x2q, n2q), Common_v7.hlsl:326-345
__device__ __forceinline__ float Jacobian(f3 r_x1, f3 q_x1, f3 x2q, f3 n2q) {
    const f3 vq = x2q - q_x1, vr = x2q - r_x1;
    const float cosPhi2q = fabsf(dot3(normalize3(-vq), normalize3(n2q)));
    const float cosPhi2r = fabsf(dot3(normalize3(-vr), normalize3(n2q)));
    const float len2_vq = dot3(vq, vq), len2_vr = dot3(vr, vr);
    return (cosPhi2q / cosPhi2r) * (len2_vr / len2_vq);
}
// GetRandomPixelCircleWeighted, Common_v7.hlsl:203-244 (pow(u, spatial_exponent = 1) is the identity)
__device__ __forceinline__ uint32_t RandomPixel(uint32_t w, uint32_t h, uint32_t x, uint32_t y, uint2& seed) {
    int newX, newY;
    do {
        const float u = RandomFloat(seed);
        const float r = (float)RS_RADIUS * u;
        const float angle = RandomFloat(seed) * 6.2831853f;
        float sn, cs; d_sincos(angle, &sn, &cs);
        const int offsetX = (int)(cs * r), offsetY = (int)(sn * r);
        newX = (int)x + offsetX; newY = (int)y + offsetY;
        while (newX < 0 || newX >= (int)w) { if (newX < 0) newX = -newX; else newX = 2 * (int)w - newX - 2; }
        while (newY < 0 || newY >= (int)h) { if (newY < 0) newY = -newY; else newY = 2 * (int)h - newY - 2; }
    } while (newX == (int)x && newY == (int)y);
    return (uint32_t)newY * w + (uint32_t)newX;
}
// GetBestReprojectedPixel_d, Sampler_v7.hlsl:738-785.  D11: outside the image = no candidate.
__device__ __forceinline__ bool Reproject(const SceneData& S, const rtx_camera_params* cam, f3 worldPos, uint32_t objID, uint32_t n_inst,
                                          uint32_t& tp) {
    if (objID >= n_inst) return false;
    const rtx_instance_props* ip = S.props + objID;
    const float4 lp = mul44(ip->objectToWorldInverse, worldPos.x, worldPos.y, worldPos.z, 1.0f);
    const float4 pw = mul44(ip->prevObjectToWorld, lp.x, lp.y, lp.z, lp.w);
    const float4 vp = mul44(cam->prevView, pw.x, pw.y, pw.z, pw.w);
    const float4 cp = mul44(cam->prevProjection, vp.x, vp.y, vp.z, vp.w);
    if (cp.w <= 0.0f) return false;
    const float ndcx = cp.x / cp.w, ndcy = cp.y / cp.w;
    const float uvx = ndcx * 0.5f + 0.5f;
    float uvy = ndcy * 0.5f + 0.5f;
    uvy = 1.0f - uvy;
    const float fx = rintf(uvx * (float)S.width), fy = rintf(uvy * (float)S.height);
    if (!(fx >= 0.0f && fx < (float)S.width && fy >= 0.0f && fy < (float)S.height)) return false;
    tp = (uint32_t)(int)fy * S.width + (uint32_t)(int)fx;
    return true;
}

struct VisRay { f3 o, d; float tmax; };
// VisibilityCheck(x1, n1, normalize(x2 - x1), length(x2 - x1)), Sampler_v7.hlsl:86-104,167-168
__device__ __forceinline__ VisRay make_vis_ray(f3 x1, f3 n1, f3 x2) {
    const f3 dd = x2 - x1;
    VisRay r;
    r.d = normalize3(dd);
    const float dist = length3(dd);
    r.o = x1 + normalize3(n1) * RTX_S_BIAS;
    r.tmax = fmaxf(dist - 10.0f * RTX_S_BIAS, 2.0f * RTX_S_BIAS);
    return r;
}
__device__ __forceinline__ float vbit(uint32_t mask, uint32_t slot) { return (mask >> slot) & 1u ? 0.0f : 1.0f; }

struct RsView {
    float4 *di_cur, *di_last, *gi_cur, *gi_last, *sd_cur, *sd_last;
    uint4* cand; float4* tmp; uint32_t* vmask; uint32_t n;
};

// ------------------------------------------------------------------------------------------------ pass 1 tail
// Pass_init_di_v7.hlsl:112-189: reservoir (M = 1, W from the shadowed p_hat), reservoir_GI (M = 1, W_GI), SampleData
__global__ void __launch_bounds__(RS_BLOCK)
k_restir_store(StateView st, RsView R) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= R.n) return;
    const float4 res = st.at(SP_RESULT, p);
    Res di = zero_res(), gi = zero_res();
    SD sd; sd.x1 = sd.n1 = sd.o = sd.L1 = mk3(0, 0, 0); sd.mID = 0u; sd.objID = 0u; sd.kind = 0u;
    if (res.w == 3.0f) {                    // primary hit on an emitter: L1 = half3(Ke), no sampling (:103-106,129-137)
        sd.kind = 1u; sd.L1 = xyz(res);
        sd.mID = __float_as_uint(st.at(SP_X1, p).w) & 0xFFFFu;
        sd.objID = __float_as_uint(st.at(SP_DI_L2, p).w);
    } else if (res.w == 2.0f) {
        sd.kind = 2u;
        const float4 a0 = st.at(SP_X1, p), b0 = st.at(SP_DI_X2, p), b2 = st.at(SP_DI_L2, p), c0 = st.at(SP_GI_XN, p), c2 = st.at(SP_GI_E3, p);
        sd.x1 = xyz(a0); sd.mID = __float_as_uint(a0.w) & 0xFFFFu;
        sd.n1 = normalize3(xyz(st.at(SP_N1, p))); sd.o = xyz(st.at(SP_O, p)); sd.objID = __float_as_uint(b2.w);
        di.x = xyz(b0); di.w_sum = b0.w; di.n = xyz(st.at(SP_DI_N2, p)); di.W = st.at(SP_DI_R, p).w; di.L = xyz(b2); di.M = 1u;
        // reservoir_GI.xn / nn are written by UpdateReservoir_GI only (Path_Sampler_v7.hlsl:175-183,251-259: xn, normalize(nn))
        gi.w_sum = c0.w; gi.W = c2.w; gi.L = xyz(c2); gi.M = 1u;
        if (st.at(SP_GI_SC, p).z != 0.0f) { gi.x = xyz(c0); gi.n = normalize3(xyz(st.at(SP_GI_NN, p))); }
    }
    store_res(R.di_cur, R.n, p, di); store_res(R.gi_cur, R.n, p, gi); store_sd(R.sd_cur, R.n, p, sd);
}

// ------------------------------------------------------------------------------------------------ temporal reuse (RayGen2)
struct TemporalSetup { bool okDI, okGI; uint32_t tp; };
__device__ __forceinline__ TemporalSetup temporal_setup(const SceneData& S, const rtx_camera_params* cam, const RsView& R, const SD& sc,
                                                       uint32_t n_inst, Res& rl, Res& gl) {
    TemporalSetup t; t.okDI = t.okGI = false; t.tp = 0;
    if (!Reproject(S, cam, sc.x1, sc.objID, n_inst, t.tp)) return t;
    const f3 init_orig = mul43(cam->viewI, 0.0f, 0.0f, 0.0f, 1.0f);
    rl = load_res(R.di_last, R.n, t.tp); gl = load_res(R.gi_last, R.n, t.tp);
    const SD sl = load_sd(R.sd_last, R.n, t.tp);
    const bool common = !nz3(sl.L1) && !RejectDistance(sc.x1, sl.x1, init_orig, 0.1f) && sl.mID == sc.mID;      // Pass_temp_di_v7.hlsl:88-107
    t.okDI = common && IsValidDI(rl) && (rl.x.x != 0.0f && rl.x.y != 0.0f && rl.x.z != 0.0f);
    t.okGI = common && !(gl.w_sum > RS_WSUM_THRESHOLD) && IsValidGI(gl);
    return t;
}

__global__ void __launch_bounds__(RS_BLOCK, RTX_RS_MINB)
k_temporal_a(RsView R, SceneData S, const rtx_camera_params* __restrict__ cam, uint32_t n_inst, RayQueue q) {
    const uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
    bool e0 = false, e1 = false; VisRay r0, r1;
    r0.o = r0.d = r1.o = r1.d = mk3(0, 0, 0); r0.tmax = r1.tmax = 0.0f;
    if (pix < R.n) {
        const SD sc = load_sd(R.sd_cur, R.n, pix);
        if (sc.kind == 2u) {
            Res rl, gl;
            const TemporalSetup t = temporal_setup(S, cam, R, sc, n_inst, rl, gl);
            const MatOpt m = load_matopt(S, sc.mID, nullptr);
            if (t.okDI) {
                e0 = lin(ReconnectDI(S, sc.x1, sc.n1, rl.x, rl.n, rl.L, sc.o, m)) > 0.0f;
                r0 = make_vis_ray(sc.x1, sc.n1, rl.x);
            }
            if (t.okGI) {
                e1 = lin(ReconnectGI(S, sc.x1, sc.n1, gl.x, gl.L, sc.o, m)) > 0.0f;
                r1 = make_vis_ray(sc.x1, sc.n1, gl.x);
            }
        }
    }
    // the queue's trace order is recorded like in the E0 stages (wave_dev.cuh push_ray2)
    push_ray2(q, e0, e0 && ray_is_heavy(S, r0.o, 0.0f, r0.d, r0.tmax), r0.o, 0.0f, r0.d, r0.tmax, pix * 16u + 0u);
    push_ray2(q, e1, e1 && ray_is_heavy(S, r1.o, 0.0f, r1.d, r1.tmax), r1.o, 0.0f, r1.d, r1.tmax, pix * 16u + 1u);
}

__global__ void __launch_bounds__(RS_BLOCK, RTX_RS_MINB)
k_temporal_b(RsView R, SceneData S, const rtx_camera_params* __restrict__ cam, uint32_t n_inst, uint32_t frame) {
    const uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= R.n) return;
    const SD sc = load_sd(R.sd_cur, R.n, pix);
    if (sc.kind != 2u) return;
    Res rl, gl;
    const TemporalSetup t = temporal_setup(S, cam, R, sc, n_inst, rl, gl);
    if (!t.okDI && !t.okGI) return;
    const uint32_t vm = R.vmask[pix];
    const MatOpt m = load_matopt(S, sc.mID, nullptr);
    uint2 seed = init_seed(pix % S.width, pix / S.width, 2u, frame);
    if (t.okDI) {                                                                   // Pass_temp_di_v7.hlsl:112-153
        Res rc = load_res(R.di_cur, R.n, pix);
        const float cM = (float)mcap(RS_TEMPORAL_M_CAP, rc.M), nM = (float)mcap(RS_TEMPORAL_M_CAP, rl.M);
        const float M_sum = cM + nM;
        float mi_c = cM / M_sum;                                                    // MIS_v7.hlsl:63-71
        { const float m_num = cM, m_den = m_num + (M_sum - cM); if (m_den > 0.0f) mi_c += (nM / M_sum) * (m_num / m_den); }
        float mi_t = 0.0f;                                                          // :73-80
        { const float m_num = M_sum - cM, m_den = m_num + cM; if (m_den > 0.0f) mi_t = ((nM / M_sum) * m_num) / m_den; }
        if (length3(rl.n) == 0.0f) { mi_c = 1.0f; mi_t = 0.0f; }
        const float w_c = (mi_c * lin(ReconnectDI(S, sc.x1, sc.n1, rc.x, rc.n, rc.L, sc.o, m))) * rc.W;
        const float w_t = (mi_t * (lin(ReconnectDI(S, sc.x1, sc.n1, rl.x, rl.n, rl.L, sc.o, m)) * vbit(vm, 0u))) * rl.W;
        rc.M = mcap(RS_TEMPORAL_M_CAP, rc.M) + mcap(RS_TEMPORAL_M_CAP, rl.M);
        rc.w_sum = w_c + w_t;                                                       // UpdateReservoir, Reservoir_v7.hlsl:57-80
        if (RandomFloat(seed) < w_t / rc.w_sum) { rc.x = rl.x; rc.n = rl.n; rc.L = rl.L; }
        const float p_hat = lin(ReconnectDI(S, sc.x1, sc.n1, rc.x, rc.n, rc.L, sc.o, m));
        rc.W = GetW(rc.w_sum, p_hat);
        store_res(R.di_cur, R.n, pix, rc);
    }
    if (t.okGI) {                                                                   // :156-198
        Res gc = load_res(R.gi_cur, R.n, pix);
        const float cM = (float)mcap(RS_TEMPORAL_M_CAP, gc.M), nM = (float)mcap(RS_TEMPORAL_M_CAP, gl.M);
        const float M_sum = cM + nM;
        float mi_c = cM / M_sum;                                                    // MIS_GI_v6.hlsl:79-95
        { const float m_num = cM, m_den = m_num + (M_sum - cM); if (m_den > 0.0f) mi_c += (nM / M_sum) * (m_num / m_den); }
        float mi_t = 0.0f;                                                          // :97-112
        { const float m_num = M_sum - cM, m_den = m_num + cM; if (m_den > 0.0f) mi_t = ((nM / M_sum) * m_num) / m_den; }
        const f3 f_c = ReconnectGI(S, sc.x1, sc.n1, gc.x, gc.L, sc.o, m);
        const float w_c = (mi_c * lin(f_c)) * gc.W;
        const f3 f_t = ReconnectGI(S, sc.x1, sc.n1, gl.x, gl.L, sc.o, m) * vbit(vm, 1u);
        const float w_t = (mi_t * lin(f_t)) * gl.W;
        gc.M = mcap(RS_TEMPORAL_M_CAP, gc.M) + mcap(RS_TEMPORAL_M_CAP, gl.M);
        gc.w_sum = w_c + w_t;                                                       // UpdateReservoir_GI, Reservoir_v7.hlsl:30-53
        if (RandomFloat(seed) < w_t / gc.w_sum) { gc.x = gl.x; gc.n = gl.n; gc.L = gl.L; }
        gc.W = GetW(gc.w_sum, lin(ReconnectGI(S, sc.x1, sc.n1, gc.x, gc.L, sc.o, m)));
        store_res(R.gi_cur, R.n, pix, gc);
    }
}

// ------------------------------------------------------------------------------------------------ spatial reuse (RayGen3)
// neighbour search, Pass_spat_di_v7.hlsl:84-196.  Writes the candidate lists and the RNG state that follows them.
__global__ void __launch_bounds__(RS_BLOCK, RTX_RS_MINB)
k_spatial_a(RsView R, SceneData S, const rtx_camera_params* __restrict__ cam, uint32_t frame, RayQueue q) {
    const uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = pix < R.n;
    SD sc; sc.kind = 0u; sc.x1 = sc.n1 = sc.o = sc.L1 = mk3(0, 0, 0); sc.mID = sc.objID = 0u;
    uint32_t candDI[3] = {0u, 0u, 0u}, candGI[3] = {0u, 0u, 0u}, nDI = 0u, nGI = 0u;
    Res can = zero_res(), cang = zero_res();
    MatOpt m;
    if (live) sc = load_sd(R.sd_cur, R.n, pix);
    const bool active = live && sc.kind == 2u;
    if (active) {
        const uint32_t x = pix % S.width, y = pix / S.width;
        const f3 init_orig = mul43(cam->viewI, 0.0f, 0.0f, 0.0f, 1.0f);
        uint2 seed = init_seed(x, y, 3u, frame);
        m = load_matopt(S, sc.mID, nullptr);
        for (uint32_t attempt = 0; attempt < RS_MAX_TRIES && nDI < RS_CANDIDATES; attempt++) {      // :105-134
            const uint32_t r = RandomPixel(S.width, S.height, x, y, seed);
            const SD sr = load_sd(R.sd_cur, R.n, r);
            const bool ok = !(dot3(sc.n1, sr.n1) < 0.9f) && !RejectDistance(sc.x1, sr.x1, init_orig, 0.1f) &&
                            IsValidDI(load_res(R.di_cur, R.n, r)) && !nz3(sr.L1) && sr.kind == 2u && sr.mID == sc.mID;
            if (ok) candDI[nDI++] = r;
        }
        for (uint32_t attempt = 0; attempt < RS_MAX_TRIES && nGI < RS_CANDIDATES; attempt++) {      // :144-186
            const uint32_t r = RandomPixel(S.width, S.height, x, y, seed);
            const SD sr = load_sd(R.sd_cur, R.n, r);
            const Res gr = load_res(R.gi_cur, R.n, r);
            const bool ok = m.Pr > 0.3f && !RejectDistance(sc.x1, sr.x1, init_orig, 0.1f) &&
                            !(dot3(normalize3(gr.x - sc.x1), sc.n1) < 0.0f) && !(gr.w_sum > RS_WSUM_THRESHOLD) && IsValidGI(gr) &&
                            !RejectJacobian(Jacobian(sr.x1, sc.x1, gr.x, gr.n), RS_J_THRESHOLD) && !nz3(sr.L1) && sr.kind == 2u &&
                            sr.mID == sc.mID;
            if (ok) candGI[nGI++] = r;
        }
        R.cand[pix] = make_uint4(candDI[0], candDI[1], candDI[2], nDI | (nGI << 8));
        R.cand[(size_t)R.n + pix] = make_uint4(candGI[0], candGI[1], candGI[2], 0u);
        R.cand[(size_t)2 * R.n + pix] = make_uint4(seed.x, seed.y, 0u, 0u);
        can = load_res(R.di_cur, R.n, pix); cang = load_res(R.gi_cur, R.n, pix);
    }
    // visibility rays of the pairwise-MIS terms: slots 0-2 canonical DI sample seen from DI neighbour j (MIS_v7.hlsl:24),
    // 3-5 canonical GI sample seen from GI neighbour j (MIS_GI_v6.hlsl:30), 6-8 GI neighbour v's sample seen from here (:316-325)
#pragma unroll 1
    for (uint32_t slot = 0; slot < 9u; slot++) {
        bool e = false; VisRay r; r.o = r.d = mk3(0, 0, 0); r.tmax = 0.0f;
        const uint32_t k = slot % 3u;
        if (active) {
            if (slot < 3u) {
                if (k < nDI) {
                    const SD sn = load_sd(R.sd_cur, R.n, candDI[k]);
                    e = lin(ReconnectDI(S, sn.x1, sn.n1, can.x, can.n, can.L, sn.o, m)) > 0.0f;
                    r = make_vis_ray(sn.x1, sn.n1, can.x);
                }
            } else if (slot < 6u) {
                if (k < nGI) {
                    const SD sn = load_sd(R.sd_cur, R.n, candGI[k]);
                    e = lin(ReconnectGI(S, sn.x1, sn.n1, cang.x, cang.L, sn.o, m)) > 0.0f;
                    r = make_vis_ray(sn.x1, sn.n1, cang.x);
                }
            } else if (k < nGI) {
                const Res gn = load_res(R.gi_cur, R.n, candGI[k]);
                e = lin(ReconnectGI(S, sc.x1, sc.n1, gn.x, gn.L, sc.o, m)) > 0.0f;
                r = make_vis_ray(sc.x1, sc.n1, gn.x);
            }
        }
        push_ray2(q, e, e && ray_is_heavy(S, r.o, 0.0f, r.d, r.tmax), r.o, 0.0f, r.d, r.tmax, pix * 16u + slot);
    }
}

// pairwise MIS + reservoir merges (Pass_spat_di_v7.hlsl:198-341), GI final shade (:367-381), emits the DI winner's ray (:343-352)
__global__ void __launch_bounds__(RS_BLOCK, RTX_RS_MINB)
k_spatial_b(RsView R, SceneData S, RayQueue q) {
    const uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
    bool e = false; VisRay vr; vr.o = vr.d = mk3(0, 0, 0); vr.tmax = 0.0f;
    if (pix < R.n) {
        const SD sc = load_sd(R.sd_cur, R.n, pix);
        if (sc.kind == 0u) {                                        // miss: radiance 0, nothing to reuse next frame (D1)
            store_res(R.di_last, R.n, pix, zero_res()); store_res(R.gi_last, R.n, pix, zero_res()); store_sd(R.sd_last, R.n, pix, sc);
            R.tmp[pix] = make_float4(0, 0, 0, 0); R.tmp[(size_t)R.n + pix] = make_float4(0, 0, 0, 0.0f);
        } else if (sc.kind == 1u) {                                 // :458-463; the *_last buffers keep their contents
            R.tmp[pix] = make_float4(0, 0, 0, 0); R.tmp[(size_t)R.n + pix] = f4(sc.L1, 1.0f);
        } else {
            const MatOpt m = load_matopt(S, sc.mID, nullptr);
            const uint4 c0 = R.cand[pix], c1 = R.cand[(size_t)R.n + pix], c2 = R.cand[(size_t)2 * R.n + pix];
            const uint32_t candDI[3] = {c0.x, c0.y, c0.z}, candGI[3] = {c1.x, c1.y, c1.z};
            const uint32_t nDI = c0.w & 0xffu, nGI = (c0.w >> 8) & 0xffu;
            uint2 seed = make_uint2(c2.x, c2.y);
            const uint32_t vm = R.vmask[pix];
            Res rc = load_res(R.di_cur, R.n, pix), gc = load_res(R.gi_cur, R.n, pix);
            const Res can = rc, cang = gc;
            const float CAPf_c = (float)mcap(RS_SPATIAL_M_CAP, can.M), CAPf_cg = (float)mcap(RS_SPATIAL_M_CAP, cang.M);
            float M_sum_DI = CAPf_c, M_sum_GI = CAPf_cg;                                            // :95-96,129,181
            for (uint32_t j = 0; j < nDI; j++) M_sum_DI += (float)mcap(RS_SPATIAL_M_CAP, __float_as_uint(R.di_cur[(size_t)2 * R.n + candDI[j]].w));
            for (uint32_t j = 0; j < nGI; j++) M_sum_GI += (float)mcap(RS_SPATIAL_M_CAP, __float_as_uint(R.gi_cur[(size_t)2 * R.n + candGI[j]].w));
            // ---- DI: GenPairwiseMIS_canonical (MIS_v7.hlsl:2-37); p_from[j] (unshadowed) is reused by the non-canonical weights
            const float p_c = lin(ReconnectDI(S, sc.x1, sc.n1, can.x, can.n, can.L, sc.o, m));
            float p_from[3] = {0.0f, 0.0f, 0.0f};
            float mi_c;
            {
                const float c_M_max = M_sum_DI - CAPf_c, c_m_num = CAPf_c * p_c;
                mi_c = CAPf_c / M_sum_DI;
                for (uint32_t j = 0; j < nDI; j++) {
                    const SD sn = load_sd(R.sd_cur, R.n, candDI[j]);
                    const float n_M_min = (float)mcap(RS_SPATIAL_M_CAP, __float_as_uint(R.di_cur[(size_t)2 * R.n + candDI[j]].w));
                    p_from[j] = lin(ReconnectDI(S, sn.x1, sn.n1, can.x, can.n, can.L, sn.o, m));
                    const float m_den = c_m_num + (c_M_max * (p_from[j] * vbit(vm, j)));
                    if (m_den > 0.0f) mi_c += (n_M_min / M_sum_DI) * (c_m_num / m_den);
                }
            }
            const float w_c = (mi_c * p_c) * can.W;                                                 // :210-222
            // ---- GI: GenPairwiseMIS_canonical_GI (MIS_GI_v6.hlsl:2-40)
            const f3 f_c = ReconnectGI(S, sc.x1, sc.n1, cang.x, cang.L, sc.o, m);
            const float p_c_gi = lin(f_c);
            float pg_from[3] = {0.0f, 0.0f, 0.0f}, jg[3] = {0.0f, 0.0f, 0.0f};
            float mi_c_gi;
            {
                const float c_M_max = M_sum_GI - CAPf_cg, c_m_num = CAPf_cg * p_c_gi;
                float m_c = CAPf_cg / M_sum_GI;
                for (uint32_t j = 0; j < nGI; j++) {
                    const SD sn = load_sd(R.sd_cur, R.n, candGI[j]);
                    const float n_M_min = (float)mcap(RS_SPATIAL_M_CAP, __float_as_uint(R.gi_cur[(size_t)2 * R.n + candGI[j]].w));
                    jg[j] = Jacobian(sc.x1, sn.x1, cang.x, cang.n);
                    const f3 fr = ReconnectGI(S, sn.x1, sn.n1, cang.x, cang.L, sn.o, m);
                    pg_from[j] = lin(fr);
                    const float p_hat_from = lin(fr * vbit(vm, 3u + j)) * jg[j];
                    const float m_den = c_m_num + (c_M_max * p_hat_from);
                    if (m_den > 0.0f) m_c += (n_M_min / M_sum_GI) * (c_m_num / m_den);
                }
                mi_c_gi = fminf(fmaxf(m_c, 0.0f), 1.0f);
            }
            const float w_c_gi = (mi_c_gi * p_c_gi) * cang.W;                                       // :236-240
            rc.M = mcap(RS_SPATIAL_M_CAP, can.M); rc.w_sum = w_c;
            gc.M = mcap(RS_SPATIAL_M_CAP, cang.M); gc.w_sum = w_c_gi;
            for (uint32_t v = 0; v < nDI; v++) {                                                    // :252-290
                const Res rn = load_res(R.di_cur, R.n, candDI[v]);
                float mi_s = 0.0f;                                                                  // MIS_v7.hlsl:40-61
                {
                    const float m_num = (M_sum_DI - CAPf_c) * p_from[v];
                    const float m_den = m_num + (CAPf_c * p_c);
                    if (m_den > 0.0f) mi_s = ((float)mcap(RS_SPATIAL_M_CAP, rn.M) / M_sum_DI) * (m_num / m_den);
                }
                const float w_s = (mi_s * lin(ReconnectDI(S, sc.x1, sc.n1, rn.x, rn.n, rn.L, sc.o, m))) * rn.W;
                rc.M += mcap(RS_SPATIAL_M_CAP, rn.M);
                rc.w_sum += w_s;
                if (RandomFloat(seed) < w_s / rc.w_sum) { rc.x = rn.x; rc.n = rn.n; rc.L = rn.L; }
            }
            for (uint32_t v = 0; v < nGI; v++) {                                                    // :293-341
                const Res gn = load_res(R.gi_cur, R.n, candGI[v]);
                const SD sn = load_sd(R.sd_cur, R.n, candGI[v]);
                float mi_s = 0.0f;                                                                  // MIS_GI_v6.hlsl:43-76
                {
                    const float p_hat_from = pg_from[v] * jg[v];
                    const float m_num = (M_sum_GI - CAPf_cg) * p_hat_from;
                    const float m_den = m_num + (CAPf_cg * p_c_gi);
                    if (m_den > 0.0f) mi_s = fminf(fmaxf(((float)mcap(RS_SPATIAL_M_CAP, gn.M) / M_sum_GI) * (m_num / m_den), 0.0f), 1.0f);
                }
                const float j_gi = Jacobian(sn.x1, sc.x1, gn.x, gn.n);
                const f3 f_gi = ReconnectGI(S, sc.x1, sc.n1, gn.x, gn.L, sc.o, m) * vbit(vm, 6u + v);
                const float w_s = ((mi_s * lin(f_gi)) * gn.W) * j_gi;
                if (j_gi != 0.0f) {
                    gc.M += mcap(RS_SPATIAL_M_CAP, gn.M);
                    gc.w_sum += w_s;
                    if (RandomFloat(seed) < w_s / gc.w_sum) { gc.x = gn.x; gc.n = gn.n; gc.L = gn.L; }
                }
            }
            // DI winner: its visibility ray decides W (:343-353); GI is final here (:367-381)
            const f3 rdi = ReconnectDI(S, sc.x1, sc.n1, rc.x, rc.n, rc.L, sc.o, m);
            const float f_g = lin(rdi);
            e = f_g > 0.0f;
            vr = make_vis_ray(sc.x1, sc.n1, rc.x);
            const f3 f_fin = ReconnectGI(S, sc.x1, sc.n1, gc.x, gc.L, sc.o, m);
            gc.W = GetW(gc.w_sum, lin(f_fin));
            store_res(R.di_last, R.n, pix, rc); store_res(R.gi_last, R.n, pix, gc); store_sd(R.sd_last, R.n, pix, sc);
            R.tmp[pix] = f4(rdi, f_g);
            R.tmp[(size_t)R.n + pix] = f4(f_fin * gc.W, 2.0f);
        }
    }
    push_ray2(q, e, e && ray_is_heavy(S, vr.o, 0.0f, vr.d, vr.tmax), vr.o, 0.0f, vr.d, vr.tmax, pix * 16u + 9u);
}

// W of the DI reservoir, final colour, accumulation F20 (Pass_spat_di_v7.hlsl:343-365,383-404)
__global__ void __launch_bounds__(RS_BLOCK)
k_spatial_c(RsView R, float4* __restrict__ accum) {
    const uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= R.n) return;
    const float4 t0 = R.tmp[pix], t1 = R.tmp[(size_t)R.n + pix];
    f3 C = xyz(t1);
    if (t1.w == 2.0f) {
        const float p_hat = t0.w * vbit(R.vmask[pix], 9u);
        const float w_sum = R.di_last[pix].w;
        const float W = GetW(w_sum, p_hat);
        R.di_last[(size_t)R.n + pix].w = W;
        C = xyz(t0) * W + xyz(t1);
    }
    if (!any_nan_inf(C)) {
        float4 a = accum[pix];
        a.x += C.x; a.y += C.y; a.z += C.z; a.w += 1.0f;
        accum[pix] = a;
    }
}

// The occlusion bits of a traced shadow queue (pid = pixel * 16 + slot) are set by the any-hit launch itself (trace.cu: vmask); this
// adds the queue's size to the shadow-ray counter
__global__ void k_count_shadow_rays(const uint32_t* __restrict__ n_ptr, unsigned long long* ray_counters) {
    atomicAdd(&ray_counters[1], (unsigned long long)*n_ptr);
}

__global__ void __launch_bounds__(RS_BLOCK)
k_restir_dump(RsView R, float* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R.n) return;
    const Res r = load_res(R.di_last, R.n, i), g = load_res(R.gi_last, R.n, i);
    const SD s = load_sd(R.sd_last, R.n, i);
    float* o = out + (size_t)40 * i;
    for (int k = 0; k < 40; k++) o[k] = 0.0f;
    o[0] = r.x.x; o[1] = r.x.y; o[2] = r.x.z; o[3] = r.w_sum; o[4] = r.n.x; o[5] = r.n.y; o[6] = r.n.z; o[7] = r.W;
    o[8] = r.L.x; o[9] = r.L.y; o[10] = r.L.z; o[11] = (float)r.M;
    o[12] = g.x.x; o[13] = g.x.y; o[14] = g.x.z; o[15] = g.w_sum; o[16] = g.n.x; o[17] = g.n.y; o[18] = g.n.z; o[19] = g.W;
    o[20] = g.L.x; o[21] = g.L.y; o[22] = g.L.z; o[23] = (float)g.M;
    o[24] = s.x1.x; o[25] = s.x1.y; o[26] = s.x1.z; o[27] = (float)s.mID; o[28] = s.n1.x; o[29] = s.n1.y; o[30] = s.n1.z;
    o[31] = (float)s.objID; o[32] = s.o.x; o[33] = s.o.y; o[34] = s.o.z; o[35] = (float)s.kind; o[36] = s.L1.x; o[37] = s.L1.y; o[38] = s.L1.z;
}

// ------------------------------------------------------------------------------------------------ host side
cudaError_t restir_alloc(RestirBuffers* R, uint32_t width, uint32_t height) {
    const size_t n = (size_t)width * height;
    R->n = (uint32_t)n;
    for (int i = 0; i < 2; i++) {
        CKE(cudaMalloc((void**)&R->di[i], n * RS_RES_PLANES * 16));
        CKE(cudaMalloc((void**)&R->gi[i], n * RS_RES_PLANES * 16));
        CKE(cudaMalloc((void**)&R->sd[i], n * RS_SD_PLANES * 16));
    }
    CKE(cudaMalloc((void**)&R->cand, n * 3 * 16));
    CKE(cudaMalloc((void**)&R->tmp, n * 2 * 16));
    CKE(cudaMalloc((void**)&R->vmask, n * 4));
    const size_t cap = n * RS_MAX_RAYS_PER_PIXEL;
    CKE(cudaMalloc((void**)&R->q.o_tmin, cap * 16));
    CKE(cudaMalloc((void**)&R->q.d_tmax, cap * 16));
    CKE(cudaMalloc((void**)&R->q.pid, cap * 4));
    CKE(cudaMalloc((void**)&R->q.order, cap * 4));
    R->q.cap = (uint32_t)cap;
    CKE(cudaMalloc((void**)&R->q_count, 12 * 4));          // queue i: count [i], heavy rays [4 + i], the others [8 + i]
    R->q.count = R->q_count;
    return cudaSuccess;
}

void restir_free(RestirBuffers* R) {
    for (int i = 0; i < 2; i++) {
        if (R->di[i]) cudaFree(R->di[i]);
        if (R->gi[i]) cudaFree(R->gi[i]);
        if (R->sd[i]) cudaFree(R->sd[i]);
    }
    void* ptrs[] = {R->cand, R->tmp, R->vmask, R->q.o_tmin, R->q.d_tmax, R->q.pid, R->q.order, R->q_count};
    for (void* p : ptrs) if (p) cudaFree(p);
    *R = RestirBuffers();
}

// the reference's buffers start zero-filled (committed resources): a zero reservoir is invalid, so frame 0 reuses nothing
cudaError_t restir_clear(RestirBuffers& R, cudaStream_t stream) {
    const size_t n = R.n;
    for (int i = 0; i < 2; i++) {
        CKE(cudaMemsetAsync(R.di[i], 0, n * RS_RES_PLANES * 16, stream));
        CKE(cudaMemsetAsync(R.gi[i], 0, n * RS_RES_PLANES * 16, stream));
        CKE(cudaMemsetAsync(R.sd[i], 0, n * RS_SD_PLANES * 16, stream));
    }
    return cudaSuccess;
}

static RsView view_of(RestirBuffers& R) {
    RsView v;
    v.di_cur = R.di[0]; v.di_last = R.di[1]; v.gi_cur = R.gi[0]; v.gi_last = R.gi[1]; v.sd_cur = R.sd[0]; v.sd_last = R.sd[1];
    v.cand = R.cand; v.tmp = R.tmp; v.vmask = R.vmask; v.n = R.n;
    return v;
}

cudaError_t restir_store_pass1(RestirBuffers& R, WaveBuffers& B, const SceneData& S, cudaStream_t stream, uint64_t* launches) {
    StateView st{B.state, R.n};
    k_restir_store<<<(R.n + RS_BLOCK - 1) / RS_BLOCK, RS_BLOCK, 0, stream>>>(st, view_of(R));
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t restir_reuse_passes(RestirBuffers& R, WaveBuffers& B, const SceneData& S, const SceneAS& AS, uint32_t frame_index,
                                cudaStream_t stream, uint64_t* launches) {
    const RsView V = view_of(R);
    const unsigned grid = (R.n + RS_BLOCK - 1) / RS_BLOCK;
    uint64_t L = 0;
    const bool lpt = S.heavy_valid != 0u;
    auto trace_queue = [&](const RayQueue& q) -> cudaError_t {
        CKE(launch_trace(AS, q.o_tmin, q.d_tmax, q.count, 0, B.cursor, B.hit_a, B.hit_inst, true, nullptr, stream, 1, q.order, q.n_heavy, q.cap,
                         q.pid, nullptr, R.vmask));
        k_count_shadow_rays<<<1, 1, 0, stream>>>(q.count, B.ray_counters);
        L += 2;
        return cudaGetLastError();
    };
    auto use_queue = [&](RayQueue& q, int i) {
        q.count = R.q_count + i; q.n_heavy = R.q_count + 4 + i; q.n_light = R.q_count + 8 + i;
        q.order = lpt ? R.q.order : nullptr;
    };
    CKE(cudaMemsetAsync(R.q_count, 0, 12 * 4, stream));
    CKE(cudaMemsetAsync(R.vmask, 0, (size_t)R.n * 4, stream));
    RayQueue q = R.q;
    // ---- RayGen2
    use_queue(q, 0);
    k_temporal_a<<<grid, RS_BLOCK, 0, stream>>>(V, S, B.cam, AS.n_instances, q);
    CKE(trace_queue(q));
    k_temporal_b<<<grid, RS_BLOCK, 0, stream>>>(V, S, B.cam, AS.n_instances, frame_index);
    // ---- RayGen3
    CKE(cudaMemsetAsync(R.vmask, 0, (size_t)R.n * 4, stream));
    use_queue(q, 1);
    k_spatial_a<<<grid, RS_BLOCK, 0, stream>>>(V, S, B.cam, frame_index, q);
    CKE(trace_queue(q));
    use_queue(q, 2);
    k_spatial_b<<<grid, RS_BLOCK, 0, stream>>>(V, S, q);
    CKE(trace_queue(q));
    k_spatial_c<<<grid, RS_BLOCK, 0, stream>>>(V, B.accum);
    L += 5;
    if (launches) *launches += L;
    return cudaGetLastError();
}

cudaError_t restir_dump(RestirBuffers& R, cudaStream_t stream, float* host_out) {
    float* d = nullptr;
    CKE(cudaMalloc((void**)&d, (size_t)R.n * 40 * 4));
    k_restir_dump<<<(R.n + RS_BLOCK - 1) / RS_BLOCK, RS_BLOCK, 0, stream>>>(view_of(R), d);
    cudaError_t e = cudaMemcpyAsync(host_out, d, (size_t)R.n * 40 * 4, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(d);
    return e;
}

}  // namespace rtx
