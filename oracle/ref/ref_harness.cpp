// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under royaltracer-dx_b200/ may include, link or call this.
//
// ref_harness.cpp: the DXR runtime's part in front of the reference's own shader text (oracle/ref/make_ref.py generates the .inc
// files from /root/reference/Pathtracer/shaders/*.hlsl; they are reference text and are never committed).  It
//   * binds the reference's global resources (rdn/Renderer.cpp:953-1008: t0-t6, u0-u7, b0) to caller-owned arrays,
//   * runs RayGen / RayGen2 / RayGen3 one pixel at a time (DispatchRays W x H x 1, rdn/Renderer.cpp:611-673),
//   * implements TraceRay as: ask the ray caster (a callback — tests pass the oracle's orc_trace, since the reference has no source
//     for traversal or the ray/triangle test: that contract is the oracle's), then run the reference's
//     ClosestHit / Miss / ShadowClosestHit / ShadowMiss on the payload, with the per-instance vertex / index buffers bound as the hit
//     group record does (rdn/Renderer.cpp:983-1008,1621-1629).
// Locals the reference leaves uninitialised are zero here (hlsl_shim.h), out-of-bounds buffer reads return zeros.
#include "hlsl_shim.h"

#include <vector>

namespace hlsl {
thread_local DxrState g_dxr;
namespace ref_hit {
#include "hit.inc"
}
namespace ref_miss {
#include "miss.inc"
}
namespace ref_shadow {
#include "shadow.inc"
}
namespace ref_rg1 {
#include "rg1.inc"
}
namespace ref_rg2 {
#include "rg2.inc"
}
namespace ref_rg3 {
#include "rg3.inc"
}
}  // namespace hlsl

// the reference's #defines (PI, bounces, beta, ...) are still active below: this file avoids those identifiers
using namespace hlsl;

namespace {

typedef void (*trace_fn)(void* scene, const void* rays, uint32_t n, void* hits, int any_hit, int mode);
struct HitRec { float t, u, v; uint32_t prim, inst; };
struct RayRec { float o[3], tmin, d[3], tmax; };

struct Env {
    trace_fn trace = nullptr; void* trace_scene = nullptr; int trace_mode = 1;
    std::vector<const void*> verts; std::vector<uint32_t> n_verts; std::vector<const void*> idx; std::vector<uint32_t> n_idx;
    std::vector<uint32_t> inst_model;
    uint64_t closest_rays = 0, shadow_rays = 0;
    uint32_t w = 0, h = 0;
    std::vector<ref_rg1::Reservoir_DI> res_cur, res_last;
    std::vector<ref_rg1::Reservoir_GI> gi_cur, gi_last;
    std::vector<ref_rg1::SampleData> s_cur, s_last;
    std::vector<float4> permanent, output;
} E;

void cast(const RayDesc& ray, int any_hit, HitRec* h) {
    RayRec r = {{ray.Origin.x, ray.Origin.y, ray.Origin.z}, ray.TMin, {ray.Direction.x, ray.Direction.y, ray.Direction.z}, ray.TMax};
    E.trace(E.trace_scene, &r, 1, h, any_hit, E.trace_mode);
}

// hit group 0 / miss 0
void trace_closest(const RayDesc& ray, void* payload) {
    E.closest_rays++;
    HitRec h;
    cast(ray, 0, &h);
    if (h.inst == 0xFFFFFFFFu) {
        ref_miss::Miss(*reinterpret_cast<ref_miss::HitInfo*>(payload));
        return;
    }
    g_dxr.instance_id = h.inst; g_dxr.primitive_index = h.prim; g_dxr.ray_t = h.t;
    g_dxr.world_origin = ray.Origin; g_dxr.world_direction = ray.Direction;
    const uint32_t m = E.inst_model[h.inst];       // the hit group record of instance i binds ITS model's buffers to t2 / t1
    ref_hit::BTriVertex.data = reinterpret_cast<const ref_hit::STriVertex*>(E.verts[m]); ref_hit::BTriVertex.count = E.n_verts[m];
    ref_hit::indices.data = reinterpret_cast<const int*>(E.idx[m]); ref_hit::indices.count = E.n_idx[m];
    ref_hit::Attributes a; a.bary = float2(h.u, h.v);
    ref_hit::ClosestHit(*reinterpret_cast<ref_hit::HitInfo*>(payload), a);
}
// hit group 1 / miss 1
void trace_shadow(const RayDesc& ray, void* payload) {
    E.shadow_rays++;
    HitRec h;
    cast(ray, 1, &h);
    ref_shadow::Attributes a;
    if (h.inst == 0xFFFFFFFFu) ref_shadow::ShadowMiss(*reinterpret_cast<ref_shadow::ShadowHitInfo*>(payload));
    else ref_shadow::ShadowClosestHit(*reinterpret_cast<ref_shadow::ShadowHitInfo*>(payload), a);
}

static_assert(sizeof(ref_rg1::HitInfo) == sizeof(ref_hit::HitInfo) && sizeof(ref_rg1::HitInfo) == sizeof(ref_miss::HitInfo), "payload layouts differ");
static_assert(sizeof(ref_rg1::Material) == 128 && sizeof(ref_rg1::InstanceProperties) == 384 && sizeof(ref_rg1::LightTriangle) == 80 &&
              sizeof(ref_rg1::STriVertex) == 28, "struct layouts differ from the host's (SURVEY.md 8a S1-S7)");

#define BIND_RAYGEN(NS)                                                                                                                    \
    {                                                                                                                                      \
        NS::g_Reservoirs_current.data = reinterpret_cast<NS::Reservoir_DI*>(E.res_cur.data()); NS::g_Reservoirs_current.count = E.res_cur.size();          \
        NS::g_Reservoirs_last.data = reinterpret_cast<NS::Reservoir_DI*>(E.res_last.data()); NS::g_Reservoirs_last.count = E.res_last.size();              \
        NS::g_Reservoirs_current_gi.data = reinterpret_cast<NS::Reservoir_GI*>(E.gi_cur.data()); NS::g_Reservoirs_current_gi.count = E.gi_cur.size();      \
        NS::g_Reservoirs_last_gi.data = reinterpret_cast<NS::Reservoir_GI*>(E.gi_last.data()); NS::g_Reservoirs_last_gi.count = E.gi_last.size();          \
        NS::g_sample_current.data = reinterpret_cast<NS::SampleData*>(E.s_cur.data()); NS::g_sample_current.count = E.s_cur.size();                        \
        NS::g_sample_last.data = reinterpret_cast<NS::SampleData*>(E.s_last.data()); NS::g_sample_last.count = E.s_last.size();                            \
        NS::gPermanentData.data = E.permanent.data(); NS::gPermanentData.w = E.w; NS::gPermanentData.h = E.h;                              \
        NS::gOutput.data = E.output.data(); NS::gOutput.w = E.w; NS::gOutput.h = E.h; NS::gOutput.layers = 1;                              \
    }
#define BIND_SCENE(NS)                                                                                                                     \
    {                                                                                                                                      \
        NS::instanceProps.data = reinterpret_cast<const NS::InstanceProperties*>(props); NS::instanceProps.count = n_inst;                 \
        NS::materialIDs.data = material_ids; NS::materialIDs.count = n_ids;                                                                \
        NS::materials.data = reinterpret_cast<const NS::Material*>(materials); NS::materials.count = n_mat;                                \
        NS::g_EmissiveTriangles.data = reinterpret_cast<const NS::LightTriangle*>(lights); NS::g_EmissiveTriangles.count = n_lights;       \
    }
#define BIND_CAMERA(NS)                                                                                                                    \
    {                                                                                                                                      \
        memcpy(&NS::view, c, 64); memcpy(&NS::projection, c + 16, 64); memcpy(&NS::viewI, c + 32, 64); memcpy(&NS::projectionI, c + 48, 64); \
        memcpy(&NS::prevView, c + 64, 64); memcpy(&NS::prevProjection, c + 80, 64); NS::time = c[96];                                       \
    }

}  // namespace

extern "C" {

void ref_set_tracer(trace_fn fn, void* scene, int mode) {
    E.trace = fn; E.trace_scene = scene; E.trace_mode = mode;
    g_dxr.trace_closest = trace_closest; g_dxr.trace_shadow = trace_shadow;
}

// t1/t2 per model (bound per instance at hit time), t3..t6
void ref_set_scene(uint32_t n_models, const void* const* verts, const uint32_t* n_verts, const void* const* idx, const uint32_t* n_idx,
                   uint32_t n_inst, const uint32_t* inst_model, const void* props, const uint32_t* material_ids, uint32_t n_ids,
                   const void* materials, uint32_t n_mat, const void* lights, uint32_t n_lights) {
    E.verts.assign(verts, verts + n_models); E.n_verts.assign(n_verts, n_verts + n_models);
    E.idx.assign(idx, idx + n_models); E.n_idx.assign(n_idx, n_idx + n_models);
    E.inst_model.assign(inst_model, inst_model + n_inst);
    BIND_SCENE(ref_hit) BIND_SCENE(ref_rg1) BIND_SCENE(ref_rg2) BIND_SCENE(ref_rg3)
}

// b0: the 512-byte CameraParams block (view, projection, viewI, projectionI, prevView, prevProjection, time)
void ref_set_camera(const float* c) { BIND_CAMERA(ref_rg1) BIND_CAMERA(ref_rg2) BIND_CAMERA(ref_rg3) }
// uint(time) is the frame seed (Pass_init_di_v7.hlsl:76-77); the oracle puts the global sample index there (deviation D2)
void ref_set_time(float t) { ref_rg1::time = t; ref_rg2::time = t; ref_rg3::time = t; }

// u0..u7 for a W x H dispatch, zero-filled
void ref_alloc_frame(uint32_t w, uint32_t h) {
    E.w = w; E.h = h;
    // MapPixelID pads the last tile column: allocate for whole 4x4 tiles
    const size_t n = (size_t)((w + 3) / 4) * ((h + 3) / 4) * 16;
    E.res_cur.assign(n, ref_rg1::Reservoir_DI()); E.res_last.assign(n, ref_rg1::Reservoir_DI());
    E.gi_cur.assign(n, ref_rg1::Reservoir_GI()); E.gi_last.assign(n, ref_rg1::Reservoir_GI());
    E.s_cur.assign(n, ref_rg1::SampleData()); E.s_last.assign(n, ref_rg1::SampleData());
    E.permanent.assign((size_t)w * h, float4()); E.output.assign((size_t)w * h, float4());
    BIND_RAYGEN(ref_rg1) BIND_RAYGEN(ref_rg2) BIND_RAYGEN(ref_rg3)
}

// one thread of DispatchRays(W, H, 1) of RayGen (pass 1), RayGen2 (2) or RayGen3 (3)
void ref_raygen(int pass, uint32_t x, uint32_t y) {
    g_dxr.launch_index = uint3(x, y, 0); g_dxr.launch_dims = uint3(E.w, E.h, 1);
    if (pass == 1) ref_rg1::RayGen();
    else if (pass == 2) ref_rg2::RayGen2();
    else ref_rg3::RayGen3();
}
void ref_dispatch(int pass) {
    for (uint32_t y = 0; y < E.h; y++)
        for (uint32_t x = 0; x < E.w; x++) ref_raygen(pass, x, y);
}
void ref_ray_counts(uint64_t* closest, uint64_t* shadow, int reset) {
    *closest = E.closest_rays; *shadow = E.shadow_rays;
    if (reset) E.closest_rays = E.shadow_rays = 0;
}

// per-pixel buffers in the oracle's dump layout (oracle/rtx_oracle.cpp orc_frames_dump), 40 floats per pixel, row-major:
// [0..3] x2,w_sum [4..7] n2,W [8..10] L2 [11] M | [12..15] xn,w_sum [16..19] nn,W [20..22] E3 [23] M |
// [24..26] x1 [27] mID [28..30] n1 [31] objID [32..34] o [35] - [36..38] L1 ; which = 0: *_current, 1: *_last
void ref_frames_dump(int which, float* out) {
    for (uint32_t y = 0; y < E.h; y++)
        for (uint32_t x = 0; x < E.w; x++) {
            const uint32_t k = ref_rg1::MapPixelID(uint2(E.w, E.h), uint2(x, y));
            const ref_rg1::Reservoir_DI& r = which ? E.res_last[k] : E.res_cur[k];
            const ref_rg1::Reservoir_GI& g = which ? E.gi_last[k] : E.gi_cur[k];
            const ref_rg1::SampleData& s = which ? E.s_last[k] : E.s_cur[k];
            float* o = out + 40 * ((size_t)y * E.w + x);
            memset(o, 0, 40 * sizeof(float));
            o[0] = r.x2.x; o[1] = r.x2.y; o[2] = r.x2.z; o[3] = r.w_sum; o[4] = r.n2.x; o[5] = r.n2.y; o[6] = r.n2.z; o[7] = r.W;
            o[8] = r.L2.x; o[9] = r.L2.y; o[10] = r.L2.z; o[11] = (float)r.M;
            o[12] = g.xn.x; o[13] = g.xn.y; o[14] = g.xn.z; o[15] = g.w_sum; o[16] = g.nn.x; o[17] = g.nn.y; o[18] = g.nn.z; o[19] = g.W;
            o[20] = g.E3.x; o[21] = g.E3.y; o[22] = g.E3.z; o[23] = (float)g.M;
            o[24] = s.x1.x; o[25] = s.x1.y; o[26] = s.x1.z; o[27] = (float)s.mID; o[28] = s.n1.x; o[29] = s.n1.y; o[30] = s.n1.z;
            o[31] = (float)s.objID; o[32] = s.o.x; o[33] = s.o.y; o[34] = s.o.z; o[36] = s.L1.x; o[37] = s.L1.y; o[38] = s.L1.z;
        }
}
// gPermanentData (u1): float4 per pixel, row-major; gOutput slice 0 as float4 in [0,1] (the UNORM store is the caller's)
void ref_read_permanent(float* out) { memcpy(out, E.permanent.data(), E.permanent.size() * 16); }
void ref_write_permanent(const float* in) { memcpy(E.permanent.data(), in, E.permanent.size() * 16); }
void ref_read_output(float* out) { memcpy(out, E.output.data(), E.output.size() * 16); }

// ---- known-answer entry points on the reference's leaf functions (each call = the reference's own function body)
void ref_kat_rng(uint32_t sx, uint32_t sy, uint32_t n, float* out, uint32_t* seed_out) {        // Common_v7.hlsl:119-138
    uint2 seed(sx, sy);
    for (uint32_t i = 0; i < n; i++) out[i] = ref_rg1::RandomFloat(seed);
    seed_out[0] = seed.x; seed_out[1] = seed.y;
}
uint32_t ref_kat_map_pixel(uint32_t w, uint32_t h, uint32_t x, uint32_t y) { return ref_rg1::MapPixelID(uint2(w, h), uint2(x, y)); }   // :173-198
void ref_kat_srgb(const float* rgb, uint32_t n, float* out) {                                     // Common_v7.hlsl:353-376
    for (uint32_t i = 0; i < n; i++) {
        const float3 r = ref_rg1::sRGBGammaCorrection(float3(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]));
        out[3 * i] = r.x; out[3 * i + 1] = r.y; out[3 * i + 2] = r.z;
    }
}
// op: 0 EvaluateBRDF(0) 1 EvaluateBRDF(1) 2 BRDF_PDF(0) 3 BRDF_PDF(1) 4 CalculateStrategyProbabilities
//     5 SampleBRDF(strategy 0) 6 SampleBRDF(strategy 1) 7 SelectSamplingStrategy (out[0] = strategy, out[1] = probability)
// the material is read like the raygen does (MaterialOptimized from materials[mat_id], Pass_init_di_v7.hlsl:109-112)
void ref_kat_bsdf(int op, uint32_t mat_id, const float* n, const float* in, const float* o, uint32_t* seed_io, float* out4) {
    using namespace ref_rg1;
    MaterialOptimized m = {materials[mat_id].Kd, materials[mat_id].Pr_Pm_Ps_Pc, materials[mat_id].Ks, materials[mat_id].Ke, mat_id};
    const float3 N(n[0], n[1], n[2]), I(in[0], in[1], in[2]), O(o[0], o[1], o[2]);
    uint2 seed(seed_io[0], seed_io[1]);
    float3 r(0, 0, 0); float w = 0;
    if (op == 0 || op == 1) r = EvaluateBRDF(op, m, N, I, O);
    else if (op == 2 || op == 3) r.x = BRDF_PDF(op - 2, m, N, I, O);
    else if (op == 4) { const float2 p = CalculateStrategyProbabilities(m, O, N); r.x = p.x; r.y = p.y; }
    else if (op == 5 || op == 6) { float3 s, org; SampleBRDF(op - 5, m, O, N, N, s, org, float3(0, 0, 0), seed); r = s; }
    else if (op == 7) { float prob = 0; r.x = (float)SelectSamplingStrategy(m, O, N, seed, prob); r.y = prob; }
    out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = w;
    seed_io[0] = seed.x; seed_io[1] = seed.y;
}
// UpdateReservoir / UpdateReservoir_GI (Reservoir_v7.hlsl:30-80): state = (w_sum, M), returns accepted
int ref_kat_update_reservoir(int gi, float* w_sum, float* M, float wi, float M_in, uint32_t* seed_io) {
    using namespace ref_rg1;
    uint2 seed(seed_io[0], seed_io[1]);
    bool acc;
    if (gi) { Reservoir_GI r = {float3(0, 0, 0), *w_sum, float3(0, 0, 0), 0.0f, half3(0, 0, 0), (uint16_t)*M};
              acc = UpdateReservoir_GI(r, wi, M_in, float3(1, 2, 3), float3(0, 1, 0), float3(1, 1, 1), seed); *w_sum = r.w_sum; *M = (float)r.M; }
    else { Reservoir_DI r = {float3(0, 0, 0), *w_sum, float3(0, 0, 0), 0.0f, half3(0, 0, 0), (uint16_t)*M};
           acc = UpdateReservoir(r, wi, M_in, float3(1, 2, 3), float3(0, 1, 0), float3(1, 1, 1), seed); *w_sum = r.w_sum; *M = (float)r.M; }
    seed_io[0] = seed.x; seed_io[1] = seed.y;
    return acc ? 1 : 0;
}
// Estimator E0 = what RayGen3 shades when no temporal or spatial candidate is accepted: pass 1's reservoirs combined
// exactly as Pass_spat_di_v7.hlsl:64,96-102 (matOpt from sdata.mID) and :334-372 do — every operation below is a call into the
// reference's own functions; W_DI / W_GI are the ones RayGen stored (Pass_init_di_v7.hlsl:166-181).  Emitter pixels return L1.
// out: float4 per pixel, row-major (C.rgb, 1 = sampled / 3 = emitter)
void ref_e0(float* out) {
    using namespace ref_rg3;
    for (uint32_t y = 0; y < E.h; y++)
        for (uint32_t x = 0; x < E.w; x++) {
            const uint32_t pixelIdx = MapPixelID(uint2(E.w, E.h), uint2(x, y));
            SampleData sdata_current = g_sample_current[pixelIdx];
            float* o = out + 4 * ((size_t)y * E.w + x);
            if (sdata_current.L1.x == 0.0f && sdata_current.L1.y == 0.0f && sdata_current.L1.z == 0.0f) {
                uint mID = sdata_current.mID;
                MaterialOptimized matOpt = {materials[mID].Kd, materials[mID].Pr_Pm_Ps_Pc, materials[mID].Ks, materials[mID].Ke, mID};
                Reservoir_DI reservoir_current = g_Reservoirs_current[pixelIdx];
                Reservoir_GI reservoir_current_gi = g_Reservoirs_current_gi[pixelIdx];
                float3 accumulation = ReconnectDI(sdata_current.x1, sdata_current.n1, reservoir_current.x2, reservoir_current.n2, reservoir_current.L2,
                                                  sdata_current.o, matOpt) * reservoir_current.W;
                float3 f_gi_final = GetP_Hat_GI(sdata_current.x1, sdata_current.n1, reservoir_current_gi.xn, reservoir_current_gi.nn,
                                                reservoir_current_gi.E3, sdata_current.o, matOpt, false);
                accumulation += f_gi_final * reservoir_current_gi.W;
                o[0] = accumulation.x; o[1] = accumulation.y; o[2] = accumulation.z; o[3] = 1.0f;
            } else {
                o[0] = sdata_current.L1.x; o[1] = sdata_current.L1.y; o[2] = sdata_current.L1.z; o[3] = 3.0f;
            }
        }
}

// the compile-time path configuration this library was generated with (Common_v7.hlsl:8-11)
void ref_config(uint32_t* out4) { out4[0] = bounces; out4[1] = nee_samples; out4[2] = nee_samples_DI; out4[3] = bsdf_samples_DI; }

}  // extern "C"
