// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under royaltracer-dx_b200/ may include, link or call this.
//
// ref_legacy_harness.cpp: the DXR runtime's part in front of the reference's FIRST estimator: include/RayGen.hlsl
// (path loop with Russian roulette + accumulation), include/Hit.hlsl (ClosestHit: RIS over RIS_M light candidates, one shadow ray, BSDF
// sample, MIS on emitter hits), include/Miss.hlsl and include/ShadowRay.hlsl, generated into oracle/_ref/gen_legacy/ by
// oracle/ref/make_ref.py --legacy (reference text, never committed).  Same division of labour as ref_harness.cpp: TraceRay asks the
// oracle's ray caster, then runs the reference's own hit / miss shaders on the payload; here ClosestHit itself calls TraceRay for its
// shadow ray (hit group 1 / miss 1), so the DXR system values of the outer hit stay in place across the nested call.
// The entry points carry the names of ref_harness.cpp's so that oracle/ref/ref.py drives both libraries.
#include "hlsl_shim.h"

#include <vector>

namespace hlsl {
thread_local DxrState g_dxr;
namespace leg_hit {
#include "leg_hit.inc"
}
namespace leg_miss {
#include "leg_miss.inc"
}
namespace ref_shadow {
#include "shadow.inc"
}
namespace leg_rg {
#include "leg_rg.inc"
}
}  // namespace hlsl

// the reference's #defines (PI, EPSILON, s_bias, RIS_M ...) are still active below: this file avoids those identifiers
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
        leg_miss::Miss(*reinterpret_cast<leg_miss::HitInfo*>(payload));
        return;
    }
    g_dxr.instance_id = h.inst; g_dxr.primitive_index = h.prim; g_dxr.ray_t = h.t;
    g_dxr.world_origin = ray.Origin; g_dxr.world_direction = ray.Direction;
    const uint32_t m = E.inst_model[h.inst];       // the hit group record of instance i binds ITS model's buffers to t2 / t1
    leg_hit::BTriVertex.data = reinterpret_cast<const leg_hit::STriVertex*>(E.verts[m]); leg_hit::BTriVertex.count = E.n_verts[m];
    leg_hit::indices.data = reinterpret_cast<const int*>(E.idx[m]); leg_hit::indices.count = E.n_idx[m];
    leg_hit::Attributes a; a.bary = float2(h.u, h.v);
    leg_hit::ClosestHit(*reinterpret_cast<leg_hit::HitInfo*>(payload), a);
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

static_assert(sizeof(leg_rg::HitInfo) == sizeof(leg_hit::HitInfo) && sizeof(leg_rg::HitInfo) == sizeof(leg_miss::HitInfo), "payload layouts differ");
static_assert(sizeof(leg_hit::ShadowHitInfo) == sizeof(ref_shadow::ShadowHitInfo), "shadow payload layouts differ");
static_assert(sizeof(leg_hit::STriVertex) == 28, "vertex layout differs from the host's (SURVEY.md 8a S1)");
// The legacy text declares OLDER layouts of three host structures than the ones rdn/Renderer.cpp uploads today (the layouts of
// shaders/Common_v7.hlsl:53-97): a 196-byte Material (include/Common.hlsl:31-40: Ks, Ke unpadded, aniso_anisor, Ni, LUT[32]), a
// 256-byte InstanceProperties (include/Hit.hlsl:8-14: no inverse matrices) and a LightTriangle with another field order
// (include/Hit.hlsl:16-29).  The harness fills the legacy structs FIELD BY FIELD from the host's data (fields of the same name;
// LUT[16..31], aniso_anisor and the pads are zero and never read by this path) — the same reading the oracle's `legacy` namespace uses.
static_assert(sizeof(leg_hit::Material) == 196 && sizeof(leg_hit::InstanceProperties) == 256 && sizeof(leg_hit::LightTriangle) == 80,
              "include/Common.hlsl / include/Hit.hlsl structures changed");
struct HostMaterial { float Kd[4]; float Ks[3]; float Ni; float Ke[3]; float pad0; float Pr_Pm_Ps_Pc[4]; float LUT[16]; };
struct HostInstanceProps { float objectToWorld[16], objectToWorldInverse[16], prevObjectToWorld[16], prevObjectToWorldInverse[16],
                                 objectToWorldNormal[16], prevObjectToWorldNormal[16]; };
struct HostLight { float x[3]; float cdf; float y[3]; uint32_t instanceID; float z[3]; float weight; float emission[3]; uint32_t triCount;
                   float total_weight; float pad0[3]; };
static_assert(sizeof(HostMaterial) == 128 && sizeof(HostInstanceProps) == 384 && sizeof(HostLight) == 80, "host layouts (SURVEY.md 8a S4, S6, S7)");
std::vector<leg_hit::Material> g_materials;
std::vector<leg_hit::InstanceProperties> g_props;
std::vector<leg_hit::LightTriangle> g_lights;
const HostInstanceProps* g_host_props = nullptr; uint32_t g_n_inst = 0;
void convert_props() {          // the tests update the host's array in place between frames (moving instances)
    g_props.assign(g_n_inst, leg_hit::InstanceProperties());
    for (uint32_t i = 0; i < g_n_inst; i++) {
        memcpy(&g_props[i].objectToWorld, g_host_props[i].objectToWorld, 64);
        memcpy(&g_props[i].prevObjectToWorld, g_host_props[i].prevObjectToWorld, 64);
        memcpy(&g_props[i].objectToWorldNormal, g_host_props[i].objectToWorldNormal, 64);
        memcpy(&g_props[i].prevObjectToWorldNormal, g_host_props[i].prevObjectToWorldNormal, 64);
    }
    leg_hit::instanceProps.data = g_props.data(); leg_hit::instanceProps.count = g_n_inst;
}

}  // namespace

extern "C" {

void ref_set_tracer(trace_fn fn, void* scene, int mode) {
    E.trace = fn; E.trace_scene = scene; E.trace_mode = mode;
    g_dxr.trace_closest = trace_closest; g_dxr.trace_shadow = trace_shadow;
}

void ref_set_scene(uint32_t n_models, const void* const* verts, const uint32_t* n_verts, const void* const* idx, const uint32_t* n_idx,
                   uint32_t n_inst, const uint32_t* inst_model, const void* props, const uint32_t* material_ids, uint32_t n_ids,
                   const void* materials, uint32_t n_mat, const void* lights, uint32_t n_lights) {
    E.verts.assign(verts, verts + n_models); E.n_verts.assign(n_verts, n_verts + n_models);
    E.idx.assign(idx, idx + n_models); E.n_idx.assign(n_idx, n_idx + n_models);
    E.inst_model.assign(inst_model, inst_model + n_inst);
    g_host_props = reinterpret_cast<const HostInstanceProps*>(props); g_n_inst = n_inst;
    convert_props();
    leg_hit::materialIDs.data = material_ids; leg_hit::materialIDs.count = n_ids;
    g_materials.assign(n_mat, leg_hit::Material());
    for (uint32_t i = 0; i < n_mat; i++) {
        const HostMaterial& h = reinterpret_cast<const HostMaterial*>(materials)[i];
        leg_hit::Material& m = g_materials[i];
        m.Kd = float4(h.Kd[0], h.Kd[1], h.Kd[2], h.Kd[3]); m.Ks = float3(h.Ks[0], h.Ks[1], h.Ks[2]); m.Ke = float3(h.Ke[0], h.Ke[1], h.Ke[2]);
        m.Pr_Pm_Ps_Pc = float4(h.Pr_Pm_Ps_Pc[0], h.Pr_Pm_Ps_Pc[1], h.Pr_Pm_Ps_Pc[2], h.Pr_Pm_Ps_Pc[3]);
        m.aniso_anisor = float2(0.0f, 0.0f); m.Ni = h.Ni;
        for (int k = 0; k < 32; k++) m.LUT[k] = k < 16 ? h.LUT[k] : 0.0f;
    }
    leg_hit::materials.data = g_materials.data(); leg_hit::materials.count = n_mat;
    g_lights.assign(n_lights, leg_hit::LightTriangle());
    for (uint32_t i = 0; i < n_lights; i++) {
        const HostLight& h = reinterpret_cast<const HostLight*>(lights)[i];
        leg_hit::LightTriangle& l = g_lights[i];
        l.x = float3(h.x[0], h.x[1], h.x[2]); l.y = float3(h.y[0], h.y[1], h.y[2]); l.z = float3(h.z[0], h.z[1], h.z[2]);
        l.pad0 = l.pad1 = l.pad2 = 0.0f;
        l.instanceID = h.instanceID; l.weight = h.weight; l.triCount = h.triCount; l.total_weight = h.total_weight;
        l.emission = float3(h.emission[0], h.emission[1], h.emission[2]); l.cdf = h.cdf;
    }
    leg_hit::g_EmissiveTriangles.data = g_lights.data(); leg_hit::g_EmissiveTriangles.count = n_lights;
}

// b0: the 512-byte CameraParams block (view, projection, viewI, projectionI, prevView, prevProjection, time)
void ref_set_camera(const float* c) {
    memcpy(&leg_rg::view, c, 64); memcpy(&leg_rg::projection, c + 16, 64); memcpy(&leg_rg::viewI, c + 32, 64);
    memcpy(&leg_rg::projectionI, c + 48, 64); memcpy(&leg_rg::prevView, c + 64, 64); memcpy(&leg_rg::prevProjection, c + 80, 64);
    leg_rg::time = c[96];
}
// uint(time) is the frame seed (RayGen.hlsl:83-84); the oracle puts the global sample index there (deviation D2)
void ref_set_time(float t) { leg_rg::time = t; }

void ref_alloc_frame(uint32_t w, uint32_t h) {
    E.w = w; E.h = h;
    E.permanent.assign((size_t)w * h, float4()); E.output.assign((size_t)w * h, float4());
    leg_rg::gPermanentData.data = E.permanent.data(); leg_rg::gPermanentData.w = w; leg_rg::gPermanentData.h = h;
    leg_rg::gOutput.data = E.output.data(); leg_rg::gOutput.w = w; leg_rg::gOutput.h = h; leg_rg::gOutput.layers = 1;
}

// one thread of DispatchRays(W, H, 1) of RayGen (rdn/Renderer.cpp: the legacy pipeline has this one ray-generation shader)
void ref_raygen(int pass, uint32_t x, uint32_t y) {
    (void)pass;
    g_dxr.launch_index = uint3(x, y, 0); g_dxr.launch_dims = uint3(E.w, E.h, 1);
    leg_rg::RayGen();
}
void ref_dispatch(int pass) {
    convert_props();
    for (uint32_t y = 0; y < E.h; y++)
        for (uint32_t x = 0; x < E.w; x++) ref_raygen(pass, x, y);
}
void ref_ray_counts(uint64_t* closest, uint64_t* shadow, int reset) {
    *closest = E.closest_rays; *shadow = E.shadow_rays;
    if (reset) E.closest_rays = E.shadow_rays = 0;
}
void ref_read_permanent(float* out) { memcpy(out, E.permanent.data(), E.permanent.size() * 16); }
void ref_write_permanent(const float* in) { memcpy(E.permanent.data(), in, E.permanent.size() * 16); }
void ref_read_output(float* out) { memcpy(out, E.output.data(), E.output.size() * 16); }

// RIS_M, and the path-length cap the library was generated with (RayGen.hlsl:63; 10000000 in the reference)
void ref_config(uint32_t* out4) { out4[0] = RIS_M; out4[1] = 0; out4[2] = 0; out4[3] = 0; }

}  // extern "C"
