/* ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * C interface of the single-threaded CPU restatement of the reference's hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  Parity status: PINNED to the reference's own shader text — oracle/_ref/libref.so
 * is the HLSL of Pathtracer/shaders compiled for the CPU (oracle/ref/make_ref.py) and tests/test_ref_pins.py
 * demands bit-identical results from this library for RayGen/RayGen2/RayGen3 and their leaf
 * functions.  Unpinned (no reference source or output exists): which triangle TraceRay returns —
 * that contract is defined here (orc_trace mode 0).
 */
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Byte-compatible with the reference's GPU structs. */
typedef struct { float position[3]; float normal_material[4]; } orc_vertex;            /* S1, 28 B  src/Components/Vertex.h:25-35 */
typedef struct { float Kd[4]; float Ks[3]; float Ni; float Ke[3]; float pad0;
                 float Pr_Pm_Ps_Pc[4]; float LUT[16]; } orc_material;                     /* S4, 128 B Vertex.h:14-23 */
typedef struct { float objectToWorld[16], objectToWorldInverse[16], prevObjectToWorld[16],
                 prevObjectToWorldInverse[16], objectToWorldNormal[16],
                 prevObjectToWorldNormal[16]; } orc_instance_props;                       /* S6, 384 B Renderer.h:275-282 */
typedef struct { float x[3]; float cdf; float y[3]; uint32_t instanceID; float z[3]; float weight;
                 float emission[3]; uint32_t triCount; float total_weight; float pad0[3]; } orc_light_tri; /* S7, 80 B Renderer.h:113-124 */
typedef struct { float view[16], projection[16], viewI[16], projectionI[16], prevView[16],
                 prevProjection[16]; float time; float pad[31]; } orc_camera;             /* S8, 512 B Pass_init_di_v7.hlsl:36-45 */

typedef struct { float origin[3]; float tmin; float direction[3]; float tmax; } orc_ray;   /* RayDesc */
typedef struct { float t, u, v; uint32_t prim; uint32_t inst; } orc_hit;                   /* inst = 0xFFFFFFFF on miss */

#define ORC_FLAG_JITTER        1u   /* 2 RandomFloat draws before anything else (legacy include/RayGen.hlsl:84-85) */
#define ORC_FLAG_LAMBERT_ONLY  2u   /* strategy probabilities forced to (1,0) */
#define ORC_FLAG_LEGACY_RR     16u  /* orc_render runs the legacy estimator (include/RayGen.hlsl + include/Hit.hlsl: RIS-10 NEE with one shadow
                                       ray per bounce, MIS on emitter hits, Russian roulette after depth 3); cfg.bounces caps the path length */

typedef struct {
    uint32_t width, height;
    uint32_t bounces;          /* Common_v7.hlsl:11  (default 3) */
    uint32_t nee_samples;      /* Common_v7.hlsl:8   (default 4) */
    uint32_t nee_samples_di;   /* Common_v7.hlsl:9   (default 4) */
    uint32_t flags;
} orc_config;

typedef struct {
    uint64_t closest_rays, shadow_rays;     /* TraceRay-equivalents actually issued */
    uint64_t paths;
    uint64_t bvh_nodes_visited, tris_tested, instances_entered;   /* oracle's own BVH2 — diagnostics only */
} orc_counters;

typedef struct orc_scene orc_scene;

orc_scene* orc_scene_create(void);
void       orc_scene_destroy(orc_scene*);
/* T1: one geometry per model; returns the model id. material_id_offset replaces the float smuggled in
 * vertex.normal.w (Hit_v7.hlsl:16-17). */
int  orc_add_model(orc_scene*, const orc_vertex* v, uint32_t nv, const uint32_t* idx, uint32_t ni,
                   uint32_t material_id_offset);
void orc_set_material_ids(orc_scene*, const uint32_t* ids, uint32_t n);      /* S3 */
void orc_set_materials(orc_scene*, const orc_material* m, uint32_t n);       /* S4 */
void orc_set_instances(orc_scene*, const uint32_t* model_ids, const orc_instance_props* props, uint32_t n); /* T2,S6 */
void orc_set_lights(orc_scene*, const orc_light_tri* l, uint32_t n);         /* S7 */
void orc_build(orc_scene*);                                                   /* builds the oracle's BVH2s */

/* T3/T4. mode 0 = brute force over every triangle of every instance (the definition),
 *        mode 1 = through the oracle's BVH2 (must agree with mode 0). */
void orc_trace(orc_scene*, const orc_ray* rays, uint32_t n, orc_hit* out, int any_hit, int mode);

/* E0 estimator + accumulation F20.  Renders samples [first_sample, first_sample+n)
 * of every pixel with x % step == 0 && y % step == 0 into accum (float4 per pixel, row-major W*H). */
void orc_render(orc_scene*, const orc_config*, const orc_camera*, uint32_t first_sample, uint32_t n_samples,
                uint32_t step, float* accum, orc_counters* counters, int trace_mode);
/* same, rows [y0, y1) only; disjoint row ranges may run concurrently on several host threads */
void orc_render_rows(orc_scene*, const orc_config*, const orc_camera*, uint32_t first_sample, uint32_t n_samples,
                     uint32_t step, uint32_t y0, uint32_t y1, float* accum, orc_counters* counters, int trace_mode);
/* The reference's whole frame: pass 1 (RayGen) + temporal reuse (RayGen2) + spatial reuse, final
 * shade and accumulation (RayGen3).  orc_frames holds the per-pixel reservoir / sample buffers u2..u7 (current and last). */
typedef struct orc_frames orc_frames;
orc_frames* orc_frames_create(uint32_t width, uint32_t height);
void        orc_frames_destroy(orc_frames*);
void orc_render_frame(orc_scene*, const orc_config*, const orc_camera*, uint32_t frame_index, orc_frames*, float* accum,
                      orc_counters* counters, int trace_mode);
void orc_frames_dump(const orc_frames*, float* out40_per_pixel);   /* the *_last buffers, layout in rtx_oracle.cpp */
/* F20 output: sRGB(sum/n) -> RGBA8 */
void orc_resolve(const float* accum, uint32_t n_pixels, uint8_t* rgba8);

/* Known-answer hooks for unit tests (each returns through out[]). */
void orc_kat_rng(uint32_t sx, uint32_t sy, uint32_t n, float* out, uint32_t* seed_out);
void orc_kat_seed(uint32_t x, uint32_t y, uint32_t pass, uint32_t sample, uint32_t* seed_out);
void orc_kat_sincos(const float* x, uint32_t n, float* s, float* c);
void orc_kat_half(const float* x, uint32_t n, float* out);
void orc_kat_pow(const float* x, float y, uint32_t n, float* out);
uint32_t orc_kat_map_pixel(uint32_t w, uint32_t h, uint32_t x, uint32_t y);
/* op: 0 EvaluateBRDF(0) 1 EvaluateBRDF(1) 2 BRDF_PDF(0) 3 BRDF_PDF(1) 4 strategy probs
 *     5 SampleBRDF(strategy=flag&1) with seed -> dir, seed' ; out has 4 floats */
void orc_kat_bsdf(const orc_scene*, int op, uint32_t mat_id, const float* n, const float* in, const float* o,
                  uint32_t* seed, float* out);
void orc_kat_camera_ray(const orc_config*, const orc_camera*, uint32_t x, uint32_t y, float jx, float jy, float* o3d3);
/* per-pixel debug record of one sample: see OrcPixelDebug in rtx_oracle.cpp */
void orc_debug_pixel(orc_scene*, const orc_config*, const orc_camera*, uint32_t x, uint32_t y, uint32_t sample,
                     float* out64, int trace_mode);

#ifdef __cplusplus
}
#endif
