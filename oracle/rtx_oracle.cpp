// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under royaltracer-dx_b200/ may include, link or call this.
//
// Single-threaded CPU restatement of the reference's hot path (ML200/RoyalTracer-DX):
//   camera ray -> TraceRay -> ClosestHit/Miss -> RIS direct lighting -> SamplePathSimple (GI) ->
//   estimator E0 -> accumulation (F20).
// Every function cites the reference file:line it follows (paths relative to
// /root/reference/Pathtracer/).  Numerics contract: oracle/det_math.h.
//
// PARITY STATUS: pinned to the reference's own shader text.  oracle/_ref/libref.so = Pathtracer/shaders/*.hlsl compiled for the CPU
// (oracle/ref/make_ref.py + hlsl_shim.h); tests/test_ref_pins.py runs RayGen / RayGen2 / RayGen3 and the leaf functions from it and
// demands bit-identical results from the functions below.  The traversal + ray/triangle test has no reference source at all (closed
// D3D12 driver / RT cores): that contract is defined HERE (orc_trace mode 0 = brute force) from the DXR semantics the reference's
// call sites rely on and is the one part that remains "parity unpinned".
//
// Deliberate deviations from reference undefined behaviour (DESIGN.md §"Deviations"):
//   D1 miss => path terminates, radiance 0 (ref: Miss_v7.hlsl:3-8 leaves the payload uninitialised)
//   D2 seed uses the global sample index where the ref uses uint(time) (Pass_init_di_v7.hlsl:76-77)
//   D3 optional jitter: 2 draws before anything else (legacy include/RayGen.hlsl:84-85)
//   D4 emitter hit by a GI BSDF ray with zero/NaN contribution => path terminates
//      (ref: Path_Sampler_v7.hlsl:262-268 continues with uninitialised new_origin/new_normal)
//   D5 primary-hit emitters are accumulated like any other sample (ref bypasses accumulation,
//      Pass_spat_di_v7.hlsl:458-463; identical when every sample of the pixel hits the emitter)
//   D6 the DI visibility ray is skipped when its result cannot matter (!(f_g > EPSILON))
//   D7 material id offset is a per-model uint instead of a float in vertex.normal.w
#include "rtx_oracle.h"
#include "det_math.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace orc;

namespace {

// shaders/Common_v7.hlsl:1-3
constexpr float PI_REF = 3.1415f;
constexpr float S_BIAS = 0.00002f;
constexpr float EPS = 0.000001f;
constexpr uint32_t MISS_ID = 4294967294u;   // shaders/Miss_v7.hlsl:7

// ------------------------------------------------------------------------------------------ scene
struct Box { f3 lo, hi; };
struct Node2 { Box b; uint32_t left, right; uint32_t first, count; uint32_t axis; };   // count>0 => leaf; axis: the split axis (left = the lower side)

struct Model {
    std::vector<f3> pos;
    std::vector<f3> nrm;
    std::vector<uint32_t> idx;
    uint32_t mat_offset = 0;
    std::vector<Node2> nodes;
    std::vector<uint32_t> order;    // triangle ids in leaf order
    Box bounds;
};

struct Instance { uint32_t model; orc_instance_props p; Box wbox; };

}  // namespace

struct orc_scene {
    std::vector<Model> models;
    std::vector<uint32_t> material_ids;
    std::vector<orc_material> materials;
    std::vector<Instance> instances;
    std::vector<orc_light_tri> lights;
    std::vector<Node2> tlas; std::vector<uint32_t> tlas_order;     // BVH2 over the instances' world boxes (orc_build)
    orc_counters* ctr = nullptr;
};

namespace {

orc_counters g_dummy_ctr;

// ------------------------------------------------------------------------------------------ ray / triangle
// Contract T3/T4: hit iff TMin < t < TMax; closest = lexicographic min of
// (t, instance, primitive).  Watertight edge-function test (shear the triangle into ray space,
// exact-zero edge functions re-evaluated in binary64).
struct RayPre { f3 o, d; int kx, ky, kz; float Sx, Sy, Sz; };

inline float comp(f3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }

RayPre ray_pre(f3 o, f3 d) {
    RayPre r; r.o = o; r.d = d;
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    int kz = 0; float m = ax;
    if (ay > m) { kz = 1; m = ay; }
    if (az > m) { kz = 2; }
    int kx = (kz + 1) % 3, ky = (kx + 1) % 3;
    if (comp(d, kz) < 0.0f) std::swap(kx, ky);
    r.kx = kx; r.ky = ky; r.kz = kz;
    float dz = comp(d, kz);
    r.Sz = 1.0f / dz; r.Sx = comp(d, kx) * r.Sz; r.Sy = comp(d, ky) * r.Sz;
    return r;
}

bool tri_test(const RayPre& r, f3 v0, f3 v1, f3 v2, float tmin, float tmax, float& t, float& b1, float& b2) {
    f3 A = v0 - r.o, B = v1 - r.o, C = v2 - r.o;
    float Akz = comp(A, r.kz), Bkz = comp(B, r.kz), Ckz = comp(C, r.kz);
    float Ax = comp(A, r.kx) - r.Sx * Akz, Ay = comp(A, r.ky) - r.Sy * Akz;
    float Bx = comp(B, r.kx) - r.Sx * Bkz, By = comp(B, r.ky) - r.Sy * Bkz;
    float Cx = comp(C, r.kx) - r.Sx * Ckz, Cy = comp(C, r.ky) - r.Sy * Ckz;
    float U = Cx * By - Cy * Bx;
    float V = Ax * Cy - Ay * Cx;
    float W = Bx * Ay - By * Ax;
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
        V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
        W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    float det = (U + V) + W;
    if (det == 0.0f) return false;
    float Az = r.Sz * Akz, Bz = r.Sz * Bkz, Cz = r.Sz * Ckz;
    float T = (U * Az + V * Bz) + W * Cz;
    float rcp = 1.0f / det;
    float tt = T * rcp;
    if (!(tt > tmin && tt < tmax)) return false;
    t = tt; b1 = V * rcp; b2 = W * rcp;
    return true;
}

// conservative slab test (NaN-ignoring fminf/fmaxf; inflated far side)
inline bool box_test(const Box& b, f3 o, f3 inv, float tmin, float tmax) {
    float t0 = (b.lo.x - o.x) * inv.x, t1 = (b.hi.x - o.x) * inv.x;
    float tn = fminf(t0, t1), tf = fmaxf(t0, t1);
    t0 = (b.lo.y - o.y) * inv.y; t1 = (b.hi.y - o.y) * inv.y;
    tn = fmaxf(tn, fminf(t0, t1)); tf = fminf(tf, fmaxf(t0, t1));
    t0 = (b.lo.z - o.z) * inv.z; t1 = (b.hi.z - o.z) * inv.z;
    tn = fmaxf(tn, fminf(t0, t1)); tf = fminf(tf, fmaxf(t0, t1));
    tn = fmaxf(tn, tmin); tf = fminf(tf, tmax);
    return tn <= tf * 1.0000004f + 1e-30f;
}

inline Box box_empty() { Box b; b.lo = mk3(INFINITY, INFINITY, INFINITY); b.hi = mk3(-INFINITY, -INFINITY, -INFINITY); return b; }
inline void box_grow(Box& b, f3 p) {
    b.lo = mk3(fminf(b.lo.x, p.x), fminf(b.lo.y, p.y), fminf(b.lo.z, p.z));
    b.hi = mk3(fmaxf(b.hi.x, p.x), fmaxf(b.hi.y, p.y), fmaxf(b.hi.z, p.z));
}
inline void box_pad(Box& b, float pad) {
    b.lo = mk3(b.lo.x - pad, b.lo.y - pad, b.lo.z - pad);
    b.hi = mk3(b.hi.x + pad, b.hi.y + pad, b.hi.z + pad);
}

// Top-down BVH2 over primitive boxes: median split of the widest axis, or (sah) the binned surface-area heuristic (16 bins per axis, all
// three axes) with the median split where the bins cannot separate anything.  Children are visited nearer side first.  The oracle's hits do not depend on the tree (closest = lexicographic minimum of
// (t, instance, primitive) over everything the conservative boxes let through; tests compare with the brute-force mode); a decent tree
// only makes the CPU arm of bench.py a fairer baseline and the parity tests at full size faster.
inline float box_half_area(const Box& b) {
    const float dx = b.hi.x - b.lo.x, dy = b.hi.y - b.lo.y, dz = b.hi.z - b.lo.z;
    return (dx * dy + dy * dz) + dz * dx;
}
void build_tree(const std::vector<Box>& tb, const std::vector<f3>& cen, float pad, uint32_t leaf_max, bool sah, std::vector<uint32_t>& order,
                std::vector<Node2>& nodes) {
    const uint32_t nt = (uint32_t)tb.size();
    order.resize(nt);
    for (uint32_t t = 0; t < nt; t++) order[t] = t;
    nodes.clear();
    if (nt == 0) return;
    nodes.reserve(2 * nt);
    struct Job { uint32_t node, first, count; };
    std::vector<Job> stack;
    nodes.push_back(Node2{});
    stack.push_back({0u, 0u, nt});
    const int NB = 16;
    while (!stack.empty()) {
        Job j = stack.back(); stack.pop_back();
        Box b = box_empty(), cb = box_empty();
        for (uint32_t i = j.first; i < j.first + j.count; i++) {
            uint32_t t = order[i];
            box_grow(b, tb[t].lo); box_grow(b, tb[t].hi); box_grow(cb, cen[t]);
        }
        box_pad(b, pad);
        nodes[j.node].b = b; nodes[j.node].axis = 0;
        f3 ext = cb.hi - cb.lo;
        int wide = 0; if (ext.y > ext.x) wide = 1; if (ext.z > comp(ext, wide)) wide = 2;
        if (j.count <= leaf_max || !(comp(ext, wide) > 0.0f)) {
            Node2& n = nodes[j.node];
            n.first = j.first; n.count = j.count; n.left = n.right = 0;
            continue;
        }
        // binned SAH
        int best_axis = -1, best_split = 0; float best_cost = INFINITY;
        for (int axis = 0; sah && axis < 3; axis++) {
            const float lo = comp(cb.lo, axis), e = comp(ext, axis);
            if (!(e > 0.0f)) continue;
            const float k = (float)NB / e;
            uint32_t cnt[NB] = {0}; Box bb[NB];
            for (int q = 0; q < NB; q++) bb[q] = box_empty();
            for (uint32_t i = j.first; i < j.first + j.count; i++) {
                const uint32_t t = order[i];
                int q = (int)((comp(cen[t], axis) - lo) * k); q = q < 0 ? 0 : (q >= NB ? NB - 1 : q);
                cnt[q]++; box_grow(bb[q], tb[t].lo); box_grow(bb[q], tb[t].hi);
            }
            float la[NB]; uint32_t lc[NB]; Box acc = box_empty(); uint32_t c = 0;
            for (int q = 0; q < NB - 1; q++) { if (cnt[q]) { box_grow(acc, bb[q].lo); box_grow(acc, bb[q].hi); } c += cnt[q]; la[q] = c ? box_half_area(acc) : 0.0f; lc[q] = c; }
            acc = box_empty(); c = 0;
            for (int q = NB - 1; q >= 1; q--) {
                if (cnt[q]) { box_grow(acc, bb[q].lo); box_grow(acc, bb[q].hi); } c += cnt[q];
                const uint32_t nl = lc[q - 1];
                if (nl == 0 || c == 0) continue;
                const float cost = la[q - 1] * (float)nl + box_half_area(acc) * (float)c;
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_split = q; }       // bins [0, q) left, [q, NB) right
            }
        }
        uint32_t mid;
        int axis;
        if (best_axis >= 0) {
            axis = best_axis;
            const float lo = comp(cb.lo, axis), k = (float)NB / comp(ext, axis);
            auto it = std::stable_partition(order.begin() + j.first, order.begin() + j.first + j.count, [&](uint32_t t) {
                int q = (int)((comp(cen[t], axis) - lo) * k); q = q < 0 ? 0 : (q >= NB ? NB - 1 : q);
                return q < best_split;
            });
            mid = (uint32_t)(it - order.begin());
        } else {
            axis = wide;
            mid = j.first + j.count / 2;
            std::nth_element(order.begin() + j.first, order.begin() + mid, order.begin() + j.first + j.count, [&](uint32_t a, uint32_t c) {
                float ca = comp(cen[a], axis), cc = comp(cen[c], axis);
                return ca < cc || (ca == cc && a < c);
            });
        }
        uint32_t l = (uint32_t)nodes.size();
        nodes.push_back(Node2{}); nodes.push_back(Node2{});
        Node2& n2 = nodes[j.node];
        n2.count = 0; n2.first = 0; n2.left = l; n2.right = l + 1; n2.axis = (uint32_t)axis;
        stack.push_back({l, j.first, mid - j.first});
        stack.push_back({l + 1, mid, j.first + j.count - mid});
    }
}

void build_bvh2(Model& m) {
    uint32_t nt = (uint32_t)m.idx.size() / 3;
    std::vector<f3> cen(nt);
    std::vector<Box> tb(nt);
    Box all = box_empty();
    for (uint32_t t = 0; t < nt; t++) {
        Box b = box_empty();
        for (int k = 0; k < 3; k++) box_grow(b, m.pos[m.idx[3 * t + k]]);
        tb[t] = b;
        cen[t] = mk3(0.5f * (b.lo.x + b.hi.x), 0.5f * (b.lo.y + b.hi.y), 0.5f * (b.lo.z + b.hi.z));
        box_grow(all, b.lo); box_grow(all, b.hi);
    }
    float scale = 0.0f;
    if (nt) {
        scale = fmaxf(fmaxf(fabsf(all.lo.x), fabsf(all.hi.x)), fmaxf(fmaxf(fabsf(all.lo.y), fabsf(all.hi.y)),
                      fmaxf(fabsf(all.lo.z), fabsf(all.hi.z))));
    }
    float pad = scale * 3.0517578125e-5f + 1e-30f;   // 2^-15 of the model's scale
    m.bounds = all; box_pad(m.bounds, pad);
    // (median splits for the models: measured on the bench scene the CPU arm is bound by its shading arithmetic, the binned SAH bought
    // nothing there and triples the build time of the 10 M-triangle parity scenes)
    build_tree(tb, cen, pad, 4u, false, m.order, m.nodes);
}

struct Hit { float t, b1, b2; uint32_t prim, inst; };

inline bool better(float t, uint32_t inst, uint32_t prim, const Hit& h) {
    if (t < h.t) return true;
    if (t > h.t) return false;
    if (inst < h.inst) return true;
    if (inst > h.inst) return false;
    return prim < h.prim;
}

// returns true if (any_hit && found). Updates best.
bool trace_instance(const orc_scene& S, uint32_t ii, f3 wo, f3 wd, float tmin, float tmax, bool any_hit, int mode,
                    Hit& best, orc_counters& C) {
    const Instance& inst = S.instances[ii];
    const Model& M = S.models[inst.model];
    C.instances_entered++;
    // world -> object with the host-supplied inverse (Renderer.cpp:2091-2121); direction not renormalised,
    // so t is the same parameter in both spaces.
    f4 o4 = mul44(inst.p.objectToWorldInverse, wo.x, wo.y, wo.z, 1.0f);
    f4 d4 = mul44(inst.p.objectToWorldInverse, wd.x, wd.y, wd.z, 0.0f);
    f3 o = mk3(o4.x, o4.y, o4.z), d = mk3(d4.x, d4.y, d4.z);
    RayPre rp = ray_pre(o, d);
    uint32_t nt = (uint32_t)M.idx.size() / 3;
    auto test_tri = [&](uint32_t prim) -> bool {
        C.tris_tested++;
        float t, b1, b2;
        if (!tri_test(rp, M.pos[M.idx[3 * prim]], M.pos[M.idx[3 * prim + 1]], M.pos[M.idx[3 * prim + 2]], tmin, tmax, t, b1, b2))
            return false;
        if (any_hit) { best.t = t; best.b1 = b1; best.b2 = b2; best.prim = prim; best.inst = ii; return true; }
        if (best.inst == 0xFFFFFFFFu || better(t, ii, prim, best)) { best.t = t; best.b1 = b1; best.b2 = b2; best.prim = prim; best.inst = ii; }
        return false;
    };
    if (mode == 0) {
        for (uint32_t p = 0; p < nt; p++) if (test_tri(p)) return true;
        return false;
    }
    if (M.nodes.empty()) return false;
    f3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    uint32_t stack[128]; int sp = 0; stack[sp++] = 0;
    while (sp) {
        const Node2& n = M.nodes[stack[--sp]];
        C.bvh_nodes_visited++;
        float far = (best.inst == 0xFFFFFFFFu || any_hit) ? tmax : fminf(tmax, best.t);
        if (!box_test(n.b, o, inv, tmin, far)) continue;
        if (n.count) {
            for (uint32_t i = 0; i < n.count; i++) if (test_tri(M.order[n.first + i])) return true;
        } else if (comp(d, (int)n.axis) < 0.0f) {       // the nearer child is popped first (the left one holds the lower side of the axis)
            stack[sp++] = n.left; stack[sp++] = n.right;
        } else {
            stack[sp++] = n.right; stack[sp++] = n.left;
        }
    }
    return false;
}

bool trace(const orc_scene& S, f3 o, f3 d, float tmin, float tmax, bool any_hit, int mode, Hit& best, orc_counters& C) {
    best.t = tmax; best.b1 = best.b2 = 0.0f; best.prim = 0xFFFFFFFFu; best.inst = 0xFFFFFFFFu;
    f3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    if (mode != 0 && !S.tlas.empty()) {             // the instances through their own BVH2 (1000 of them in BASELINE config C3)
        uint32_t stack[128]; int sp = 0; stack[sp++] = 0;
        while (sp) {
            const Node2& n = S.tlas[stack[--sp]];
            float far = (best.inst == 0xFFFFFFFFu || any_hit) ? tmax : fminf(tmax, best.t);
            if (!box_test(n.b, o, inv, tmin, far)) continue;
            if (n.count) {
                for (uint32_t k = 0; k < n.count; k++) {
                    const uint32_t i = S.tlas_order[n.first + k];
                    far = (best.inst == 0xFFFFFFFFu || any_hit) ? tmax : fminf(tmax, best.t);
                    if (!box_test(S.instances[i].wbox, o, inv, tmin, far)) continue;
                    if (trace_instance(S, i, o, d, tmin, tmax, any_hit, mode, best, C)) return true;
                }
            } else if (comp(d, (int)n.axis) < 0.0f) { stack[sp++] = n.left; stack[sp++] = n.right; }
            else { stack[sp++] = n.right; stack[sp++] = n.left; }
        }
        return best.inst != 0xFFFFFFFFu;
    }
    for (uint32_t i = 0; i < (uint32_t)S.instances.size(); i++) {
        if (mode != 0) {
            float far = (best.inst == 0xFFFFFFFFu || any_hit) ? tmax : fminf(tmax, best.t);
            if (!box_test(S.instances[i].wbox, o, inv, tmin, far)) continue;
        }
        if (trace_instance(S, i, o, d, tmin, tmax, any_hit, mode, best, C)) return true;
    }
    return best.inst != 0xFFFFFFFFu;
}

// ------------------------------------------------------------------------------------------ shading types
struct MatOpt {                 // shaders/Common_v7.hlsl:62-66 — every field holds a binary16 value
    f3 Kd; float Kd_w;
    float Pr, Pm, Ps, Pc;
    f3 Ks; f3 Ke;
    uint32_t mID;
};
struct HitInfo {                // shaders/Common_v7.hlsl:35-46
    f3 hitPosition; uint32_t materialID; f3 hitNormal; float area; uint32_t objID;
};
struct ReservoirDI { f3 x2; float w_sum; f3 n2; float W; f3 L2; uint32_t M; };    // Reservoir_v7.hlsl:15-20 (L2 half3)
struct ReservoirGI { f3 xn; float w_sum; f3 nn; float W; f3 E3; uint32_t M; };    // Reservoir_v7.hlsl:22-27 (E3 half3)

struct Ctx {
    const orc_scene* S;
    orc_config cfg;
    orc_counters* C;
    int mode;
};

// OOB StructuredBuffer reads return 0
inline orc_material fetch_material(const orc_scene& S, uint32_t id) {
    if (id < S.materials.size()) return S.materials[id];
    orc_material z; memset(&z, 0, sizeof z); return z;
}
// Pass_init_di_v7.hlsl:108-111 / Sampler_v7.hlsl:71-82
inline MatOpt make_matopt(const orc_material& m, uint32_t id) {
    MatOpt o;
    o.Kd = mk3(q16(m.Kd[0]), q16(m.Kd[1]), q16(m.Kd[2])); o.Kd_w = q16(m.Kd[3]);
    o.Pr = q16(m.Pr_Pm_Ps_Pc[0]); o.Pm = q16(m.Pr_Pm_Ps_Pc[1]); o.Ps = q16(m.Pr_Pm_Ps_Pc[2]); o.Pc = q16(m.Pr_Pm_Ps_Pc[3]);
    o.Ks = mk3(q16(m.Ks[0]), q16(m.Ks[1]), q16(m.Ks[2]));
    o.Ke = mk3(q16(m.Ke[0]), q16(m.Ke[1]), q16(m.Ke[2]));
    o.mID = id;
    return o;
}

// shaders/Common_v7.hlsl:119-138
inline float RandomFloat(uint32_t seed[2]) {
    uint32_t v0 = seed[0], v1 = seed[1], sum = 0u;
    const uint32_t delta = 0x9e3779b9u;
    for (uint32_t i = 0; i < 4u; i++) {
        sum += delta;
        v0 += ((v1 << 4u) + 0xA341316Cu) ^ (v1 + sum) ^ ((v1 >> 5u) + 0xC8013EA4u);
        v1 += ((v0 << 4u) + 0xAD90777Du) ^ (v0 + sum) ^ ((v0 >> 5u) + 0x7E95761Eu);
    }
    seed[0] = v0; seed[1] = v1;
    return (float)v0 / 4294967296.0f;
}

// shaders/Pass_init_di_v7.hlsl:63-77 with uint(time) := global sample index (D2)
inline void init_seed(uint32_t x, uint32_t y, uint32_t pass, uint32_t sample, uint32_t seed[2]) {
    seed[0] = (y * 73856093u) ^ (x * 19349663u) ^ (pass * 83492791u) ^ (sample * 293803u);
    seed[1] = (x * 37623481u) ^ (y * 51964263u) ^ (pass * 68250729u) ^ (sample * 423977u);
}

// shaders/Common_v7.hlsl:173-198
inline uint32_t MapPixelID(uint32_t w, uint32_t /*h*/, uint32_t x, uint32_t y) {
    const uint32_t ts = 4;
    uint32_t tcx = (w + ts - 1) / ts;
    return ((y / ts) * tcx + (x / ts)) * (ts * ts) + ((y % ts) * ts + (x % ts));
}

// shaders/Common_v7.hlsl:151-160
inline f3 SafeMultiply(float s, f3 v) {
    f3 r = s * v;
    if (any_nan_inf(r)) return mk3(0, 0, 0);
    return r;
}
inline float SafeMultiply1(float s, float v) {   // float overload via float3 broadcast, .x taken (Appendix C.3)
    float r = s * v;
    if (isnan1(r) || isinf1(r)) return 0.0f;
    return r;
}

// ------------------------------------------------------------------------------------------ GGX (shaders/GGX_v7.hlsl)
// :1-23
inline float ESS_LUT(const Ctx& c, const MatOpt& mat, float NdotV) {
    NdotV = saturate1(NdotV);
    float thetaIdxF = NdotV * 15.0f;
    int i0 = (int)floorf(thetaIdxF);
    int i1 = std::min(i0 + 1, 15);
    float w = thetaIdxF - (float)i0;
    orc_material m = fetch_material(*c.S, mat.mID);
    float v0 = m.LUT[i0], v1 = m.LUT[i1];
    return lerp1(v0, v1, w);
}
// :26-29   pow(abs(1-c),5) = x*x*x*x*x (Appendix C.3)
inline f3 SchlickFresnel(f3 F0, float cosTheta) {
    float x = fabsf(1.0f - cosTheta);
    float p = (((x * x) * x) * x) * x;
    return saturate3(mk3(F0.x + (1.0f - F0.x) * p, F0.y + (1.0f - F0.y) * p, F0.z + (1.0f - F0.z) * p));
}
// :31-40
inline float D_GGX(float NdotH, float roughness) {
    float alpha = roughness * roughness;
    float alpha2 = alpha * alpha;
    float NdotH2 = NdotH * NdotH;
    float denom = (NdotH2 * (alpha2 - 1.0f) + 1.0f);
    return alpha2 / ((PI_REF * denom) * denom);
}
// :43-52
inline float G2_SmithGGX(float NdotV, float NdotL, float alpha) {
    float alpha2 = alpha * alpha;
    float denomA = NdotV * sqrtf(alpha2 + ((1.0f - alpha2) * NdotL) * NdotL);
    float denomB = NdotL * sqrtf(alpha2 + ((1.0f - alpha2) * NdotV) * NdotV);
    return ((2.0f * NdotL) * NdotV) / (denomA + denomB);
}
// :55-61
inline float G1_SmithGGX(float NdotV, float alpha) {
    float alpha2 = alpha * alpha;
    float denomC = sqrtf(alpha2 + ((1.0f - alpha2) * NdotV) * NdotV) + NdotV;
    return (2.0f * NdotV) / denomC;
}
// :65-76
inline void CoordinateSystem(f3 N, f3& T, f3& B) {
    if (fabsf(N.z) < 0.999f) T = normalize3(cross3(mk3(0, 0, 1), N));
    else T = normalize3(cross3(mk3(1, 0, 0), N));
    B = cross3(N, T);
}
// :93-169.  `alpha = Pr*Pr` is a half*half product (true 16-bit types, DXRHelper.h:125).
inline f3 SampleBRDF_GGX(const MatOpt& mat, f3 outgoing, f3 normal, uint32_t seed[2]) {
    float alpha = hmul(mat.Pr, mat.Pr);
    f3 N = normalize3(normal), V = normalize3(outgoing), T1, T2;
    CoordinateSystem(N, T1, T2);
    float vx = dot3(T1, V), vy = dot3(T2, V), vz = dot3(N, V);
    f3 Ve = normalize3(mk3(alpha * vx, alpha * vy, vz));
    float lensq = Ve.x * Ve.x + Ve.y * Ve.y;
    f3 T1h = (lensq > 0.0f) ? mk3(-Ve.y, Ve.x, 0.0f) * d_rsqrt(lensq) : mk3(1, 0, 0);
    f3 T2h = cross3(Ve, T1h);
    float U1 = RandomFloat(seed), U2 = RandomFloat(seed);
    float r = sqrtf(U1);
    float phi = (2.0f * PI_REF) * U2;
    float sn, cs; d_sincos(phi, &sn, &cs);
    float t1 = r * cs, t2 = r * sn;
    float s = 0.5f * (1.0f + Ve.z);
    t2 = (1.0f - s) * sqrtf(saturate1(1.0f - t1 * t1)) + s * t2;
    f3 Nh = (t1 * T1h + t2 * T2h) + sqrtf(saturate1((1.0f - t1 * t1) - t2 * t2)) * Ve;
    f3 Ne = normalize3(mk3(alpha * Nh.x, alpha * Nh.y, fmaxf(0.0f, Nh.z)));
    f3 H = (Ne.x * T1 + Ne.y * T2) + Ne.z * N;
    f3 sample = reflect3(-V, H);
    if (dot3(sample, normal) < 0.0f) sample = -sample;
    return sample;
}
// :174-206
inline f3 EvaluateBRDF_GGX(const Ctx& c, const MatOpt& mat, f3 normal, f3 incoming, f3 outgoing) {
    f3 N = normalize3(normal), V = normalize3(outgoing), L = normalize3(-incoming);
    f3 H = normalize3(V + L);
    float NdotV = dot3(N, V), NdotL = dot3(N, L), NdotH = dot3(N, H), VdotH = dot3(V, H);
    f3 F = SchlickFresnel(mat.Ks, VdotH);
    float D = D_GGX(NdotH, mat.Pr);
    float G = G2_SmithGGX(NdotV, NdotL, hmul(mat.Pr, mat.Pr));
    float denominator = (4.0f * NdotV) * NdotL;
    if (denominator < EPS) return mk3(0, 0, 0);
    f3 specular = ((F * D) * G) / denominator;
    float Ess = ESS_LUT(c, mat, NdotV);
    float kms = (1.0f - Ess) / Ess;
    f3 specular_ess = specular * mk3(1.0f + mat.Ks.x * kms, 1.0f + mat.Ks.y * kms, 1.0f + mat.Ks.z * kms);
    if (any_nan_inf(specular_ess)) return mk3(0, 0, 0);
    return specular_ess;
}
// :209-224
inline float BRDF_PDF_GGX(const MatOpt& mat, f3 normal, f3 incoming, f3 outgoing) {
    f3 N = normalize3(normal), V = normalize3(outgoing), L = normalize3(-incoming);
    f3 H = normalize3(V + L);
    float NdotH = dot3(N, H), NdotV = dot3(N, V);
    float alpha = hmul(mat.Pr, mat.Pr);
    float G1 = G1_SmithGGX(NdotV, alpha);
    float D = D_GGX(NdotH, mat.Pr);
    return (G1 * D) / (NdotV * 4.0f);
}

// ------------------------------------------------------------------------------------------ Lambert (include/Lambertian_v6.hlsl)
// :2-37
inline f3 RandomUnitVectorInHemisphere(f3 normal, uint32_t seed[2]) {
    float u1 = RandomFloat(seed), u2 = RandomFloat(seed);
    float r = sqrtf(u1);
    float theta = (2.0f * 3.14159265358979323846f) * u2;
    float sn, cs; d_sincos(theta, &sn, &cs);
    float x = r * cs, y = r * sn;
    float z = sqrtf(fmaxf(0.0f, (1.0f - x * x) - y * y));
    f3 h = normal;
    f3 up = fabsf(normal.z) < 0.999f ? mk3(0, 0, 1) : mk3(1, 0, 0);
    f3 right = normalize3(cross3(up, h));
    f3 forward = cross3(h, right);
    f3 hs = (x * right + y * forward) + z * h;
    hs = normalize3(hs);
    if (dot3(hs, normal) < 0.0f) hs = -hs;
    return hs;
}
// :51-58
inline f3 EvaluateBRDF_Lambertian(const MatOpt& mat) { return mat.Kd / PI_REF; }
// :61-64
inline float BRDF_PDF_Lambertian(f3 normal, f3 incoming) { return fmaxf(dot3(normal, -incoming), EPS) / PI_REF; }

// ------------------------------------------------------------------------------------------ shaders/BRDF_v7.hlsl
// :50-70
inline void CalculateStrategyProbabilities(const Ctx& c, const MatOpt& mat, f3 outgoing, f3 normal, float& p_d, float& p_s) {
    if (c.cfg.flags & ORC_FLAG_LAMBERT_ONLY) { p_d = 1.0f; p_s = 0.0f; return; }
    float cosTheta = dot3(normal, outgoing);
    f3 fr = SchlickFresnel(mat.Ks, cosTheta);
    p_s = fminf(1.0f, ((fr.x + fr.y) + fr.z) / 3.0f + mat.Pm);
    p_d = 1.0f - p_s;
}
// :7-48  (always consumes one RandomFloat)
inline uint32_t SelectSamplingStrategy(const Ctx& c, const MatOpt& mat, f3 outgoing, f3 normal, uint32_t seed[2], float& probability) {
    float r = RandomFloat(seed);
    if (c.cfg.flags & ORC_FLAG_LAMBERT_ONLY) { probability = 0.0f; return 0; }
    float cosTheta = dot3(normal, outgoing);
    f3 fr = SchlickFresnel(mat.Ks, cosTheta);
    float p_s = fminf(1.0f, ((fr.x + fr.y) + fr.z) / 3.0f + mat.Pm);
    probability = p_s;
    if (r <= p_s) { if (mat.Pr < 0.04f) return 0; return 1; }
    return 0;
}
// :74-88
inline f3 SampleBRDF(uint32_t strategy, const MatOpt& mat, f3 outgoing, f3 normal, uint32_t seed[2]) {
    if (strategy == 0) return RandomUnitVectorInHemisphere(normal, seed);
    return SampleBRDF_GGX(mat, outgoing, normal, seed);
}
// :91-106
inline f3 EvaluateBRDF(const Ctx& c, uint32_t strategy, const MatOpt& mat, f3 normal, f3 incidence, f3 outgoing) {
    if (strategy == 0) return EvaluateBRDF_Lambertian(mat);
    return EvaluateBRDF_GGX(c, mat, normal, incidence, outgoing);
}
// :109-124
inline float BRDF_PDF(uint32_t strategy, const MatOpt& mat, f3 normal, f3 incidence, f3 outgoing) {
    if (strategy == 0) return BRDF_PDF_Lambertian(normal, incidence);
    return BRDF_PDF_GGX(mat, normal, incidence, outgoing);
}

// the "combined lobe" pattern repeated at Sampler_v7.hlsl:123-128, 248-261, 361-374, 443-456, 601-614 and
// Path_Sampler_v7.hlsl:66-78: F = p_d*f0 + p_s*f1 (SafeMultiply'd), P likewise.
inline f3 CombinedF(const Ctx& c, const MatOpt& mat, f3 normal, f3 incidence, f3 outgoing, float p_d, float p_s) {
    f3 F1 = SafeMultiply(p_d, EvaluateBRDF(c, 0, mat, normal, incidence, outgoing));
    f3 F2 = (c.cfg.flags & ORC_FLAG_LAMBERT_ONLY) ? mk3(0, 0, 0)
                                                   : SafeMultiply(p_s, EvaluateBRDF(c, 1, mat, normal, incidence, outgoing));
    return F1 + F2;
}

// ------------------------------------------------------------------------------------------ ClosestHit (shaders/Hit_v7.hlsl:12-61)
inline void ClosestHit(const Ctx& c, f3 ro, f3 rd, const Hit& h, HitInfo& p) {
    const orc_scene& S = *c.S;
    const Instance& inst = S.instances[h.inst];
    const Model& M = S.models[inst.model];
    p.objID = h.inst;
    f3 worldOrigin = ro + h.t * rd;
    uint32_t vertId = 3 * h.prim;
    uint32_t mslot = vertId + M.mat_offset;
    uint32_t materialID = mslot < S.material_ids.size() ? S.material_ids[mslot] : 0u;
    float bary[3] = {(1.0f - h.b1) - h.b2, h.b1, h.b2};
    uint32_t i0 = M.idx[vertId], i1 = M.idx[vertId + 1], i2 = M.idx[vertId + 2];
    f3 e1 = M.pos[i1] - M.pos[i0], e2 = M.pos[i2] - M.pos[i0];
    f3 cross_a = cross3(e1, e2);
    float area_l = fabsf(length3(cross_a) * 0.5f);
    f3 flatNormal = normalize3(cross_a);
    p.area = area_l;
    f3 smooth = mk3(0, 0, 0);
    const uint32_t vi[3] = {i0, i1, i2};
    for (int i = 0; i < 3; i++) {
        f3 n = M.nrm[vi[i]];
        if (n.x != 0.0f && n.y != 0.0f && n.z != 0.0f) smooth = smooth + n * bary[i];
        else smooth = smooth + flatNormal * bary[i];
    }
    f3 normal;
    if (length3(smooth) > 0.0001f) normal = normalize3(smooth); else normal = flatNormal;
    f4 wn = mul44(inst.p.objectToWorldNormal, normal.x, normal.y, normal.z, 0.0f);
    p.hitNormal = normalize3(mk3(wn.x, wn.y, wn.z));
    p.materialID = materialID;
    p.hitPosition = worldOrigin;
}

// TraceRay closest (T3) + ClosestHit/Miss
inline bool TraceClosest(const Ctx& c, f3 o, f3 d, float tmin, float tmax, HitInfo& p) {
    c.C->closest_rays++;
    Hit h;
    if (!trace(*c.S, o, d, tmin, tmax, false, c.mode, h, *c.C)) { p.materialID = MISS_ID; return false; }
    ClosestHit(c, o, d, h, p);
    return true;
}
// TraceRay shadow (T4): returns isHit
inline bool TraceShadow(const Ctx& c, f3 o, f3 d, float tmin, float tmax) {
    c.C->shadow_rays++;
    Hit h;
    return trace(*c.S, o, d, tmin, tmax, true, c.mode, h, *c.C);
}

// ------------------------------------------------------------------------------------------ shaders/Sampler_v7.hlsl
inline float LinearizeVector(f3 v) { return length3(v); }   // :1-5

// :86-104
inline float VisibilityCheck(const Ctx& c, f3 x1, f3 n1, f3 dir, float dist) {
    f3 o = x1 + normalize3(n1) * S_BIAS;
    float tmax = fmaxf(dist - 10.0f * S_BIAS, 2.0f * S_BIAS);
    return TraceShadow(c, o, dir, 0.0f, tmax) ? 0.0f : 1.0f;
}
// :106-131
inline f3 ReconnectDI(const Ctx& c, f3 x1, f3 n1, f3 x2, f3 n2, f3 L, f3 outgoing, const MatOpt& material) {
    f3 dir = x2 - x1;
    float dist = length3(dir);
    float cosThetaX1 = fmaxf(0.0f, dot3(n1, normalize3(dir)));
    if (dot3(n2, normalize3(-dir)) < 0.0f) n2 = -n2;
    float cosThetaX2 = fmaxf(0.0f, dot3(n2, normalize3(-dir)));
    float p_d, p_s;
    CalculateStrategyProbabilities(c, material, normalize3(outgoing), n1, p_d, p_s);
    f3 F = CombinedF(c, material, n1, normalize3(-dir), normalize3(outgoing), p_d, p_s);
    return (((F * L) * cosThetaX1) * cosThetaX2) / (dist * dist);
}
// :134-161
inline f3 ReconnectGI(const Ctx& c, f3 x1, f3 n1, f3 x2, f3 /*n2*/, f3 L, f3 outgoing, const MatOpt& material1) {
    f3 dir = x2 - x1;
    float cosThetaX1 = fabsf(dot3(n1, normalize3(dir)));
    float p_d, p_s;
    CalculateStrategyProbabilities(c, material1, normalize3(outgoing), n1, p_d, p_s);
    f3 Fx1 = CombinedF(c, material1, n1, normalize3(-dir), normalize3(outgoing), p_d, p_s);
    f3 fr = (Fx1 * cosThetaX1) * L;
    if (any_nan_inf(fr)) return mk3(0, 0, 0);
    return fr;
}

// light selection :293-308 (binary search for the first cdf > u; u == 1.0 keeps index 0, Appendix C.3)
inline uint32_t SelectLight(const orc_scene& S, float randomValue) {
    int left = 0, right = (int)S.lights[0].triCount - 1, selected = 0;
    while (left <= right) {
        int mid = left + (right - left) / 2;
        if (randomValue < S.lights[mid].cdf) { selected = mid; right = mid - 1; }
        else left = mid + 1;
    }
    return (uint32_t)selected;
}

struct LightSample { f3 point, normal_l, L_norm, emission; float dist2, dist, pdf_l, cos_x_raw, cos_y_raw; };

// shared front half of SampleLightNEE (:292-346) and SampleLightNEE_GI (:529-584)
inline void SampleLightPoint(const Ctx& c, f3 origin, uint32_t seed[2], LightSample& ls) {
    const orc_scene& S = *c.S;
    float randomValue = RandomFloat(seed);
    const orc_light_tri& lt = S.lights[SelectLight(S, randomValue)];
    const float* M = S.instances[lt.instanceID].p.objectToWorld;
    f4 a = mul44(M, lt.x[0], lt.x[1], lt.x[2], 1.0f);
    f4 b = mul44(M, lt.y[0], lt.y[1], lt.y[2], 1.0f);
    f4 cc = mul44(M, lt.z[0], lt.z[1], lt.z[2], 1.0f);
    f3 x_v = mk3(a.x, a.y, a.z), y_v = mk3(b.x, b.y, b.z), z_v = mk3(cc.x, cc.y, cc.z);
    float xi1 = RandomFloat(seed), xi2 = RandomFloat(seed);
    if (xi1 + xi2 > 1.0f) { xi1 = 1.0f - xi1; xi2 = 1.0f - xi2; }
    float u = (1.0f - xi1) - xi2, v = xi1, w = xi2;
    ls.point = (u * x_v + v * y_v) + w * z_v;
    f3 L = ls.point - origin;
    ls.dist2 = dot3(L, L);
    ls.dist = sqrtf(fmaxf(ls.dist2, EPS));
    ls.L_norm = normalize3(L);
    f3 cross_l = cross3(y_v - x_v, z_v - x_v);
    f3 normal_l = normalize3(cross_l);
    if (dot3(normal_l, -ls.L_norm) < 0.0f) normal_l = -normal_l;
    ls.normal_l = normal_l;
    float area_l = fabsf(length3(cross_l) * 0.5f);
    ls.pdf_l = lt.weight / fmaxf(area_l, EPS);
    ls.emission = mk3(lt.emission[0], lt.emission[1], lt.emission[2]);
}

// :273-396 with useVisibility=false (call site :677-693)
inline void SampleLightNEE(const Ctx& c, float& pdf_light, float& pdf_bsdf, float& p_hat, uint32_t seed[2], f3 worldOrigin,
                           f3 normal, f3 outgoing, const MatOpt& material, f3& emission, f3& x2, f3& n2) {
    LightSample ls;
    SampleLightPoint(c, worldOrigin, seed, ls);
    x2 = ls.point; n2 = ls.normal_l;
    float cos_theta_x = dot3(normal, ls.L_norm);
    float cos_theta_y = dot3(ls.normal_l, -ls.L_norm);
    float G = fmaxf((cos_theta_y * cos_theta_x) / ls.dist2, EPS);
    emission = ls.emission;
    float p_d, p_s;
    f3 on = normalize3(outgoing);
    CalculateStrategyProbabilities(c, material, on, normal, p_d, p_s);
    f3 brdf_light = CombinedF(c, material, normal, -ls.L_norm, on, p_d, p_s);
    float pdf0 = (BRDF_PDF(0, material, normal, -ls.L_norm, on) * cos_theta_y) / ls.dist2;
    float P1 = SafeMultiply1(p_d, pdf0);
    float P2 = 0.0f;
    if (!(c.cfg.flags & ORC_FLAG_LAMBERT_ONLY)) {
        float pdf1 = (BRDF_PDF(1, material, normal, -ls.L_norm, on) * cos_theta_y) / ls.dist2;
        P2 = SafeMultiply1(p_s, pdf1);
    }
    float P = P1 + P2;
    p_hat = LinearizeVector((ls.emission * brdf_light) * G);   // * V with V == 1.0f
    pdf_light = fmaxf(EPS, ls.pdf_l);
    pdf_bsdf = P;
}

// :199-271
inline void SampleLightBSDF(const Ctx& c, float& pdf_light, float& pdf_bsdf, float& p_hat, uint32_t seed[2], f3 worldOrigin,
                            f3 normal, f3 outgoing, const MatOpt& material, uint32_t strategy, f3& emission, f3& x2, f3& n2) {
    f3 sample = SampleBRDF(strategy, material, outgoing, normal, seed);
    HitInfo sp;
    bool hit = TraceClosest(c, worldOrigin, sample, S_BIAS, 10000.0f, sp);
    if (!hit) { p_hat = 0.0f; return; }       // materials[MISS_ID] reads 0 => Ke == 0 => p_hat = 0
    orc_material mk = fetch_material(*c.S, sp.materialID);
    float Ke = (mk.Ke[0] + mk.Ke[1]) + mk.Ke[2];
    emission = mk3(mk.Ke[0], mk.Ke[1], mk.Ke[2]);
    x2 = sp.hitPosition; n2 = sp.hitNormal;
    if (Ke > EPS) {
        f3 L = sp.hitPosition - worldOrigin;
        float dist = length3(L);
        float dist2 = dist * dist;
        float cos_theta = dot3(sp.hitNormal, -sample);
        pdf_light = (Ke / 3.0f) / c.S->lights[0].total_weight;
        float p_d, p_s;
        f3 on = normalize3(outgoing);
        CalculateStrategyProbabilities(c, material, on, normal, p_d, p_s);
        f3 brdf = CombinedF(c, material, normal, -sample, on, p_d, p_s);
        float pdf0 = (BRDF_PDF(0, material, normal, -sample, outgoing) * cos_theta) / dist2;
        float P1 = SafeMultiply1(p_d, pdf0);
        float P2 = 0.0f;
        if (!(c.cfg.flags & ORC_FLAG_LAMBERT_ONLY)) {
            float pdf1 = (BRDF_PDF(1, material, normal, -sample, outgoing) * cos_theta) / dist2;
            P2 = SafeMultiply1(p_s, pdf1);
        }
        pdf_bsdf = P1 + P2;
        float ndot = dot3(normal, sample);
        p_hat = LinearizeVector(((((brdf * emission) * ndot) * cos_theta)) / dist2);
    } else {
        p_hat = 0.0f;
    }
}

// shaders/Reservoir_v7.hlsl:57-80
inline bool UpdateReservoir(ReservoirDI& r, float wi, f3 x2, f3 n2, f3 L2, uint32_t seed[2]) {
    r.w_sum += wi;
    if (RandomFloat(seed) < wi / r.w_sum) { r.x2 = x2; r.n2 = n2; r.L2 = q16v(L2); return true; }
    return false;
}
// shaders/Reservoir_v7.hlsl:30-53
inline bool UpdateReservoir_GI(ReservoirGI& r, float wi, f3 xn, f3 nn, f3 E3, uint32_t seed[2]) {
    r.w_sum += wi;
    if (RandomFloat(seed) < wi / r.w_sum) { r.xn = xn; r.nn = nn; r.E3 = q16v(E3); return true; }
    return false;
}

// :653-736
inline void SampleRIS(const Ctx& c, uint32_t M1, uint32_t M2, f3 outgoing, ReservoirDI& reservoir, const HitInfo& payload,
                      const MatOpt& matOpt, uint32_t seed[2]) {
    float p_strategy = 1.0f;
    uint32_t strategy = SelectSamplingStrategy(c, matOpt, outgoing, payload.hitNormal, seed, p_strategy);
    float fM1 = (float)M1, fM2 = (float)M2;
    for (uint32_t i = 0; i < M1; i++) {
        float pdf_light = 0.0f, pdf_bsdf = 0.0f, p_hat = 0.0f; f3 emission, x2, n2;
        SampleLightNEE(c, pdf_light, pdf_bsdf, p_hat, seed, payload.hitPosition, payload.hitNormal, outgoing, matOpt, emission, x2, n2);
        float mi = pdf_light / (fM1 * pdf_light + fM2 * pdf_bsdf);
        float wi = (mi * p_hat) / pdf_light;
        if (p_hat > 0.0f) UpdateReservoir(reservoir, wi, x2, n2, emission, seed);
    }
    for (uint32_t j = 0; j < M2; j++) {
        float pdf_light = 0.0f, pdf_bsdf = 0.0f, p_hat = 0.0f; f3 emission = mk3(0, 0, 0), x2 = mk3(0, 0, 0), n2 = mk3(0, 0, 0);
        SampleLightBSDF(c, pdf_light, pdf_bsdf, p_hat, seed, payload.hitPosition, payload.hitNormal, outgoing, matOpt, strategy, emission, x2, n2);
        float mi = pdf_bsdf / (fM1 * pdf_light + fM2 * pdf_bsdf);
        float wi = (mi * p_hat) / pdf_bsdf;
        if (p_hat > 0.0f) UpdateReservoir(reservoir, wi, x2, n2, emission, seed);
    }
    reservoir.M = 1;
}

// :508-647 with useVisibility=false (call site Path_Sampler_v7.hlsl:133-151)
inline f3 SampleLightNEE_GI(const Ctx& c, float& pdf_light, float& pdf_bsdf, f3& incoming, f3& x2_pos, uint32_t seed[2], f3 origin,
                            f3 normal, f3 outgoing, f3 acc_l, float acc_pdf, f3& throughput, f3& emission, const MatOpt& material) {
    LightSample ls;
    SampleLightPoint(c, origin, seed, ls);
    x2_pos = ls.point;
    float cos_theta_x = fabsf(dot3(normal, ls.L_norm));
    if (cos_theta_x < EPS) cos_theta_x = 0.0f;
    float cos_theta_y = fabsf(dot3(ls.normal_l, -ls.L_norm));
    if (cos_theta_y < EPS) cos_theta_y = 0.0f;
    float G = cos_theta_x;
    float p_d, p_s;
    f3 on = normalize3(outgoing);
    CalculateStrategyProbabilities(c, material, on, normal, p_d, p_s);
    f3 brdf_light = CombinedF(c, material, normal, -ls.L_norm, on, p_d, p_s);
    float P1 = SafeMultiply1(p_d, BRDF_PDF(0, material, normal, -ls.L_norm, on));
    float P2 = (c.cfg.flags & ORC_FLAG_LAMBERT_ONLY) ? 0.0f : SafeMultiply1(p_s, BRDF_PDF(1, material, normal, -ls.L_norm, on));
    float P = P1 + P2;
    if (cos_theta_y > 0.0f) pdf_light = (fmaxf(EPS, ls.pdf_l) * ls.dist2) / cos_theta_y;
    pdf_bsdf = P;
    incoming = -ls.L_norm;
    acc_pdf *= pdf_light;
    acc_l = acc_l * (brdf_light * G);           // * V with V == 1.0f
    throughput = brdf_light * G;
    emission = ls.emission;
    if (acc_pdf > 0.0f) return (ls.emission * acc_l) / acc_pdf;
    return mk3(0, 0, 0);
}

// :399-505.  Returns status: 0 = continue (no emitter), 1 = emitter hit, 2 = miss (D1)
inline int SampleLightBSDF_GI(const Ctx& c, float& pdf_light, float& pdf_bsdf, f3& incoming, f3& new_origin, f3& new_normal,
                              f3& new_outgoing, MatOpt& new_material, uint32_t seed[2], uint32_t strategy, f3 origin, f3 normal,
                              f3 outgoing, f3& acc_l, float& acc_pdf, f3& throughput, f3& emission, const MatOpt& material,
                              f3& contribution) {
    f3 sample = SampleBRDF(strategy, material, outgoing, normal, seed);
    HitInfo sp;
    bool hit = TraceClosest(c, origin, sample, S_BIAS, 10000.0f, sp);
    contribution = mk3(0, 0, 0);
    if (!hit) return 2;
    MatOpt mat_ke = make_matopt(fetch_material(*c.S, sp.materialID), sp.materialID);
    float p_d, p_s;
    f3 on = normalize3(outgoing);
    CalculateStrategyProbabilities(c, material, on, normal, p_d, p_s);
    f3 brdf = CombinedF(c, material, normal, -sample, on, p_d, p_s);
    float P1 = SafeMultiply1(p_d, BRDF_PDF(0, material, normal, -sample, outgoing));
    float P2 = (c.cfg.flags & ORC_FLAG_LAMBERT_ONLY) ? 0.0f : SafeMultiply1(p_s, BRDF_PDF(1, material, normal, -sample, outgoing));
    float P = P1 + P2;
    pdf_bsdf = P;
    float NdotL = dot3(normal, sample);
    incoming = -sample;
    bool emitter = (mat_ke.Ke.x != 0.0f || mat_ke.Ke.y != 0.0f || mat_ke.Ke.z != 0.0f);   // length(half3 Ke) > 0
    acc_pdf *= pdf_bsdf;
    acc_l = acc_l * (brdf * NdotL);
    throughput = brdf * NdotL;
    if (emitter) {
        f3 L = sp.hitPosition - origin;
        float dist = length3(L);
        float dist2 = dist * dist;
        float cos_theta = dot3(sp.hitNormal, -sample);
        float kesum = hadd(hadd(mat_ke.Ke.x, mat_ke.Ke.y), mat_ke.Ke.z);       // half arithmetic
        pdf_light = (((kesum / 3.0f) / c.S->lights[0].total_weight) * dist2) / cos_theta;
        emission = mat_ke.Ke;
        contribution = (mat_ke.Ke * acc_l) / acc_pdf;
        return 1;
    }
    new_origin = sp.hitPosition; new_normal = sp.hitNormal; new_outgoing = -sample; new_material = mat_ke;
    emission = mk3(0, 0, 0);
    return 0;
}

// ------------------------------------------------------------------------------------------ shaders/Path_Sampler_v7.hlsl:3-286
inline f3 SamplePathSimple(const Ctx& c, ReservoirGI& reservoir, f3 initPoint, f3 initNormal, f3 initOutgoing,
                           const MatOpt& initMaterial, uint32_t seed[2]) {
    f3 acc_f = mk3(1, 1, 1), acc_f_reconnection = mk3(1, 1, 1);
    float acc_pdf = 1.0f;
    f3 x1_shadow = mk3(0, 0, 0), x2_shadow = mk3(0, 0, 0);
    f3 acc_L = mk3(0, 0, 0);
    f3 origin = initPoint, normal = initNormal, outgoing = normalize3(initOutgoing);
    MatOpt material = initMaterial;
    const uint32_t nee = c.cfg.nee_samples;
    const float fnee = (float)nee;
    {   // :37-99
        float p_strategy;
        uint32_t strategy = SelectSamplingStrategy(c, material, outgoing, normal, seed, p_strategy);
        f3 sample = SampleBRDF(strategy, material, outgoing, normal, seed);
        HitInfo sp;
        bool hit = TraceClosest(c, origin, sample, S_BIAS, 10000.0f, sp);
        if (!hit) return mk3(0, 0, 0);                                          // D1
        orc_material hm = fetch_material(*c.S, sp.materialID);
        if (length3(mk3(hm.Ke[0], hm.Ke[1], hm.Ke[2])) > 0.0f) return mk3(0, 0, 0);   // :55-60
        f3 incoming = normalize3(-sample);
        float p_d, p_s;
        CalculateStrategyProbabilities(c, material, outgoing, normal, p_d, p_s);
        f3 F = CombinedF(c, material, normal, incoming, outgoing, p_d, p_s);
        float P1 = SafeMultiply1(p_d, BRDF_PDF(0, material, normal, incoming, outgoing));
        float P2 = (c.cfg.flags & ORC_FLAG_LAMBERT_ONLY) ? 0.0f : SafeMultiply1(p_s, BRDF_PDF(1, material, normal, incoming, outgoing));
        float P = P1 + P2;
        float NdotL = dot3(normal, sample);
        acc_pdf *= P;
        acc_f = acc_f * (F * NdotL);
        outgoing = incoming;
        material = make_matopt(hm, sp.materialID);
        normal = sp.hitNormal;
        origin = sp.hitPosition;
    }
    f3 xn = origin, nn = normalize3(normal);          // :104-106
    for (uint32_t i = 0; i < c.cfg.bounces; i++) {    // :112-269
        float p_strategy = 1.0f;
        uint32_t strategy = SelectSamplingStrategy(c, material, outgoing, normal, seed, p_strategy);
        for (uint32_t j = 0; j < nee; j++) {
            float pdf_light = 1.0f, pdf_bsdf = 1.0f;
            f3 throughput_NEE = mk3(1, 1, 1), emission_NEE = mk3(0, 0, 0), incoming_NEE, x2;
            f3 contribution = SampleLightNEE_GI(c, pdf_light, pdf_bsdf, incoming_NEE, x2, seed, origin, normal, outgoing, acc_f,
                                                acc_pdf, throughput_NEE, emission_NEE, material);
            float mi = pdf_light / (fnee * pdf_light + pdf_bsdf);
            f3 E_reconnection = ((acc_f_reconnection * mi) * emission_NEE) * throughput_NEE;
            f3 E_path = mi * contribution;
            float wi = LinearizeVector(E_path);
            acc_L = acc_L + mi * contribution;
            if (isnan1(wi) || isinf1(wi)) wi = 0.0f;
            if (UpdateReservoir_GI(reservoir, wi, xn, normalize3(nn), E_reconnection, seed)) {
                x1_shadow = origin + S_BIAS * normalize3(normal);
                x2_shadow = x2;
            }
        }
        (void)strategy;
        float pdf_light = 1.0f, pdf_bsdf = 1.0f;
        f3 throughput_BSDF = mk3(1, 1, 1), emission_BSDF = mk3(0, 0, 0), incoming_BSDF;
        f3 new_origin, new_normal, new_outgoing; MatOpt new_material;
        strategy = SelectSamplingStrategy(c, material, outgoing, normal, seed, p_strategy);
        f3 contribution;
        int st = SampleLightBSDF_GI(c, pdf_light, pdf_bsdf, incoming_BSDF, new_origin, new_normal, new_outgoing, new_material, seed,
                                    strategy, origin, normal, outgoing, acc_f, acc_pdf, throughput_BSDF, emission_BSDF, material,
                                    contribution);
        if (st == 2) break;                                                        // D1
        acc_f_reconnection = acc_f_reconnection * throughput_BSDF;
        if (length3(contribution) > 0.0f) {
            float mi = pdf_bsdf / (fnee * pdf_light + pdf_bsdf);
            f3 E_reconnection = (acc_f_reconnection * mi) * emission_BSDF;
            f3 E_path = mi * contribution;
            float wi = LinearizeVector(E_path);
            acc_L = acc_L + E_path;
            if (isnan1(wi) || isinf1(wi)) wi = 0.0f;
            UpdateReservoir_GI(reservoir, wi, xn, normalize3(nn), E_reconnection, seed);
            break;
        } else {
            if (st == 1) break;                                                    // D4
            origin = new_origin; material = new_material; outgoing = new_outgoing; normal = new_normal;
        }
    }
    f3 ds = x2_shadow - x1_shadow;
    if (nee > 0 && length3(ds) > EPS) {              // :271-283
        float len = length3(ds);
        bool isHit = TraceShadow(c, x1_shadow, normalize3(ds), 0.5f * S_BIAS, fmaxf(S_BIAS, len - (S_BIAS * 5.0f)));
        reservoir.w_sum *= isHit ? 0.0f : 1.0f;
    }
    return acc_L;
}

// ------------------------------------------------------------------------------------------ camera (Pass_init_di_v7.hlsl:59,80-95)
inline void CameraRay(const orc_config& cfg, const orc_camera& cam, uint32_t x, uint32_t y, float jx, float jy, f3& o, f3& dir) {
    float dimx = (float)cfg.width, dimy = (float)cfg.height;
    f4 o4 = mul44(cam.viewI, 0.0f, 0.0f, 0.0f, 1.0f);
    o = mk3(o4.x, o4.y, o4.z);
    float dx = (((float)x + jx) / dimx) * 2.0f - 1.0f;
    float dy = (((float)y + jy) / dimy) * 2.0f - 1.0f;
    f4 target = mul44(cam.projectionI, dx, -dy, 1.0f, 1.0f);
    f4 d4 = mul44(cam.viewI, target.x, target.y, target.z, 0.0f);
    dir = normalize3(mk3(d4.x, d4.y, d4.z));
}

struct PixelDebug {
    f3 cam_o, cam_d; uint32_t hit_inst, hit_prim; float hit_t;
    uint32_t mID; f3 x1, n1;
    ReservoirDI rdi; ReservoirGI rgi; float p_hat; f3 C; uint32_t seed_end[2]; f3 acc_L;
    uint32_t kind; f3 L1;       // 0 = miss, 1 = primary hit on an emitter (L1 = half3 Ke), 2 = sampled
};

// Pass_init_di_v7.hlsl:48-190 followed by the E0 final shade (Pass_spat_di_v7.hlsl:334-372 with no neighbours)
inline f3 RenderSample(const Ctx& c, const orc_camera& cam, uint32_t x, uint32_t y, uint32_t sample, PixelDebug* dbg) {
    uint32_t seed[2];
    init_seed(x, y, 1u, sample, seed);
    float jx = 0.0f, jy = 0.0f;
    if (c.cfg.flags & ORC_FLAG_JITTER) { jx = RandomFloat(seed); jy = RandomFloat(seed); }    // D3
    f3 origin, direction;
    CameraRay(c.cfg, cam, x, y, jx, jy, origin, direction);
    c.C->paths++;
    if (dbg) { dbg->cam_o = origin; dbg->cam_d = direction; }
    // primary ray :91-99 — trace here (not via TraceClosest) to expose the raw hit for debugging
    c.C->closest_rays++;
    Hit h;
    HitInfo payload;
    bool hit = trace(*c.S, origin, direction, 0.0001f, 10000.0f, false, c.mode, h, *c.C);
    if (dbg) { dbg->hit_inst = h.inst; dbg->hit_prim = h.prim; dbg->hit_t = hit ? h.t : -1.0f; }
    if (!hit) return mk3(0, 0, 0);                                                  // D1
    ClosestHit(c, origin, direction, h, payload);
    uint32_t mID = payload.materialID;
    orc_material fm = fetch_material(*c.S, mID);
    MatOpt matOpt = make_matopt(fm, mID);
    if (dbg) { dbg->mID = mID; dbg->x1 = payload.hitPosition; dbg->n1 = payload.hitNormal; }
    if (length3(mk3(fm.Ke[0], fm.Ke[1], fm.Ke[2])) > 0.0f) {                       // :103-106,132 ; D5 (L1 is half3)
        if (dbg) { dbg->kind = 1; dbg->L1 = matOpt.Ke; }
        return matOpt.Ke;
    }
    if (dbg) dbg->kind = 2;
    ReservoirDI reservoir; memset(&reservoir, 0, sizeof reservoir);
    ReservoirGI reservoir_GI; memset(&reservoir_GI, 0, sizeof reservoir_GI);
    f3 o = -direction;
    SampleRIS(c, c.cfg.nee_samples_di, 1u, o, reservoir, payload, matOpt, seed);    // :151-159
    f3 x1 = payload.hitPosition, n1 = normalize3(payload.hitNormal);               // :161-164
    // GetP_Hat(..., true) :166  (Sampler_v7.hlsl:163-171)
    f3 rdi = ReconnectDI(c, x1, n1, reservoir.x2, reservoir.n2, reservoir.L2, o, matOpt);
    float f_g = LinearizeVector(rdi);
    float v = 1.0f;
    if (f_g > EPS) {                                                                // D6
        f3 d21 = reservoir.x2 - x1;
        v = VisibilityCheck(c, x1, n1, normalize3(d21), length3(d21));
    }
    float p_hat = f_g * v;
    reservoir.W = (p_hat > EPS) ? reservoir.w_sum / p_hat : 0.0f;                   // GetW :183-188
    f3 acc_L = SamplePathSimple(c, reservoir_GI, payload.hitPosition, payload.hitNormal, o, matOpt, seed);   // :173
    // E0 final shade
    f3 Cdi = ReconnectDI(c, x1, n1, reservoir.x2, reservoir.n2, reservoir.L2, o, matOpt) * reservoir.W;
    f3 f_gi = ReconnectGI(c, x1, n1, reservoir_GI.xn, reservoir_GI.nn, reservoir_GI.E3, o, matOpt);   // GetP_Hat_GI(..., false)
    float p_hat_gi = LinearizeVector(f_gi);
    reservoir_GI.W = (p_hat_gi > EPS) ? reservoir_GI.w_sum / p_hat_gi : 0.0f;       // GetW_GI :190-195
    reservoir_GI.M = 1;
    f3 Cc = Cdi + f_gi * reservoir_GI.W;
    if (dbg) { dbg->rdi = reservoir; dbg->rgi = reservoir_GI; dbg->p_hat = p_hat; dbg->C = Cc; dbg->seed_end[0] = seed[0]; dbg->seed_end[1] = seed[1]; dbg->acc_L = acc_L; }
    return Cc;
}

// ============================================================================================ legacy estimator
// The reference's older single-pass path tracer: include/RayGen.hlsl:60-137 (path loop with Russian roulette after depth 3),
// include/Hit.hlsl:58-357 (everything in ClosestHit: MIS-weighted emitter hits, RIS over RIS_M = 10 light candidates with ONE
// shadow ray per bounce, BSDF sample), include/Miss.hlsl, include/BRDF.hlsl, include/GGX.hlsl, include/Lambertian.hlsl,
// include/Common.hlsl (PI 3.1415, s_bias 1e-5, EPSILON 1e-4).  Full fp32 materials (no MaterialOptimized), face-forwarded normals,
// primary direction NOT normalised (RayGen.hlsl:89).  It runs on the v7 buffers (S4 materials with LUT[16], S7 lights, S6 instance
// properties hold the same quantities as the legacy structs).  Deviations, applied to oracle and GPU alike: D2 (seed from the sample
// index), D13: the path loop is capped at cfg.bounces closest-hit rays instead of 10 000 000.
namespace legacy {
const float L_EPS = 0.0001f, L_BIAS = 0.00001f;

struct Payload {
    f3 color; f3 emission; f3 direction; f3 origin; float util_x, util_y; uint32_t seed[2]; float pdf; f3 hitNormal;
};
inline f3 Kd3(const orc_material& m) { return mk3(m.Kd[0], m.Kd[1], m.Kd[2]); }
inline f3 Ks3(const orc_material& m) { return mk3(m.Ks[0], m.Ks[1], m.Ks[2]); }
inline f3 Ke3(const orc_material& m) { return mk3(m.Ke[0], m.Ke[1], m.Ke[2]); }
inline f3 abs3(f3 v) { return mk3(fabsf(v.x), fabsf(v.y), fabsf(v.z)); }

// include/GGX.hlsl:4-27
inline float ESS_LUT(const orc_material& mat, float NdotV) {
    NdotV = saturate1(NdotV);
    float thetaIdxF = NdotV * 15.0f;
    int i0 = (int)floorf(thetaIdxF);
    int i1 = std::min(i0 + 1, 15);
    float w = thetaIdxF - (float)i0;
    return lerp1(mat.LUT[i0], mat.LUT[i1], w);
}
// :37-46 (denominator clamped, unlike v7)
inline float D_GGX(float NdotH, float roughness) {
    float alpha = roughness * roughness;
    float alpha2 = alpha * alpha;
    float NdotH2 = NdotH * NdotH;
    float denom = (NdotH2 * (alpha2 - 1.0f) + 1.0f);
    denom = fmaxf(denom, 1e-7f);
    return alpha2 / ((PI_REF * denom) * denom);
}
// :87-142: Heitz VNDF, invalid samples (below the surface) become the zero vector
inline f3 SampleBRDF_GGX(const orc_material& mat, f3 outgoing, f3 normal, uint32_t seed[2]) {
    float alpha = mat.Pr_Pm_Ps_Pc[0] * mat.Pr_Pm_Ps_Pc[0];
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
    float phi = (2.0f * PI_REF) * e1;
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
inline f3 EvaluateBRDF_GGX(const orc_material& mat, f3 normal, f3 incoming, f3 outgoing) {
    f3 N = normalize3(normal), V = normalize3(outgoing), L = normalize3(-incoming);
    f3 H = normalize3(V + L);
    float NdotV = saturate1(dot3(N, V)), NdotL = saturate1(dot3(N, L)), NdotH = saturate1(dot3(N, H)), VdotH = saturate1(dot3(V, H));
    f3 F = SchlickFresnel(Ks3(mat), VdotH);
    float D = D_GGX(NdotH, mat.Pr_Pm_Ps_Pc[0]);
    float G = G2_SmithGGX(NdotV, NdotL, mat.Pr_Pm_Ps_Pc[0] * mat.Pr_Pm_Ps_Pc[0]);
    float denominator = (4.0f * NdotV) * NdotL;
    denominator = fmaxf(denominator, 1e-7f);
    f3 specular = ((F * D) * G) / denominator;
    float Ess = ESS_LUT(mat, NdotV);
    float kms = (1.0f - Ess) / Ess;
    f3 ks = Ks3(mat);
    return specular * mk3(1.0f + ks.x * kms, 1.0f + ks.y * kms, 1.0f + ks.z * kms);
}
// :179-197
inline float BRDF_PDF_GGX(const orc_material& mat, f3 normal, f3 incoming, f3 outgoing) {
    f3 N = normalize3(normal), V = normalize3(outgoing), L = normalize3(-incoming);
    f3 H = normalize3(V + L);
    float NdotH = saturate1(dot3(N, H)), NdotV = saturate1(dot3(N, V));
    float alpha = mat.Pr_Pm_Ps_Pc[0] * mat.Pr_Pm_Ps_Pc[0];
    float G1 = G1_SmithGGX(NdotV, alpha);
    float D = D_GGX(NdotH, mat.Pr_Pm_Ps_Pc[0]);
    return (G1 * D) / (NdotV * 4.0f);
}
// include/BRDF.hlsl:10-53: strategy 1 (GGX) iff r <= p_s; clearcoat and dissolve enter the probabilities
inline uint32_t SelectSamplingStrategy(const Ctx& c, const orc_material& mat, f3 outgoing, f3 normal, uint32_t seed[2]) {
    float r = RandomFloat(seed);
    float metallic = mat.Pr_Pm_Ps_Pc[1], clearcoat = mat.Pr_Pm_Ps_Pc[3];
    float cosTheta = dot3(normal, outgoing);
    f3 fresnel = SchlickFresnel(Ks3(mat), cosTheta);
    float p_s = fminf(1.0f, (((fresnel.x + fresnel.y) + fresnel.z) / 3.0f + clearcoat) + metallic);
    if (c.cfg.flags & ORC_FLAG_LAMBERT_ONLY) return 0u;
    return r <= p_s ? 1u : 0u;
}
// BRDF.hlsl:72-106; include/Lambertian.hlsl:54-66 (pdf floor 1e-4)
inline f3 EvaluateBRDF(uint32_t strategy, const orc_material& mat, f3 normal, f3 incidence, f3 outgoing) {
    return strategy == 0u ? Kd3(mat) / PI_REF : EvaluateBRDF_GGX(mat, normal, incidence, outgoing);
}
inline float BRDF_PDF(uint32_t strategy, const orc_material& mat, f3 normal, f3 incidence, f3 outgoing) {
    return strategy == 0u ? fmaxf(dot3(normal, -incidence), 0.0001f) / PI_REF : BRDF_PDF_GGX(mat, normal, incidence, outgoing);
}

// include/Hit.hlsl:58-357
inline void ClosestHit(const Ctx& c, Payload& payload, f3 ro, f3 rd, const Hit& h) {
    const orc_scene& S = *c.S;
    const Instance& inst = S.instances[h.inst];
    const Model& M = S.models[inst.model];
    uint32_t vertId = 3 * h.prim;
    uint32_t mslot = vertId + M.mat_offset;
    uint32_t materialID = mslot < S.material_ids.size() ? S.material_ids[mslot] : 0u;
    const orc_material mat = fetch_material(S, materialID);
    float bary[3] = {(1.0f - h.b1) - h.b2, h.b1, h.b2};
    uint32_t i0 = M.idx[vertId], i1 = M.idx[vertId + 1], i2 = M.idx[vertId + 2];
    f3 e1 = M.pos[i1] - M.pos[i0], e2 = M.pos[i2] - M.pos[i0];
    f3 flatNormal = normalize3(cross3(e1, e2));
    f3 smooth = mk3(0, 0, 0);
    const uint32_t vi[3] = {i0, i1, i2};
    for (int i = 0; i < 3; i++) {
        f3 n = M.nrm[vi[i]];
        if (n.x != 0.0f && n.y != 0.0f && n.z != 0.0f) smooth = smooth + n * bary[i];
        else smooth = smooth + flatNormal * bary[i];
    }
    f3 normal;
    if (length3(smooth) > 0.0001f) normal = normalize3(smooth); else normal = flatNormal;
    f4 wn = mul44(inst.p.objectToWorldNormal, normal.x, normal.y, normal.z, 0.0f);
    normal = normalize3(mk3(wn.x, wn.y, wn.z));
    f4 wf = mul44(inst.p.objectToWorldNormal, flatNormal.x, flatNormal.y, flatNormal.z, 0.0f);
    flatNormal = normalize3(mk3(wf.x, wf.y, wf.z));
    if (dot3(normal, -payload.direction) < 0.0f) normal = -normal;                 // :108-111 face-forwarding
    if (dot3(flatNormal, -payload.direction) < 0.0f) flatNormal = -flatNormal;
    f3 worldOrigin = ro + h.t * rd;
    f3 direct = mk3(0, 0, 0), emissive = mk3(0, 0, 0);
    float pdf_sample = 1.0f;
    f3 brdf_sample = mk3(0, 0, 0);
    f3 origin = payload.origin;
    f3 incoming = -payload.direction;
    const f3 Ke = Ke3(mat);
    if (length3(Ke) > 0.0f) {                                                       // :127-171 emitter hit
        if (payload.util_y == 0.0f) { emissive = Ke; payload.util_x = 1.0f; }
        else {
            f3 L = worldOrigin - origin;
            float dist2 = fmaxf(dot3(L, L), L_EPS);
            float dist = fmaxf(sqrtf(dist2), L_EPS);
            f3 Ln = L / dist;
            float cos_e = fmaxf(L_EPS, dot3(normal, -Ln));
            const float* MM = inst.p.objectToWorld;
            f4 a = mul44(MM, M.pos[i0].x, M.pos[i0].y, M.pos[i0].z, 1.0f);
            f4 b = mul44(MM, M.pos[i1].x, M.pos[i1].y, M.pos[i1].z, 1.0f);
            f4 cc = mul44(MM, M.pos[i2].x, M.pos[i2].y, M.pos[i2].z, 1.0f);
            f3 x_v = mk3(a.x, a.y, a.z), y_v = mk3(b.x, b.y, b.z), z_v = mk3(cc.x, cc.y, cc.z);
            f3 cross_l = cross3(y_v - x_v, z_v - x_v);
            float area_l = fabsf(length3(cross_l) * 0.5f);
            float s_weight = area_l * (((Ke.x + Ke.y) + Ke.z) / 3.0f);
            float t_weight = S.lights.empty() ? 0.0f : S.lights[0].total_weight;
            float weight = s_weight / t_weight;
            float pdf_l = fmaxf(L_EPS, (weight * dist2) / cos_e);
            float weight_emissive = payload.pdf / (payload.pdf + pdf_l);
            emissive = (Ke * payload.color) * weight_emissive;
            payload.util_x = 1.0f;
        }
    } else {                                                                        // :173-331 RIS-10 NEE + BSDF sample
        f3 outgoing = -payload.direction;
        uint32_t strategy = SelectSamplingStrategy(c, mat, outgoing, normal, payload.seed);
        const int RIS_M = 10;
        float ris_weights[RIS_M], ris_luminance[RIS_M], ris_dist[RIS_M], ris_cos_theta[RIS_M], ris_pdf_brdf_light[RIS_M], ris_cdf[RIS_M], ris_pdf_l[RIS_M];
        f3 ris_f[RIS_M], ris_LDir[RIS_M];
        for (int i = 0; i < RIS_M; i++) {
            float randomValue = RandomFloat(payload.seed);
            orc_light_tri zl; memset(&zl, 0, sizeof zl);
            const orc_light_tri& lt = S.lights.empty() ? zl : S.lights[SelectLight(S, randomValue)];
            const float* MM = S.instances[lt.instanceID].p.objectToWorld;
            f4 a = mul44(MM, lt.x[0], lt.x[1], lt.x[2], 1.0f);
            f4 b = mul44(MM, lt.y[0], lt.y[1], lt.y[2], 1.0f);
            f4 cc = mul44(MM, lt.z[0], lt.z[1], lt.z[2], 1.0f);
            f3 x_v = mk3(a.x, a.y, a.z), y_v = mk3(b.x, b.y, b.z), z_v = mk3(cc.x, cc.y, cc.z);
            float xi1 = RandomFloat(payload.seed), xi2 = RandomFloat(payload.seed);
            if (xi1 + xi2 > 1.0f) { xi1 = 1.0f - xi1; xi2 = 1.0f - xi2; }
            float u = (1.0f - xi1) - xi2, v = xi1, w = xi2;
            f3 samplePoint = (u * x_v + v * y_v) + w * z_v;
            f3 L = samplePoint - (worldOrigin + L_BIAS * flatNormal);
            float dist2 = fmaxf(dot3(L, L), L_EPS);
            float dist = fmaxf(sqrtf(dist2), L_EPS);
            f3 L_norm = L / dist;
            f3 cross_l = cross3(y_v - x_v, z_v - x_v);
            f3 normal_l = normalize3(cross_l);
            float area_l = fabsf(length3(cross_l) * 0.5f);
            float cos_theta_x = fmaxf(L_EPS, dot3(normal, L_norm));
            float cos_theta_y = fmaxf(L_EPS, dot3(normal_l, -L_norm));
            float G = fmaxf((cos_theta_x * cos_theta_y) / dist2, L_EPS);
            float pdf_l = lt.weight / fmaxf(area_l, L_EPS);
            f3 emission_l = mk3(lt.emission[0], lt.emission[1], lt.emission[2]);
            f3 brdf_light = EvaluateBRDF(strategy, mat, normal, -L_norm, -payload.direction);
            float pdf_brdf_light = fmaxf(BRDF_PDF(strategy, mat, normal, -L_norm, -payload.direction), L_EPS);
            ris_f[i] = (emission_l * brdf_light) * G;
            float lum = ((emission_l.x + emission_l.y) + emission_l.z) / 3.0f;
            ris_luminance[i] = (lum * brdf_light.x) * G;                            // float3 -> float keeps .x (:268)
            ris_weights[i] = (1.0f / 10.0f) * (((lum * brdf_light.x) * G) / pdf_l);
            ris_LDir[i] = L_norm; ris_dist[i] = dist; ris_cos_theta[i] = cos_theta_y;
            ris_pdf_brdf_light[i] = pdf_brdf_light; ris_pdf_l[i] = pdf_l;
        }
        ris_cdf[0] = ris_weights[0];
        for (int i = 1; i < RIS_M; i++) ris_cdf[i] = ris_cdf[i - 1] + ris_weights[i];
        float ris_total_weight = ris_cdf[RIS_M - 1];
        float threshold = RandomFloat(payload.seed) * ris_total_weight;
        int sel = 0;
        for (int i = 0; i < RIS_M; i++) if (threshold < ris_cdf[i]) { sel = i; break; }
        float WX = fmaxf(L_EPS, (1.0f / fmaxf(L_EPS, ris_luminance[sel])) * ris_total_weight);
        bool hit = TraceShadow(c, worldOrigin + L_BIAS * flatNormal, ris_LDir[sel], L_BIAS, fabsf(ris_dist[sel]) - L_BIAS);
        float visible = hit ? 0.0f : 1.0f;
        direct = (ris_f[sel] * visible) * WX;
        float pdf_l_sa = fmaxf(L_EPS, ((ris_pdf_l[sel] * ris_dist[sel]) * ris_dist[sel]) / ris_cos_theta[sel]);
        float weight_light = pdf_l_sa / (pdf_l_sa + ris_pdf_brdf_light[sel]);
        direct = direct * (payload.color * weight_light);
        f3 sample = strategy == 0u ? RandomUnitVectorInHemisphere(normal, payload.seed) : SampleBRDF_GGX(mat, outgoing, normal, payload.seed);
        payload.direction = sample;
        payload.origin = worldOrigin + L_BIAS * flatNormal;
        incoming = -payload.direction;
        if (length3(payload.direction) < 0.01f) payload.util_x = 1.1f;              // :320-323 invalid sample ends the path
        else {
            pdf_sample = fmaxf(BRDF_PDF(strategy, mat, normal, incoming, outgoing), 0.0001f);
            brdf_sample = EvaluateBRDF(strategy, mat, normal, incoming, outgoing);
        }
    }
    payload.emission = payload.emission + (abs3(direct) + abs3(emissive));          // :337
    payload.color = ((payload.color * brdf_sample) * dot3(normal, -incoming)) / pdf_sample;   // :342
    payload.pdf = pdf_sample;
    payload.hitNormal = normal;
}

// include/RayGen.hlsl:60-137 for one sample
inline f3 RenderSample(const Ctx& c, const orc_camera& cam, uint32_t x, uint32_t y, uint32_t sample) {
    Payload p;
    init_seed(x, y, 2u, sample, p.seed);                                            // uint(samples + 1) = 2; D2
    float jx = RandomFloat(p.seed), jy = RandomFloat(p.seed);
    float dimx = (float)c.cfg.width, dimy = (float)c.cfg.height;
    f4 o4 = mul44(cam.viewI, 0.0f, 0.0f, 0.0f, 1.0f);
    float dx = (((float)x + jx) / dimx) * 2.0f - 1.0f;
    float dy = (((float)y + jy) / dimy) * 2.0f - 1.0f;
    f4 target = mul44(cam.projectionI, dx, -dy, 1.0f, 1.0f);
    f4 d4 = mul44(cam.viewI, target.x, target.y, target.z, 0.0f);
    p.color = mk3(1, 1, 1); p.emission = mk3(0, 0, 0);
    p.origin = mk3(o4.x, o4.y, o4.z); p.direction = mk3(d4.x, d4.y, d4.z);          // not normalised (:89)
    p.pdf = 1.0f; p.hitNormal = mk3(0, 0, 0);
    c.C->paths++;
    for (uint32_t yb = 0; yb < c.cfg.bounces; yb++) {                               // D13
        p.util_x = 0.0f; p.util_y = (float)yb;
        c.C->closest_rays++;
        Hit h;
        const f3 ro = p.origin, rd = p.direction;
        if (trace(*c.S, ro, rd, 0.0001f, 10000.0f, false, c.mode, h, *c.C)) ClosestHit(c, p, ro, rd, h);
        else {                                                                       // include/Miss.hlsl:9-11
            p.emission = p.emission + mk3(0.0f, 0.0f, 0.0f) * p.color;
            p.util_x = 1.0f;
        }
        if (p.util_x >= 1.0f) break;
        if (yb > 3u) {                                                              // :118-130 Russian roulette
            float max_throughput = fmaxf(p.color.x, fmaxf(p.color.y, p.color.z));
            float q = fminf(fmaxf(max_throughput, 0.05f), 1.0f);
            float random = RandomFloat(p.seed);
            if (random > q) break;
            p.color = p.color * (1.0f / q);
        }
    }
    return p.emission;
}
}  // namespace legacy

// ============================================================================================ ReSTIR reuse
// Passes 2 and 3 of the reference's frame: shaders/Pass_temp_di_v7.hlsl:46-204 (RayGen2), shaders/Pass_spat_di_v7.hlsl:46-464
// (RayGen3), shaders/MIS_v7.hlsl, include/MIS_GI_v6.hlsl, shaders/Common_v7.hlsl:203-350, shaders/Sampler_v7.hlsl:738-785.
// Per-pixel buffers are indexed linearly (MapPixelID is a storage order only, F3).
struct SData { f3 x1; uint32_t mID; f3 L1; f3 n1; f3 o; uint32_t objID; uint32_t kind; };   // Reservoir_v7.hlsl:2-11; mID is uint16_t

// GetP_Hat / GetP_Hat_GI (Sampler_v7.hlsl:163-181).  Deviation D6': the visibility ray is traced only when the unshadowed
// value is > 0 (the product with V is 0 or NaN otherwise, whatever V is) — this also defines the ray count.
inline float GetP_Hat(const Ctx& c, f3 x1, f3 n1, f3 x2, f3 n2, f3 L2, f3 o, const MatOpt& m, bool use_visibility) {
    float f_g = LinearizeVector(ReconnectDI(c, x1, n1, x2, n2, L2, o, m));
    float v = 1.0f;
    if (use_visibility && f_g > 0.0f) { f3 d = x2 - x1; v = VisibilityCheck(c, x1, n1, normalize3(d), length3(d)); }
    return f_g * v;
}
inline f3 GetP_Hat_GI(const Ctx& c, f3 x1, f3 n1, f3 x2, f3 n2, f3 L2, f3 o, const MatOpt& m, bool use_visibility) {
    f3 f_g = ReconnectGI(c, x1, n1, x2, n2, L2, o, m);
    float v = 1.0f;
    if (use_visibility && LinearizeVector(f_g) > 0.0f) { f3 d = x2 - x1; v = VisibilityCheck(c, x1, n1, normalize3(d), length3(d)); }
    return f_g * v;
}
inline float GetW(float w_sum, float p_hat) { return p_hat > EPS ? w_sum / p_hat : 0.0f; }           // :183-195
// Common_v7.hlsl:268-283, 286-307; length(half3) compared with 0 = "any component non-zero" (D10)
inline bool nz3(f3 v) { return v.x != 0.0f || v.y != 0.0f || v.z != 0.0f; }
inline bool RejectDistance(f3 x1, f3 x2, f3 camPos, float threshold) {
    float d1 = length3(x1 - camPos), d2 = length3(x2 - camPos);
    float rel = fabsf(d1 - d2) / fmaxf(d1, d2);
    return rel > threshold;
}
inline bool RejectJacobian(float J, float threshold) { return J > threshold || J < 1.0f / threshold || isnan1(J) || isinf1(J); }
inline bool IsValidReservoir(const ReservoirDI& r) { return length3(r.n2) > 0.0f && nz3(r.L2) && r.w_sum > 0.0f && r.M > 0u; }
inline bool IsValidReservoir_GI(const ReservoirGI& r) { return r.w_sum > 0.0f && r.M > 0u; }
// Common_v7.hlsl:326-345
inline float Jacobian_Reconnection(const SData& r, const SData& q, f3 x2q, f3 n2q) {
    f3 vq = x2q - q.x1, vr = x2q - r.x1;
    float cosPhi2q = fabsf(dot3(normalize3(-vq), normalize3(n2q)));
    float cosPhi2r = fabsf(dot3(normalize3(-vr), normalize3(n2q)));
    float len2_vq = dot3(vq, vq), len2_vr = dot3(vr, vr);
    return (cosPhi2q / cosPhi2r) * (len2_vr / len2_vq);
}
// Common_v7.hlsl:203-244 (pow(u, spatial_exponent = 1) is the identity); returns the pixel
inline void GetRandomPixelCircleWeighted(uint32_t radius, uint32_t w, uint32_t h, uint32_t x, uint32_t y, uint32_t seed[2], int& px, int& py) {
    int newX, newY;
    do {
        float u = RandomFloat(seed);
        float r = (float)radius * u;
        float angle = RandomFloat(seed) * 6.2831853f;
        float sn, cs; d_sincos(angle, &sn, &cs);
        int offsetX = (int)(cs * r), offsetY = (int)(sn * r);
        newX = (int)x + offsetX; newY = (int)y + offsetY;
        while (newX < 0 || newX >= (int)w) { if (newX < 0) newX = -newX; else newX = 2 * (int)w - newX - 2; }
        while (newY < 0 || newY >= (int)h) { if (newY < 0) newY = -newY; else newY = 2 * (int)h - newY - 2; }
    } while (newX == (int)x && newY == (int)y);
    px = newX; py = newY;
}
// Sampler_v7.hlsl:738-785.  D11: a reprojection that lands outside the image is "no candidate" (the reference only tests
// for the (-1,-1) sentinel and lets other out-of-range coordinates alias through MapPixelID / read out of bounds).
inline bool GetBestReprojectedPixel(const orc_scene& S, const orc_camera& cam, f3 worldPos, uint32_t w, uint32_t h, uint32_t objID, int& px, int& py) {
    if (objID >= S.instances.size()) return false;
    const orc_instance_props& ip = S.instances[objID].p;
    f4 lp = mul44(ip.objectToWorldInverse, worldPos.x, worldPos.y, worldPos.z, 1.0f);
    f4 pw = mul44(ip.prevObjectToWorld, lp.x, lp.y, lp.z, lp.w);
    f4 vp = mul44(cam.prevView, pw.x, pw.y, pw.z, pw.w);
    f4 cp = mul44(cam.prevProjection, vp.x, vp.y, vp.z, vp.w);
    if (cp.w <= 0.0f) return false;
    float ndcx = cp.x / cp.w, ndcy = cp.y / cp.w;
    float uvx = ndcx * 0.5f + 0.5f, uvy = ndcy * 0.5f + 0.5f;
    uvy = 1.0f - uvy;
    float fx = rintf(uvx * (float)w), fy = rintf(uvy * (float)h);      // round(): to nearest even
    if (!(fx >= 0.0f && fx < (float)w && fy >= 0.0f && fy < (float)h)) return false;
    px = (int)fx; py = (int)fy;
    return true;
}
inline MatOpt matopt_of(const orc_scene& S, uint32_t mID) { return make_matopt(fetch_material(S, mID), mID); }
inline uint32_t mcap(uint32_t cap, uint32_t M) { return M < cap ? M : cap; }

}  // namespace
struct orc_frames {
    uint32_t w = 0, h = 0;
    std::vector<ReservoirDI> cur, last;
    std::vector<ReservoirGI> gcur, glast;
    std::vector<SData> scur, slast;
};
namespace {

// RayGen2, Pass_temp_di_v7.hlsl:46-204
void TemporalPixel(const Ctx& c, const orc_camera& cam, orc_frames& F, uint32_t x, uint32_t y, uint32_t frame) {
    const uint32_t W = F.w, H = F.h, pix = y * W + x;
    ReservoirDI rc = F.cur[pix]; ReservoirGI gc = F.gcur[pix]; const SData sc = F.scur[pix];
    if (sc.kind != 2u) return;
    f4 o4 = mul44(cam.viewI, 0.0f, 0.0f, 0.0f, 1.0f);
    const f3 init_orig = mk3(o4.x, o4.y, o4.z);
    uint32_t seed[2]; init_seed(x, y, 2u, frame, seed);
    int px, py;
    if (!GetBestReprojectedPixel(*c.S, cam, sc.x1, W, H, sc.objID, px, py)) return;
    const uint32_t tp = (uint32_t)py * W + (uint32_t)px;
    const ReservoirDI rl = F.last[tp]; const ReservoirGI gl = F.glast[tp]; const SData sl = F.slast[tp];
    const bool okDI = !nz3(sl.L1) && IsValidReservoir(rl) && !RejectDistance(sc.x1, sl.x1, init_orig, 0.1f) &&
                      (rl.x2.x != 0.0f && rl.x2.y != 0.0f && rl.x2.z != 0.0f) && sl.mID == sc.mID;
    const bool okGI = !nz3(sl.L1) && !(gl.w_sum > 5.0f) && !RejectDistance(sc.x1, sl.x1, init_orig, 0.1f) &&
                      IsValidReservoir_GI(gl) && sl.mID == sc.mID;
    const MatOpt m = matopt_of(*c.S, sc.mID);
    const uint32_t CAP = 16u;
    if (okDI) {
        const float cM = (float)mcap(CAP, rc.M), nM = (float)mcap(CAP, rl.M);
        const float M_sum = cM + nM;
        float mi_c = cM / M_sum;                                                    // MIS_v7.hlsl:63-71
        { float m_num = cM, m_den = m_num + (M_sum - cM); if (m_den > 0.0f) mi_c += (nM / M_sum) * (m_num / m_den); }
        float mi_t = 0.0f;                                                          // :73-80
        { float m_num = M_sum - cM, m_den = m_num + cM; if (m_den > 0.0f) mi_t = ((nM / M_sum) * m_num) / m_den; }
        if (length3(rl.n2) == 0.0f) { mi_c = 1.0f; mi_t = 0.0f; }
        float w_c = (mi_c * GetP_Hat(c, sc.x1, sc.n1, rc.x2, rc.n2, rc.L2, sc.o, m, false)) * rc.W;
        float w_t = (mi_t * GetP_Hat(c, sc.x1, sc.n1, rl.x2, rl.n2, rl.L2, sc.o, m, true)) * rl.W;
        rc.M = mcap(CAP, rc.M); rc.w_sum = w_c;
        rc.M += mcap(CAP, rl.M);
        UpdateReservoir(rc, w_t, rl.x2, rl.n2, rl.L2, seed);
        float p_hat = GetP_Hat(c, sc.x1, sc.n1, rc.x2, rc.n2, rc.L2, sc.o, m, false);
        rc.W = GetW(rc.w_sum, p_hat);
    }
    if (okGI) {
        const float cM = (float)mcap(CAP, gc.M), nM = (float)mcap(CAP, gl.M);
        const float M_sum = cM + nM;
        float mi_c = cM / M_sum;                                                    // MIS_GI_v6.hlsl:79-95
        { float m_num = cM, m_den = m_num + (M_sum - cM); if (m_den > 0.0f) mi_c += (nM / M_sum) * (m_num / m_den); }
        float mi_t = 0.0f;                                                          // :97-112
        { float m_num = M_sum - cM, m_den = m_num + cM; if (m_den > 0.0f) mi_t = ((nM / M_sum) * m_num) / m_den; }
        f3 f_c = GetP_Hat_GI(c, sc.x1, sc.n1, gc.xn, gc.nn, gc.E3, sc.o, m, false);
        float w_c = (mi_c * LinearizeVector(f_c)) * gc.W;
        f3 f_t = GetP_Hat_GI(c, sc.x1, sc.n1, gl.xn, gl.nn, gl.E3, sc.o, m, true);
        float w_t = (mi_t * LinearizeVector(f_t)) * gl.W;
        gc.M = mcap(CAP, gc.M); gc.w_sum = w_c;
        gc.M += mcap(CAP, gl.M);
        UpdateReservoir_GI(gc, w_t, gl.xn, gl.nn, gl.E3, seed);
        gc.W = GetW(gc.w_sum, LinearizeVector(GetP_Hat_GI(c, sc.x1, sc.n1, gc.xn, gc.nn, gc.E3, sc.o, m, false)));
    }
    F.cur[pix] = rc; F.gcur[pix] = gc;
}

// RayGen3, Pass_spat_di_v7.hlsl:46-464 up to the accumulation; returns the frame's radiance sample of the pixel
f3 SpatialPixel(const Ctx& c, const orc_camera& cam, orc_frames& F, uint32_t x, uint32_t y, uint32_t frame) {
    const uint32_t W = F.w, H = F.h, pix = y * W + x;
    const SData sc = F.scur[pix];
    if (sc.kind == 0u) { F.last[pix] = ReservoirDI(); F.glast[pix] = ReservoirGI(); F.slast[pix] = sc; return mk3(0, 0, 0); }   // D1
    if (sc.kind == 1u) return sc.L1;                            // :458-463, accumulated (D5); the *_last buffers keep their old contents
    f4 o4 = mul44(cam.viewI, 0.0f, 0.0f, 0.0f, 1.0f);
    const f3 init_orig = mk3(o4.x, o4.y, o4.z);
    uint32_t seed[2]; init_seed(x, y, 3u, frame, seed);
    const MatOpt m = matopt_of(*c.S, sc.mID);
    const uint32_t NC = 3u, TRIES = 9u, CAP = 128u;
    uint32_t candDI[3], candGI[3]; uint32_t nDI = 0, nGI = 0;
    float M_sum_DI = (float)mcap(CAP, F.cur[pix].M), M_sum_GI = (float)mcap(CAP, F.gcur[pix].M);
    for (uint32_t attempt = 0; attempt < TRIES && nDI < NC; attempt++) {                            // :105-134
        int px, py; GetRandomPixelCircleWeighted(20u, W, H, x, y, seed, px, py);
        const uint32_t r = (uint32_t)py * W + (uint32_t)px;
        const SData& sr = F.scur[r];
        bool ok = !(dot3(sc.n1, sr.n1) < 0.9f) && !RejectDistance(sc.x1, sr.x1, init_orig, 0.1f) && IsValidReservoir(F.cur[r]) &&
                  !nz3(sr.L1) && sr.kind == 2u && sr.mID == sc.mID;
        if (ok) { candDI[nDI++] = r; M_sum_DI += (float)mcap(CAP, F.cur[r].M); }
    }
    for (uint32_t attempt = 0; attempt < TRIES && nGI < NC; attempt++) {                            // :144-186
        int px, py; GetRandomPixelCircleWeighted(20u, W, H, x, y, seed, px, py);
        const uint32_t r = (uint32_t)py * W + (uint32_t)px;
        const SData& sr = F.scur[r]; const ReservoirGI& gr = F.gcur[r];
        bool ok = m.Pr > 0.3f && !RejectDistance(sc.x1, sr.x1, init_orig, 0.1f) &&
                  !(dot3(normalize3(gr.xn - sc.x1), sc.n1) < 0.0f) && !(gr.w_sum > 5.0f) && IsValidReservoir_GI(gr) &&
                  !RejectJacobian(Jacobian_Reconnection(sr, sc, gr.xn, gr.nn), 5.0f) && !nz3(sr.L1) && sr.kind == 2u && sr.mID == sc.mID;
        if (ok) { candGI[nGI++] = r; M_sum_GI += (float)mcap(CAP, gr.M); }
    }
    ReservoirDI rc = F.cur[pix]; ReservoirGI gc = F.gcur[pix];
    const ReservoirDI can = rc; const ReservoirGI cang = gc;
    // GenPairwiseMIS_canonical, MIS_v7.hlsl:2-37
    float mi_c;
    {
        float c_M_min = (float)mcap(CAP, can.M), c_M_max = M_sum_DI - c_M_min;
        float p_c = GetP_Hat(c, sc.x1, sc.n1, can.x2, can.n2, can.L2, sc.o, m, false);
        float c_m_num = c_M_min * p_c;
        mi_c = c_M_min / M_sum_DI;
        for (uint32_t j = 0; j < nDI; j++) {
            const SData& sn = F.scur[candDI[j]];
            float n_M_min = (float)mcap(CAP, F.cur[candDI[j]].M);
            float p_from = GetP_Hat(c, sn.x1, sn.n1, can.x2, can.n2, can.L2, sn.o, m, true);
            float m_den = c_m_num + (c_M_max * p_from);
            if (m_den > 0.0f) mi_c += (n_M_min / M_sum_DI) * (c_m_num / m_den);
        }
    }
    float w_c = (mi_c * GetP_Hat(c, sc.x1, sc.n1, can.x2, can.n2, can.L2, sc.o, m, false)) * can.W;
    // GenPairwiseMIS_canonical_GI, MIS_GI_v6.hlsl:2-40
    float mi_c_gi;
    {
        float c_M_min = (float)mcap(CAP, cang.M), c_M_max = M_sum_GI - c_M_min;
        float p_c = LinearizeVector(GetP_Hat_GI(c, sc.x1, sc.n1, cang.xn, cang.nn, cang.E3, sc.o, m, false));
        float c_m_num = c_M_min * p_c;
        float m_c = c_M_min / M_sum_GI;
        for (uint32_t j = 0; j < nGI; j++) {
            const SData& sn = F.scur[candGI[j]];
            float n_M_min = (float)mcap(CAP, F.gcur[candGI[j]].M);
            float j_gi = Jacobian_Reconnection(sc, sn, cang.xn, cang.nn);
            float p_from = LinearizeVector(GetP_Hat_GI(c, sn.x1, sn.n1, cang.xn, cang.nn, cang.E3, sn.o, m, true)) * j_gi;
            float m_den = c_m_num + (c_M_max * p_from);
            if (m_den > 0.0f) m_c += (n_M_min / M_sum_GI) * (c_m_num / m_den);
        }
        mi_c_gi = fminf(fmaxf(m_c, 0.0f), 1.0f);
    }
    f3 f_c = GetP_Hat_GI(c, sc.x1, sc.n1, cang.xn, cang.nn, cang.E3, sc.o, m, false);
    float w_c_gi = (mi_c_gi * LinearizeVector(f_c)) * cang.W;
    rc.M = mcap(CAP, can.M); rc.w_sum = w_c;
    gc.M = mcap(CAP, cang.M); gc.w_sum = w_c_gi;
    for (uint32_t v = 0; v < nDI; v++) {                                                            // :252-290
        const uint32_t sp = candDI[v]; const SData& sn = F.scur[sp]; const ReservoirDI& rn = F.cur[sp];
        float mi_s = 0.0f;                                                                          // MIS_v7.hlsl:40-61
        {
            float c_M_min = (float)mcap(CAP, can.M);
            float p_c = GetP_Hat(c, sc.x1, sc.n1, can.x2, can.n2, can.L2, sc.o, m, false);
            float p_from = GetP_Hat(c, sn.x1, sn.n1, can.x2, can.n2, can.L2, sn.o, m, false);
            float m_num = (M_sum_DI - c_M_min) * p_from;
            float m_den = m_num + (c_M_min * p_c);
            if (m_den > 0.0f) mi_s = ((float)mcap(CAP, rn.M) / M_sum_DI) * (m_num / m_den);
        }
        float w_s = (mi_s * GetP_Hat(c, sc.x1, sc.n1, rn.x2, rn.n2, rn.L2, sc.o, m, false)) * rn.W;
        rc.M += mcap(CAP, rn.M);
        UpdateReservoir(rc, w_s, rn.x2, rn.n2, rn.L2, seed);
    }
    for (uint32_t v = 0; v < nGI; v++) {                                                            // :293-341
        const uint32_t sp = candGI[v]; const SData& sn = F.scur[sp]; const ReservoirGI& gn = F.gcur[sp];
        float mi_s = 0.0f;                                                                          // MIS_GI_v6.hlsl:43-76
        {
            float c_M_min = (float)mcap(CAP, cang.M);
            float p_c = LinearizeVector(GetP_Hat_GI(c, sc.x1, sc.n1, cang.xn, cang.nn, cang.E3, sc.o, m, false));
            float j = Jacobian_Reconnection(sc, sn, cang.xn, cang.nn);
            float p_from = LinearizeVector(GetP_Hat_GI(c, sn.x1, sn.n1, cang.xn, cang.nn, cang.E3, sn.o, m, false)) * j;
            float m_num = (M_sum_GI - c_M_min) * p_from;
            float m_den = m_num + (c_M_min * p_c);
            if (m_den > 0.0f) mi_s = fminf(fmaxf(((float)mcap(CAP, gn.M) / M_sum_GI) * (m_num / m_den), 0.0f), 1.0f);
        }
        float j_gi = Jacobian_Reconnection(sn, sc, gn.xn, gn.nn);
        f3 f_gi = GetP_Hat_GI(c, sc.x1, sc.n1, gn.xn, gn.nn, gn.E3, sc.o, m, true);
        float w_s = ((mi_s * LinearizeVector(f_gi)) * gn.W) * j_gi;
        if (j_gi != 0.0f) { gc.M += mcap(CAP, gn.M); UpdateReservoir_GI(gc, w_s, gn.xn, gn.nn, gn.E3, seed); }
    }
    float p_hat = GetP_Hat(c, sc.x1, sc.n1, rc.x2, rc.n2, rc.L2, sc.o, m, true);                    // :343-353
    rc.W = GetW(rc.w_sum, p_hat);
    f3 accumulation = ReconnectDI(c, sc.x1, sc.n1, rc.x2, rc.n2, rc.L2, sc.o, m) * rc.W;
    f3 f_fin = GetP_Hat_GI(c, sc.x1, sc.n1, gc.xn, gc.nn, gc.E3, sc.o, m, false);                   // :367-381
    gc.W = GetW(gc.w_sum, LinearizeVector(f_fin));
    accumulation = accumulation + f_fin * gc.W;
    F.last[pix] = rc; F.glast[pix] = gc; F.slast[pix] = sc;                                         // :434-436
    return accumulation;
}

// shaders/Common_v7.hlsl:353-376
inline float srgb1(float c) {
    if (c <= 0.0031308f) return 12.92f * c;
    return 1.055f * d_pow(c, 1.0f / 2.4f) - 0.055f;
}

}  // namespace

// ============================================================================================ C interface
extern "C" {

orc_scene* orc_scene_create(void) { return new orc_scene(); }
void orc_scene_destroy(orc_scene* s) { delete s; }

int orc_add_model(orc_scene* s, const orc_vertex* v, uint32_t nv, const uint32_t* idx, uint32_t ni, uint32_t material_id_offset) {
    Model m;
    m.pos.resize(nv); m.nrm.resize(nv);
    for (uint32_t i = 0; i < nv; i++) {
        m.pos[i] = mk3(v[i].position[0], v[i].position[1], v[i].position[2]);
        m.nrm[i] = mk3(v[i].normal_material[0], v[i].normal_material[1], v[i].normal_material[2]);
    }
    m.idx.assign(idx, idx + (ni / 3) * 3);
    m.mat_offset = material_id_offset;
    s->models.push_back(std::move(m));
    return (int)s->models.size() - 1;
}
void orc_set_material_ids(orc_scene* s, const uint32_t* ids, uint32_t n) { s->material_ids.assign(ids, ids + n); }
void orc_set_materials(orc_scene* s, const orc_material* m, uint32_t n) { s->materials.assign(m, m + n); }
void orc_set_instances(orc_scene* s, const uint32_t* model_ids, const orc_instance_props* props, uint32_t n) {
    s->instances.resize(n);
    for (uint32_t i = 0; i < n; i++) { s->instances[i].model = model_ids[i]; s->instances[i].p = props[i]; }
}
void orc_set_lights(orc_scene* s, const orc_light_tri* l, uint32_t n) { s->lights.assign(l, l + n); }

void orc_build(orc_scene* s) {
    for (auto& m : s->models) if (m.nodes.empty()) build_bvh2(m);       // (models never change after orc_add_model; orc_build runs again after every instance update)
    for (auto& in : s->instances) {
        const Model& m = s->models[in.model];
        Box w = box_empty();
        if (!m.idx.empty()) {
            for (int k = 0; k < 8; k++) {
                f3 p = mk3((k & 1) ? m.bounds.hi.x : m.bounds.lo.x, (k & 2) ? m.bounds.hi.y : m.bounds.lo.y, (k & 4) ? m.bounds.hi.z : m.bounds.lo.z);
                f4 q = mul44(in.p.objectToWorld, p.x, p.y, p.z, 1.0f);
                box_grow(w, mk3(q.x, q.y, q.z));
            }
            float sc = fmaxf(fmaxf(fabsf(w.lo.x), fabsf(w.hi.x)), fmaxf(fmaxf(fabsf(w.lo.y), fabsf(w.hi.y)), fmaxf(fabsf(w.lo.z), fabsf(w.hi.z))));
            box_pad(w, sc * 3.0517578125e-5f + 1e-30f);
        }
        in.wbox = w;
    }
    s->tlas.clear(); s->tlas_order.clear();
    if (s->instances.size() > 4) {
        std::vector<Box> ib; std::vector<f3> ic;
        for (const auto& in : s->instances) {
            Box w = in.wbox;
            if (!(w.lo.x <= w.hi.x)) { w.lo = mk3(0, 0, 0); w.hi = mk3(0, 0, 0); }      // an instance of an empty model
            ib.push_back(w);
            ic.push_back(mk3(0.5f * (w.lo.x + w.hi.x), 0.5f * (w.lo.y + w.hi.y), 0.5f * (w.lo.z + w.hi.z)));
        }
        build_tree(ib, ic, 0.0f, 2u, true, s->tlas_order, s->tlas);
    }
}

void orc_trace(orc_scene* s, const orc_ray* rays, uint32_t n, orc_hit* out, int any_hit, int mode) {
    orc_counters C; memset(&C, 0, sizeof C);
    for (uint32_t i = 0; i < n; i++) {
        Hit h;
        f3 o = mk3(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]);
        f3 d = mk3(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2]);
        bool hit = trace(*s, o, d, rays[i].tmin, rays[i].tmax, any_hit != 0, mode, h, C);
        if (hit) { out[i].t = h.t; out[i].u = h.b1; out[i].v = h.b2; out[i].prim = h.prim; out[i].inst = h.inst; }
        else { out[i].t = rays[i].tmax; out[i].u = out[i].v = 0.0f; out[i].prim = 0xFFFFFFFFu; out[i].inst = 0xFFFFFFFFu; }
    }
}

void orc_render_rows(orc_scene* s, const orc_config* cfg, const orc_camera* cam, uint32_t first_sample, uint32_t n_samples, uint32_t step,
                     uint32_t y0, uint32_t y1, float* accum, orc_counters* counters, int trace_mode);

void orc_render(orc_scene* s, const orc_config* cfg, const orc_camera* cam, uint32_t first_sample, uint32_t n_samples, uint32_t step,
                float* accum, orc_counters* counters, int trace_mode) {
    orc_render_rows(s, cfg, cam, first_sample, n_samples, step, 0, cfg->height, accum, counters, trace_mode);
}

// rows [y0, y1) only: the scene is read-only, so disjoint row ranges may run on different host threads
void orc_render_rows(orc_scene* s, const orc_config* cfg, const orc_camera* cam, uint32_t first_sample, uint32_t n_samples, uint32_t step,
                     uint32_t y0, uint32_t y1, float* accum, orc_counters* counters, int trace_mode) {
    orc_counters local; memset(&local, 0, sizeof local);
    Ctx c; c.S = s; c.cfg = *cfg; c.C = counters ? counters : &local; c.mode = trace_mode;
    if (step == 0) step = 1;
    if (y1 > cfg->height) y1 = cfg->height;
    y0 = ((y0 + step - 1) / step) * step;
    for (uint32_t y = y0; y < y1; y += step)
        for (uint32_t x = 0; x < cfg->width; x += step) {
            float* px = accum + 4 * ((size_t)y * cfg->width + x);
            for (uint32_t k = 0; k < n_samples; k++) {
                f3 C = (cfg->flags & ORC_FLAG_LEGACY_RR) ? legacy::RenderSample(c, *cam, x, y, first_sample + k)
                                                         : RenderSample(c, *cam, x, y, first_sample + k, nullptr);
                // Pass_spat_di_v7.hlsl:388-404: drop non-finite samples, sum += C, n += 1
                if (!any_nan_inf(C)) { px[0] += C.x; px[1] += C.y; px[2] += C.z; px[3] += 1.0f; }
            }
        }
}

orc_frames* orc_frames_create(uint32_t w, uint32_t h) {
    orc_frames* F = new orc_frames(); F->w = w; F->h = h;
    const size_t n = (size_t)w * h;
    ReservoirDI zr; memset(&zr, 0, sizeof zr); ReservoirGI zg; memset(&zg, 0, sizeof zg); SData zs; memset(&zs, 0, sizeof zs);
    F->cur.assign(n, zr); F->last.assign(n, zr); F->gcur.assign(n, zg); F->glast.assign(n, zg); F->scur.assign(n, zs); F->slast.assign(n, zs);
    return F;
}
void orc_frames_destroy(orc_frames* F) { delete F; }

// One frame of the reference's dispatch sequence (rdn/Renderer.cpp:611-673): RayGen (pass 1) for every pixel, RayGen2
// (temporal reuse), RayGen3 (spatial reuse, final shade, accumulation F20).  uint(time) := frame_index (D2).
void orc_render_frame(orc_scene* s, const orc_config* cfg, const orc_camera* cam, uint32_t frame_index, orc_frames* F, float* accum,
                      orc_counters* counters, int trace_mode) {
    orc_counters local; memset(&local, 0, sizeof local);
    Ctx c; c.S = s; c.cfg = *cfg; c.C = counters ? counters : &local; c.mode = trace_mode;
    const uint32_t W = cfg->width, H = cfg->height;
    for (uint32_t y = 0; y < H; y++)
        for (uint32_t x = 0; x < W; x++) {
            PixelDebug d; memset(&d, 0, sizeof d); d.hit_inst = 0xFFFFFFFFu;
            RenderSample(c, *cam, x, y, frame_index, &d);
            const size_t pix = (size_t)y * W + x;
            ReservoirDI zr; memset(&zr, 0, sizeof zr); ReservoirGI zg; memset(&zg, 0, sizeof zg); SData sd; memset(&sd, 0, sizeof sd);
            sd.kind = d.kind;
            if (d.kind == 1u) { sd.mID = d.mID & 0xFFFFu; sd.L1 = d.L1; sd.objID = d.hit_inst; }       // Pass_init_di_v7.hlsl:129-137
            if (d.kind == 2u) {
                sd.x1 = d.x1; sd.n1 = normalize3(d.n1); sd.o = -d.cam_d; sd.mID = d.mID & 0xFFFFu; sd.objID = d.hit_inst;   // :161-164
                zr = d.rdi; zg = d.rgi;
            }
            F->cur[pix] = zr; F->gcur[pix] = zg; F->scur[pix] = sd;
        }
    for (uint32_t y = 0; y < H; y++)
        for (uint32_t x = 0; x < W; x++) TemporalPixel(c, *cam, *F, x, y, frame_index);
    for (uint32_t y = 0; y < H; y++)
        for (uint32_t x = 0; x < W; x++) {
            f3 C = SpatialPixel(c, *cam, *F, x, y, frame_index);
            float* px = accum + 4 * ((size_t)y * W + x);
            if (!any_nan_inf(C)) { px[0] += C.x; px[1] += C.y; px[2] += C.z; px[3] += 1.0f; }      // Pass_spat_di_v7.hlsl:388-404
        }
}

// the *_last buffers after a frame, 40 floats per pixel:
// [0..3] x2,w_sum [4..7] n2,W [8..10] L2 [11] M | [12..15] xn,w_sum [16..19] nn,W [20..22] E3 [23] M |
// [24..26] x1 [27] mID [28..30] n1 [31] objID [32..34] o [35] kind [36..38] L1
void orc_frames_dump(const orc_frames* F, float* out) {
    const size_t n = (size_t)F->w * F->h;
    for (size_t i = 0; i < n; i++) {
        float* o = out + 40 * i; memset(o, 0, 40 * sizeof(float));
        const ReservoirDI& r = F->last[i]; const ReservoirGI& g = F->glast[i]; const SData& s = F->slast[i];
        auto put3 = [&](int k, f3 v) { o[k] = v.x; o[k + 1] = v.y; o[k + 2] = v.z; };
        put3(0, r.x2); o[3] = r.w_sum; put3(4, r.n2); o[7] = r.W; put3(8, r.L2); o[11] = (float)r.M;
        put3(12, g.xn); o[15] = g.w_sum; put3(16, g.nn); o[19] = g.W; put3(20, g.E3); o[23] = (float)g.M;
        put3(24, s.x1); o[27] = (float)s.mID; put3(28, s.n1); o[31] = (float)s.objID; put3(32, s.o); o[35] = (float)s.kind; put3(36, s.L1);
    }
}

void orc_resolve(const float* accum, uint32_t n_pixels, uint8_t* rgba8) {
    for (uint32_t i = 0; i < n_pixels; i++) {
        const float* a = accum + 4 * (size_t)i;
        float n = a[3];
        float c[3] = {a[0] / n, a[1] / n, a[2] / n};          // :405 averagedColor = sum / frameCount
        bool nan = isnan1(c[0]) || isnan1(c[1]) || isnan1(c[2]);
        bool inf = isinf1(c[0]) || isinf1(c[1]) || isinf1(c[2]);
        if (nan) { c[0] = 1; c[1] = 0; c[2] = 1; }             // :429-430
        if (inf) { c[0] = 0; c[1] = 1; c[2] = 1; }             // :431-432
        for (int k = 0; k < 3; k++) {
            float v = srgb1(c[k]);                              // :440
            v = saturate1(v);                                   // RGBA8_UNORM store clamps, then round-to-nearest
            rgba8[4 * i + k] = (uint8_t)(int)(v * 255.0f + 0.5f);
        }
        rgba8[4 * i + 3] = 255;
    }
}

void orc_kat_rng(uint32_t sx, uint32_t sy, uint32_t n, float* out, uint32_t* seed_out) {
    uint32_t seed[2] = {sx, sy};
    for (uint32_t i = 0; i < n; i++) out[i] = RandomFloat(seed);
    seed_out[0] = seed[0]; seed_out[1] = seed[1];
}
void orc_kat_seed(uint32_t x, uint32_t y, uint32_t pass, uint32_t sample, uint32_t* seed_out) { init_seed(x, y, pass, sample, seed_out); }
void orc_kat_sincos(const float* x, uint32_t n, float* s, float* c) { for (uint32_t i = 0; i < n; i++) d_sincos(x[i], &s[i], &c[i]); }
void orc_kat_half(const float* x, uint32_t n, float* out) { for (uint32_t i = 0; i < n; i++) out[i] = q16(x[i]); }
void orc_kat_pow(const float* x, float y, uint32_t n, float* out) { for (uint32_t i = 0; i < n; i++) out[i] = d_pow(x[i], y); }
uint32_t orc_kat_map_pixel(uint32_t w, uint32_t h, uint32_t x, uint32_t y) { return MapPixelID(w, h, x, y); }

void orc_kat_bsdf(const orc_scene* s, int op, uint32_t mat_id, const float* n, const float* in, const float* o, uint32_t* seed, float* out) {
    orc_counters C; memset(&C, 0, sizeof C);
    Ctx c; c.S = s; memset(&c.cfg, 0, sizeof c.cfg); c.C = &C; c.mode = 0;
    MatOpt m = make_matopt(fetch_material(*s, mat_id), mat_id);
    f3 N = mk3(n[0], n[1], n[2]), I = mk3(in[0], in[1], in[2]), O = mk3(o[0], o[1], o[2]);
    out[0] = out[1] = out[2] = out[3] = 0.0f;
    switch (op) {
        case 0: case 1: { f3 r = EvaluateBRDF(c, (uint32_t)op, m, N, I, O); out[0] = r.x; out[1] = r.y; out[2] = r.z; break; }
        case 2: case 3: out[0] = BRDF_PDF((uint32_t)(op - 2), m, N, I, O); break;
        case 4: CalculateStrategyProbabilities(c, m, O, N, out[0], out[1]); break;
        case 5: case 6: { f3 r = SampleBRDF((uint32_t)(op - 5), m, O, N, seed); out[0] = r.x; out[1] = r.y; out[2] = r.z; break; }
        case 7: { float p; out[0] = (float)SelectSamplingStrategy(c, m, O, N, seed, p); out[1] = p; break; }
        default: break;
    }
}

void orc_kat_camera_ray(const orc_config* cfg, const orc_camera* cam, uint32_t x, uint32_t y, float jx, float jy, float* o3d3) {
    f3 o, d; CameraRay(*cfg, *cam, x, y, jx, jy, o, d);
    o3d3[0] = o.x; o3d3[1] = o.y; o3d3[2] = o.z; o3d3[3] = d.x; o3d3[4] = d.y; o3d3[5] = d.z;
}

// out64 layout (floats; uints bit-cast): see tests/orc.py PIXEL_DEBUG_FIELDS
void orc_debug_pixel(orc_scene* s, const orc_config* cfg, const orc_camera* cam, uint32_t x, uint32_t y, uint32_t sample, float* out, int trace_mode) {
    orc_counters C; memset(&C, 0, sizeof C);
    Ctx c; c.S = s; c.cfg = *cfg; c.C = &C; c.mode = trace_mode;
    PixelDebug d; memset(&d, 0, sizeof d);
    d.hit_inst = d.hit_prim = 0xFFFFFFFFu;
    f3 Cc = RenderSample(c, *cam, x, y, sample, &d);
    auto put3 = [&](int i, f3 v) { out[i] = v.x; out[i + 1] = v.y; out[i + 2] = v.z; };
    auto putu = [&](int i, uint32_t u) { memcpy(&out[i], &u, 4); };
    memset(out, 0, 64 * sizeof(float));
    put3(0, d.cam_o); put3(3, d.cam_d); putu(6, d.hit_inst); putu(7, d.hit_prim); out[8] = d.hit_t; putu(9, d.mID);
    put3(10, d.x1); put3(13, d.n1);
    put3(16, d.rdi.x2); out[19] = d.rdi.w_sum; put3(20, d.rdi.n2); out[23] = d.rdi.W; put3(24, d.rdi.L2);
    put3(27, d.rgi.xn); out[30] = d.rgi.w_sum; put3(31, d.rgi.nn); out[34] = d.rgi.W; put3(35, d.rgi.E3);
    out[38] = d.p_hat; put3(39, Cc); putu(42, d.seed_end[0]); putu(43, d.seed_end[1]); put3(44, d.acc_L);
    putu(47, (uint32_t)C.closest_rays); putu(48, (uint32_t)C.shadow_rays);
}

}  // extern "C"
