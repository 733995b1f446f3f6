// traverse.cuh — closest-hit / any-hit traversal of the two-level compressed 8-wide BVH.
//
// Replaces the fixed-function TraceRay of the reference (no source exists: D3D12 driver + RT cores;
// call sites shaders/Pass_init_di_v7.hlsl:91-99, Sampler_v7.hlsl:86-104,223-229,428-434,
// Path_Sampler_v7.hlsl:45-52,271-283).  Contract:
//   * opaque triangles, no culling (two-sided), instance mask 0xFF;
//   * a triangle is hit iff TMin < t < TMax (strict), t measured along the un-renormalised object-space
//     direction so it equals the world-space parameter;
//   * closest hit = lexicographic minimum of (t, instance, primitive) — traversal-order independent;
//   * any-hit returns as soon as one triangle is hit.
// The ray/triangle test is the watertight edge-function test; its operation order is part of the contract
// (hit IDs are compared bit-exactly against the CPU oracle).  Box tests only have to be conservative.
#pragma once
#include "common.cuh"

namespace rtx {

#define RTX_STACK_SIZE 40
// A full stack drops the entry (wrong image, no memory fault) and raises g_stack_overflow, which every API call checks.
__device__ unsigned int g_stack_overflow;
#define RTX_PUSH(v)                                                   \
    do {                                                              \
        if (sp < RTX_STACK_SIZE) stack[sp++] = (v);                   \
        else g_stack_overflow = 1u;                                   \
    } while (0)

__device__ __forceinline__ uint32_t sign_extend_s8x4(uint32_t x) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(0u), "r"(0xba98u));
    return d;
}
__device__ __forceinline__ uint32_t extract_byte(uint32_t x, int i) { return (x >> (i * 8)) & 0xffu; }
// byte i of x as the float 1 + q/256 (mantissa bits 15..22): full-rate integer ops instead of the quarter-rate I2F
__device__ __forceinline__ float byte_as_unit_float(uint32_t x, int i) {
    return __uint_as_float(0x3F800000u | (((x >> (i * 8)) & 0xffu) << 15));
}

struct RaySpace {          // the ray in the space currently traversed + derived constants
    float ox, oy, oz;
    float ix, iy, iz;      // approximate reciprocal direction, clamped away from 0 (box tests only: conservative, not reproducible)
    float Sx, Sy, Sz;      // watertight shear constants (BLAS space only)
    uint32_t k;            // kx | ky<<2 | kz<<4
    uint32_t octinv4;      // (dx>=0 | dy>=0 <<1 | dz>=0 <<2) * 0x01010101
};

__device__ __forceinline__ float sel3(float x, float y, float z, uint32_t k) { return k == 0 ? x : (k == 1 ? y : z); }

__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// box-test part: origin, reciprocal direction, octant
__device__ __forceinline__ void setup_box(RaySpace& r, float ox, float oy, float oz, float dx, float dy, float dz) {
    r.ox = ox; r.oy = oy; r.oz = oz;
    const float tiny = 1e-20f;
    r.ix = rcp_approx(fabsf(dx) > tiny ? dx : copysignf(tiny, dx));
    r.iy = rcp_approx(fabsf(dy) > tiny ? dy : copysignf(tiny, dy));
    r.iz = rcp_approx(fabsf(dz) > tiny ? dz : copysignf(tiny, dz));
    const uint32_t oct = (dx >= 0.0f ? 1u : 0u) | (dy >= 0.0f ? 2u : 0u) | (dz >= 0.0f ? 4u : 0u);
    r.octinv4 = oct * 0x01010101u;
}
// triangle-test part (same rule as the oracle: first maximum in x,y,z order; swap kx,ky if d[kz] < 0;
// Sz = 1/d[kz] (IEEE), Sx = d[kx]*Sz, Sy = d[ky]*Sz)
__device__ __forceinline__ void setup_tri(RaySpace& r, float dx, float dy, float dz) {
    const float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
    uint32_t kz = 0; float m = ax;
    if (ay > m) { kz = 1; m = ay; }
    if (az > m) { kz = 2; }
    uint32_t kx = (kz == 2u) ? 0u : kz + 1u, ky = (kx == 2u) ? 0u : kx + 1u;
    const float dkz = sel3(dx, dy, dz, kz);
    if (dkz < 0.0f) { const uint32_t t = kx; kx = ky; ky = t; }
    r.k = kx | (ky << 2) | (kz << 4);
    r.Sz = 1.0f / dkz;
    r.Sx = sel3(dx, dy, dz, kx) * r.Sz;
    r.Sy = sel3(dx, dy, dz, ky) * r.Sz;
}

// Watertight ray/triangle test; identical operation order to oracle/rtx_oracle.cpp tri_test().
__device__ __forceinline__ bool tri_test(const RaySpace& r, float4 a, float4 b, float4 c, float tmin, float tmax,
                                         float& t, float& b1, float& b2) {
    const uint32_t kx = r.k & 3u, ky = (r.k >> 2) & 3u, kz = r.k >> 4;
    float Ax_ = a.x - r.ox, Ay_ = a.y - r.oy, Az_ = a.z - r.oz;
    float Bx_ = b.x - r.ox, By_ = b.y - r.oy, Bz_ = b.z - r.oz;
    float Cx_ = c.x - r.ox, Cy_ = c.y - r.oy, Cz_ = c.z - r.oz;
    float Akz = sel3(Ax_, Ay_, Az_, kz), Bkz = sel3(Bx_, By_, Bz_, kz), Ckz = sel3(Cx_, Cy_, Cz_, kz);
    float Ax = sel3(Ax_, Ay_, Az_, kx) - r.Sx * Akz, Ay = sel3(Ax_, Ay_, Az_, ky) - r.Sy * Akz;
    float Bx = sel3(Bx_, By_, Bz_, kx) - r.Sx * Bkz, By = sel3(Bx_, By_, Bz_, ky) - r.Sy * Bkz;
    float Cx = sel3(Cx_, Cy_, Cz_, kx) - r.Sx * Ckz, Cy = sel3(Cx_, Cy_, Cz_, ky) - r.Sy * Ckz;
    float U = Cx * By - Cy * Bx;
    float V = Ax * Cy - Ay * Cx;
    float W = Bx * Ay - By * Ax;
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        U = (float)(__dsub_rn(__dmul_rn((double)Cx, (double)By), __dmul_rn((double)Cy, (double)Bx)));
        V = (float)(__dsub_rn(__dmul_rn((double)Ax, (double)Cy), __dmul_rn((double)Ay, (double)Cx)));
        W = (float)(__dsub_rn(__dmul_rn((double)Bx, (double)Ay), __dmul_rn((double)By, (double)Ax)));
    }
    const bool neg = (U < 0.0f) | (V < 0.0f) | (W < 0.0f);
    const bool pos = (U > 0.0f) | (V > 0.0f) | (W > 0.0f);
    float det = (U + V) + W;
    if ((neg & pos) | (det == 0.0f)) return false;
    float Az = r.Sz * Akz, Bz = r.Sz * Bkz, Cz = r.Sz * Ckz;
    float T = (U * Az + V * Bz) + W * Cz;
    float rcp = 1.0f / det;
    float tt = T * rcp;
    if (!(tt > tmin && tt < tmax)) return false;
    t = tt; b1 = V * rcp; b2 = W * rcp;
    return true;
}

// Intersects the 8 quantised child boxes of one node; returns the hit mask in the traversal's group format:
// inner children set bit 24 + (slot ^ octinv) (front-to-back priority), leaf children set their unary triangle
// bits at their offset in the node's primitive range.
__device__ __forceinline__ uint32_t intersect_node(const RaySpace& r, uint4 n0, uint4 n1, uint4 n2, uint4 n3, uint4 n4,
                                                   float tmin, float tmax) {
    const float px = __uint_as_float(n0.x), py = __uint_as_float(n0.y), pz = __uint_as_float(n0.z);
    const uint32_t e = n0.w;
    const float ax = __uint_as_float((e & 0xffu) << 23) * r.ix;
    const float ay = __uint_as_float(((e >> 8) & 0xffu) << 23) * r.iy;
    const float az = __uint_as_float(((e >> 16) & 0xffu) << 23) * r.iz;
    // plane = p + q * 2^e  ->  t = (1 + q/256) * (256 a) + (b - 256 a); the rounding of (b - 256 a) is 2^-16 of a grid cell
    const float ax2 = ax * 256.0f, ay2 = ay * 256.0f, az2 = az * 256.0f;
    const float bx = (px - r.ox) * r.ix - ax2, by = (py - r.oy) * r.iy - ay2, bz = (pz - r.oz) * r.iz - az2;
    const bool nx = !(r.octinv4 & 1u), ny = !(r.octinv4 & 2u), nz = !(r.octinv4 & 4u);
    uint32_t hitmask = 0;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const uint32_t meta4 = h ? n1.w : n1.z;
        const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        const uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
        const uint32_t bit_index4 = (meta4 ^ (r.octinv4 & inner_mask4)) & 0x1f1f1f1fu;
        const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
        const uint32_t qlox = h ? n2.y : n2.x, qloy = h ? n2.w : n2.z, qloz = h ? n3.y : n3.x;
        const uint32_t qhix = h ? n3.w : n3.z, qhiy = h ? n4.y : n4.x, qhiz = h ? n4.w : n4.z;
        const uint32_t xn = nx ? qhix : qlox, xf = nx ? qlox : qhix;
        const uint32_t yn = ny ? qhiy : qloy, yf = ny ? qloy : qhiy;
        const uint32_t zn = nz ? qhiz : qloz, zf = nz ? qloz : qhiz;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float t0x = __fmaf_rn(byte_as_unit_float(xn, j), ax2, bx);
            float t0y = __fmaf_rn(byte_as_unit_float(yn, j), ay2, by);
            float t0z = __fmaf_rn(byte_as_unit_float(zn, j), az2, bz);
            float t1x = __fmaf_rn(byte_as_unit_float(xf, j), ax2, bx);
            float t1y = __fmaf_rn(byte_as_unit_float(yf, j), ay2, by);
            float t1z = __fmaf_rn(byte_as_unit_float(zf, j), az2, bz);
            float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, tmin));
            float tf = fminf(fminf(t1x, t1y), fminf(t1z, tmax));
            if (tn <= tf) {
                hitmask |= extract_byte(child_bits4, j) << extract_byte(bit_index4, j);
            }
        }
    }
    return hitmask;
}

struct HitRec {
    float t, b1, b2;
    uint32_t prim, inst;
};

__device__ __forceinline__ bool hit_better(float t, uint32_t inst, uint32_t prim, const HitRec& h) {
    if (t < h.t) return true;
    if (t > h.t) return false;
    if (inst < h.inst) return true;
    if (inst > h.inst) return false;
    return prim < h.prim;
}

// Traversal state of one ray (registers) — lets a persistent warp retire and replace single lanes.
struct Trav {
    RaySpace r;                 // current space
    float wox, woy, woz, wdx, wdy, wdz;   // world-space ray
    float wix, wiy, wiz; uint32_t woct4;   // world-space box-test constants (restored when a BLAS is left)
    float tmin, tmax;
    HitRec h;
    const uint4* nodes; const float4* prims;
    uint2 G;
    uint32_t cur_inst;
    int sp;
    bool in_blas;
};

__device__ __forceinline__ void trav_init(Trav& T, const SceneAS& S, float4 o_tmin, float4 d_tmax) {
    T.wox = o_tmin.x; T.woy = o_tmin.y; T.woz = o_tmin.z; T.wdx = d_tmax.x; T.wdy = d_tmax.y; T.wdz = d_tmax.z;
    T.tmin = o_tmin.w; T.tmax = d_tmax.w;
    setup_box(T.r, T.wox, T.woy, T.woz, T.wdx, T.wdy, T.wdz);
    T.wix = T.r.ix; T.wiy = T.r.iy; T.wiz = T.r.iz; T.woct4 = T.r.octinv4;
    T.h.t = d_tmax.w; T.h.b1 = 0.0f; T.h.b2 = 0.0f; T.h.prim = 0xFFFFFFFFu; T.h.inst = 0xFFFFFFFFu;
    T.nodes = S.tlas_nodes; T.prims = S.inst_recs;
    T.G = make_uint2(0u, 0x80000000u);
    T.cur_inst = 0; T.sp = 0; T.in_blas = false;
}

// One traversal step: at most one node intersection, then the node's leaf primitives, then a pop.
// Returns true when the ray is finished.
template <bool ANY_HIT, bool STATS>
__device__ __forceinline__ bool trav_step(Trav& T, const SceneAS& S, uint2* stack, unsigned int* c_nodes, unsigned int* c_tris,
                                          unsigned int* c_insts, int postpone_th) {
    uint2 G = T.G, Gt;
    int sp = T.sp;
    bool fresh = false;
    if (G.y & 0xff000000u) {
        fresh = true;
        const uint32_t bit = 31u - __clz(G.y);
        G.y &= ~(1u << bit);
        const uint32_t imask = G.y & 0xffu;   // low byte carries the node's imask for relative indexing
        if (G.y & 0xff000000u) RTX_PUSH(G);
        const uint32_t slot = (bit - 24u) ^ (T.r.octinv4 & 0xffu);
        const uint32_t rel = __popc(imask & ~(0xffffffffu << slot));
        const uint4* np = T.nodes + (size_t)(G.x + rel) * 5;
        const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
        if (STATS) (*c_nodes)++;
        const uint32_t hm = intersect_node(T.r, n0, n1, n2, n3, n4, T.tmin, ANY_HIT ? T.tmax : T.h.t);
        G.x = n1.x;
        Gt.x = n1.y;
        G.y = (hm & 0xff000000u) | (n0.w >> 24);
        Gt.y = hm & 0x00ffffffu;
    } else {
        Gt = G;
        G = make_uint2(0u, 0u);
    }

    while (Gt.y != 0u) {
        // Triangle postponing: when only a few lanes of the warp still have leaf primitives to test, park the group on the
        // stack and go on with node tests; it is tested when popped (never parked twice), together with more lanes.
        if (fresh && T.in_blas && (G.y & 0xff000000u) && __popc(__activemask()) < postpone_th) {
            const uint2 keep = G;
            RTX_PUSH(Gt);
            G = keep;
            break;
        }
        const uint32_t bit = 31u - __clz(Gt.y);
        Gt.y &= ~(1u << bit);
        if (T.in_blas) {
            const float4* tp = T.prims + (size_t)(Gt.x + bit) * 3;
            const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
            if (STATS) (*c_tris)++;
            float t, b1, b2;
            if (tri_test(T.r, a, b, c, T.tmin, T.tmax, t, b1, b2)) {
                const uint32_t prim = __float_as_uint(a.w);
                if (ANY_HIT) {
                    T.h.t = t; T.h.b1 = b1; T.h.b2 = b2; T.h.prim = prim; T.h.inst = T.cur_inst;
                    return true;
                }
                if (T.h.inst == 0xFFFFFFFFu || hit_better(t, T.cur_inst, prim, T.h)) {
                    T.h.t = t; T.h.b1 = b1; T.h.b2 = b2; T.h.prim = prim; T.h.inst = T.cur_inst;
                }
            }
        } else {
            // instance leaf: enter the BLAS.  Save what is left of this level, then a sentinel.
            const float4* ip = T.prims + (size_t)(Gt.x + bit) * 4;
            const float4 r0 = __ldg(ip), r1 = __ldg(ip + 1), r2 = __ldg(ip + 2), r3 = __ldg(ip + 3);
            if (STATS) (*c_insts)++;
            if (Gt.y) RTX_PUSH(Gt);
            if (G.y & 0xff000000u) RTX_PUSH(G);
            RTX_PUSH(make_uint2(0xffffffffu, 0u));
            const float tox = ((r0.x * T.wox + r0.y * T.woy) + r0.z * T.woz) + r0.w * 1.0f;
            const float toy = ((r1.x * T.wox + r1.y * T.woy) + r1.z * T.woz) + r1.w * 1.0f;
            const float toz = ((r2.x * T.wox + r2.y * T.woy) + r2.z * T.woz) + r2.w * 1.0f;
            const float tdx = ((r0.x * T.wdx + r0.y * T.wdy) + r0.z * T.wdz) + r0.w * 0.0f;
            const float tdy = ((r1.x * T.wdx + r1.y * T.wdy) + r1.z * T.wdz) + r1.w * 0.0f;
            const float tdz = ((r2.x * T.wdx + r2.y * T.wdy) + r2.z * T.wdz) + r2.w * 0.0f;
            setup_box(T.r, tox, toy, toz, tdx, tdy, tdz);
            setup_tri(T.r, tdx, tdy, tdz);
            const BlasRef br = S.blas[__float_as_uint(r3.x)];
            T.nodes = br.nodes; T.prims = br.tris;
            T.cur_inst = __float_as_uint(r3.y);
            T.in_blas = true;
            G = make_uint2(0u, 0x80000000u);
            break;
        }
    }

    if ((G.y & 0xff000000u) == 0u) {
        for (;;) {   // pop
            if (sp == 0) { T.sp = 0; return true; }
            G = stack[--sp];
            if (G.x == 0xffffffffu && G.y == 0u) {   // sentinel: back to the TLAS / world space
                T.r.ox = T.wox; T.r.oy = T.woy; T.r.oz = T.woz; T.r.ix = T.wix; T.r.iy = T.wiy; T.r.iz = T.wiz; T.r.octinv4 = T.woct4;
                T.nodes = S.tlas_nodes; T.prims = S.inst_recs; T.in_blas = false;
                continue;
            }
            break;
        }
    }
    T.G = G; T.sp = sp;
    return false;
}

}  // namespace rtx
