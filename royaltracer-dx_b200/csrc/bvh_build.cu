// bvh_build.cu — GPU builder of the compressed 8-wide BVH (BLAS over triangles, TLAS over instances).
//
// Replaces ID3D12GraphicsCommandList4::BuildRaytracingAccelerationStructure as driven by
// rdn/nv_helpers_dx12/BottomLevelASGenerator.cpp:97-117,235 and TopLevelASGenerator.cpp:181-199,240
// (the reference only describes geometry to the driver; the driver's builder has no source).
//
// Pipeline (all on the context stream, no host data structures):
//   1. primitive boxes + scene bounds        (k_tri_boxes, atomics on order-preserving uints)
//   2. 63-bit Morton codes of box centres     (k_morton)  + radix sort (cub::DeviceRadixSort — plumbing)
//   3. binary radix tree over the sorted keys (k_lbvh, Karras 2012) and bottom-up box fit (k_fit)
//   4. surface-area-guided collapse to 8-wide nodes: every wide node repeatedly opens the child with the largest
//      surface area until it has 8 children; subtrees of <= 3 primitives become leaves; children are assigned to
//      octant slots so that (slot ^ ray octant) is a front-to-back order; boxes are quantised to 8 bits per plane
//      in the node's local power-of-two grid (k_collapse, one launch per tree level).
// Boxes are padded by 2^-15 of the model's coordinate scale so that the (bit-exact, contract-defining)
// ray/triangle test never reports a hit the conservative box tests culled.
#include <cub/device/device_radix_sort.cuh>

#include "trace.h"

namespace rtx {

#define CKE(call)                            \
    do {                                     \
        cudaError_t e__ = (call);            \
        if (e__ != cudaSuccess) return e__;  \
    } while (0)

#define MAX_LEAF 3

struct BuildCounters {
    unsigned int nodes, prims, tasks_out;
    unsigned int bounds[6];   // order-preserving encodings: lo xyz (min), hi xyz (max)
};

__device__ __forceinline__ unsigned int enc_f(float f) {
    unsigned int b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __host__ __forceinline__ float dec_f(unsigned int e) {
    unsigned int b = (e & 0x80000000u) ? (e & 0x7fffffffu) : ~e;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f; memcpy(&f, &b, 4); return f;
#endif
}

__global__ void k_init_counters(BuildCounters* c) {
    c->nodes = 1; c->prims = 0; c->tasks_out = 0;
    c->bounds[0] = c->bounds[1] = c->bounds[2] = 0xffffffffu;
    c->bounds[3] = c->bounds[4] = c->bounds[5] = 0u;
}

__device__ __forceinline__ void reduce_bounds(BuildCounters* c, float lx, float ly, float lz, float hx, float hy, float hz, bool valid) {
    unsigned int v[6];
    v[0] = valid ? enc_f(lx) : 0xffffffffu; v[1] = valid ? enc_f(ly) : 0xffffffffu; v[2] = valid ? enc_f(lz) : 0xffffffffu;
    v[3] = valid ? enc_f(hx) : 0u; v[4] = valid ? enc_f(hy) : 0u; v[5] = valid ? enc_f(hz) : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; k++) v[k] = min(v[k], __shfl_xor_sync(0xffffffffu, v[k], o));
#pragma unroll
        for (int k = 3; k < 6; k++) v[k] = max(v[k], __shfl_xor_sync(0xffffffffu, v[k], o));
    }
    if ((threadIdx.x & 31) == 0) {
        for (int k = 0; k < 3; k++) atomicMin(&c->bounds[k], v[k]);
        for (int k = 3; k < 6; k++) atomicMax(&c->bounds[k], v[k]);
    }
}

__global__ void k_tri_boxes(const uint8_t* __restrict__ verts, const uint32_t* __restrict__ idx, uint32_t n,
                            float4* __restrict__ lo, float4* __restrict__ hi, BuildCounters* c) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = t < n;
    float lx = 0, ly = 0, lz = 0, hx = 0, hy = 0, hz = 0;
    if (valid) {
        const float* a = (const float*)(verts + (size_t)idx[3 * t] * 28);
        const float* b = (const float*)(verts + (size_t)idx[3 * t + 1] * 28);
        const float* d = (const float*)(verts + (size_t)idx[3 * t + 2] * 28);
        lx = fminf(a[0], fminf(b[0], d[0])); hx = fmaxf(a[0], fmaxf(b[0], d[0]));
        ly = fminf(a[1], fminf(b[1], d[1])); hy = fmaxf(a[1], fmaxf(b[1], d[1]));
        lz = fminf(a[2], fminf(b[2], d[2])); hz = fmaxf(a[2], fmaxf(b[2], d[2]));
        lo[t] = make_float4(lx, ly, lz, 0.0f);
        hi[t] = make_float4(hx, hy, hz, 0.0f);
    }
    reduce_bounds(c, lx, ly, lz, hx, hy, hz, valid);
}

__global__ void k_box_bounds(const float4* __restrict__ lo, const float4* __restrict__ hi, uint32_t n, BuildCounters* c) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = t < n;
    float4 l = valid ? lo[t] : make_float4(0, 0, 0, 0), h = valid ? hi[t] : make_float4(0, 0, 0, 0);
    reduce_bounds(c, l.x, l.y, l.z, h.x, h.y, h.z, valid);
}

__device__ __forceinline__ unsigned long long expand21(unsigned long long v) {
    v &= 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

__global__ void k_morton(const float4* __restrict__ lo, const float4* __restrict__ hi, uint32_t n, const BuildCounters* c,
                         unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    float blx = dec_f(c->bounds[0]), bly = dec_f(c->bounds[1]), blz = dec_f(c->bounds[2]);
    float ex = dec_f(c->bounds[3]) - blx, ey = dec_f(c->bounds[4]) - bly, ez = dec_f(c->bounds[5]) - blz;
    float4 l = lo[t], h = hi[t];
    float cx = 0.5f * (l.x + h.x), cy = 0.5f * (l.y + h.y), cz = 0.5f * (l.z + h.z);
    float fx = ex > 0.0f ? (cx - blx) / ex : 0.0f, fy = ey > 0.0f ? (cy - bly) / ey : 0.0f, fz = ez > 0.0f ? (cz - blz) / ez : 0.0f;
    unsigned long long qx = (unsigned long long)fminf(fmaxf(fx * 2097152.0f, 0.0f), 2097151.0f);
    unsigned long long qy = (unsigned long long)fminf(fmaxf(fy * 2097152.0f, 0.0f), 2097151.0f);
    unsigned long long qz = (unsigned long long)fminf(fmaxf(fz * 2097152.0f, 0.0f), 2097151.0f);
    keys[t] = (expand21(qx) << 2) | (expand21(qy) << 1) | expand21(qz);
    vals[t] = t;
}

// ---- binary radix tree (Karras 2012).  Node ids: internal i in [0, n-2], leaf j -> (n-1)+j.
__device__ __forceinline__ int lcp(const unsigned long long* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    unsigned long long a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll(a ^ b);
}

__global__ void k_lbvh(const unsigned long long* __restrict__ keys, int n, int2* __restrict__ child, int2* __restrict__ range,
                       int* __restrict__ parent) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (lcp(keys, n, i, i + 1) - lcp(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = lcp(keys, n, i, i - d);
    int lmax = 2;
    while (lcp(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (lcp(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = lcp(keys, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) / 2;
        if (lcp(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int first = min(i, j), last = max(i, j);
    int left = (first == gamma) ? (n - 1 + gamma) : gamma;
    int right = (last == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
    child[i] = make_int2(left, right);
    range[i] = make_int2(first, last);
    parent[left] = i;
    parent[right] = i;
    if (i == 0) parent[0] = -1;
}

__global__ void k_fit(const float4* __restrict__ plo, const float4* __restrict__ phi, const uint32_t* __restrict__ vals, int n,
                      const int2* __restrict__ child, const int* __restrict__ parent, unsigned int* __restrict__ flags,
                      float4* __restrict__ nlo, float4* __restrict__ nhi, const BuildCounters* c) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    float s = 0.0f;
    for (int k = 0; k < 6; k++) s = fmaxf(s, fabsf(dec_f(c->bounds[k])));
    const float pad = s * 3.0517578125e-5f + 1e-30f;
    uint32_t p = vals[j];
    float4 l = plo[p], h = phi[p];
    l.x -= pad; l.y -= pad; l.z -= pad; h.x += pad; h.y += pad; h.z += pad;
    int id = n - 1 + j;
    nlo[id] = l; nhi[id] = h;
    int par = parent[id];
    while (par >= 0) {
        __threadfence();
        unsigned int old = atomicAdd(&flags[par], 1u);
        if (old == 0u) return;
        int2 ch = child[par];
        float4 al = __ldcg(&nlo[ch.x]), ah = __ldcg(&nhi[ch.x]), bl = __ldcg(&nlo[ch.y]), bh = __ldcg(&nhi[ch.y]);
        l = make_float4(fminf(al.x, bl.x), fminf(al.y, bl.y), fminf(al.z, bl.z), 0.0f);
        h = make_float4(fmaxf(ah.x, bh.x), fmaxf(ah.y, bh.y), fmaxf(ah.z, bh.z), 0.0f);
        nlo[par] = l; nhi[par] = h;
        id = par;
        par = parent[id];
    }
}

// ---- collapse to 8-wide compressed nodes
struct CollapseArgs {
    const int2* child; const int2* range;
    const float4* nlo; const float4* nhi;
    const uint32_t* vals;
    int n;
    uint4* out_nodes; float4* out_prims;
    BuildCounters* ctr;
};

__device__ __forceinline__ float box_area(float4 l, float4 h) {
    float dx = h.x - l.x, dy = h.y - l.y, dz = h.z - l.z;
    return 2.0f * (dx * dy + dy * dz + dz * dx);
}

struct TriSource {       // 3 x float4 per primitive
    const uint8_t* verts; const uint32_t* idx;
    __device__ __forceinline__ void write(float4* dst, uint32_t prim) const {
        const float* a = (const float*)(verts + (size_t)idx[3 * prim] * 28);
        const float* b = (const float*)(verts + (size_t)idx[3 * prim + 1] * 28);
        const float* d = (const float*)(verts + (size_t)idx[3 * prim + 2] * 28);
        dst[0] = make_float4(a[0], a[1], a[2], __uint_as_float(prim));
        dst[1] = make_float4(b[0], b[1], b[2], 0.0f);
        dst[2] = make_float4(d[0], d[1], d[2], 0.0f);
    }
};
struct RecSource {       // 4 x float4 per primitive
    const float4* recs;
    __device__ __forceinline__ void write(float4* dst, uint32_t prim) const {
        for (int k = 0; k < 4; k++) dst[k] = recs[(size_t)prim * 4 + k];
    }
};

__device__ __forceinline__ int exp_for_extent(float ext) {
    // smallest e with 255 * 2^e >= ext (conservatively one step larger when in doubt); biased into [1,254]
    float q = ext / 255.0f;
    int e = 0;
    if (q > 0.0f) { frexpf(q, &e); } else e = -126;
    int biased = e + 127;
    return max(1, min(254, biased));
}

template <int PRIM_F4, typename Source>
__device__ void emit_node(const CollapseArgs& A, const Source& src, const int* cid, int cnt, float4 plo, float4 phi,
                          uint32_t out_idx, uint2* tasks_out) {
    const int n = A.n;
    // slot assignment: greedy on cost[c][s] = dot(centre_c - centre_parent, sign vector of s)
    float ccx[8], ccy[8], ccz[8];
    const float pcx = 0.5f * (plo.x + phi.x), pcy = 0.5f * (plo.y + phi.y), pcz = 0.5f * (plo.z + phi.z);
    for (int c = 0; c < cnt; c++) {
        float4 l = A.nlo[cid[c]], h = A.nhi[cid[c]];
        ccx[c] = 0.5f * (l.x + h.x) - pcx; ccy[c] = 0.5f * (l.y + h.y) - pcy; ccz[c] = 0.5f * (l.z + h.z) - pcz;
    }
    int slot_child[8];
    for (int s = 0; s < 8; s++) slot_child[s] = -1;
    unsigned child_done = 0, slot_done = 0;
    for (int it = 0; it < cnt; it++) {
        float best = -INFINITY; int bc = -1, bs = -1;
        for (int c = 0; c < cnt; c++) {
            if (child_done & (1u << c)) continue;
            for (int s = 0; s < 8; s++) {
                if (slot_done & (1u << s)) continue;
                float cost = ((s & 1) ? ccx[c] : -ccx[c]) + ((s & 2) ? ccy[c] : -ccy[c]) + ((s & 4) ? ccz[c] : -ccz[c]);
                if (cost > best) { best = cost; bc = c; bs = s; }
            }
        }
        slot_child[bs] = bc; child_done |= 1u << bc; slot_done |= 1u << bs;
    }
    // classify + count
    unsigned imask = 0; int n_inner = 0, n_prims = 0;
    int cnt_of[8];
    for (int s = 0; s < 8; s++) {
        cnt_of[s] = 0;
        int c = slot_child[s];
        if (c < 0) continue;
        int id = cid[c];
        int pc = (id >= n - 1) ? 1 : (A.range[id].y - A.range[id].x + 1);
        if (pc > MAX_LEAF) { imask |= 1u << s; n_inner++; }
        else { cnt_of[s] = pc; n_prims += pc; }
    }
    const uint32_t child_base = n_inner ? atomicAdd(&A.ctr->nodes, (unsigned)n_inner) : 0u;
    const uint32_t prim_base = n_prims ? atomicAdd(&A.ctr->prims, (unsigned)n_prims) : 0u;
    const uint32_t task_base = n_inner ? atomicAdd(&A.ctr->tasks_out, (unsigned)n_inner) : 0u;
    // quantisation grid
    const int ex = exp_for_extent(phi.x - plo.x), ey = exp_for_extent(phi.y - plo.y), ez = exp_for_extent(phi.z - plo.z);
    const float sx = __uint_as_float((unsigned)ex << 23), sy = __uint_as_float((unsigned)ey << 23), sz = __uint_as_float((unsigned)ez << 23);
    unsigned char meta[8], qlo[3][8], qhi[3][8];
    int inner_rank = 0, prim_off = 0;
    for (int s = 0; s < 8; s++) {
        meta[s] = 0;
        for (int a = 0; a < 3; a++) { qlo[a][s] = 0; qhi[a][s] = 0; }
        int c = slot_child[s];
        if (c < 0) continue;
        int id = cid[c];
        float4 l = A.nlo[id], h = A.nhi[id];
        const float lov[3] = {l.x, l.y, l.z}, hiv[3] = {h.x, h.y, h.z}, pv[3] = {plo.x, plo.y, plo.z}, sv[3] = {sx, sy, sz};
        for (int a = 0; a < 3; a++) {
            int ql = (int)floorf((lov[a] - pv[a]) / sv[a]);
            int qh = (int)ceilf((hiv[a] - pv[a]) / sv[a]);
            ql = max(0, min(255, ql)); qh = max(0, min(255, qh));
            while (ql > 0 && pv[a] + (float)ql * sv[a] > lov[a]) ql--;
            while (qh < 255 && pv[a] + (float)qh * sv[a] < hiv[a]) qh++;
            qlo[a][s] = (unsigned char)ql; qhi[a][s] = (unsigned char)qh;
        }
        if (imask & (1u << s)) {
            meta[s] = (unsigned char)((1u << 5) | (24u + (unsigned)s));
            tasks_out[task_base + inner_rank] = make_uint2((unsigned)id, child_base + inner_rank);
            inner_rank++;
        } else {
            int pc = cnt_of[s];
            meta[s] = (unsigned char)((((1u << pc) - 1u) << 5) | (unsigned)prim_off);
            int first = (id >= n - 1) ? (id - (n - 1)) : A.range[id].x;
            for (int k = 0; k < pc; k++)
                src.write(A.out_prims + (size_t)(prim_base + prim_off + k) * PRIM_F4, A.vals[first + k]);
            prim_off += pc;
        }
    }
    auto pack4 = [](const unsigned char* b) { return (unsigned)b[0] | ((unsigned)b[1] << 8) | ((unsigned)b[2] << 16) | ((unsigned)b[3] << 24); };
    uint4* o = A.out_nodes + (size_t)out_idx * 5;
    o[0] = make_uint4(__float_as_uint(plo.x), __float_as_uint(plo.y), __float_as_uint(plo.z),
                      (unsigned)ex | ((unsigned)ey << 8) | ((unsigned)ez << 16) | (imask << 24));
    o[1] = make_uint4(child_base, prim_base, pack4(meta), pack4(meta + 4));
    o[2] = make_uint4(pack4(qlo[0]), pack4(qlo[0] + 4), pack4(qlo[1]), pack4(qlo[1] + 4));
    o[3] = make_uint4(pack4(qlo[2]), pack4(qlo[2] + 4), pack4(qhi[0]), pack4(qhi[0] + 4));
    o[4] = make_uint4(pack4(qhi[1]), pack4(qhi[1] + 4), pack4(qhi[2]), pack4(qhi[2] + 4));
}

template <int PRIM_F4, typename Source>
__global__ void k_collapse(CollapseArgs A, Source src, const uint2* __restrict__ tasks_in, unsigned n_in, uint2* tasks_out) {
    unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_in) return;
    const int n = A.n;
    const uint2 task = tasks_in[t];
    const int root = (int)task.x;
    int cid[8]; int cnt = 2;
    cid[0] = A.child[root].x; cid[1] = A.child[root].y;
    while (cnt < 8) {
        int best = -1; float bestA = -1.0f;
        for (int c = 0; c < cnt; c++) {
            int id = cid[c];
            if (id >= n - 1) continue;
            if (A.range[id].y - A.range[id].x + 1 <= MAX_LEAF) continue;
            float a = box_area(A.nlo[id], A.nhi[id]);
            if (a > bestA) { bestA = a; best = c; }
        }
        if (best < 0) break;
        int id = cid[best];
        cid[best] = A.child[id].x;
        cid[cnt++] = A.child[id].y;
    }
    emit_node<PRIM_F4, Source>(A, src, cid, cnt, A.nlo[root], A.nhi[root], task.y, tasks_out);
}

// n <= MAX_LEAF: one node, one leaf child holding everything
template <int PRIM_F4, typename Source>
__global__ void k_single_node(Source src, const float4* __restrict__ plo, const float4* __restrict__ phi, int n, uint4* out_nodes,
                              float4* out_prims, BuildCounters* c) {
    if (threadIdx.x || blockIdx.x) return;
    float s = 0.0f;
    for (int k = 0; k < 6; k++) s = fmaxf(s, fabsf(dec_f(c->bounds[k])));
    const float pad = s * 3.0517578125e-5f + 1e-30f;
    float4 l = make_float4(INFINITY, INFINITY, INFINITY, 0), h = make_float4(-INFINITY, -INFINITY, -INFINITY, 0);
    for (int i = 0; i < n; i++) {
        l.x = fminf(l.x, plo[i].x - pad); l.y = fminf(l.y, plo[i].y - pad); l.z = fminf(l.z, plo[i].z - pad);
        h.x = fmaxf(h.x, phi[i].x + pad); h.y = fmaxf(h.y, phi[i].y + pad); h.z = fmaxf(h.z, phi[i].z + pad);
    }
    const int ex = exp_for_extent(h.x - l.x), ey = exp_for_extent(h.y - l.y), ez = exp_for_extent(h.z - l.z);
    unsigned meta0 = (((1u << n) - 1u) << 5) | 0u;
    for (int i = 0; i < n; i++) src.write(out_prims + (size_t)i * PRIM_F4, (uint32_t)i);
    out_nodes[0] = make_uint4(__float_as_uint(l.x), __float_as_uint(l.y), __float_as_uint(l.z), (unsigned)ex | ((unsigned)ey << 8) | ((unsigned)ez << 16));
    out_nodes[1] = make_uint4(0u, 0u, meta0, 0u);
    out_nodes[2] = make_uint4(0u, 0u, 0u, 0u);               // qlo_x, qlo_y = 0
    out_nodes[3] = make_uint4(0u, 0u, 0xffu, 0u);            // qlo_z = 0, qhi_x[0] = 255
    out_nodes[4] = make_uint4(0xffu, 0u, 0xffu, 0u);         // qhi_y[0] = qhi_z[0] = 255
    c->nodes = 1; c->prims = (unsigned)n;
}

struct Scratch {
    void* p = nullptr;
    ~Scratch() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
};

template <int PRIM_F4, typename Source>
static cudaError_t build_generic(const float4* d_plo, const float4* d_phi, uint32_t n, Source src, BuildCounters* d_ctr, Bvh8* out,
                                 cudaStream_t stream) {
    const int TB = 256;
    auto grid = [&](size_t k) { return (unsigned)((k + TB - 1) / TB); };
    const uint32_t max_nodes = n <= MAX_LEAF ? 1u : n;
    Scratch nodes_s, prims_s;
    CKE(nodes_s.alloc((size_t)max_nodes * 80));
    CKE(prims_s.alloc((size_t)n * PRIM_F4 * 16));
    uint4* d_nodes = (uint4*)nodes_s.p;
    float4* d_prims = (float4*)prims_s.p;
    BuildCounters h_ctr;

    if (n <= MAX_LEAF) {
        k_single_node<PRIM_F4, Source><<<1, 32, 0, stream>>>(src, d_plo, d_phi, (int)n, d_nodes, d_prims, d_ctr);
    } else {
        Scratch keys_a, keys_b, vals_a, vals_b, child_s, range_s, parent_s, flags_s, nlo_s, nhi_s, tasks_a, tasks_b, tmp_s;
        CKE(keys_a.alloc((size_t)n * 8)); CKE(keys_b.alloc((size_t)n * 8));
        CKE(vals_a.alloc((size_t)n * 4)); CKE(vals_b.alloc((size_t)n * 4));
        CKE(child_s.alloc((size_t)n * 8)); CKE(range_s.alloc((size_t)n * 8));
        CKE(parent_s.alloc((size_t)2 * n * 4)); CKE(flags_s.alloc((size_t)n * 4));
        CKE(nlo_s.alloc((size_t)2 * n * 16)); CKE(nhi_s.alloc((size_t)2 * n * 16));
        CKE(tasks_a.alloc((size_t)n * 8)); CKE(tasks_b.alloc((size_t)n * 8));
        size_t tmp_bytes = 0;
        CKE(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (unsigned long long*)keys_a.p, (unsigned long long*)keys_b.p,
                                            (uint32_t*)vals_a.p, (uint32_t*)vals_b.p, (int)n, 0, 63, stream));
        CKE(tmp_s.alloc(tmp_bytes));
        k_morton<<<grid(n), TB, 0, stream>>>(d_plo, d_phi, n, d_ctr, (unsigned long long*)keys_a.p, (uint32_t*)vals_a.p);
        CKE(cub::DeviceRadixSort::SortPairs(tmp_s.p, tmp_bytes, (unsigned long long*)keys_a.p, (unsigned long long*)keys_b.p,
                                            (uint32_t*)vals_a.p, (uint32_t*)vals_b.p, (int)n, 0, 63, stream));
        k_lbvh<<<grid(n - 1), TB, 0, stream>>>((unsigned long long*)keys_b.p, (int)n, (int2*)child_s.p, (int2*)range_s.p, (int*)parent_s.p);
        CKE(cudaMemsetAsync(flags_s.p, 0, (size_t)n * 4, stream));
        k_fit<<<grid(n), TB, 0, stream>>>(d_plo, d_phi, (uint32_t*)vals_b.p, (int)n, (int2*)child_s.p, (int*)parent_s.p,
                                          (unsigned int*)flags_s.p, (float4*)nlo_s.p, (float4*)nhi_s.p, d_ctr);
        CollapseArgs A;
        A.child = (int2*)child_s.p; A.range = (int2*)range_s.p; A.nlo = (float4*)nlo_s.p; A.nhi = (float4*)nhi_s.p;
        A.vals = (uint32_t*)vals_b.p; A.n = (int)n; A.out_nodes = d_nodes; A.out_prims = d_prims; A.ctr = d_ctr;
        uint2 root_task = make_uint2(0u, 0u);
        CKE(cudaMemcpyAsync(tasks_a.p, &root_task, sizeof root_task, cudaMemcpyHostToDevice, stream));
        unsigned n_in = 1;
        uint2* tin = (uint2*)tasks_a.p; uint2* tout = (uint2*)tasks_b.p;
        while (n_in > 0) {
            k_collapse<PRIM_F4, Source><<<grid(n_in), TB, 0, stream>>>(A, src, tin, n_in, tout);
            CKE(cudaMemcpyAsync(&h_ctr, d_ctr, sizeof h_ctr, cudaMemcpyDeviceToHost, stream));
            CKE(cudaStreamSynchronize(stream));
            n_in = h_ctr.tasks_out;
            CKE(cudaMemsetAsync(&d_ctr->tasks_out, 0, sizeof(unsigned), stream));
            uint2* t = tin; tin = tout; tout = t;
        }
        CKE(cudaGetLastError());
    }
    CKE(cudaMemcpyAsync(&h_ctr, d_ctr, sizeof h_ctr, cudaMemcpyDeviceToHost, stream));
    CKE(cudaStreamSynchronize(stream));
    CKE(cudaGetLastError());
    out->n_nodes = h_ctr.nodes; out->n_prims = h_ctr.prims;
    // right-size the node array (child indices are relative to the array start, so a plain copy is valid)
    CKE(cudaMalloc((void**)&out->nodes, (size_t)out->n_nodes * 80));
    CKE(cudaMemcpyAsync(out->nodes, d_nodes, (size_t)out->n_nodes * 80, cudaMemcpyDeviceToDevice, stream));
    out->prims = d_prims; prims_s.p = nullptr;
    float s = 0.0f;
    for (int k = 0; k < 6; k++) s = fmaxf(s, fabsf(dec_f(h_ctr.bounds[k])));
    const float pad = s * 3.0517578125e-5f + 1e-30f;
    for (int k = 0; k < 3; k++) { out->lo[k] = dec_f(h_ctr.bounds[k]) - pad; out->hi[k] = dec_f(h_ctr.bounds[3 + k]) + pad; }
    CKE(cudaStreamSynchronize(stream));
    return cudaSuccess;
}

cudaError_t build_blas(const uint8_t* d_vertices, uint32_t n_vertices, const uint32_t* d_indices, uint32_t n_tris, Bvh8* out,
                       cudaStream_t stream) {
    (void)n_vertices;
    cudaEvent_t e0, e1;
    CKE(cudaEventCreate(&e0)); CKE(cudaEventCreate(&e1));
    CKE(cudaEventRecord(e0, stream));
    Scratch plo, phi, ctr;
    CKE(plo.alloc((size_t)n_tris * 16)); CKE(phi.alloc((size_t)n_tris * 16)); CKE(ctr.alloc(sizeof(BuildCounters)));
    k_init_counters<<<1, 1, 0, stream>>>((BuildCounters*)ctr.p);
    if (n_tris) k_tri_boxes<<<(n_tris + 255) / 256, 256, 0, stream>>>(d_vertices, d_indices, n_tris, (float4*)plo.p, (float4*)phi.p, (BuildCounters*)ctr.p);
    TriSource src{d_vertices, d_indices};
    cudaError_t e = build_generic<3, TriSource>((float4*)plo.p, (float4*)phi.p, n_tris, src, (BuildCounters*)ctr.p, out, stream);
    if (e == cudaSuccess) {
        cudaEventRecord(e1, stream); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&out->build_ms, e0, e1);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return e;
}

cudaError_t build_tlas(const float4* d_inst_recs, const float4* d_box_lo, const float4* d_box_hi, uint32_t n_instances, Bvh8* out,
                       cudaStream_t stream) {
    Scratch ctr;
    CKE(ctr.alloc(sizeof(BuildCounters)));
    k_init_counters<<<1, 1, 0, stream>>>((BuildCounters*)ctr.p);
    if (n_instances) k_box_bounds<<<(n_instances + 255) / 256, 256, 0, stream>>>(d_box_lo, d_box_hi, n_instances, (BuildCounters*)ctr.p);
    RecSource src{d_inst_recs};
    return build_generic<4, RecSource>(d_box_lo, d_box_hi, n_instances, src, (BuildCounters*)ctr.p, out, stream);
}

void free_bvh(Bvh8* b) {
    if (b->nodes) cudaFree(b->nodes);
    if (b->prims) cudaFree(b->prims);
    b->nodes = nullptr; b->prims = nullptr; b->n_nodes = b->n_prims = 0;
}

// ---- instance records: rows of world->object from objectToWorldInverse (binding t3, Renderer.cpp:2091-2121) and the
// world box of each instance (8 corners of the padded BLAS box through the desc's 3x4 objectToWorld).
__global__ void k_instance_records(const rtx_instance_desc* __restrict__ descs, const rtx_instance_props* __restrict__ props,
                                   const BlasBounds* __restrict__ bounds, uint32_t n, float4* __restrict__ recs,
                                   float4* __restrict__ lo, float4* __restrict__ hi) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* m = props[i].objectToWorldInverse;
    recs[4 * i + 0] = make_float4(m[0], m[4], m[8], m[12]);
    recs[4 * i + 1] = make_float4(m[1], m[5], m[9], m[13]);
    recs[4 * i + 2] = make_float4(m[2], m[6], m[10], m[14]);
    const uint32_t blas = (uint32_t)descs[i].blas;
    recs[4 * i + 3] = make_float4(__uint_as_float(blas), __uint_as_float(i), 0.0f, 0.0f);
    const BlasBounds b = bounds[blas];
    float l[3] = {INFINITY, INFINITY, INFINITY}, h[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int k = 0; k < 8; k++) {
        float x = (k & 1) ? b.hi[0] : b.lo[0], y = (k & 2) ? b.hi[1] : b.lo[1], z = (k & 4) ? b.hi[2] : b.lo[2];
        for (int r = 0; r < 3; r++) {
            const float* t = descs[i].transform[r];
            float v = t[0] * x + t[1] * y + t[2] * z + t[3];
            l[r] = fminf(l[r], v); h[r] = fmaxf(h[r], v);
        }
    }
    float s = 0.0f;
    for (int r = 0; r < 3; r++) s = fmaxf(s, fmaxf(fabsf(l[r]), fabsf(h[r])));
    const float pad = s * 3.0517578125e-5f + 1e-30f;
    lo[i] = make_float4(l[0] - pad, l[1] - pad, l[2] - pad, 0.0f);
    hi[i] = make_float4(h[0] + pad, h[1] + pad, h[2] + pad, 0.0f);
}

cudaError_t launch_instance_records(const rtx_instance_desc* d_descs, const rtx_instance_props* d_props, const BlasBounds* d_bounds,
                                    uint32_t n, float4* d_recs, float4* d_lo, float4* d_hi, cudaStream_t stream) {
    if (n) k_instance_records<<<(n + 127) / 128, 128, 0, stream>>>(d_descs, d_props, d_bounds, n, d_recs, d_lo, d_hi);
    return cudaGetLastError();
}

}  // namespace rtx
