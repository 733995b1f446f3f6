// wave_dev.cuh — device helpers shared by the wavefront stages (wavefront.cu) and the ReSTIR reuse passes (restir.cu).
#pragma once
#include "wavefront.h"

namespace rtx {

struct StateView {
    float4* base; uint32_t n;
    uint2* seed;          // the path's RNG state in its own 8-byte plane: every stage updates it, and as the .w of two float4 planes it cost a
                          // 32-byte read and a 32-byte write per path and stage to move 8 bytes (E0 stages only; nullptr elsewhere)
#ifdef RTX_STATE_AOS
    // one 288-byte record per path (9 whole 32-byte sectors): a thread's accesses use every byte of the sectors it touches whatever the
    // order of the path ids in its warp — queues that are not in pixel order (two-ended LPT queues, material bins) cost no extra sectors
    __device__ __forceinline__ float4& at(int plane, uint32_t pid) const { return base[(size_t)pid * NSTATE + plane]; }
#else
    __device__ __forceinline__ float4& at(int plane, uint32_t pid) const { return base[(size_t)plane * n + pid]; }
#endif
};

__device__ __forceinline__ f3 xyz(float4 v) { return mk3(v.x, v.y, v.z); }
__device__ __forceinline__ float4 f4(f3 v, float w) { return make_float4(v.x, v.y, v.z, w); }
__device__ __forceinline__ float4 f4u(f3 v, uint32_t w) { return make_float4(v.x, v.y, v.z, __uint_as_float(w)); }

// Warp-level queue compaction: one atomicAdd per warp claims a contiguous run of slots.  Must be reached by all 32 lanes.
__device__ __forceinline__ void push_ray(const RayQueue& q, bool emit, f3 o, float tmin, f3 d, float tmax, uint32_t pid) {
    const unsigned mask = __ballot_sync(0xffffffffu, emit);
    if (mask == 0u) return;
    const unsigned lane = threadIdx.x & 31u;
    const int leader = __ffs(mask) - 1;
    unsigned base = 0;
    if ((int)lane == leader) base = atomicAdd(q.count, (unsigned)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (emit) {
        const unsigned slot = base + __popc(mask & ((1u << lane) - 1u));
        q.o_tmin[slot] = f4(o, tmin);
        q.d_tmax[slot] = f4(d, tmax);
        q.pid[slot] = pid;
    }
}

// Cost class of a ray: does its segment cross the bounds of the scene's large instances?  Only the ORDER of a queue depends on it (every
// path draws its own random numbers, results do not depend on queue order), so the slab test may be approximate.
__device__ __forceinline__ bool ray_is_heavy(const SceneData& S, f3 o, float tmin, f3 d, float tmax) {
    if (!S.heavy_valid) return true;
    float ix, iy, iz;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ix) : "f"(d.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iy) : "f"(d.y));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iz) : "f"(d.z));
    const float ax = (S.heavy_lo[0] - o.x) * ix, bx = (S.heavy_hi[0] - o.x) * ix;
    const float ay = (S.heavy_lo[1] - o.y) * iy, by = (S.heavy_hi[1] - o.y) * iy;
    const float az = (S.heavy_lo[2] - o.z) * iz, bz = (S.heavy_hi[2] - o.z) * iz;
    const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), tmin));
    const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tmax));
    return !(tn > tf);            // NaN (a direction component of 0 inside the slab) counts as heavy
}

// Records `slot` in the queue's trace order: heavy rays from entry 0 upwards, the others from entry cap-1 downwards (one atomic per warp
// and class).  mask = ballot(emit); must be reached by all 32 lanes.
__device__ __forceinline__ void record_order(const RayQueue& q, unsigned mask, bool emit, bool heavy, unsigned slot) {
    const unsigned lane = threadIdx.x & 31u, below = (1u << lane) - 1u;
    const unsigned mh = __ballot_sync(0xffffffffu, emit && heavy), ml = mask & ~mh;
    unsigned bh = 0, bl = 0;
    if (mh) { const int l = __ffs(mh) - 1; if ((int)lane == l) bh = atomicAdd(q.n_heavy, (unsigned)__popc(mh)); bh = __shfl_sync(0xffffffffu, bh, l); }
    if (ml) { const int l = __ffs(ml) - 1; if ((int)lane == l) bl = atomicAdd(q.n_light, (unsigned)__popc(ml)); bl = __shfl_sync(0xffffffffu, bl, l); }
    if (emit) q.order[heavy ? bh + __popc(mh & below) : q.cap - 1u - (bl + __popc(ml & below))] = slot;
}

// push_ray that also records the ray's slot in the queue's trace order (wavefront.h RayQueue).  Must be reached by all 32 lanes.
__device__ __forceinline__ void push_ray2(const RayQueue& q, bool emit, bool heavy, f3 o, float tmin, f3 d, float tmax, uint32_t pid) {
    const unsigned mask = __ballot_sync(0xffffffffu, emit);
    if (mask == 0u) return;
    const unsigned lane = threadIdx.x & 31u, below = (1u << lane) - 1u;
    const int leader = __ffs(mask) - 1;
    unsigned base = 0;
    if ((int)lane == leader) base = atomicAdd(q.count, (unsigned)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    const unsigned slot = base + __popc(mask & below);
    if (emit) {
        q.o_tmin[slot] = f4(o, tmin);
        q.d_tmax[slot] = f4(d, tmax);
        q.pid[slot] = pid;
    }
    if (q.order) record_order(q, mask, emit, heavy, slot);
}

}  // namespace rtx
