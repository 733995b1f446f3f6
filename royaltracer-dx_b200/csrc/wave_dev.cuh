// wave_dev.cuh — device helpers shared by the wavefront stages (wavefront.cu) and the ReSTIR reuse passes (restir.cu).
#pragma once
#include "wavefront.h"

namespace rtx {

struct StateView {
    float4* base; uint32_t n;
    __device__ __forceinline__ float4& at(int plane, uint32_t pid) const { return base[(size_t)plane * n + pid]; }
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

}  // namespace rtx
