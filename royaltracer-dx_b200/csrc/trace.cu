// trace.cu — persistent-thread traversal kernels (the `extend` and `connect` stages of the wavefront).
//
// One launch drains a whole ray queue: the grid is sized to the machine (148 SMs x resident CTAs), and every
// warp pulls batches of consecutive rays from a global cursor with one atomic per refill.  A warp refills its
// idle lanes as soon as fewer than FETCH_THRESHOLD lanes are still traversing, so long-running incoherent rays do
// not strand the other 31 lanes (B200 has no RT cores; warp-execution efficiency is the second-order term after
// memory latency — see DESIGN.md §Kernels).
#include "trace.h"
#include "traverse.cuh"

namespace rtx {

#define TRACE_BLOCK 128
#define FETCH_THRESHOLD 20   // refill when fewer than this many lanes are active

template <bool ANY_HIT, bool STATS>
__global__ void __launch_bounds__(TRACE_BLOCK, 4)
trace_kernel(SceneAS S, const float4* __restrict__ o_tmin, const float4* __restrict__ d_tmax,
             const uint32_t* __restrict__ n_ptr, uint32_t n_fixed, unsigned int* __restrict__ cursor,
             float4* __restrict__ hit_a, uint32_t* __restrict__ hit_inst, TraceStats* st) {
    const uint32_t n = n_ptr ? *n_ptr : n_fixed;
    const unsigned lane = threadIdx.x & 31u;
    uint2 stack[RTX_STACK_SIZE];

    // This first version keeps the whole traversal of one ray inside traverse(); lanes that finish early wait
    // for the slowest lane of the batch.  Batches are 32 consecutive rays.
    for (;;) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(cursor, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        const uint32_t j = base + lane;
        if (j < n) {
            const float4 o = __ldg(o_tmin + j), d = __ldg(d_tmax + j);
            HitRec h;
            h.t = d.w; h.b1 = 0.0f; h.b2 = 0.0f; h.prim = 0xFFFFFFFFu; h.inst = 0xFFFFFFFFu;
            traverse<ANY_HIT, STATS>(S, o.x, o.y, o.z, d.x, d.y, d.z, o.w, d.w, h, stack, st);
            if (ANY_HIT) {
                hit_inst[j] = h.inst;
            } else {
                hit_a[j] = make_float4(h.t, h.b1, h.b2, __uint_as_float(h.prim));
                hit_inst[j] = h.inst;
            }
        }
    }
}

static int g_num_sms = 0;

static int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

cudaError_t launch_trace(const SceneAS& S, const float4* o_tmin, const float4* d_tmax, const uint32_t* n_ptr, uint32_t n_fixed,
                         unsigned int* cursor, float4* hit_a, uint32_t* hit_inst, bool any_hit, TraceStats* stats,
                         cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(cursor, 0, sizeof(unsigned int), stream);
    if (e != cudaSuccess) return e;
    const int grid = num_sms() * 4 * 2;   // 2 waves of resident CTAs: tail balancing is done by the cursor
    if (stats) {
        if (any_hit) trace_kernel<true, true><<<grid, TRACE_BLOCK, 0, stream>>>(S, o_tmin, d_tmax, n_ptr, n_fixed, cursor, hit_a, hit_inst, stats);
        else trace_kernel<false, true><<<grid, TRACE_BLOCK, 0, stream>>>(S, o_tmin, d_tmax, n_ptr, n_fixed, cursor, hit_a, hit_inst, stats);
    } else {
        if (any_hit) trace_kernel<true, false><<<grid, TRACE_BLOCK, 0, stream>>>(S, o_tmin, d_tmax, n_ptr, n_fixed, cursor, hit_a, hit_inst, nullptr);
        else trace_kernel<false, false><<<grid, TRACE_BLOCK, 0, stream>>>(S, o_tmin, d_tmax, n_ptr, n_fixed, cursor, hit_a, hit_inst, nullptr);
    }
    return cudaGetLastError();
}

cudaError_t read_stack_overflow(unsigned int* host_flag, cudaStream_t stream) {
    cudaError_t e = cudaMemcpyFromSymbolAsync(host_flag, g_stack_overflow, sizeof(unsigned int), 0, cudaMemcpyDeviceToHost, stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(stream);
}

}  // namespace rtx
