// trace.cu — persistent-thread traversal kernels (the `extend` and `connect` stages of the wavefront).
//
// One launch drains a whole ray queue: the grid is sized to the machine (148 SMs x resident CTAs), and every
// warp pulls batches of consecutive rays from a global cursor with one atomic per refill.  A warp refills its
// idle lanes as soon as fewer than FETCH_THRESHOLD lanes are still traversing, so long-running incoherent rays do
// not strand the other 31 lanes (B200 has no RT cores; warp-execution efficiency is the second-order term after
// memory latency — see DESIGN.md §Kernels).
#include <stdlib.h>

#include "trace.h"
#include "traverse.cuh"

namespace rtx {

#define TRACE_BLOCK 128
#ifndef RTX_COOP_TRI
#define RTX_COOP_TRI 0         // warp-cooperative triangle testing (traverse.cuh); 0 = per-lane leaf loop
#endif
#ifndef RTX_TRACE_MINB
#define RTX_TRACE_MINB 6       // resident CTAs per SM the register budget is set for
#endif
#define FETCH_THRESHOLD 20     // refill the warp's idle lanes when fewer than this many lanes are still traversing
#ifndef RTX_FETCH_CHUNK
#define RTX_FETCH_CHUNK 32     // rays a warp claims from the global cursor with one atomic
#endif

template <bool ANY_HIT, bool STATS>
__global__ void __launch_bounds__(TRACE_BLOCK, RTX_TRACE_MINB)
trace_kernel(SceneAS S, const float4* __restrict__ o_tmin, const float4* __restrict__ d_tmax,
             const uint32_t* __restrict__ n_ptr, uint32_t n_fixed, unsigned int* __restrict__ cursor,
             float4* __restrict__ hit_a, uint32_t* __restrict__ hit_inst, TraceStats* st, int fetch_th) {
    const uint32_t n = n_ptr ? *n_ptr : n_fixed;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    uint2 stack[RTX_STACK_SIZE];
#if RTX_COOP_TRI
    __shared__ CoopShared sh;
    const unsigned wbase = threadIdx.x & ~31u;
#endif
    __shared__ uint8_t s_perm[2048];
    fill_perm_table(s_perm, threadIdx.x, blockDim.x);
    __syncthreads();
    Trav T;
    bool active = false;
    bool exhausted = (S.n_instances == 0u);
    uint32_t j = 0;
    // the warp's private pool of claimed rays [pool_next, pool_end) and one chunk claimed ahead of need (its atomic
    // is in flight while the warp traverses): all warp-uniform
    uint32_t pool_next = 0, pool_end = 0, ahead = 0;
    bool have_ahead = false;
    unsigned int c_nodes = 0, c_tris = 0, c_insts = 0;

    for (;;) {
        // ---- refill idle lanes from the pool
        const unsigned idle = __ballot_sync(0xffffffffu, !active);
        if (idle && !exhausted) {
            uint32_t need = (uint32_t)__popc(idle);
            const uint32_t rank = (uint32_t)__popc(idle & lt_mask);
            uint32_t take = min(need, pool_end - pool_next);
            uint32_t mine = pool_next + rank;
            pool_next += take;
            if (take < need) {                                   // pool empty: switch to the chunk claimed ahead (or claim one now)
                if (!have_ahead) { if (lane == 0) ahead = atomicAdd(cursor, (unsigned)RTX_FETCH_CHUNK); }
                const uint32_t base = __shfl_sync(0xffffffffu, ahead, 0);
                have_ahead = false;
                if (base >= n) { exhausted = true; pool_next = pool_end = 0; need = take; }
                else {
                    pool_end = min(base + (uint32_t)RTX_FETCH_CHUNK, n);
                    const uint32_t take2 = min(need - take, pool_end - base);
                    if (rank >= take) mine = base + (rank - take);
                    pool_next = base + take2;
                    need = take + take2;
                }
            }
            if (!active && rank < need) {
                j = mine;
                trav_init(T, S, __ldg(o_tmin + j), __ldg(d_tmax + j));
                active = true;
            }
            if (!exhausted && !have_ahead && pool_end - pool_next < 32u) {   // claim the next chunk now, use it later
                if (lane == 0) ahead = atomicAdd(cursor, (unsigned)RTX_FETCH_CHUNK);
                have_ahead = true;
            }
        }
        const unsigned act = __ballot_sync(0xffffffffu, active);
        if (act == 0u) {
            if (exhausted) break;
            continue;
        }
        // ---- traverse until too few lanes are left (or to the end once the queue is drained)
        const int threshold = exhausted ? 1 : fetch_th;
#if RTX_COOP_TRI
        do {
            uint32_t leaf_base = 0u, leaf_bits = 0u, leaf_W = 0u;
            if (active) trav_node<ANY_HIT, STATS>(T, S, s_perm, stack, leaf_base, leaf_bits, leaf_W, &c_nodes);
            const bool has = active && T.blas_sp >= 0 && leaf_bits != 0u;
            const bool found = coop_triangles<ANY_HIT, STATS>(T, sh, has, leaf_base, leaf_bits, leaf_W, lane, lt_mask, wbase, &c_tris);
            if (active) {
                bool done;
                if (ANY_HIT && found) { T.h.inst = T.cur_inst; done = true; }
                else done = trav_finish<STATS>(T, S, sh, threadIdx.x, stack, leaf_base, T.blas_sp >= 0 ? 0u : leaf_bits, leaf_W, &c_insts);
                if (done) {
                    active = false;
                    if (!ANY_HIT) hit_a[j] = make_float4(T.h.t, T.h.b1, T.h.b2, __uint_as_float(T.h.prim));
                    hit_inst[j] = T.h.inst;
                }
            }
        } while (__popc(__ballot_sync(0xffffffffu, active)) >= threshold);
#else
        do {
            if (active) {
                if (trav_step<ANY_HIT, STATS>(T, S, s_perm, stack, &c_nodes, &c_tris, &c_insts)) {
                    active = false;
                    if (!ANY_HIT) hit_a[j] = make_float4(T.h.t, T.h.b1, T.h.b2, __uint_as_float(T.h.prim));
                    hit_inst[j] = T.h.inst;
                }
            }
        } while (__popc(__ballot_sync(0xffffffffu, active)) >= threshold);
#endif
    }
    if (S.n_instances == 0u) {   // empty scene: everything misses
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            if (!ANY_HIT) hit_a[i] = make_float4(__ldg(d_tmax + i).w, 0.0f, 0.0f, __uint_as_float(0xFFFFFFFFu));
            hit_inst[i] = 0xFFFFFFFFu;
        }
    }
    if (STATS) {
        atomicAdd(&st->nodes, (unsigned long long)c_nodes);
        atomicAdd(&st->tris, (unsigned long long)c_tris);
        atomicAdd(&st->insts, (unsigned long long)c_insts);
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
    static int fetch_th = -1, waves = -1;
    if (fetch_th < 0) {   // tuning knobs (defaults are the measured optimum on C2, see profiles/)
        const char* e = getenv("RTX_FETCH_TH"); fetch_th = e ? atoi(e) : FETCH_THRESHOLD;
        e = getenv("RTX_TRACE_WAVES"); waves = e ? atoi(e) : 1;
    }
    const int grid = num_sms() * RTX_TRACE_MINB * waves;   // persistent: resident CTAs (x waves); the cursor balances the tail
    if (stats) {
        if (any_hit) trace_kernel<true, true><<<grid, TRACE_BLOCK, 0, stream>>>(S, o_tmin, d_tmax, n_ptr, n_fixed, cursor, hit_a, hit_inst, stats, fetch_th);
        else trace_kernel<false, true><<<grid, TRACE_BLOCK, 0, stream>>>(S, o_tmin, d_tmax, n_ptr, n_fixed, cursor, hit_a, hit_inst, stats, fetch_th);
    } else {
        if (any_hit) trace_kernel<true, false><<<grid, TRACE_BLOCK, 0, stream>>>(S, o_tmin, d_tmax, n_ptr, n_fixed, cursor, hit_a, hit_inst, nullptr, fetch_th);
        else trace_kernel<false, false><<<grid, TRACE_BLOCK, 0, stream>>>(S, o_tmin, d_tmax, n_ptr, n_fixed, cursor, hit_a, hit_inst, nullptr, fetch_th);
    }
    return cudaGetLastError();
}

cudaError_t read_stack_overflow(unsigned int* host_flag, cudaStream_t stream) {
    cudaError_t e = cudaMemcpyFromSymbolAsync(host_flag, g_stack_overflow, sizeof(unsigned int), 0, cudaMemcpyDeviceToHost, stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(stream);
}

}  // namespace rtx
