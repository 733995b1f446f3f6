// trace.cu — persistent-thread traversal kernels (the `extend` and `connect` stages of the wavefront).
//
// One launch drains a whole ray queue: the grid is sized to the machine (148 SMs x resident CTAs), and every
// warp pulls batches of consecutive rays from a global cursor with one atomic per refill.  A warp refills its
// idle lanes as soon as fewer than FETCH_THRESHOLD lanes are still traversing, so long-running incoherent rays do
// not strand the other 31 lanes (B200 has no RT cores; warp-execution efficiency is the second-order term after
// memory latency — see DESIGN.md §Kernels).
#include <stdlib.h>

#include <algorithm>
using std::max;

#include "trace.h"
#include "traverse.cuh"

namespace rtx {

#ifndef TRACE_BLOCK
#define TRACE_BLOCK 128
#endif
#ifndef RTX_TRACE_MINB
#define RTX_TRACE_MINB (1024 / TRACE_BLOCK)   // resident CTAs per SM the register budget is set for (64 registers, 1024 threads)
#endif
#define FETCH_THRESHOLD 20     // refill the warp's idle lanes when fewer than this many lanes are still traversing
#ifndef RTX_FETCH_CHUNK
#define RTX_FETCH_CHUNK 32     // rays a warp claims from the global cursor with one atomic
#endif
#ifndef RTX_FETCH_MIN
#define RTX_FETCH_MIN 32u      // guided self-scheduling (claims shrinking to this many rays as the queue runs out) measured slower than
#endif                         // constant claims: 8.94 against 8.82 ms per C2 pass with a minimum of 4; kept as a build knob
#ifndef RTX_SCHED_DEFAULT
#define RTX_SCHED_DEFAULT 0x060808   // th_tri | th_inst << 8 | th_node << 16: run the triangle (instance) phase when >= th lanes are parked,
#endif                               // or whenever fewer than th_node lanes have node work left (sweep: profiles/r01_s4_sched_sweep.txt)

template <bool ANY_HIT, bool STATS>
__global__ void __launch_bounds__(TRACE_BLOCK, RTX_TRACE_MINB)
trace_kernel(SceneAS S, const float4* __restrict__ o_tmin, const float4* __restrict__ d_tmax,
             const uint32_t* __restrict__ n_ptr, uint32_t n_fixed, unsigned int* __restrict__ cursor,
             float4* __restrict__ hit_a, uint32_t* __restrict__ hit_inst, TraceStats* st, int fetch_th, int sched) {
    const uint32_t n = n_ptr ? *n_ptr : n_fixed;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    uint2 stack[RTX_STACK_SIZE];
    __shared__ uint8_t s_perm[2048];
    __shared__ float4 s_cold[4][TRACE_BLOCK];
    fill_perm_table(s_perm, threadIdx.x, blockDim.x);
    __syncthreads();
    const TravCold C = {&s_cold[0][threadIdx.x], &s_cold[1][threadIdx.x], &s_cold[2][threadIdx.x], &s_cold[3][threadIdx.x]};
    Trav T;
    uint32_t pend = PEND_NONE, leaf_base = 0u, leaf_bits = 0u, leaf_W = 0u;    // a parked leaf group (traverse.cuh)
    const int th_tri = sched & 0xff, th_inst = (sched >> 8) & 0xff, th_node = (sched >> 16) & 0xff;
    bool active = false;
    bool exhausted = (S.n_instances == 0u);
    // the warp's private pool of claimed rays [pool_next, pool_end) and one chunk claimed ahead of need (its atomic
    // is in flight while the warp traverses): all warp-uniform
    uint32_t pool_next = 0, pool_end = 0, ahead = 0, ahead_sz = 0, claim_sz = RTX_FETCH_CHUNK;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    bool have_ahead = false;
    unsigned int c_nodes = 0, c_tris = 0, c_insts = 0;

#ifdef RTX_TRACE_TIMELINE
    unsigned long long tl_t0, tl_tex = 0, tl_rays = 0; unsigned tl_steps = 0, tl_drain_steps = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tl_t0));
#endif
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
                const uint32_t sz = have_ahead ? ahead_sz : claim_sz;
                if (!have_ahead) { if (lane == 0) ahead = atomicAdd(cursor, (unsigned)sz); }
                const uint32_t base = __shfl_sync(0xffffffffu, ahead, 0);
                have_ahead = false;
                if (base >= n) {
                    exhausted = true; pool_next = pool_end = 0; need = take;
#ifdef RTX_TRACE_TIMELINE
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tl_tex));
#endif
                }
                else {
                    pool_end = min(base + sz, n);
                    // optional guided self-scheduling (RTX_FETCH_MIN < RTX_FETCH_CHUNK): the claims shrink as the queue runs out
                    claim_sz = min((uint32_t)RTX_FETCH_CHUNK, max(RTX_FETCH_MIN, (n - pool_end) / (2u * n_warps)));
                    const uint32_t take2 = min(need - take, pool_end - base);
                    if (rank >= take) mine = base + (rank - take);
                    pool_next = base + take2;
                    need = take + take2;
                }
            }
            if (!active && rank < need) {
                trav_init(T, C, S, __ldg(o_tmin + mine), __ldg(d_tmax + mine), mine);
                active = true;
            }
            if (!exhausted && !have_ahead && pool_end - pool_next < 32u) {   // claim the next chunk now, use it later
                if (lane == 0) ahead = atomicAdd(cursor, (unsigned)claim_sz);
                ahead_sz = claim_sz;
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
        do {
            if (active && pend == PEND_NONE) {
                trav_node<ANY_HIT, STATS>(T, S, s_perm, stack, leaf_base, leaf_bits, leaf_W, &c_nodes);
                if (leaf_bits) pend = T.blas_sp >= 0 ? PEND_TRI : PEND_INST;
            }
            const unsigned m_tri = __ballot_sync(0xffffffffu, pend == PEND_TRI), m_inst = __ballot_sync(0xffffffffu, pend == PEND_INST);
            const unsigned m_node = __ballot_sync(0xffffffffu, active && pend == PEND_NONE);
            const bool flush = __popc(m_node) < th_node;
            const bool do_tri = m_tri && (flush || __popc(m_tri) >= th_tri);       // warp-uniform
            const bool do_inst = m_inst && (flush || __popc(m_inst) >= th_inst);
            // ONE divergent region for the three kinds of lanes, so that their instruction streams can interleave (a lane that pops does
            // not wait for the triangle loads of its neighbours); the pop/finish tail is therefore written out in both branches
#define RTX_FINISH(FOUND)                                                                                      \
            {                                                                                                  \
                bool done;                                                                                     \
                if (ANY_HIT && (FOUND)) { T.hinst = T.cur_inst; done = true; }                                 \
                else done = trav_pop(T, C, S, stack);                                                          \
                if (done) {                                                                                    \
                    active = false;                                                                            \
                    const uint32_t j = __float_as_uint(C.wo->w);                                               \
                    if (!ANY_HIT) { const float4 h = *C.hit; hit_a[j] = make_float4(T.ht, h.x, h.y, h.z); }    \
                    hit_inst[j] = T.hinst;                                                                     \
                }                                                                                              \
            }
            if (pend == PEND_TRI) {
                if (do_tri) {
                    const bool found = trav_tris<ANY_HIT, STATS>(T, C, leaf_base, leaf_bits, leaf_W, &c_tris);
                    pend = PEND_NONE;
                    RTX_FINISH(found)
                }
            } else if (pend == PEND_INST) {
                if (do_inst) {
                    trav_enter_instance<STATS>(T, C, S, stack, leaf_base, leaf_bits, leaf_W, &c_insts);
                    pend = PEND_NONE;
                }
            } else if (active) {
                RTX_FINISH(false)
            }
#undef RTX_FINISH
#ifdef RTX_TRACE_TIMELINE
            tl_steps++; if (exhausted) tl_drain_steps++;
#endif
        } while (__popc(__ballot_sync(0xffffffffu, active)) >= threshold);
    }
#ifdef RTX_TRACE_TIMELINE
    {
        unsigned long long tl_t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tl_t1));
        const unsigned wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        if (lane == 0 && (wid % 37u) == 0u && n > 100000u)
            printf("TL %u n %u start %llu exhausted_at %llu end %llu steps %u drain_steps %u\n", wid, n, tl_t0 % 100000000ull,
                   tl_tex ? tl_tex - tl_t0 : 0ull, tl_t1 - tl_t0, tl_steps, tl_drain_steps);
    }
#endif
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
                         cudaStream_t stream, int grid_share) {
    cudaError_t e = cudaMemsetAsync(cursor, 0, sizeof(unsigned int), stream);
    if (e != cudaSuccess) return e;
    static int fetch_th = -1, waves = -1, sched = 0;
    if (fetch_th < 0) {   // tuning knobs (defaults are the measured optimum on C2, see profiles/)
        const char* e = getenv("RTX_FETCH_TH"); fetch_th = e ? atoi(e) : FETCH_THRESHOLD;
        e = getenv("RTX_SCHED"); sched = e ? (int)strtol(e, nullptr, 0) : RTX_SCHED_DEFAULT;
        e = getenv("RTX_TRACE_WAVES"); waves = e ? atoi(e) : 1;
    }
    // persistent: the resident CTAs of the machine (x waves), or this launch's share of them when several traversals run concurrently
    // (wave_render_pass: the parts' kernels are then co-resident and one part's drain phase overlaps the other's work)
    const int grid = max(num_sms(), num_sms() * RTX_TRACE_MINB * waves / max(grid_share, 1));
    if (stats) {
        if (any_hit) trace_kernel<true, true><<<grid, TRACE_BLOCK, 0, stream>>>(S, o_tmin, d_tmax, n_ptr, n_fixed, cursor, hit_a, hit_inst, stats, fetch_th, sched);
        else trace_kernel<false, true><<<grid, TRACE_BLOCK, 0, stream>>>(S, o_tmin, d_tmax, n_ptr, n_fixed, cursor, hit_a, hit_inst, stats, fetch_th, sched);
    } else {
        if (any_hit) trace_kernel<true, false><<<grid, TRACE_BLOCK, 0, stream>>>(S, o_tmin, d_tmax, n_ptr, n_fixed, cursor, hit_a, hit_inst, nullptr, fetch_th, sched);
        else trace_kernel<false, false><<<grid, TRACE_BLOCK, 0, stream>>>(S, o_tmin, d_tmax, n_ptr, n_fixed, cursor, hit_a, hit_inst, nullptr, fetch_th, sched);
    }
    return cudaGetLastError();
}

cudaError_t read_stack_overflow(unsigned int* host_flag, cudaStream_t stream) {
    cudaError_t e = cudaMemcpyFromSymbolAsync(host_flag, g_stack_overflow, sizeof(unsigned int), 0, cudaMemcpyDeviceToHost, stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(stream);
}

}  // namespace rtx
