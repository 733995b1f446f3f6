// restir.h — host interface of the ReSTIR reuse passes (the reference's RayGen2 / RayGen3).
#pragma once
#include "wavefront.h"

namespace rtx {

// Per-pixel buffers u2..u7 of the reference (rdn/Renderer.cpp:1331-1577) as structure-of-arrays float4 planes
// (storage order is the build's choice); n = width * height entries per plane.
//   reservoir (DI and GI alike, Reservoir_v7.hlsl:15-27): plane 0 = (x2|xn, w_sum), 1 = (n2|nn, W), 2 = (L2|E3 as binary16 values, bits(M))
//   sample data (Reservoir_v7.hlsl:2-11): plane 0 = (x1, bits(mID16 | kind << 16)), 1 = (n1, bits(objID)), 2 = (o, 0), 3 = (L1, 0)
//   kind: 0 = the primary ray missed, 1 = primary hit on an emitter, 2 = sampled
enum { RS_RES_PLANES = 3, RS_SD_PLANES = 4, RS_MAX_RAYS_PER_PIXEL = 9 };

struct RestirBuffers {
    uint32_t n = 0;
    float4* di[2] = {nullptr, nullptr};      // [0] current, [1] last
    float4* gi[2] = {nullptr, nullptr};
    float4* sd[2] = {nullptr, nullptr};
    uint4* cand = nullptr;                   // 3 planes: DI candidates + counts, GI candidates, seed after the candidate search
    float4* tmp = nullptr;                   // 2 planes: (ReconnectDI vector, f_g), (GI contribution, kind)
    uint32_t* vmask = nullptr;               // occlusion bits of the pixel's visibility rays
    RayQueue q;                              // shadow-ray queue, capacity RS_MAX_RAYS_PER_PIXEL * n
    uint32_t* q_count = nullptr;             // 12 counters: ray count, heavy rays, other rays of the temporal / spatial A / spatial B queues
};

cudaError_t restir_alloc(RestirBuffers* R, uint32_t width, uint32_t height);
void restir_free(RestirBuffers* R);
cudaError_t restir_clear(RestirBuffers& R, cudaStream_t stream);

// pass 1 tail: packs the wavefront path state of a finished 1-spp pass into the current reservoir / sample buffers
cudaError_t restir_store_pass1(RestirBuffers& R, WaveBuffers& B, const SceneData& S, cudaStream_t stream, uint64_t* launches);
// passes 2 and 3 + accumulation (F20) into B.accum
cudaError_t restir_reuse_passes(RestirBuffers& R, WaveBuffers& B, const SceneData& S, const SceneAS& AS, uint32_t frame_index,
                                cudaStream_t stream, uint64_t* launches);
// the *_last buffers, 40 floats per pixel (layout in restir.cu: restir_dump)
cudaError_t restir_dump(RestirBuffers& R, cudaStream_t stream, float* host_out);

}  // namespace rtx
