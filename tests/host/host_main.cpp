// host_main.cpp — a C++ host that drives the engine the way the reference's Renderer drives DXR (rdn/Renderer.cpp:44-103 OnInit,
// :431-452 OnUpdate, :468-506 OnRender), through rdx::Renderer (royaltracer-dx_b200/host/Renderer.{h,cpp}) and the C ABI only:
//   CreateVB(path) per OBJ (LoadAssets, :362-370,1973-2072) -> AddInstance -> OnInit -> per frame { SetInstanceTransform (:444-449),
//   OnUpdate, OnRenderFrame } -> CRC of gPermanentData and of the reservoir buffers, ray counters.
// usage: host_main W H frames xforms.bin model0.obj [model1.obj ...]     (xforms.bin: 16 floats per model = its XMMATRIX)
// With RTX_HOST_NCCL_ID=<file> RTX_HOST_RANK=r RTX_HOST_WORLD=n it joins an n-rank job (rank 0 writes the NCCL id file) and renders
// the samples s = r (mod n) of the E0 estimator with the per-pass reduce instead of ReSTIR frames.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "Renderer.h"

static uint32_t crc32_of(const void* data, size_t n) {
    static uint32_t table[256];
    if (!table[1])
        for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; }
    uint32_t c = 0xFFFFFFFFu;
    const unsigned char* p = (const unsigned char*)data;
    for (size_t i = 0; i < n; i++) c = table[(c ^ p[i]) & 0xffu] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

int main(int argc, char** argv) {
    if (argc < 6) { fprintf(stderr, "usage: host_main W H frames xforms.bin model.obj...\n"); return 2; }
    const uint32_t W = (uint32_t)atoi(argv[1]), H = (uint32_t)atoi(argv[2]);
    const int frames = atoi(argv[3]);
    const int n_models = argc - 5;
    std::vector<rdx::XMMATRIX> xf(n_models);
    FILE* f = fopen(argv[4], "rb");
    if (!f || fread(xf.data(), sizeof(rdx::XMMATRIX), n_models, f) != (size_t)n_models) { fprintf(stderr, "cannot read %s\n", argv[4]); return 2; }
    fclose(f);
    const char* id_file = getenv("RTX_HOST_NCCL_ID");
    const int rank = getenv("RTX_HOST_RANK") ? atoi(getenv("RTX_HOST_RANK")) : 0, world = getenv("RTX_HOST_WORLD") ? atoi(getenv("RTX_HOST_WORLD")) : 1;
    try {
        rdx::Renderer r(W, H);
        r.device = rank;
        r.flags = world > 1 ? 0u : RTX_FLAG_RESTIR;
        for (int m = 0; m < n_models; m++) {
            const uint32_t model = r.CreateVB(std::string(argv[5 + m]));      // OBJ/MTL ingest + ESS LUTs (src/Util/ObjLoader.h:393-495)
            r.AddInstance(model, rdx::XMMatrixIdentity());
        }
        r.OnInit();
        if (world > 1) {
            unsigned char id[128];
            if (rank == 0) {
                if (rtx_comm_unique_id(id) != RTX_OK) throw std::runtime_error(rtx_last_error());
                FILE* o = fopen((std::string(id_file) + ".tmp").c_str(), "wb"); fwrite(id, 1, 128, o); fclose(o);
                rename((std::string(id_file) + ".tmp").c_str(), id_file);
            } else {
                FILE* i = nullptr;
                for (int t = 0; t < 600 && !(i = fopen(id_file, "rb")); t++) std::this_thread::sleep_for(std::chrono::milliseconds(50));
                if (!i || fread(id, 1, 128, i) != 128) throw std::runtime_error("no NCCL id file");
                fclose(i);
            }
            r.InitComm(id, rank, world);
        }
        for (int fr = 0; fr < frames; fr++) {
            for (int m = 0; m < n_models; m++) r.SetInstanceTransform((uint32_t)m, xf[m]);     // rdn/Renderer.cpp:444-449
            r.OnUpdate();
            if (world > 1) r.OnRender((uint32_t)(fr * world + rank), 1);
            else r.OnRenderFrame((uint32_t)fr);
            rtx_counters c;
            if (rtx_get_counters(r.Context(), &c) != RTX_OK) throw std::runtime_error(rtx_last_error());
            std::vector<float> acc;
            if (world > 1 && rank == 0) { acc.resize((size_t)W * H * 4); if (rtx_read_reduced_accum(r.Context(), acc.data()) != RTX_OK) throw std::runtime_error(rtx_last_error()); }
            else r.ReadAccumulation(acc);
            uint32_t rcrc = 0;
            if (world == 1) {
                std::vector<float> rs((size_t)W * H * 40);
                if (rtx_read_restir(r.Context(), rs.data()) != RTX_OK) throw std::runtime_error(rtx_last_error());
                rcrc = crc32_of(rs.data(), rs.size() * 4);
            }
            printf("frame %d rank %d closest %llu shadow %llu accum_crc %u restir_crc %u\n", fr, rank, (unsigned long long)c.closest_rays,
                   (unsigned long long)c.shadow_rays, crc32_of(acc.data(), acc.size() * 4), rcrc);
        }
        std::vector<uint8_t> img;
        r.ReadOutput(img);
        printf("output_crc %u lights %zu\n", crc32_of(img.data(), img.size()), r.EmissiveTriangles().size());
    } catch (const std::exception& e) {
        fprintf(stderr, "host_main: %s\n", e.what());
        return 1;
    }
    return 0;
}
