// Renderer.h — C++ host side above the C ABI, mirroring the slices of the reference's rdn/Renderer.{h,cpp} that
// define the inputs and the call sequence of the hot path:
//   CreateVB / model ingest           rdn/Renderer.cpp:1973-2072   -> Renderer::CreateVB
//   CreateAccelerationStructures      :893-946                      -> Renderer::CreateAccelerationStructures
//   CollectEmissiveTriangles          :2123-2213                    -> Renderer::CollectEmissiveTriangles
//   UpdateInstancePropertiesBuffer    :2091-2121                    -> Renderer::UpdateInstancePropertiesBuffer
//   UpdateCameraBuffer                :1722-1768                    -> Renderer::UpdateCameraBuffer
//   OnInit / OnUpdate / OnRender      :44-103, :431-452, :468-506   -> same names
// Matrices follow the reference's memory conventions: XMMATRIX = row-major, row-vector (p' = p * M).
// Everything D3D12 / Win32 / Streamline in the reference's Renderer is out of scope.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/rtx_b200.h"

namespace rdx {

struct XMMATRIX { float m[4][4]; };       // row-major, row-vector convention (DirectXMath)

XMMATRIX XMMatrixIdentity();
XMMATRIX XMMatrixMultiply(const XMMATRIX& a, const XMMATRIX& b);
XMMATRIX XMMatrixTranspose(const XMMATRIX& a);
XMMATRIX XMMatrixInverse(const XMMATRIX& a);
XMMATRIX XMMatrixPerspectiveFovRH(float fovY, float aspect, float zn, float zf);
XMMATRIX XMMatrixRotationY(float angle);
XMMATRIX XMMatrixTranslation(float x, float y, float z);
XMMATRIX XMMatrixScaling(float x, float y, float z);
// glm::lookAt(eye, center, up) (right-handed) in glm's column-major memory, memcpy'd into an XMMATRIX exactly like
// rdn/Renderer.cpp:1726-1727 does.
XMMATRIX LookAtAsCopied(const float eye[3], const float center[3], const float up[3]);

// F23, src/Util/ObjLoader.h:294-387 with a fixed-seed generator (the reference seeds from std::random_device)
void GenerateEssLUT(rtx_material& mat, uint32_t seed = 12345u);

class Renderer {
public:
    Renderer(uint32_t width, uint32_t height);
    ~Renderer();

    // rtx_config fields the BASELINE configs vary; set before OnInit
    uint32_t bounces = 3, nee_samples = 4, nee_samples_di = 4, flags = 0, samples_per_pass = 1;
    int device = 0;

    // scene description (what loadObjFile + LoadAssets produce): global material list / per-face-vertex ids
    uint32_t AddMaterial(const rtx_material& m);
    // one model = one BLAS; material_ids has 3 entries per triangle (global material indices).  Returns the model index.
    uint32_t CreateVB(const std::vector<rtx_vertex>& vertices, const std::vector<uint32_t>& indices,
                      const std::vector<uint32_t>& material_ids);
    uint32_t AddInstance(uint32_t model, const XMMATRIX& objectToWorld);
    void SetInstanceTransform(uint32_t instance, const XMMATRIX& objectToWorld);
    void SetCamera(const float eye[3], const float center[3], const float up[3]);

    void OnInit();                       // uploads, BLAS/TLAS builds, light list, buffers
    void OnUpdate();                     // camera CB + instance properties (+ TLAS refit)
    void OnRender(uint32_t first_sample, uint32_t n_samples);
    void OnRenderFrame(uint32_t frame_index);   // the reference's full frame incl. ReSTIR reuse (flags |= RTX_FLAG_RESTIR)
    // CreateVB(std::string) of the reference: OBJ/MTL ingest through ObjLoader.h, materials appended to the global list
    uint32_t CreateVB(const std::string& obj_path);
    // multi-GPU (after OnInit): join the job's communicator; OnRender then ends with the per-pass reduce to rank 0
    void InitComm(const void* nccl_unique_id128, int rank, int world);
    void ReadAccumulation(std::vector<float>& rgba32f);
    void ReadOutput(std::vector<uint8_t>& rgba8);

    // the slices, exposed for tests
    void CollectEmissiveTriangles();
    void UpdateInstancePropertiesBuffer();
    void UpdateCameraBuffer();

    const std::vector<rtx_light_triangle>& EmissiveTriangles() const { return m_emissiveTriangles; }
    const std::vector<rtx_instance_props>& InstanceProperties() const { return m_instanceProperties; }
    const rtx_camera_params& CameraBuffer() const { return m_camera; }
    rtx_ctx* Context() { return m_ctx; }

private:
    struct Model { std::vector<rtx_vertex> vertices; std::vector<uint32_t> indices; uint32_t materialIDOffset; uint32_t rtxModel; };
    uint32_t m_width, m_height;
    float m_aspectRatio;
    std::vector<Model> m_models;
    std::vector<rtx_material> m_materials;
    std::vector<uint32_t> m_materialIDs;
    std::vector<std::pair<uint32_t, XMMATRIX>> m_instances;      // (model index, objectToWorld)
    std::vector<rtx_instance_props> m_instanceProperties;
    std::vector<rtx_light_triangle> m_emissiveTriangles;
    rtx_camera_params m_camera;
    XMMATRIX m_prevViewMatrix, m_prevProjMatrix;
    float m_eye[3] = {-1.5f, 1.5f, 3.5f}, m_center[3] = {0.f, 1.f, 0.f}, m_up[3] = {0.f, 1.f, 0.f};   // rdn/Renderer.cpp:47-48
    rtx_ctx* m_ctx = nullptr;
    bool m_first = true;
    int m_rank = 0, m_world = 1;
    void Check(int status, const char* what);
};

}  // namespace rdx

// C wrappers of the slices (used by the Python tests / bench through ctypes)
extern "C" {
// in: n XMMATRIX (16 floats each, row-major row-vector).  out: props (prev* = current on first call) and TLAS descs.
void rdx_instance_properties(const float* xm, const uint32_t* model_ids, uint32_t n, rtx_instance_props* props, rtx_instance_desc* descs);
// F21.  models are given as flat arrays; returns the number of light triangles written (<= cap); if cap is too small returns the needed count.
uint32_t rdx_collect_emissive_triangles(uint32_t n_instances, const uint32_t* inst_model, uint32_t n_models,
                                        const rtx_vertex* const* verts, const uint32_t* const* indices, const uint32_t* n_indices,
                                        const uint32_t* material_id_offsets, const uint32_t* material_ids, const rtx_material* materials,
                                        rtx_light_triangle* out, uint32_t cap);
void rdx_camera_params(const float* eye, const float* center, const float* up, float fovy_deg, float aspect, float zn, float zf,
                       rtx_camera_params* out);
void rdx_generate_ess_lut(rtx_material* mat, uint32_t seed);
}
