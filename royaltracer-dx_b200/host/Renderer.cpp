// Renderer.cpp — see Renderer.h.  Host-side preparation of the hot path's inputs, written against the C ABI.
#include "Renderer.h"
#include "ObjLoader.h"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <stdexcept>

namespace rdx {

// ------------------------------------------------------------------------------------------ DirectXMath-shaped helpers
XMMATRIX XMMatrixIdentity() {
    XMMATRIX r; memset(&r, 0, sizeof r);
    r.m[0][0] = r.m[1][1] = r.m[2][2] = r.m[3][3] = 1.0f;
    return r;
}
XMMATRIX XMMatrixMultiply(const XMMATRIX& a, const XMMATRIX& b) {
    XMMATRIX r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            float s = 0.0f;
            for (int k = 0; k < 4; k++) s += a.m[i][k] * b.m[k][j];
            r.m[i][j] = s;
        }
    return r;
}
XMMATRIX XMMatrixTranspose(const XMMATRIX& a) {
    XMMATRIX r;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.m[i][j] = a.m[j][i];
    return r;
}
// general 4x4 inverse (Gauss-Jordan with partial pivoting in binary64, rounded once to binary32)
XMMATRIX XMMatrixInverse(const XMMATRIX& a) {
    double w[4][8];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { w[i][j] = a.m[i][j]; w[i][4 + j] = (i == j) ? 1.0 : 0.0; }
    for (int c = 0; c < 4; c++) {
        int p = c;
        for (int r = c + 1; r < 4; r++) if (fabs(w[r][c]) > fabs(w[p][c])) p = r;
        if (p != c) for (int j = 0; j < 8; j++) std::swap(w[p][j], w[c][j]);
        double d = w[c][c];
        if (d == 0.0) { XMMATRIX z; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) z.m[i][j] = INFINITY; return z; }   // XMMatrixInverse returns an infinite matrix
        for (int j = 0; j < 8; j++) w[c][j] /= d;
        for (int r = 0; r < 4; r++) {
            if (r == c) continue;
            double f = w[r][c];
            if (f != 0.0) for (int j = 0; j < 8; j++) w[r][j] -= f * w[c][j];
        }
    }
    XMMATRIX r;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.m[i][j] = (float)w[i][4 + j];
    return r;
}
XMMATRIX XMMatrixPerspectiveFovRH(float fovY, float aspect, float zn, float zf) {
    float s = sinf(0.5f * fovY), c = cosf(0.5f * fovY);
    float h = c / s, w = h / aspect, fr = zf / (zn - zf);
    XMMATRIX r; memset(&r, 0, sizeof r);
    r.m[0][0] = w; r.m[1][1] = h; r.m[2][2] = fr; r.m[2][3] = -1.0f; r.m[3][2] = fr * zn;
    return r;
}
XMMATRIX XMMatrixRotationY(float angle) {
    float s = sinf(angle), c = cosf(angle);
    XMMATRIX r = XMMatrixIdentity();
    r.m[0][0] = c; r.m[0][2] = -s; r.m[2][0] = s; r.m[2][2] = c;
    return r;
}
XMMATRIX XMMatrixTranslation(float x, float y, float z) { XMMATRIX r = XMMatrixIdentity(); r.m[3][0] = x; r.m[3][1] = y; r.m[3][2] = z; return r; }
XMMATRIX XMMatrixScaling(float x, float y, float z) { XMMATRIX r = XMMatrixIdentity(); r.m[0][0] = x; r.m[1][1] = y; r.m[2][2] = z; return r; }

static void norm3(float v[3]) { float l = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); v[0] /= l; v[1] /= l; v[2] /= l; }
static void cross3(const float a[3], const float b[3], float r[3]) { r[0] = a[1] * b[2] - a[2] * b[1]; r[1] = a[2] * b[0] - a[0] * b[2]; r[2] = a[0] * b[1] - a[1] * b[0]; }
static float dot3(const float a[3], const float b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

XMMATRIX LookAtAsCopied(const float eye[3], const float center[3], const float up[3]) {
    float f[3] = {center[0] - eye[0], center[1] - eye[1], center[2] - eye[2]}; norm3(f);
    float s[3]; cross3(f, up, s); norm3(s);
    float u[3]; cross3(s, f, u);
    // glm column-major: g[col][row]; memcpy into XMMATRIX.m[row'][col'] = g[row'][col'] (same linear memory)
    XMMATRIX r = XMMatrixIdentity();
    r.m[0][0] = s[0]; r.m[1][0] = s[1]; r.m[2][0] = s[2];
    r.m[0][1] = u[0]; r.m[1][1] = u[1]; r.m[2][1] = u[2];
    r.m[0][2] = -f[0]; r.m[1][2] = -f[1]; r.m[2][2] = -f[2];
    r.m[3][0] = -dot3(s, eye); r.m[3][1] = -dot3(u, eye); r.m[3][2] = dot3(f, eye);
    return r;
}

// ------------------------------------------------------------------------------------------ F23 ESS LUT (ObjLoader.h:139-387)
namespace {
struct V3 { float x, y, z; };
inline V3 v3(float x, float y, float z) { V3 r = {x, y, z}; return r; }
inline V3 add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 mul(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline V3 normalize(V3 a) { float l = sqrtf(dot(a, a)); return v3(a.x / l, a.y / l, a.z / l); }
const float kPi = 3.14159265359f;   // ObjLoader.h:22

float G1_SmithGGX(float NdotV, float alpha) {           // :81-86
    float a2 = alpha * alpha;
    float d = sqrtf(a2 + (1.0f - a2) * NdotV * NdotV) + NdotV;
    return 2.0f * NdotV / std::max(d, 1e-7f);
}
float G2_SmithGGX(float NdotV, float NdotL, float alpha) {   // :89-94
    float a2 = alpha * alpha;
    float dA = NdotV * sqrtf(a2 + (1.0f - a2) * NdotL * NdotL);
    float dB = NdotL * sqrtf(a2 + (1.0f - a2) * NdotV * NdotV);
    return 2.0f * NdotL * NdotV / (dA + dB);
}
void CoordinateSystem(V3 N, V3& T1, V3& T2) {           // :97-104
    if (fabsf(N.z) < 0.999f) T1 = normalize(cross(v3(0, 0, 1), N)); else T1 = normalize(cross(v3(1, 0, 0), N));
    T2 = cross(N, T1);
}
V3 SampleGGX(float roughness, V3 outgoing, V3 normal, float e0, float e1) {   // :107-179
    float alpha = roughness * roughness;
    V3 N = normalize(normal), V = normalize(outgoing), T1, T2;
    CoordinateSystem(N, T1, T2);
    V3 Vh = normalize(v3(dot(T1, V), dot(T2, V), dot(N, V)));
    V3 Vs = normalize(v3(alpha * Vh.x, alpha * Vh.y, Vh.z));
    float lensq = Vs.x * Vs.x + Vs.y * Vs.y;
    V3 T1h, T2h;
    if (lensq > 0.0f) { float inv = 1.0f / sqrtf(lensq); T1h = normalize(v3(-Vs.y * inv, Vs.x * inv, 0.0f)); T2h = cross(Vs, T1h); }
    else { T1h = v3(1, 0, 0); T2h = v3(0, 1, 0); }
    float r = sqrtf(e0), phi = 2.0f * kPi * e1;
    float x = r * cosf(phi), y = r * sinf(phi);
    float z = sqrtf(std::max(0.0f, 1.0f - x * x - y * y));
    V3 Nhs = normalize(add(add(mul(T1h, x), mul(T2h, y)), mul(Vs, z)));
    V3 Nh = normalize(v3(alpha * Nhs.x, alpha * Nhs.y, std::max(0.0f, Nhs.z)));
    V3 H = normalize(add(add(mul(T1, Nh.x), mul(T2, Nh.y)), mul(N, Nh.z)));
    V3 negV = mul(V, -1.0f);
    V3 L = add(negV, mul(H, -2.0f * dot(negV, H)));       // XMVector3Reflect(I, N) = I - 2 dot(I,N) N
    return normalize(L);
}
float ComputeEss(V3 N, V3 V, float roughness, int numSamples, uint32_t& state) {   // :294-328
    auto next = [&]() {                                    // xorshift32 -> [0,1)
        state ^= state << 13; state ^= state >> 17; state ^= state << 5;
        return (float)(state >> 8) * (1.0f / 16777216.0f);
    };
    float Ess = 0.0f;
    for (int i = 0; i < numSamples; i++) {
        float u1 = next(), u2 = next();
        V3 L = SampleGGX(roughness, V, N, u1, u2);
        if (dot(N, L) <= 0.0f) continue;
        float NdotL = fabsf(dot(normalize(N), normalize(L)));
        // EvaluateBRDF_GGX (:183-195) with F = 1: G2 / max(4 NdotV NdotL, 1e-7)
        V3 Vn = normalize(V), Ln = normalize(L), Nn = normalize(N);
        float NdotV = std::max(dot(Nn, Vn), 0.0f), NdotLc = std::max(dot(Nn, Ln), 0.0f);
        float brdf = G2_SmithGGX(NdotV, NdotLc, roughness * roughness) / std::max(4.0f * NdotV * NdotLc, 1e-7f);
        // BRDF_PDF_GGX (:198-216): G1 / max(4 NdotV, 1e-7)
        float pdf = G1_SmithGGX(NdotV, roughness * roughness) / std::max(NdotV * 4.0f, 1e-7f);
        pdf = std::max(pdf, 1e-7f);
        if (brdf > 0.0f) Ess += (NdotL * brdf) / pdf;
    }
    return numSamples > 0 ? Ess / numSamples : 0.0f;
}
}  // namespace

void GenerateEssLUT(rtx_material& mat, uint32_t seed) {   // :351-387
    const float EPS = 0.04f;
    uint32_t state = seed ? seed : 1u;
    for (int i = 0; i < 16; i++) {
        float cosT = EPS + (float)i / 15.0f * (1.0f - EPS);
        float sinT = sqrtf(std::max(EPS, 1.0f - cosT * cosT));
        mat.LUT[i] = ComputeEss(v3(0, 0, 1), v3(sinT, 0, cosT), mat.Pr_Pm_Ps_Pc[0], 16000, state);
    }
}

// ------------------------------------------------------------------------------------------ slices as free functions
static void instance_properties(const XMMATRIX& xm, const rtx_instance_props* prev, rtx_instance_props* cur, rtx_instance_desc* desc,
                                uint32_t index, uint32_t model) {
    rtx_instance_props p;
    XMMATRIX inv = XMMatrixInverse(xm);
    XMMATRIX upper = xm;                                     // Renderer.cpp:2106-2117
    upper.m[0][3] = upper.m[1][3] = upper.m[2][3] = 0.0f;
    upper.m[3][0] = upper.m[3][1] = upper.m[3][2] = 0.0f; upper.m[3][3] = 1.0f;
    XMMATRIX nrm = XMMatrixTranspose(XMMatrixInverse(upper));
    memcpy(p.objectToWorld, xm.m, 64);
    memcpy(p.objectToWorldInverse, inv.m, 64);
    memcpy(p.objectToWorldNormal, nrm.m, 64);
    if (prev) {
        memcpy(p.prevObjectToWorld, prev->objectToWorld, 64);
        memcpy(p.prevObjectToWorldInverse, prev->objectToWorldInverse, 64);
        memcpy(p.prevObjectToWorldNormal, prev->objectToWorldNormal, 64);
    } else {
        memcpy(p.prevObjectToWorld, xm.m, 64); memcpy(p.prevObjectToWorldInverse, inv.m, 64); memcpy(p.prevObjectToWorldNormal, nrm.m, 64);
    }
    *cur = p;
    if (desc) {                                              // TopLevelASGenerator.cpp:181-199
        XMMATRIX t = XMMatrixTranspose(xm);
        memcpy(desc->transform, t.m, 48);
        desc->instance_id_mask = (index & 0xFFFFFFu) | (0xFFu << 24);
        desc->hit_group_flags = ((2u * index) & 0xFFFFFFu);
        desc->blas = model;
    }
}

static float triangle_weight(const float* v0, const float* v1, const float* v2, const float* ke) {   // Renderer.cpp:2217-2233
    float e1[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]}, e2[3] = {v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]}, c[3];
    cross3(e1, e2, c);
    float area = 0.5f * sqrtf(dot3(c, c));
    return area * ((ke[0] + ke[1] + ke[2]) / 3.0f);
}

static void collect_emissive(uint32_t n_instances, const uint32_t* inst_model, const rtx_vertex* const* verts, const uint32_t* const* indices,
                             const uint32_t* n_indices, const uint32_t* offsets, const uint32_t* material_ids, const rtx_material* materials,
                             std::vector<rtx_light_triangle>& out) {
    out.clear();
    for (uint32_t ii = 0; ii < n_instances; ii++) {
        uint32_t mi = inst_model[ii];
        uint32_t ntri = n_indices[mi] / 3;
        for (uint32_t t = 0; t < ntri; t++) {
            uint32_t m0 = material_ids[offsets[mi] + 3 * t], m1 = material_ids[offsets[mi] + 3 * t + 1], m2 = material_ids[offsets[mi] + 3 * t + 2];
            if (m0 != m1 || m0 != m2) continue;               // :2155-2158
            const rtx_material& mat = materials[m0];
            if (mat.Ke[0] + mat.Ke[1] + mat.Ke[2] > 0.0f) {
                const rtx_vertex &a = verts[mi][indices[mi][3 * t]], &b = verts[mi][indices[mi][3 * t + 1]], &c = verts[mi][indices[mi][3 * t + 2]];
                rtx_light_triangle lt; memset(&lt, 0, sizeof lt);
                memcpy(lt.x, a.position, 12); memcpy(lt.y, b.position, 12); memcpy(lt.z, c.position, 12);
                lt.instanceID = ii;
                lt.weight = triangle_weight(a.position, b.position, c.position, mat.Ke);
                memcpy(lt.emission, mat.Ke, 12);
                out.push_back(lt);
            }
        }
    }
    // descending by weight; stable so that ties keep scene order (the reference's std::sort leaves ties unspecified)
    std::stable_sort(out.begin(), out.end(), [](const rtx_light_triangle& a, const rtx_light_triangle& b) { return a.weight > b.weight; });
    float total = 0.0f;
    for (auto& t : out) total += t.weight;
    float cum = 0.0f;
    for (auto& t : out) { t.weight /= total; cum += t.weight; t.cdf = cum; t.total_weight = total; }
    if (!out.empty()) out.back().cdf = 1.0f;
    for (auto& t : out) t.triCount = (uint32_t)out.size();      // CreateEmissiveTrianglesBuffer :2241-2243
}

static void camera_params(const float* eye, const float* center, const float* up, float fovy_deg, float aspect, float zn, float zf,
                          const XMMATRIX* prevView, const XMMATRIX* prevProj, rtx_camera_params* out, XMMATRIX* viewOut, XMMATRIX* projOut) {
    XMMATRIX view = LookAtAsCopied(eye, center, up);
    float fov = fovy_deg * 3.141592654f / 180.0f;             // XM_PI
    XMMATRIX proj = XMMatrixPerspectiveFovRH(fov, aspect, zn, zf);
    XMMATRIX viewI = XMMatrixInverse(view), projI = XMMatrixInverse(proj);
    memset(out, 0, sizeof *out);
    memcpy(out->view, view.m, 64); memcpy(out->projection, proj.m, 64);
    memcpy(out->viewI, viewI.m, 64); memcpy(out->projectionI, projI.m, 64);
    memcpy(out->prevView, prevView ? prevView->m : view.m, 64);
    memcpy(out->prevProjection, prevProj ? prevProj->m : proj.m, 64);
    out->time = 0.0f;      // the seed uses the global sample index instead of wall-clock time (DESIGN.md deviation D2)
    if (viewOut) *viewOut = view;
    if (projOut) *projOut = proj;
}

// ------------------------------------------------------------------------------------------ Renderer
// (left out of librdx_prep.so, the engine-free build of the host-side data preparation that the CPU arms of bench.py and the
// Python scene builders load: -DRDX_PREP_ONLY)
#ifndef RDX_PREP_ONLY
Renderer::Renderer(uint32_t width, uint32_t height) : m_width(width), m_height(height), m_aspectRatio((float)width / (float)height) {
    memset(&m_camera, 0, sizeof m_camera);
    m_prevViewMatrix = XMMatrixIdentity(); m_prevProjMatrix = XMMatrixIdentity();
}
Renderer::~Renderer() { if (m_ctx) rtx_destroy(m_ctx); }

void Renderer::Check(int status, const char* what) {
    if (status != RTX_OK) throw std::runtime_error(std::string(what) + ": " + rtx_last_error());   // ThrowIfFailed
}

uint32_t Renderer::AddMaterial(const rtx_material& m) { m_materials.push_back(m); return (uint32_t)m_materials.size() - 1; }

uint32_t Renderer::CreateVB(const std::vector<rtx_vertex>& vertices, const std::vector<uint32_t>& indices, const std::vector<uint32_t>& material_ids) {
    if (material_ids.size() != indices.size()) throw std::logic_error("CreateVB: one material id per face-vertex expected");
    Model m; m.vertices = vertices; m.indices = indices;
    m.materialIDOffset = (uint32_t)m_materialIDs.size();     // m_materialIDOffsets, Renderer.cpp:1989-1996
    for (auto& v : m.vertices) v.normal_material[3] = (float)m.materialIDOffset;   // the reference's float smuggling, kept for layout fidelity
    m_materialIDs.insert(m_materialIDs.end(), material_ids.begin(), material_ids.end());
    m.rtxModel = 0;
    m_models.push_back(std::move(m));
    return (uint32_t)m_models.size() - 1;
}
// CreateVB(std::string name), rdn/Renderer.cpp:1973-2072: loadObjFile appends this model's material block to the global list
uint32_t Renderer::CreateVB(const std::string& obj_path) {
    ObjModel o = loadObjFile(obj_path, (uint32_t)m_materials.size());
    if (!o.error.empty()) throw std::logic_error("CreateVB: " + o.error);
    m_materials.insert(m_materials.end(), o.materials.begin(), o.materials.end());
    return CreateVB(o.vertices, o.indices, o.material_ids);
}
uint32_t Renderer::AddInstance(uint32_t model, const XMMATRIX& o2w) { m_instances.push_back({model, o2w}); return (uint32_t)m_instances.size() - 1; }
void Renderer::SetInstanceTransform(uint32_t i, const XMMATRIX& o2w) { m_instances.at(i).second = o2w; }
void Renderer::SetCamera(const float eye[3], const float center[3], const float up[3]) {
    memcpy(m_eye, eye, 12); memcpy(m_center, center, 12); memcpy(m_up, up, 12);
}

void Renderer::CollectEmissiveTriangles() {
    std::vector<uint32_t> inst_model; std::vector<const rtx_vertex*> v; std::vector<const uint32_t*> idx; std::vector<uint32_t> ni, off;
    for (auto& in : m_instances) inst_model.push_back(in.first);
    for (auto& m : m_models) { v.push_back(m.vertices.data()); idx.push_back(m.indices.data()); ni.push_back((uint32_t)m.indices.size()); off.push_back(m.materialIDOffset); }
    collect_emissive((uint32_t)m_instances.size(), inst_model.data(), v.data(), idx.data(), ni.data(), off.data(), m_materialIDs.data(),
                     m_materials.data(), m_emissiveTriangles);
}

void Renderer::UpdateInstancePropertiesBuffer() {
    std::vector<rtx_instance_props> prev = m_instanceProperties;
    m_instanceProperties.resize(m_instances.size());
    for (size_t i = 0; i < m_instances.size(); i++)
        instance_properties(m_instances[i].second, i < prev.size() ? &prev[i] : nullptr, &m_instanceProperties[i], nullptr, (uint32_t)i, m_instances[i].first);
}

void Renderer::UpdateCameraBuffer() {
    XMMATRIX view, proj;
    camera_params(m_eye, m_center, m_up, 60.0f, m_aspectRatio, 0.1f, 1000.0f, m_first ? nullptr : &m_prevViewMatrix,
                  m_first ? nullptr : &m_prevProjMatrix, &m_camera, &view, &proj);
    m_prevViewMatrix = view; m_prevProjMatrix = proj;
}

void Renderer::OnInit() {
    rtx_config cfg; memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = sizeof cfg; cfg.device = device; cfg.width = m_width; cfg.height = m_height;
    cfg.bounces = bounces; cfg.nee_samples = nee_samples; cfg.nee_samples_di = nee_samples_di; cfg.flags = flags;
    cfg.samples_per_pass = samples_per_pass; cfg.stream = nullptr;
    Check(rtx_create(&cfg, &m_ctx), "rtx_create");
    for (auto& m : m_models)                                  // CreateBottomLevelAS per model
        Check(rtx_upload_model(m_ctx, m.vertices.data(), (uint32_t)m.vertices.size(), m.indices.data(), (uint32_t)m.indices.size(),
                               m.materialIDOffset, &m.rtxModel), "rtx_upload_model");
    Check(rtx_set_material_ids(m_ctx, m_materialIDs.data(), (uint32_t)m_materialIDs.size()), "rtx_set_material_ids");
    Check(rtx_set_materials(m_ctx, m_materials.data(), (uint32_t)m_materials.size()), "rtx_set_materials");
    CollectEmissiveTriangles();
    Check(rtx_set_emissive_triangles(m_ctx, m_emissiveTriangles.data(), (uint32_t)m_emissiveTriangles.size()), "rtx_set_emissive_triangles");
    OnUpdate();
}

void Renderer::OnUpdate() {
    UpdateCameraBuffer();
    UpdateInstancePropertiesBuffer();
    std::vector<rtx_instance_desc> descs(m_instances.size());
    for (size_t i = 0; i < m_instances.size(); i++) {
        rtx_instance_props tmp;
        instance_properties(m_instances[i].second, nullptr, &tmp, &descs[i], (uint32_t)i, m_models[m_instances[i].first].rtxModel);
    }
    Check(rtx_set_instances(m_ctx, descs.data(), m_instanceProperties.data(), (uint32_t)descs.size()), "rtx_set_instances");   // TLAS refit, Renderer.cpp:594
    Check(rtx_set_camera(m_ctx, &m_camera), "rtx_set_camera");
    m_first = false;
}

void Renderer::OnRender(uint32_t first_sample, uint32_t n_samples) {
    Check(rtx_render_pass(m_ctx, first_sample, n_samples), "rtx_render_pass");
    if (m_world > 1) Check(rtx_reduce_accum(m_ctx), "rtx_reduce_accum");   // one ncclReduce of gPermanentData per pass
    Check(rtx_synchronize(m_ctx), "rtx_synchronize");          // WaitForPreviousFrame, Renderer.cpp:717-735
}
// multi-GPU: every rank runs one Renderer on its own device over the replicated scene and renders the samples s = rank (mod world)
void Renderer::InitComm(const void* nccl_unique_id128, int rank, int world) {
    Check(rtx_comm_init(m_ctx, nccl_unique_id128, rank, world), "rtx_comm_init");
    m_rank = rank; m_world = world;
}
// PopulateCommandList's three DispatchRays (RayGen, RayGen2, RayGen3), rdn/Renderer.cpp:611-673; needs flags |= RTX_FLAG_RESTIR
void Renderer::OnRenderFrame(uint32_t frame_index) {
    Check(rtx_render_frame(m_ctx, frame_index), "rtx_render_frame");
    Check(rtx_synchronize(m_ctx), "rtx_synchronize");
}
void Renderer::ReadAccumulation(std::vector<float>& out) { out.resize((size_t)m_width * m_height * 4); Check(rtx_read_accum(m_ctx, out.data()), "rtx_read_accum"); }
void Renderer::ReadOutput(std::vector<uint8_t>& out) { out.resize((size_t)m_width * m_height * 4); Check(rtx_read_output(m_ctx, out.data()), "rtx_read_output"); }
#endif  // RDX_PREP_ONLY

}  // namespace rdx

// ------------------------------------------------------------------------------------------ C wrappers
extern "C" {
void rdx_instance_properties(const float* xm, const uint32_t* model_ids, uint32_t n, rtx_instance_props* props, rtx_instance_desc* descs) {
    for (uint32_t i = 0; i < n; i++) {
        rdx::XMMATRIX m; memcpy(m.m, xm + 16 * i, 64);
        rdx::instance_properties(m, nullptr, &props[i], descs ? &descs[i] : nullptr, i, model_ids[i]);
    }
}
uint32_t rdx_collect_emissive_triangles(uint32_t n_instances, const uint32_t* inst_model, uint32_t n_models, const rtx_vertex* const* verts,
                                        const uint32_t* const* indices, const uint32_t* n_indices, const uint32_t* material_id_offsets,
                                        const uint32_t* material_ids, const rtx_material* materials, rtx_light_triangle* out, uint32_t cap) {
    (void)n_models;
    std::vector<rtx_light_triangle> v;
    rdx::collect_emissive(n_instances, inst_model, verts, indices, n_indices, material_id_offsets, material_ids, materials, v);
    if (v.size() <= cap && out) memcpy(out, v.data(), v.size() * sizeof(rtx_light_triangle));
    return (uint32_t)v.size();
}
void rdx_camera_params(const float* eye, const float* center, const float* up, float fovy_deg, float aspect, float zn, float zf, rtx_camera_params* out) {
    rdx::camera_params(eye, center, up, fovy_deg, aspect, zn, zf, nullptr, nullptr, out, nullptr, nullptr);
}
void rdx_generate_ess_lut(rtx_material* mat, uint32_t seed) { rdx::GenerateEssLUT(*mat, seed); }
}
