// ObjLoader.cpp — see ObjLoader.h.
#include "ObjLoader.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <fstream>
#include <map>
#include <sstream>
#include <unordered_map>

#include "Renderer.h"

namespace rdx {
namespace {

struct MtlRec { std::string name; float Kd[3] = {0, 0, 0}, Ks[3] = {0, 0, 0}, Ke[3] = {0, 0, 0}; float d = 1.0f, Pr = 0, Pm = 0, Ps = 0, Pc = 0; bool has_d = false; };

std::string trim(const std::string& s) {
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
std::string dir_of(const std::string& path) {
    size_t p = path.find_last_of("/\\");
    return p == std::string::npos ? std::string() : path.substr(0, p + 1);
}
void read3(std::istringstream& ss, float* v) { for (int i = 0; i < 3; i++) if (!(ss >> v[i])) { v[i] = 0.0f; ss.clear(); } }   // missing = 0 (tinyobj parseReal3)

bool load_mtl(const std::string& path, std::vector<MtlRec>& out) {
    std::ifstream f(path);
    if (!f) return false;
    std::string line;
    MtlRec* cur = nullptr;
    while (std::getline(f, line)) {
        line = trim(line);
        if (line.empty() || line[0] == '#') continue;
        std::istringstream ss(line);
        std::string key; ss >> key;
        if (key == "newmtl") { out.emplace_back(); cur = &out.back(); std::string rest; std::getline(ss, rest); cur->name = trim(rest); continue; }
        if (!cur) continue;
        if (key == "Kd") read3(ss, cur->Kd);
        else if (key == "Ks") read3(ss, cur->Ks);
        else if (key == "Ke") read3(ss, cur->Ke);
        else if (key == "d") { ss >> cur->d; cur->has_d = true; }
        else if (key == "Tr") { float tr; if ((ss >> tr) && !cur->has_d) cur->d = 1.0f - tr; }   // `d` wins over `Tr` (tinyobj 1.0.5)
        else if (key == "Pr") ss >> cur->Pr;
        else if (key == "Pm") ss >> cur->Pm;
        else if (key == "Ps") ss >> cur->Ps;
        else if (key == "Pc") ss >> cur->Pc;
    }
    return true;
}

struct FaceVert { int v, vn; };
// "a", "a/b", "a//c", "a/b/c"; negative indices are relative to the current count
bool parse_face_vertex(const std::string& tok, int nv, int nn, FaceVert& out) {
    int idx[3] = {0, 0, 0}; int k = 0; size_t pos = 0;
    while (k < 3 && pos <= tok.size()) {
        size_t slash = tok.find('/', pos);
        std::string part = tok.substr(pos, slash == std::string::npos ? std::string::npos : slash - pos);
        idx[k++] = part.empty() ? 0 : atoi(part.c_str());
        if (slash == std::string::npos) break;
        pos = slash + 1;
    }
    auto fix = [](int i, int n) { return i > 0 ? i - 1 : (i < 0 ? n + i : -1); };
    out.v = fix(idx[0], nv); out.vn = fix(idx[2], nn);
    return out.v >= 0 && out.v < nv;
}

struct PosKey {
    uint32_t b[3];
    bool operator==(const PosKey& o) const { return b[0] == o.b[0] && b[1] == o.b[1] && b[2] == o.b[2]; }
};
struct PosHash { size_t operator()(const PosKey& k) const { return (size_t)k.b[0] * 73856093u ^ (size_t)k.b[1] * 19349663u ^ (size_t)k.b[2] * 83492791u; } };
PosKey key_of(const float* p) {      // XMVector3Equal: +0 == -0 (Vertex.h:32-34)
    PosKey k;
    for (int i = 0; i < 3; i++) { float v = p[i] == 0.0f ? 0.0f : p[i]; memcpy(&k.b[i], &v, 4); }
    return k;
}

}  // namespace

ObjModel loadObjFile(const std::string& inputfile, uint32_t materialOffset, const std::string& material_search_path, uint32_t lut_seed) {
    ObjModel M;
    std::ifstream f(inputfile);
    if (!f) { M.error = "cannot open " + inputfile; return M; }
    const std::string mtl_dir = material_search_path.empty() ? dir_of(inputfile) : material_search_path;
    std::vector<float> pos, nrm;
    std::vector<MtlRec> mtls;
    std::map<std::string, int> mtl_index;
    struct Face { std::vector<FaceVert> fv; int mat; };
    std::vector<Face> faces;
    int cur_mat = -1;
    std::string line;
    while (std::getline(f, line)) {
        line = trim(line);
        if (line.empty() || line[0] == '#') continue;
        std::istringstream ss(line);
        std::string key; ss >> key;
        if (key == "v") { float v[3] = {0, 0, 0}; ss >> v[0] >> v[1] >> v[2]; pos.insert(pos.end(), v, v + 3); }
        else if (key == "vn") { float v[3] = {0, 0, 0}; ss >> v[0] >> v[1] >> v[2]; nrm.insert(nrm.end(), v, v + 3); }
        else if (key == "f") {
            Face fc; fc.mat = cur_mat;
            std::string tok;
            bool ok = true;
            while (ss >> tok) { FaceVert fv; if (!parse_face_vertex(tok, (int)pos.size() / 3, (int)nrm.size() / 3, fv)) { ok = false; break; } fc.fv.push_back(fv); }
            if (ok && fc.fv.size() >= 3) faces.push_back(std::move(fc));
        } else if (key == "usemtl") {
            std::string rest; std::getline(ss, rest); rest = trim(rest);
            auto it = mtl_index.find(rest);
            cur_mat = it == mtl_index.end() ? -1 : it->second;
        } else if (key == "mtllib") {
            std::string name;
            while (ss >> name) {                                   // several file names are allowed (tinyobj 1.0.4)
                const size_t before = mtls.size();
                if (load_mtl(mtl_dir + name, mtls)) {
                    for (size_t i = before; i < mtls.size(); i++) if (!mtl_index.count(mtls[i].name)) mtl_index[mtls[i].name] = (int)i;
                    break;
                }
            }
        }
    }
    // ---- materials: [default, mtl...] (src/Util/ObjLoader.h:415-441)
    rtx_material def; memset(&def, 0, sizeof def);
    def.Kd[0] = def.Kd[1] = def.Kd[2] = def.Kd[3] = 1.0f; def.Ks[0] = def.Ks[1] = def.Ks[2] = 1.0f; def.Ni = 1.0f; def.Pr_Pm_Ps_Pc[0] = 1.0f;
    M.materials.push_back(def); M.material_names.push_back("");
    const uint32_t first = materialOffset + 1u;
    for (const MtlRec& m : mtls) {
        rtx_material t; memset(&t, 0, sizeof t);
        t.Kd[0] = m.Kd[0]; t.Kd[1] = m.Kd[1]; t.Kd[2] = m.Kd[2]; t.Kd[3] = m.d;
        t.Ks[0] = m.Ks[0]; t.Ks[1] = m.Ks[1]; t.Ks[2] = m.Ks[2]; t.Ni = 1.0f;
        t.Ke[0] = m.Ke[0]; t.Ke[1] = m.Ke[1]; t.Ke[2] = m.Ke[2];
        t.Pr_Pm_Ps_Pc[0] = m.Pr; t.Pr_Pm_Ps_Pc[1] = m.Pm; t.Pr_Pm_Ps_Pc[2] = m.Ps; t.Pr_Pm_Ps_Pc[3] = m.Pc;
        GenerateEssLUT(t, lut_seed);
        M.materials.push_back(t); M.material_names.push_back(m.name);
    }
    // ---- geometry (:444-491)
    std::unordered_map<PosKey, uint32_t, PosHash> unique;
    auto emit_vertex = [&](const FaceVert& fv) {
        const float* p = &pos[3 * (size_t)fv.v];
        const PosKey k = key_of(p);
        auto it = unique.find(k);
        if (it == unique.end()) {
            rtx_vertex v; memset(&v, 0, sizeof v);
            v.position[0] = p[0]; v.position[1] = p[1]; v.position[2] = p[2];
            if (fv.vn >= 0 && (size_t)fv.vn * 3 + 2 < nrm.size()) { v.normal_material[0] = nrm[3 * fv.vn]; v.normal_material[1] = nrm[3 * fv.vn + 1]; v.normal_material[2] = nrm[3 * fv.vn + 2]; }
            it = unique.emplace(k, (uint32_t)M.vertices.size()).first;
            M.vertices.push_back(v);
        }
        M.indices.push_back(it->second);
    };
    auto emit_tri = [&](const Face& fc, int a, int b, int c) {
        const uint32_t id = (uint32_t)((int)first + fc.mat);      // mat == -1 -> the default material
        for (int k = 0; k < 3; k++) M.material_ids.push_back(id);
        emit_vertex(fc.fv[a]); emit_vertex(fc.fv[b]); emit_vertex(fc.fv[c]);
    };
    for (const Face& fc : faces) {
        const int n = (int)fc.fv.size();
        if (n == 3) emit_tri(fc, 0, 1, 2);
        else if (n == 4) {                                        // tiny_obj_loader.h:1511-1600: the shorter diagonal
            const float* v0 = &pos[3 * (size_t)fc.fv[0].v]; const float* v1 = &pos[3 * (size_t)fc.fv[1].v];
            const float* v2 = &pos[3 * (size_t)fc.fv[2].v]; const float* v3 = &pos[3 * (size_t)fc.fv[3].v];
            const float e02[3] = {v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]}, e13[3] = {v3[0] - v1[0], v3[1] - v1[1], v3[2] - v1[2]};
            const float sqr02 = e02[0] * e02[0] + e02[1] * e02[1] + e02[2] * e02[2], sqr13 = e13[0] * e13[0] + e13[1] * e13[1] + e13[2] * e13[2];
            if (sqr02 < sqr13) { emit_tri(fc, 0, 1, 2); emit_tri(fc, 0, 2, 3); }
            else { emit_tri(fc, 0, 1, 3); emit_tri(fc, 1, 2, 3); }
        } else {
            for (int k = 1; k + 1 < n; k++) emit_tri(fc, 0, k, k + 1);
        }
    }
    return M;
}

}  // namespace rdx

// ---- C interface for the Python harness
extern "C" {
void* rdx_obj_load(const char* path, uint32_t material_offset, const char* mtl_dir, uint32_t lut_seed) {
    rdx::ObjModel* m = new rdx::ObjModel(rdx::loadObjFile(path, material_offset, mtl_dir ? mtl_dir : "", lut_seed));
    return m;
}
void rdx_obj_free(void* h) { delete (rdx::ObjModel*)h; }
const char* rdx_obj_error(void* h) { return ((rdx::ObjModel*)h)->error.c_str(); }
// counts: [vertices, indices, material_ids, materials]
void rdx_obj_counts(void* h, uint32_t* out4) {
    rdx::ObjModel* m = (rdx::ObjModel*)h;
    out4[0] = (uint32_t)m->vertices.size(); out4[1] = (uint32_t)m->indices.size(); out4[2] = (uint32_t)m->material_ids.size(); out4[3] = (uint32_t)m->materials.size();
}
void rdx_obj_copy(void* h, rtx_vertex* v, uint32_t* idx, uint32_t* mids, rtx_material* mats) {
    rdx::ObjModel* m = (rdx::ObjModel*)h;
    if (v) memcpy(v, m->vertices.data(), m->vertices.size() * sizeof(rtx_vertex));
    if (idx) memcpy(idx, m->indices.data(), m->indices.size() * 4);
    if (mids) memcpy(mids, m->material_ids.data(), m->material_ids.size() * 4);
    if (mats) memcpy(mats, m->materials.data(), m->materials.size() * sizeof(rtx_material));
}
}
