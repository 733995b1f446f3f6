// ObjLoader.h — OBJ/MTL ingest of the plugin surface, mirroring the reference's
// ObjLoader::loadObjFile (src/Util/ObjLoader.h:393-495) on top of an own minimal Wavefront parser (the reference parses with
// tiny_obj_loader, lib/tiny_obj_loader.h; this file restates the behaviour loadObjFile relies on, nothing else):
//   * materials: every model contributes the block [default, mtl_0, mtl_1, ...] to the global list; the default is
//     Kd = (1,1,1,1), Ks = (1,1,1), Pr_Pm_Ps_Pc = (1,0,0,0), LUT = 0 (:415-417, Vertex.h:14-23); an MTL material becomes
//     Kd = (Kd, d), Pr_Pm_Ps_Pc = (Pr, Pm, Ps, Pc), Ke, Ks, Ni = 1, LUT = GenerateEssLUT (fixed seed; F23) (:420-441);
//   * material ids: one entry per face-vertex = the face's `usemtl` index + materialOffset, or the default when the face has
//     none (tinyobj id -1) (:452-460);
//   * vertices: de-duplicated by POSITION only — the first normal seen for a position wins (Vertex.h:32-34,48), missing
//     normals are (0,0,0) (:470-476);
//   * polygons: triangles as they are; quads split along the shorter diagonal, larger polygons as a fan
//     (tiny_obj_loader.h:1510-1600 with triangulate = true; its ear clipping of concave polygons is not restated).
// The per-model material-id offset is returned as an integer instead of being smuggled in vertex.normal.w (deviation D7).
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/rtx_b200.h"

namespace rdx {

struct ObjModel {
    std::vector<rtx_vertex> vertices;
    std::vector<uint32_t> indices;
    std::vector<uint32_t> material_ids;      // 3 per triangle, global material indices
    std::vector<rtx_material> materials;     // this model's block [default, mtl...]
    std::vector<std::string> material_names; // "" for the default
    std::string error;                       // non-empty: the load failed
};

// materialOffset: size of the global material list before this model (the reference's *materialOffset on entry).
ObjModel loadObjFile(const std::string& inputfile, uint32_t materialOffset, const std::string& material_search_path = "", uint32_t lut_seed = 12345u);

}  // namespace rdx
