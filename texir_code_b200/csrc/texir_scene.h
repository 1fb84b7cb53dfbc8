// texir_scene.h -- the scene handle behind `texir_scene*` of the C ABI, shared by the translation units of
// libtexir_b200.so (kernels, device BVH builder, wide-walk variant, peer-memory collectives).
#pragma once
#include <stdint.h>

#include <vector>

#include "../../include/texir_b200.h"
#include "texir_bvh.h"

struct texir_scene {
    texir::SceneView view{};
    texir_scene_info info{};
    void* d_nodes2 = nullptr;
    void* d_tris2 = nullptr;
    void* d_qnodes2 = nullptr;
    void* d_nodes8 = nullptr;
    void* d_tris8 = nullptr;
    void* d_uvs = nullptr;
    void* d_slot_of_prim = nullptr;  // original primitive id -> BVH2 slot
    void* d_tex = nullptr;
    void* d_tex_rgbe = nullptr;      // optional 4-byte RGBE texels (texir_scene_set_texture_rgbe)
    void* d_wide = nullptr;          // optional warp-cooperative wide layout (texir_wide.cu)
    int64_t geom_bytes = 0, tex_bytes = 0;
    float lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};  // scene box
    // host copy of the mesh: the 8-wide layout is built from it when texir_scene_set_bvh(scene, 8) first asks for it
    std::vector<float> h_verts;
    std::vector<int32_t> h_tris;
    int builder = 0;  // TEXIR_BUILDER_HOST / TEXIR_BUILDER_DEVICE: who built the tree that is resident
};
