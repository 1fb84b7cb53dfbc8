// bvh_build.h -- host-side BVH builder (C++17, std::thread).  Replaces the Embree BVH that
// Open3D builds lazily inside RaycastingScene.add_triangles / first cast_rays
// (reference call sites: models/mat_nvdiffrast.py:98-101, models/tracer_o3d_irt.py:85-89).
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include "texir_bvh.h"

namespace texir {

struct HostBVH {
    // BVH2 (layout: texir_bvh.h)
    std::vector<float4> nodes2;
    std::vector<float4> tris2;  // 3 per slot
    std::vector<uint32_t> prim2;
    // the same BVH2 with 32-byte nodes: child boxes on a global 16-bit grid (layout: texir_bvh.h)
    std::vector<uint4> qnodes2;  // 2 per node
    float qmin[3], qstep[3];
    // BVH8 compressed wide (layout: texir_bvh.h)
    std::vector<uint4> nodes8;  // 5 per node
    std::vector<float4> tris8;
    std::vector<uint32_t> prim8;
    int depth2 = 0, depth8 = 0;
    int top_nodes2 = 0;  // nodes2[0 .. top_nodes2) are the top levels in breadth-first order
    double sah2 = 0.0, sah8 = 0.0;
    double build_seconds = 0.0;
    float scene_lo[3], scene_hi[3];
};

// verts: [nv][3] float32, tris: [nt][3] int32.  Returns false and fills err on bad input.
// want8 = false skips the 8-wide layout (it is only walked after texir_scene_set_bvh(scene, 8)).
bool build_bvh(const float* verts, int64_t nv, const int32_t* tris, int64_t nt, int threads, HostBVH& out,
               std::string& err, bool want8 = true);

// The same binned-SAH tree built on the current CUDA device (bvh_build_gpu.cu): the BVH2 / QBVH2 node arrays, the
// slot-ordered triangle records and the primitive -> slot map, all device memory owned by the caller on success.
// Inputs are HOST arrays (they are uploaded).  Returns false with `err` set when the mesh is invalid or the tree
// cannot be built on the device (deeper than the traversal stack, whole scene in one leaf): the host builder then
// decides.
struct DeviceBVH {
    void* d_nodes2 = nullptr;        // float4 [4 * num_nodes]
    void* d_qnodes2 = nullptr;       // uint4  [2 * num_nodes]
    void* d_tris2 = nullptr;         // float4 [3 * nt]
    void* d_slot_of_prim = nullptr;  // uint32 [nt]
    int64_t num_nodes = 0;
    int top_nodes = 0;               // nodes [0, top_nodes) are the top levels, breadth first (the rest depth first)
    int depth = 0;
    double sah = 0.0, seconds = 0.0;
    float qmin[3], qstep[3];
    float scene_lo[3], scene_hi[3];
};
bool build_bvh_device(const float* verts_host, int64_t nv, const int32_t* tris_host, int64_t nt, void* stream, DeviceBVH& out,
                      std::string& err);

}  // namespace texir
