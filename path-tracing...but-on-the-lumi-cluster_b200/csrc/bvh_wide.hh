// Host side of the GPU acceleration layout: flattens the reference's link-table BVHs
// (bvh.hh:35-67, built by bvh.cc:43-229), or builds BLASes from the triangles, into compressed
// 8-wide nodes (and 4-wide float nodes for the older kernels) with pre-gathered triangle
// vertices, and builds one static TLAS over the static instances.
#pragma once
#include "../../include/ptgpu.h"
#include "pt_scene.cuh"

#include <string>
#include <vector>

namespace pt {

struct WideSubBox { float lo[3], hi[3]; };

struct WideBlasInfo
{
    uint32_t ref_node_offset, ref_node_count;   // the reference bvh handle this was built from
    ptgpu_mesh mesh;                            // derived; checked against every instance using it
    uint32_t max_stack;                         // stack entries a traversal of this BLAS can need
    // object-space boxes of the subtrees six levels down (<= 64): the world box of a per-frame instance
    // is the union of THEIR transformed corners, much tighter under rotation than the transformed root box
    std::vector<WideSubBox> sub_boxes;
};

struct WideScene
{
    std::vector<WideNode> nodes;        // all BLASes, concatenated
    std::vector<float4> tris;           // 3 per triangle slot
    std::vector<WideBlas> blas;
    std::vector<WideBlasInfo> blas_info;
    std::vector<WideInstance> instances; // static instances
    std::vector<WideNode> tlas;         // static TLAS
    uint32_t max_stack = 0;             // worst case over TLAS + any BLAS (+1 for the exit marker)

    // compressed 8-wide layout (80-byte nodes, 8-bit child boxes, octant-ordered slots)
    std::vector<float4> cw_nodes;       // 5 per node
    std::vector<float4> cw_tris;        // 3 per triangle
    std::vector<uint32_t> cw_inst_index;
    std::vector<uint32_t> cw_blas_root; // per BLAS: root node index
    uint32_t cw_tlas_root = 0;
    uint32_t cw_max_stack = 0;          // group-stack entries a traversal can need
    bool from_meshes = false;           // BLASes built from triangles (ptgpu_upload_meshes), not recovered
};

// The static instances without instancing: one compressed 8-wide BVH over their world-space triangles
// (bvh_wide.cu, "flat static scene"). Node and triangle indices are already relocated by the bases the
// caller passed, so the arrays can be appended to the device copies of cw_nodes / cw_tris.
struct FlatScene
{
    std::vector<float4> nodes;          // 5 per node; node 0 = root; [0, n_top) = the top tree, breadth-first
    std::vector<float4> tris;           // 3 per triangle: world-space vertices; v0.w = primitive id, v1.w = instance | mirrored << 31
    uint32_t depth = 0, n_top = 0;
    size_t n_tris = 0;
    float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    double build_seconds = 0.0;
};

constexpr int CW_WIDTH = 8;             // children per node
constexpr int CW_LEAF_MAX = 1;          // triangles per leaf child (the encoding allows 3; measured 1: 245 ms, 2: 251, 3: 255 on frame 520)
#ifndef CW_OPTIMAL_COLLAPSE
#define CW_OPTIMAL_COLLAPSE 1        // cost-optimal (dynamic programming) collapse into 8-wide nodes; 0 = greedy
#endif
constexpr int CW_STACK = 48;            // traversal stack entries (uint2 each, pt_cwbvh.cuh)

constexpr int WIDE_LEAF_MAX = 4;        // triangles per leaf
constexpr int WIDE_STACK = 96;          // traversal stack entries (pt_wide.cuh)

bool build_wide_scene(
    const ptgpu_bvh_node* nodes, size_t n_nodes, const ptgpu_bvh_link* links,
    const uint32_t* indices, size_t n_indices, const ptgpu_float3* pos, size_t n_verts,
    const ptgpu_tlas_instance* instances, size_t n_static,
    WideScene& out, std::string& err,
    const ptgpu_mesh* meshes = nullptr, size_t n_meshes = 0);  // nodes == nullptr: build every BLAS from `meshes`

bool build_flat_scene(const WideScene& ws, const uint32_t* indices, const ptgpu_float3* pos,
                      const ptgpu_tlas_instance* instances, size_t n_static,
                      uint32_t node_base, uint32_t tri_base, FlatScene& out, std::string& err);

uint64_t verify_flat_scene(const FlatScene& fs, uint32_t node_base, uint32_t tri_base, std::string& err);

// Fills the traversal record of one instance (static or per-frame). False if its BLAS is unknown.
bool make_wide_instance(const WideScene& ws, const ptgpu_tlas_instance& inst, uint32_t ref_index, WideInstance& out);

// Structural self-check of a flattened scene (CPU tests). Returns the number of violations.
uint64_t verify_wide_scene(const WideScene& ws, size_t n_static, std::string& err);

} // namespace pt
