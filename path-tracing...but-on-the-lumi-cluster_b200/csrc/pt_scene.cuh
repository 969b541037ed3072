// Device-side scene description shared by all kernels.
#pragma once
#include "pt_math.cuh"
#include "pt_shade.cuh"

// Compressed 8-wide nodes: the empty child slots of a node carry an INVERTED box (lo 255, hi 0: no ray
// passes the slab test) and the meta byte of one of the node's real children, so the box test needs no
// per-child validity mask: a child that passes sets bit (1 << its index). Needs one triangle per leaf child.

namespace pt {

// ---- reference-layout records read on the device (layouts: include/ptgpu.h) -------------------

// tlas_instance (bvh.hh:73-79), 160 B = 10 x 16 B
struct RefInstance
{
    uint32_t blas_node_count, blas_node_offset;               // bvh blas
    uint32_t vertex_count, triangle_count, index_offset, base_vertex; // mesh m
    uint32_t pad[2];
    float4 transform[4];      // columns; [3] = translation (math.hh:330-338)
    float4 inv_transform[4];
};
static_assert(sizeof(RefInstance) == 160, "tlas_instance layout");

// subframe (scene.hh:27-35), 160 B
struct RefSubframe
{
    uint32_t tlas_node_count, tlas_node_offset;
    uint32_t pad0[2];
    // camera (scene.hh:7-18) @16
    float4 orient[3];         // mat3 columns (16-byte float3s)
    float4 position;          // @64
    float aspect_ratio, inv_focal_length, focal_distance, aperture_angle; // @80
    int32_t aperture_polygon; float aperture_radius; uint32_t pad1[2];    // @96
    // directional_light (scene.hh:20-25) @112
    float4 light_dir;
    float4 light_color;
    float cos_solid_angle; uint32_t pad2[3];
};
static_assert(sizeof(RefSubframe) == 160, "subframe layout");

// ---- GPU traversal layout (built by bvh_build.cc from the reference arrays) -------------------

// 4-wide BVH node, 128 B = 8 x 16 B. Child boxes are stored SoA so that one LDG.128 brings the
// same bound of all four children. child[i]:
//   0xFFFFFFFF            empty slot
//   bit31 = 0             inner node index (into the same BVH's node array, relative)
//   bit31 = 1             leaf: bits 0..26 = first triangle slot (relative), bits 27..30 = count-1
struct WideNode
{
    float4 lox, loy, loz, hix, hiy, hiz;
    uint4 child;
    uint4 pad;
};
static_assert(sizeof(WideNode) == 128, "wide node");

// One BLAS in the wide layout.
struct WideBlas
{
    uint32_t node_offset;    // first WideNode
    uint32_t tri_offset;     // first triangle slot (3 float4 per slot in wide_tris)
    uint32_t node_count;
    uint32_t tri_count;
    float4 lo, hi;           // object-space bounds (root)
};

// Per-instance record for the wide traversal, 128 B.
struct WideInstance
{
    float4 inv0, inv1, inv2;  // rows of the 3x4 world->object matrix (so a transform is 3 dot4s)
    float4 lo, hi;            // world-space AABB of the instance (as build_tlas computes, bvh.cc:262-278)
    uint32_t blas;            // index into WideBlas
    uint32_t ref_instance;    // index into the reference instance array (for shading)
    uint32_t cw_root;         // root node of the instance's BLAS in the compressed 8-wide array
    uint32_t pad[9];
};
static_assert(sizeof(WideInstance) == 128, "wide instance");

struct Scene
{
    // reference layout (static BLAS region followed, in links mode, by the per-frame TLAS region)
    const float2* ref_nodes;        // 3 float2 per bvh_node
    const uint2* ref_links;
    const uint32_t* indices;
    const float4* pos;
    const float4* normal;
    const float4* albedo;
    const float4* material;
    // per triangle of the index buffer (triangle k = indices[3k..3k+2] + its mesh's base vertex), the nine
    // attribute vectors shade_hit interpolates, gathered once at upload: n0 n1 n2 a0 a1 a2 m0 m1 m2
    const float4* shade_tris;
    const RefInstance* instances;   // [0,n_static) static, then this frame's dynamic instances
    const RefSubframe* subframes;
    // wide layout
    const WideNode* wnodes;
    const float4* wtris;            // 3 float4 per triangle slot: p0,p1,p2 ; p0.w = original primitive id (bits)
    const WideBlas* wblas;
    const WideInstance* winst;      // parallel to `instances`
    const WideNode* wtlas;          // static TLAS over static instances (leaf slot = instance index)
    const uint2* dyn_range;         // per subframe: dynamic instance set {prefix p, a | len << 20} (ptgpu_api.cu)
    const float4* dyn_union;        // per subframe: lo, hi of the world box around that set
    // compressed 8-wide layout (bvh_wide.cu: build_cw_*, pt_cwbvh.cuh)
    const float4* cwnodes;          // 5 float4 (80 B) per node, all BLASes then the static TLAS
    const float4* cwtris;           // 3 float4 per triangle, leaf order; p0.w = primitive id (bits)
    const uint32_t* cw_inst_index;  // TLAS leaf order -> instance index
    uint32_t cw_tlas_root;
    // flat static scene (bvh_wide.cu): root of the one BVH over the world-space triangles of all static
    // instances, in cwnodes / cwtris behind the BLASes; 0xFFFFFFFF = walk the static TLAS + BLASes instead
    uint32_t flat_root;
    uint32_t flat_top;              // nodes [flat_root, flat_root + flat_top) are its top levels, breadth-first
    uint32_t dyn_first;             // flat scene: a query enters the per-frame instances before the static world
    uint32_t cw_magic;              // 0x47000000 (2^15 as a float), see u8f_axis (pt_cwbvh.cuh)
    // ray-sort grid (pt_wave.cuh): cell = (p - key_lo) * key_scale, 32 x 8 x 32 cells over the static scene
    float key_lo[3], key_scale[3];
    uint32_t n_static;
    uint32_t n_subframes;
    // config
    int32_t width, height, max_bounces, samples_per_subframe;
    uint32_t student_id;
};

// Per-path view of one subframe (scene.hh:27-35), pulled into registers once.
struct SubframeCtx
{
    uint32_t index;
    uint32_t tlas_count, tlas_offset;  // reference TLAS handle (links mode)
    Light light;
};

// ---- event counters (instrumented links mode) ---------------------------------------------------
struct Counters
{
    unsigned long long v[16];
};

} // namespace pt
