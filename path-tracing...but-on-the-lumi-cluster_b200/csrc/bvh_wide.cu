// Host-only: flatten the reference BVHs into the GPU layout (see bvh_wide.hh).
//
// Input is what crosses the seam (main.cc:29-37): `bvh_node` boxes in BFS order
// (bvh.cc:145-168) and eight direction-specific link tables per BVH (bvh.cc:170-229). The tree
// topology is only in the links, so it is recovered from the octant-0 table, whose position
// (links[8*node_offset + 0*node_count]) does not depend on the still unknown node_count:
//   - an inner node's `accept` is its LAST stored child (octant 0 reverses every node,
//     bvh.cc:181), siblings are contiguous in BFS numbering, each child's `cancel` is the
//     previous sibling and the first stored child's `cancel` is the parent's `cancel`;
//   - a leaf has bit 31 of `accept` set, the rest is the triangle index (bvh.cc:176-177).
// The binary SAH tree with multi-way terminal nodes (bvh.cc:113-142) is then collapsed into
// 4-wide nodes whose leaves hold up to WIDE_LEAF_MAX pre-gathered triangles.
#include "bvh_wide.hh"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <functional>

namespace pt {
namespace {

struct Box
{
    float lo[3], hi[3];
    void reset() { for(int a = 0; a < 3; ++a) { lo[a] = FLT_MAX; hi[a] = -FLT_MAX; } }
    void grow(const Box& b) { for(int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    float area() const
    {
        float x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2];
        return x * y + y * z + z * x;
    }
};

// Generic source tree: children of node i are [first, first+count); count == 0 marks a leaf.
struct TNode
{
    Box box;
    uint32_t first = 0, count = 0;
    uint32_t payload = 0; // leaf: triangle index (BLAS) or instance index (TLAS)
};

Box box_of(const ptgpu_bvh_node& n)
{
    Box b;
    b.lo[0] = n.min_x; b.lo[1] = n.min_y; b.lo[2] = n.min_z;
    b.hi[0] = n.max_x; b.hi[1] = n.max_y; b.hi[2] = n.max_z;
    return b;
}

// Recover one reference BVH starting at node offset `off`. Returns its node count (0 on error).
uint32_t recover_tree(const ptgpu_bvh_node* nodes, const ptgpu_bvh_link* links, size_t n_nodes, size_t off,
                      std::vector<TNode>& tree, std::string& err)
{
    const ptgpu_bvh_link* T = links + 8 * off; // octant-0 table
    const size_t avail = n_nodes - off;
    tree.clear();
    tree.resize(1);
    std::vector<uint32_t> todo{0};
    uint32_t max_index = 0;
    while(!todo.empty())
    {
        uint32_t i = todo.back(); todo.pop_back();
        if(i >= avail) { err = "link walks past the node array"; return 0; }
        max_index = std::max(max_index, i);
        if(tree.size() <= i) tree.resize(i + 1);
        tree[i].box = box_of(nodes[off + i]);
        const ptgpu_bvh_link l = T[i];
        if(l.accept & 0x80000000u)
        {
            tree[i].count = 0;
            tree[i].payload = l.accept & 0x7FFFFFFFu;
            continue;
        }
        uint32_t last = l.accept, c = last;
        if(c <= i || c >= avail) { err = "inner node with a non-forward child"; return 0; }
        while(c > i + 1 && T[c].cancel == c - 1) --c;
        if(T[c].cancel != l.cancel) { err = "sibling chain does not end in the parent's cancel link"; return 0; }
        tree[i].first = c;
        tree[i].count = last - c + 1;
        for(uint32_t k = c; k <= last; ++k) todo.push_back(k);
    }
    if(tree.size() != (size_t)max_index + 1) { err = "BFS numbering is not dense"; return 0; }
    return max_index + 1;
}

// ---- collapse into 4-wide nodes -------------------------------------------------------------------

struct Item
{
    Box box;
    int32_t src = -1;                  // source inner node still to be expanded / recursed into, or -1
    std::vector<uint32_t> leaves;      // leaf payloads when src < 0
};

struct Collapser
{
    const std::vector<TNode>& tree;
    int leaf_max;
    std::vector<WideNode>& out_nodes;
    size_t node_base;                                  // index of this BVH's first node in out_nodes
    std::function<uint32_t(const std::vector<uint32_t>&)> emit_leaf; // returns the encoded child word
    uint32_t max_stack = 0;

    bool all_leaf_children(const TNode& n) const
    {
        for(uint32_t k = 0; k < n.count; ++k) if(tree[n.first + k].count != 0) return false;
        return true;
    }

    // The items a source node turns into when opened up.
    void open(uint32_t src, std::vector<Item>& items) const
    {
        const TNode& n = tree[src];
        if(all_leaf_children(n))
        {   // terminal multi-way node: leaves are sorted along the split axis (bvh.cc:120), so
            // consecutive chunks are spatially coherent
            const uint32_t chunks = (n.count + leaf_max - 1) / leaf_max;
            uint32_t k = 0;
            for(uint32_t c = 0; c < chunks; ++c)
            {
                uint32_t size = (n.count - k + (chunks - c) - 1) / (chunks - c);
                Item it; it.box.reset();
                for(uint32_t j = 0; j < size; ++j, ++k)
                {
                    it.leaves.push_back(tree[n.first + k].payload);
                    it.box.grow(tree[n.first + k].box);
                }
                items.push_back(std::move(it));
            }
            return;
        }
        for(uint32_t k = 0; k < n.count; ++k)
        {
            const TNode& c = tree[n.first + k];
            Item it; it.box = c.box;
            if(c.count == 0) it.leaves.push_back(c.payload);
            else it.src = (int32_t)(n.first + k);
            items.push_back(std::move(it));
        }
    }

    size_t opened_size(uint32_t src) const
    {
        const TNode& n = tree[src];
        if(all_leaf_children(n)) return (n.count + leaf_max - 1) / leaf_max;
        return n.count;
    }

    // Builds the wide node for source node `src`; returns (index relative to node_base, stack bound).
    std::pair<uint32_t, uint32_t> build(uint32_t src)
    {
        std::vector<Item> items;
        open(src, items);
        // a terminal node with more than 4 leaf chunks (> 16 leaves): nest
        if(items.size() > 4) return build_overflow(items);
        for(;;)
        {
            int best = -1; float best_area = -1.0f;
            for(size_t i = 0; i < items.size(); ++i)
            {
                if(items[i].src < 0) continue;
                if(items.size() - 1 + opened_size((uint32_t)items[i].src) > 4) continue;
                float a = items[i].box.area();
                if(a > best_area) { best_area = a; best = (int)i; }
            }
            if(best < 0) break;
            uint32_t s = (uint32_t)items[best].src;
            items.erase(items.begin() + best);
            open(s, items);
        }
        return emit(items);
    }

    std::pair<uint32_t, uint32_t> emit(std::vector<Item>& items)
    {
        const uint32_t self = (uint32_t)(out_nodes.size() - node_base);
        out_nodes.emplace_back();
        uint32_t child[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
        Box boxes[4];
        uint32_t deepest = 0;
        const uint32_t k = (uint32_t)items.size();
        for(uint32_t i = 0; i < k; ++i)
        {
            boxes[i] = items[i].box;
            if(items[i].src >= 0)
            {
                auto r = build((uint32_t)items[i].src);
                child[i] = r.first;
                deepest = std::max(deepest, r.second);
            }
            else child[i] = emit_leaf(items[i].leaves);
        }
        WideNode& n = out_nodes[node_base + self];
        float* lox = &n.lox.x; float* loy = &n.loy.x; float* loz = &n.loz.x;
        float* hix = &n.hix.x; float* hiy = &n.hiy.x; float* hiz = &n.hiz.x;
        uint32_t* ch = &n.child.x;
        for(uint32_t i = 0; i < 4; ++i)
        {
            if(i < k)
            {
                lox[i] = boxes[i].lo[0]; loy[i] = boxes[i].lo[1]; loz[i] = boxes[i].lo[2];
                hix[i] = boxes[i].hi[0]; hiy[i] = boxes[i].hi[1]; hiz[i] = boxes[i].hi[2];
            }
            else
            {   // empty slot: an inverted box no ray can hit (NaN-free)
                lox[i] = loy[i] = loz[i] = FLT_MAX;
                hix[i] = hiy[i] = hiz[i] = -FLT_MAX;
            }
            ch[i] = child[i];
        }
        n.pad = make_uint4(0, 0, 0, 0);
        // the traversal keeps the nearest child in a register and pushes the other k-1
        const uint32_t bound = std::max(k > 0 ? k - 1 : 0u, (k > 0 ? k - 1 : 0u) + deepest);
        return {self, bound};
    }

    // More than four items at one level: nest them pairwise by proximity until four remain.
    std::pair<uint32_t, uint32_t> build_overflow(std::vector<Item>& items)
    {
        // Split the list (already spatially ordered) into four consecutive groups; groups with more
        // than one item become nested nodes built from synthetic item lists.
        const size_t n = items.size();
        std::vector<std::vector<Item>> groups(4);
        for(size_t i = 0; i < n; ++i) groups[i * 4 / n].push_back(std::move(items[i]));
        const uint32_t self = (uint32_t)(out_nodes.size() - node_base);
        out_nodes.emplace_back();
        uint32_t child[4]; Box boxes[4]; uint32_t deepest = 0;
        for(int g = 0; g < 4; ++g)
        {
            boxes[g].reset();
            for(auto& it : groups[g]) boxes[g].grow(it.box);
            if(groups[g].size() == 1 && groups[g][0].src < 0) child[g] = emit_leaf(groups[g][0].leaves);
            else if(groups[g].size() == 1)
            {
                auto r = build((uint32_t)groups[g][0].src);
                child[g] = r.first; deepest = std::max(deepest, r.second);
            }
            else
            {
                std::pair<uint32_t, uint32_t> r = groups[g].size() > 4 ? build_overflow(groups[g]) : emit(groups[g]);
                child[g] = r.first; deepest = std::max(deepest, r.second);
            }
        }
        WideNode& nd = out_nodes[node_base + self];
        float* lox = &nd.lox.x; float* loy = &nd.loy.x; float* loz = &nd.loz.x;
        float* hix = &nd.hix.x; float* hiy = &nd.hiy.x; float* hiz = &nd.hiz.x;
        uint32_t* ch = &nd.child.x;
        for(int i = 0; i < 4; ++i)
        {
            lox[i] = boxes[i].lo[0]; loy[i] = boxes[i].lo[1]; loz[i] = boxes[i].lo[2];
            hix[i] = boxes[i].hi[0]; hiy[i] = boxes[i].hi[1]; hiz[i] = boxes[i].hi[2];
            ch[i] = child[i];
        }
        nd.pad = make_uint4(0, 0, 0, 0);
        return {self, 3u + deepest};
    }
};

// Wide BVH of a whole source tree. A single-leaf tree becomes one node with one child.
uint32_t collapse_tree(const std::vector<TNode>& tree, int leaf_max, std::vector<WideNode>& out,
                       const std::function<uint32_t(const std::vector<uint32_t>&)>& emit_leaf)
{
    Collapser c{tree, leaf_max, out, out.size(), emit_leaf};
    if(tree[0].count == 0)
    {
        std::vector<Item> items(1);
        items[0].box = tree[0].box;
        items[0].leaves.push_back(tree[0].payload);
        return c.emit(items).second + 1;
    }
    return c.build(0).second + 1;
}

// ---- small full-sweep SAH builder for the static TLAS (n ~ 10^3) ----------------------------------

struct Prim { Box box; uint32_t id; };

void build_sah(std::vector<Prim>& prims, size_t begin, size_t end, std::vector<TNode>& tree, uint32_t self)
{
    // children are appended contiguously so that [first, first+count) addressing works
    const size_t n = end - begin;
    Box b; b.reset();
    for(size_t i = begin; i < end; ++i) b.grow(prims[i].box);
    tree[self].box = b;
    if(n == 1)
    {
        tree[self].count = 0;
        tree[self].payload = prims[begin].id;
        return;
    }
    float best_cost = FLT_MAX; int best_axis = 0; size_t best_split = begin + n / 2;
    std::vector<float> right_area(n);
    for(int axis = 0; axis < 3; ++axis)
    {
        std::sort(prims.begin() + begin, prims.begin() + end, [axis](const Prim& x, const Prim& y) {
            float cx = x.box.lo[axis] + x.box.hi[axis], cy = y.box.lo[axis] + y.box.hi[axis];
            return cx < cy || (cx == cy && x.id < y.id);
        });
        Box acc; acc.reset();
        for(size_t i = n; i-- > 1;) { acc.grow(prims[begin + i].box); right_area[i] = acc.area(); }
        acc.reset();
        for(size_t i = 1; i < n; ++i)
        {
            acc.grow(prims[begin + i - 1].box);
            float cost = acc.area() * (float)i + right_area[i] * (float)(n - i);
            if(cost < best_cost) { best_cost = cost; best_axis = axis; best_split = begin + i; }
        }
    }
    std::sort(prims.begin() + begin, prims.begin() + end, [best_axis](const Prim& x, const Prim& y) {
        float cx = x.box.lo[best_axis] + x.box.hi[best_axis], cy = y.box.lo[best_axis] + y.box.hi[best_axis];
        return cx < cy || (cx == cy && x.id < y.id);
    });
    const uint32_t first = (uint32_t)tree.size();
    tree.emplace_back(); tree.emplace_back();
    tree[self].first = first; tree[self].count = 2;
    build_sah(prims, begin, best_split, tree, first);
    build_sah(prims, best_split, end, tree, first + 1);
}

// World-space bounds of an instance exactly as build_tlas derives them (bvh.cc:262-278):
// the BLAS root box's eight corners through `transform`.
void instance_world_box(const ptgpu_tlas_instance& inst, const float lo[3], const float hi[3], float out_lo[3], float out_hi[3])
{
    const ptgpu_float4* c = inst.transform.r;
    for(int a = 0; a < 8; ++a)
    {
        // bvh.cc:271: bounds[a&1].x, bounds[a&2?0:1].y, bounds[a&4?0:1].z
        float x = (a & 1) ? hi[0] : lo[0];
        float y = (a & 2) ? lo[1] : hi[1];
        float z = (a & 4) ? lo[2] : hi[2];
        float v[3] = {
            c[0].x * x + c[1].x * y + c[2].x * z + c[3].x,
            c[0].y * x + c[1].y * y + c[2].y * z + c[3].y,
            c[0].z * x + c[1].z * y + c[2].z * z + c[3].z};
        for(int k = 0; k < 3; ++k)
        {
            out_lo[k] = a == 0 ? v[k] : std::min(out_lo[k], v[k]);
            out_hi[k] = a == 0 ? v[k] : std::max(out_hi[k], v[k]);
        }
    }
}

} // namespace

bool make_wide_instance(const WideScene& ws, const ptgpu_tlas_instance& inst, uint32_t ref_index, WideInstance& out)
{
    int found = -1;
    for(size_t b = 0; b < ws.blas_info.size(); ++b)
        if(ws.blas_info[b].ref_node_offset == inst.blas.node_offset) { found = (int)b; break; }
    if(found < 0) return false;
    const WideBlasInfo& bi = ws.blas_info[found];
    if(bi.ref_node_count != inst.blas.node_count || bi.mesh.index_offset != inst.m.index_offset ||
       bi.mesh.base_vertex_offset != inst.m.base_vertex_offset || bi.mesh.triangle_count != inst.m.triangle_count)
        return false;
    const ptgpu_float4* r = inst.inv_transform.r; // columns
    memset(&out, 0, sizeof(out));
    out.inv0 = make_float4(r[0].x, r[1].x, r[2].x, r[3].x);
    out.inv1 = make_float4(r[0].y, r[1].y, r[2].y, r[3].y);
    out.inv2 = make_float4(r[0].z, r[1].z, r[2].z, r[3].z);
    const WideBlas& wb = ws.blas[found];
    float lo[3] = {wb.lo.x, wb.lo.y, wb.lo.z}, hi[3] = {wb.hi.x, wb.hi.y, wb.hi.z}, wlo[3], whi[3];
    instance_world_box(inst, lo, hi, wlo, whi);
    out.lo = make_float4(wlo[0], wlo[1], wlo[2], 0.0f);
    out.hi = make_float4(whi[0], whi[1], whi[2], 0.0f);
    out.blas = (uint32_t)found;
    out.ref_instance = ref_index;
    return true;
}

bool build_wide_scene(
    const ptgpu_bvh_node* nodes, size_t n_nodes, const ptgpu_bvh_link* links,
    const uint32_t* indices, size_t n_indices, const ptgpu_float3* pos, size_t n_verts,
    const ptgpu_tlas_instance* instances, size_t n_static,
    WideScene& out, std::string& err)
{
    out = WideScene();
    // 1. every BVH in the static region, in build order (= mesh load order, scene.cc:42-47)
    size_t off = 0, index_cursor = 0, vertex_cursor = 0;
    std::vector<TNode> tree;
    while(off < n_nodes)
    {
        uint32_t count = recover_tree(nodes, links, n_nodes, off, tree, err);
        if(count == 0) { err = "BVH at node " + std::to_string(off) + ": " + err; return false; }
        uint32_t tri_count = 0;
        for(const TNode& t : tree) if(t.count == 0) tri_count++;
        // the i-th BLAS belongs to the i-th mesh: indices and vertices are appended in load order
        // (mesh.cc:118-119, 255-259), and every vertex is referenced by at least one index
        ptgpu_mesh m;
        m.triangle_count = tri_count;
        m.index_offset = (uint32_t)index_cursor;
        m.base_vertex_offset = (uint32_t)vertex_cursor;
        if(index_cursor + 3 * (size_t)tri_count > n_indices) { err = "index buffer shorter than the BLAS leaves imply"; return false; }
        uint32_t vmax = 0;
        for(size_t i = 0; i < 3 * (size_t)tri_count; ++i) vmax = std::max(vmax, indices[index_cursor + i]);
        m.vertex_count = vmax + 1;
        if(vertex_cursor + m.vertex_count > n_verts) { err = "vertex buffer shorter than the indices imply"; return false; }

        WideBlas wb{};
        wb.node_offset = (uint32_t)out.nodes.size();
        wb.tri_offset = (uint32_t)(out.tris.size() / 3);
        const uint32_t tri_base = wb.tri_offset;
        auto emit_leaf = [&](const std::vector<uint32_t>& prims) -> uint32_t {
            const uint32_t first = (uint32_t)(out.tris.size() / 3) - tri_base;
            for(uint32_t p : prims)
            {
                const uint32_t* ix = indices + m.index_offset + 3 * (size_t)p;
                for(int k = 0; k < 3; ++k)
                {
                    const ptgpu_float3& v = pos[m.base_vertex_offset + ix[k]];
                    float w = 0.0f;
                    if(k == 0) memcpy(&w, &p, 4);
                    out.tris.push_back(make_float4(v.x, v.y, v.z, w));
                }
            }
            return 0x80000000u | ((uint32_t)(prims.size() - 1) << 27) | first;
        };
        for(const TNode& t : tree)
            if(t.count == 0 && t.payload >= tri_count) { err = "leaf payload beyond the triangle count"; return false; }
        uint32_t stack = collapse_tree(tree, WIDE_LEAF_MAX, out.nodes, emit_leaf);
        wb.node_count = (uint32_t)out.nodes.size() - wb.node_offset;
        wb.tri_count = (uint32_t)(out.tris.size() / 3) - wb.tri_offset;
        if(wb.tri_count != tri_count) { err = "triangle count mismatch after collapse"; return false; }
        if(wb.tri_count >= (1u << 27)) { err = "BLAS too large for the leaf encoding"; return false; }
        wb.lo = make_float4(tree[0].box.lo[0], tree[0].box.lo[1], tree[0].box.lo[2], 0.0f);
        wb.hi = make_float4(tree[0].box.hi[0], tree[0].box.hi[1], tree[0].box.hi[2], 0.0f);
        out.blas.push_back(wb);
        out.blas_info.push_back(WideBlasInfo{(uint32_t)off, count, m, stack});
        out.max_stack = std::max(out.max_stack, stack);
        off += count;
        index_cursor += 3 * (size_t)tri_count;
        vertex_cursor += m.vertex_count;
    }
    if(index_cursor != n_indices || vertex_cursor != n_verts)
    {
        err = "mesh buffers do not line up with the BLAS sequence (indices " + std::to_string(index_cursor) + "/" +
            std::to_string(n_indices) + ", vertices " + std::to_string(vertex_cursor) + "/" + std::to_string(n_verts) + ")";
        return false;
    }

    // 2. static instances
    out.instances.resize(n_static);
    for(size_t i = 0; i < n_static; ++i)
        if(!make_wide_instance(out, instances[i], (uint32_t)i, out.instances[i]))
        {
            err = "static instance " + std::to_string(i) + " does not match any recovered BLAS/mesh";
            return false;
        }

    // 3. static TLAS, built once: a subframe's TLAS in the reference differs only by a handful of
    //    dynamic instances (scene.cc:666-674), which the kernels test separately
    std::vector<Prim> prims(n_static);
    for(size_t i = 0; i < n_static; ++i)
    {
        const WideInstance& wi = out.instances[i];
        prims[i].box.lo[0] = wi.lo.x; prims[i].box.lo[1] = wi.lo.y; prims[i].box.lo[2] = wi.lo.z;
        prims[i].box.hi[0] = wi.hi.x; prims[i].box.hi[1] = wi.hi.y; prims[i].box.hi[2] = wi.hi.z;
        prims[i].id = (uint32_t)i;
    }
    std::vector<TNode> ttree(1);
    ttree.reserve(2 * n_static);
    build_sah(prims, 0, n_static, ttree, 0);
    auto emit_inst = [](const std::vector<uint32_t>& ids) -> uint32_t { return 0x80000000u | ids[0]; };
    uint32_t tstack = collapse_tree(ttree, 1, out.tlas, emit_inst);
    out.max_stack += tstack + 1 /* exit marker */ + 16 /* dynamic instances */;
    if(out.max_stack > WIDE_STACK)
    {
        err = "traversal stack bound " + std::to_string(out.max_stack) + " exceeds WIDE_STACK " + std::to_string(WIDE_STACK);
        return false;
    }
    return true;
}

// Structural check of a flattened scene, used by the CPU test-suite: every triangle of every BLAS is
// reachable exactly once, child boxes enclose what is below them, every static instance is a TLAS
// leaf exactly once. Returns the number of violations.
uint64_t verify_wide_scene(const WideScene& ws, size_t n_static, std::string& err)
{
    uint64_t bad = 0;
    auto fail = [&](const std::string& m) { if(bad++ == 0) err = m; };
    for(size_t b = 0; b < ws.blas.size(); ++b)
    {
        const WideBlas& wb = ws.blas[b];
        std::vector<uint8_t> seen(wb.tri_count, 0);
        std::vector<uint32_t> prim_seen(wb.tri_count, 0);
        struct Todo { uint32_t node; Box box; bool has_box; };
        std::vector<Todo> todo{{0u, Box(), false}};
        while(!todo.empty())
        {
            Todo t = todo.back(); todo.pop_back();
            if(t.node >= wb.node_count) { fail("child index outside the BLAS"); continue; }
            const WideNode& n = ws.nodes[wb.node_offset + t.node];
            const float* lox = &n.lox.x; const float* loy = &n.loy.x; const float* loz = &n.loz.x;
            const float* hix = &n.hix.x; const float* hiy = &n.hiy.x; const float* hiz = &n.hiz.x;
            const uint32_t* ch = &n.child.x;
            for(int i = 0; i < 4; ++i)
            {
                if(ch[i] == 0xFFFFFFFFu) continue;
                Box cb;
                cb.lo[0] = lox[i]; cb.lo[1] = loy[i]; cb.lo[2] = loz[i];
                cb.hi[0] = hix[i]; cb.hi[1] = hiy[i]; cb.hi[2] = hiz[i];
                if(t.has_box)
                    for(int a = 0; a < 3; ++a)
                        if(cb.lo[a] < t.box.lo[a] || cb.hi[a] > t.box.hi[a]) fail("child box not inside its parent's slot box");
                if(!(ch[i] & 0x80000000u)) { todo.push_back({ch[i], cb, true}); continue; }
                const uint32_t first = ch[i] & 0x07FFFFFFu, count = ((ch[i] >> 27) & 0xFu) + 1u;
                if(count > (uint32_t)WIDE_LEAF_MAX || first + count > wb.tri_count) { fail("leaf range outside the BLAS"); continue; }
                for(uint32_t k = 0; k < count; ++k)
                {
                    if(seen[first + k]++) fail("triangle slot reachable twice");
                    const float4* v = &ws.tris[3 * (size_t)(wb.tri_offset + first + k)];
                    uint32_t prim; memcpy(&prim, &v[0].w, 4);
                    if(prim >= wb.tri_count) fail("primitive id out of range"); else prim_seen[prim]++;
                    for(int c = 0; c < 3; ++c)
                    {
                        const float p[3] = {v[c].x, v[c].y, v[c].z};
                        for(int a = 0; a < 3; ++a) if(p[a] < cb.lo[a] || p[a] > cb.hi[a]) fail("triangle vertex outside its leaf box");
                    }
                }
            }
        }
        for(uint32_t k = 0; k < wb.tri_count; ++k)
            if(seen[k] != 1 || prim_seen[k] != 1) { fail("BLAS " + std::to_string(b) + ": triangle missing or duplicated"); break; }
    }
    // TLAS
    std::vector<uint32_t> inst_seen(n_static, 0);
    std::vector<uint32_t> todo{0u};
    while(!todo.empty())
    {
        uint32_t ni = todo.back(); todo.pop_back();
        if(ni >= ws.tlas.size()) { fail("TLAS child outside the array"); continue; }
        const uint32_t* ch = &ws.tlas[ni].child.x;
        for(int i = 0; i < 4; ++i)
        {
            if(ch[i] == 0xFFFFFFFFu) continue;
            if(!(ch[i] & 0x80000000u)) { todo.push_back(ch[i]); continue; }
            uint32_t id = ch[i] & 0x7FFFFFFFu;
            if(id >= n_static) fail("TLAS leaf beyond the static instances"); else inst_seen[id]++;
        }
    }
    for(size_t i = 0; i < n_static; ++i) if(inst_seen[i] != 1) { fail("static instance missing from the TLAS or duplicated"); break; }
    return bad;
}

} // namespace pt
