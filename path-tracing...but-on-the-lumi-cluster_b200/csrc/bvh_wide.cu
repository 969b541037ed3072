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
// The binary SAH tree with multi-way terminal nodes (bvh.cc:113-142) is then collapsed twice: into
// compressed 8-wide nodes with one triangle per leaf child (the default traversal; cost-optimal
// collapse, OptimalCollapser) and into 4-wide float nodes whose leaves hold up to WIDE_LEAF_MAX
// pre-gathered triangles (tile kernel / megakernel). Without reference BVH arrays
// (ptgpu_upload_meshes) the binary tree of each mesh is built here instead (build_mesh_tree).
// One static TLAS over tight world boxes of the static instances completes the scene.
#include "bvh_wide.hh"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <chrono>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>

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

// ---- SAH builder: the static TLAS (n ~ 10^3, full sweep) and, for ptgpu_upload_meshes, the BLASes --------

struct Prim { Box box; uint32_t id; };

// Split of a large range by 32 centroid bins per axis (O(n) per level instead of three sorts): partitions
// prims[begin, end) in place and returns the split position (never begin or end).
size_t binned_split(std::vector<Prim>& prims, size_t begin, size_t end)
{
    const size_t n = end - begin;
    size_t best_split = begin + n / 2;
    constexpr int BINS = 32;
    float clo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, chi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for(size_t i = begin; i < end; ++i)
        for(int a = 0; a < 3; ++a)
        {
            const float c = prims[i].box.lo[a] + prims[i].box.hi[a];
            clo[a] = std::min(clo[a], c); chi[a] = std::max(chi[a], c);
        }
    float best_cost = FLT_MAX; int best_axis = -1, best_bin = 0;
    for(int a = 0; a < 3; ++a)
    {
        const float ext = chi[a] - clo[a];
        if(!(ext > 0.0f)) continue;
        const float scale = (float)BINS / ext;
        Box bb[BINS]; uint32_t cnt[BINS] = {};
        for(int k = 0; k < BINS; ++k) bb[k].reset();
        for(size_t i = begin; i < end; ++i)
        {
            const int k = std::min(BINS - 1, (int)((prims[i].box.lo[a] + prims[i].box.hi[a] - clo[a]) * scale));
            bb[k].grow(prims[i].box); cnt[k]++;
        }
        float right_area[BINS]; uint32_t right_cnt[BINS];
        Box acc; acc.reset(); uint32_t c = 0;
        for(int k = BINS - 1; k > 0; --k) { if(cnt[k]) acc.grow(bb[k]); c += cnt[k]; right_area[k] = c ? acc.area() : 0.0f; right_cnt[k] = c; }
        acc.reset(); c = 0;
        for(int k = 1; k < BINS; ++k)
        {
            if(cnt[k - 1]) acc.grow(bb[k - 1]);
            c += cnt[k - 1];
            if(c == 0 || right_cnt[k] == 0) continue;
            const float cost = acc.area() * (float)c + right_area[k] * (float)right_cnt[k];
            if(cost < best_cost) { best_cost = cost; best_axis = a; best_bin = k; }
        }
    }
    if(best_axis >= 0)
    {
        const int a = best_axis;
        const float scale = (float)BINS / (chi[a] - clo[a]), lo = clo[a];
        auto mid = std::partition(prims.begin() + begin, prims.begin() + end, [&](const Prim& x) {
            return std::min(BINS - 1, (int)((x.box.lo[a] + x.box.hi[a] - lo) * scale)) < best_bin;
        });
        best_split = (size_t)(mid - prims.begin());
    }
    if(best_axis < 0 || best_split == begin || best_split == end)
    {   // all centroids coincide: halve the range in primitive order
        std::sort(prims.begin() + begin, prims.begin() + end, [](const Prim& x, const Prim& y) { return x.id < y.id; });
        best_split = begin + n / 2;
    }
    return best_split;
}

// Binary SAH tree over `prims`, single-primitive leaves. Ranges of at most `sweep_below` primitives are
// split by the full sweep (every split position on every axis, as bvh.cc:43-140 does); larger ranges by
// 32 centroid bins per axis, which costs O(n) per level instead of three sorts.
void build_sah(std::vector<Prim>& prims, size_t begin, size_t end, std::vector<TNode>& tree, uint32_t self,
               size_t sweep_below = SIZE_MAX)
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
    size_t best_split = begin + n / 2;
    if(n <= sweep_below)
    {
        float best_cost = FLT_MAX; int best_axis = 0;
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
    }
    else best_split = binned_split(prims, begin, end);
    const uint32_t first = (uint32_t)tree.size();
    tree.emplace_back(); tree.emplace_back();
    tree[self].first = first; tree[self].count = 2;
    build_sah(prims, begin, best_split, tree, first, sweep_below);
    build_sah(prims, best_split, end, tree, first + 1, sweep_below);
}

// the BLAS of one mesh straight from its triangles (no reference bvh arrays).
void build_mesh_tree(const uint32_t* indices, const ptgpu_float3* pos, const ptgpu_mesh& m, std::vector<TNode>& tree)
{
    std::vector<Prim> prims(m.triangle_count);
    for(uint32_t t = 0; t < m.triangle_count; ++t)
    {
        Box b; b.reset();
        for(int k = 0; k < 3; ++k)
        {
            const ptgpu_float3& v = pos[m.base_vertex_offset + indices[m.index_offset + 3 * (size_t)t + k]];
            const float c[3] = {v.x, v.y, v.z};
            for(int a = 0; a < 3; ++a) { b.lo[a] = std::min(b.lo[a], c[a]); b.hi[a] = std::max(b.hi[a], c[a]); }
        }
        prims[t].box = b; prims[t].id = t;
    }
    tree.clear();
    tree.resize(1);
    tree.reserve(2 * (size_t)m.triangle_count);
    build_sah(prims, 0, prims.size(), tree, 0, 256);
}

// ---- compressed 8-wide BVH (after Ylitie, Karras, Laine 2017) --------------------------------------
//
// Node = 80 bytes = 5 x 16-byte loads:
//   n0  p.x p.y p.z | e.x e.y e.z imask      origin of the quantisation grid, per-axis exponent, inner mask
//   n1  child_base | tri_base | meta[0..3] | meta[4..7]
//   n2  qlo.x[0..7] | qlo.y[0..7]             child boxes, 8 bits per bound: lo = p + q * 2^e
//   n3  qlo.z[0..7] | qhi.x[0..7]
//   n4  qhi.y[0..7] | qhi.z[0..7]
// meta[slot]: inner child 0b001_11sss (24 + slot); leaf 0b{unary count}_{offset from tri_base}; 0 = empty —
// an empty slot holds an inverted box (never hit) and a copy of a real child's meta (pt_scene.cuh).
// Children sit in the slot whose octant direction best matches their offset from the node centre, so
// a ray visits slots in the order (slot XOR ray octant), high to low, without sorting distances.

struct W8Child
{
    Box box;
    int inner = -1;                 // index into the W8 node list, or -1
    std::vector<uint32_t> leaves;   // payloads when inner < 0
};
struct W8Node { Box box; std::vector<W8Child> ch; int stub = -1; };   // stub >= 0: placeholder for the root of separately emitted subtree `stub`

struct GItem { Box box; int32_t src = -1; std::vector<uint32_t> leaves; int nested = -1; };

struct CollapserW
{
    const std::vector<TNode>& tree;
    int width, leaf_max;
    std::vector<W8Node>& out;

    bool all_leaf_children(const TNode& n) const
    {
        for(uint32_t k = 0; k < n.count; ++k) if(tree[n.first + k].count != 0) return false;
        return true;
    }
    size_t opened_size(uint32_t src) const
    {
        const TNode& n = tree[src];
        return all_leaf_children(n) ? (n.count + leaf_max - 1) / leaf_max : n.count;
    }
    void open(uint32_t src, std::vector<GItem>& items) const
    {
        const TNode& n = tree[src];
        if(all_leaf_children(n))
        {
            const uint32_t chunks = (n.count + leaf_max - 1) / leaf_max;
            uint32_t k = 0;
            for(uint32_t c = 0; c < chunks; ++c)
            {
                uint32_t size = (n.count - k + (chunks - c) - 1) / (chunks - c);
                GItem it; it.box.reset();
                for(uint32_t j = 0; j < size; ++j, ++k) { it.leaves.push_back(tree[n.first + k].payload); it.box.grow(tree[n.first + k].box); }
                items.push_back(std::move(it));
            }
            return;
        }
        for(uint32_t k = 0; k < n.count; ++k)
        {
            const TNode& c = tree[n.first + k];
            GItem it; it.box = c.box;
            if(c.count == 0) it.leaves.push_back(c.payload); else it.src = (int32_t)(n.first + k);
            items.push_back(std::move(it));
        }
    }
    // node from an explicit item list (more items than `width` are nested in consecutive groups)
    int from_items(std::vector<GItem>& items, const Box& box)
    {
        if((int)items.size() > width)
        {
            std::vector<GItem> grouped;
            const size_t n = items.size();
            size_t i = 0;
            for(int g = 0; g < width; ++g)
            {
                size_t end = n * (g + 1) / width;
                if(end - i == 1) grouped.push_back(std::move(items[i]));
                else if(end > i)
                {
                    std::vector<GItem> sub(std::make_move_iterator(items.begin() + i), std::make_move_iterator(items.begin() + end));
                    Box b; b.reset(); for(auto& it : sub) b.grow(it.box);
                    GItem it; it.box = b; it.nested = from_items(sub, b);
                    grouped.push_back(std::move(it));
                }
                i = end;
            }
            items.swap(grouped);
        }
        const int self = (int)out.size();
        out.emplace_back();
        out[self].box = box;
        std::vector<W8Child> ch;
        for(auto& it : items)
        {
            W8Child c; c.box = it.box;
            if(it.nested >= 0) c.inner = it.nested;
            else if(it.src >= 0) c.inner = build((uint32_t)it.src);
            else c.leaves = std::move(it.leaves);
            ch.push_back(std::move(c));
        }
        out[self].ch = std::move(ch);
        return self;
    }
    int build(uint32_t src)
    {
        std::vector<GItem> items;
        open(src, items);
        if((int)items.size() <= width)
            for(;;)
            {   // open the largest child while the node still has room
                int best = -1; float best_area = -1.0f;
                for(size_t i = 0; i < items.size(); ++i)
                {
                    if(items[i].src < 0) continue;
                    if(items.size() - 1 + opened_size((uint32_t)items[i].src) > (size_t)width) continue;
                    float a = items[i].box.area();
                    if(a > best_area) { best_area = a; best = (int)i; }
                }
                if(best < 0) break;
                uint32_t s = (uint32_t)items[best].src;
                items.erase(items.begin() + best);
                open(s, items);
            }
        return from_items(items, tree[src].box);
    }
};

// Cost-optimal collapse for single-primitive leaves (the dynamic program of Ylitie, Karras and Laine,
// "Efficient incoherent ray traversal on GPUs through compressed wide BVHs", section 4.1). Every wide
// node costs its surface area (one eight-child box test per visit, whatever its fill) and the leaves
// are a given, so the program minimises the summed area of the wide nodes:
//     C(n, 1)      = A(n) + D(n, 8)                          n becomes a wide node
//     C(n, i > 1)  = min(D(n, i), C(n, i - 1))               n's subtree as a forest of at most i roots
//     D(n, j)      = min over 0 < k < j of C(left, k) + C(right, j - k)
// The greedy "open the largest child while there is room" rule it replaces filled 4.85 of 8 slots.
struct OptimalCollapser
{
    static constexpr int W = CW_WIDTH;
    struct BNode { Box box; int32_t left = -1, right = -1; uint32_t payload = 0; };
    std::vector<BNode> b;           // strictly binary copy of the source tree
    std::vector<float> cost;        // (W - 1) per node: C(n, 1..7)
    std::vector<uint8_t> split;     // W per node: best k of D(n, 2..8) at [j - 1]
    std::vector<W8Node>& out;

    explicit OptimalCollapser(std::vector<W8Node>& o) : out(o) {}

    // children [lo, hi) of a source node with more than two children become a balanced binary subtree
    int32_t binarise_range(const std::vector<TNode>& tree, const std::vector<int32_t>& kids, size_t lo, size_t hi)
    {
        if(hi - lo == 1) return kids[lo];
        const int32_t self = (int32_t)b.size();
        b.emplace_back();
        const size_t mid = (lo + hi) / 2;
        const int32_t l = binarise_range(tree, kids, lo, mid), r = binarise_range(tree, kids, mid, hi);
        b[self].left = l; b[self].right = r;
        b[self].box = b[l].box; b[self].box.grow(b[r].box);
        return self;
    }
    int32_t import(const std::vector<TNode>& tree, uint32_t src)
    {
        const TNode& t = tree[src];
        const int32_t self = (int32_t)b.size();
        b.emplace_back();
        b[self].box = t.box;
        if(t.count == 0) { b[self].payload = t.payload; return self; }
        std::vector<int32_t> kids;
        for(uint32_t k = 0; k < t.count; ++k) kids.push_back(import(tree, t.first + k));
        if(kids.size() == 1)
        {   // a chain link: keep the child, under the parent's box
            const Box box = b[self].box;
            b[self] = b[kids[0]];
            b[self].box = box;
            return self;
        }
        const size_t mid = kids.size() / 2;
        const int32_t l = binarise_range(tree, kids, 0, mid), r = binarise_range(tree, kids, mid, kids.size());
        b[self].left = l; b[self].right = r;
        return self;
    }
    bool leaf(int32_t n) const { return b[n].left < 0; }
    float C(int32_t n, int i) const { return leaf(n) ? 0.0f : cost[(size_t)n * (W - 1) + (i - 1)]; }

    void solve(int32_t n)
    {   // post-order without recursion on the call stack of deep trees
        std::vector<std::pair<int32_t, bool>> todo{{n, false}};
        while(!todo.empty())
        {
            auto [v, done] = todo.back(); todo.pop_back();
            if(leaf(v)) continue;
            if(!done)
            {
                todo.push_back({v, true});
                todo.push_back({b[v].left, false});
                todo.push_back({b[v].right, false});
                continue;
            }
            const int32_t l = b[v].left, r = b[v].right;
            float D[W + 1];
            for(int j = 2; j <= W; ++j)
            {
                float best = FLT_MAX; int best_k = 1;
                for(int k = 1; k < j; ++k)
                {
                    const float c = C(l, std::min(k, W - 1)) + C(r, std::min(j - k, W - 1));
                    if(c < best) { best = c; best_k = k; }
                }
                D[j] = best;
                split[(size_t)v * W + (j - 1)] = (uint8_t)best_k;
            }
            float* c = &cost[(size_t)v * (W - 1)];
            c[0] = b[v].box.area() + D[W];
            for(int i = 2; i <= W - 1; ++i) c[i - 1] = std::min(D[i], c[i - 2]);
        }
    }
    // the roots of n's subtree as a forest of at most j trees
    void collect(int32_t n, int j, std::vector<int32_t>& roots) const
    {
        if(leaf(n) || j == 1) { roots.push_back(n); return; }
        if(j <= W - 1 && C(n, j) == C(n, j - 1) ) { collect(n, j - 1, roots); return; }
        const int k = split[(size_t)n * W + (j - 1)];
        collect(b[n].left, std::min(k, W - 1), roots);
        collect(b[n].right, std::min(j - k, W - 1), roots);
    }
    int emit(int32_t n)
    {
        const int self = (int)out.size();
        out.emplace_back();
        out[self].box = b[n].box;
        std::vector<int32_t> roots;
        if(leaf(n)) roots.push_back(n);
        else
        {   // n is a wide node: its eight slots go to the two subtrees as D(n, 8) decided
            const int k = split[(size_t)n * W + (W - 1)];
            collect(b[n].left, std::min(k, W - 1), roots);
            collect(b[n].right, std::min(W - k, W - 1), roots);
        }
        std::vector<W8Child> ch;
        for(int32_t r : roots)
        {
            W8Child c; c.box = b[r].box;
            if(leaf(r)) c.leaves.push_back(b[r].payload); else c.inner = emit(r);
            ch.push_back(std::move(c));
        }
        out[self].ch = std::move(ch);
        return self;
    }
    int build(const std::vector<TNode>& tree)
    {
        b.reserve(2 * tree.size());
        const int32_t root = import(tree, 0);
        cost.assign(b.size() * (W - 1), 0.0f);
        split.assign(b.size() * W, 1);
        solve(root);
        return emit(root);
    }
};

inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

struct CwEmitter
{
    const std::vector<W8Node>& w8;
    std::vector<float4>& nodes;          // 5 per node
    std::function<uint32_t(const std::vector<uint32_t>&)> emit_leaf; // appends leaf payloads, returns first index
    uint32_t max_depth = 0;
    std::string err;
    std::vector<std::pair<uint32_t, int>> stubs;   // (node index, subtree) of every placeholder met

    // Fills node `cw` (already allocated) from w8[wi]; returns the depth below.
    uint32_t emit(int wi, uint32_t cw)
    {
        const W8Node& n = w8[wi];
        if(n.stub >= 0) { stubs.push_back({cw, n.stub}); return 0; }
        const int k = (int)n.ch.size();
        if(k > CW_WIDTH) { err = "node wider than 8"; return 0; }
        // greedy slot assignment: child with the best alignment to a free slot's octant direction first
        float cx[3]; for(int a = 0; a < 3; ++a) cx[a] = 0.5f * (n.box.lo[a] + n.box.hi[a]);
        int slot_of[CW_WIDTH]; bool slot_used[CW_WIDTH] = {false}; bool child_done[CW_WIDTH] = {false};
        for(int it = 0; it < k; ++it)
        {
            float best = -FLT_MAX; int bc = -1, bs = -1;
            for(int c = 0; c < k; ++c)
            {
                if(child_done[c]) continue;
                float dx[3]; for(int a = 0; a < 3; ++a) dx[a] = 0.5f * (n.ch[c].box.lo[a] + n.ch[c].box.hi[a]) - cx[a];
                for(int s = 0; s < CW_WIDTH; ++s)
                {
                    if(slot_used[s]) continue;
                    float score = ((s & 4) ? dx[0] : -dx[0]) + ((s & 2) ? dx[1] : -dx[1]) + ((s & 1) ? dx[2] : -dx[2]);
                    if(score > best) { best = score; bc = c; bs = s; }
                }
            }
            slot_of[bc] = bs; slot_used[bs] = true; child_done[bc] = true;
        }
        int child_in_slot[CW_WIDTH]; for(int s = 0; s < CW_WIDTH; ++s) child_in_slot[s] = -1;
        for(int c = 0; c < k; ++c) child_in_slot[slot_of[c]] = c;

        // quantisation grid
        int e[3]; float scale[3];
        for(int a = 0; a < 3; ++a)
        {
            float ext = n.box.hi[a] - n.box.lo[a];
            int ex = ext > 0.0f ? (int)std::ceil(std::log2((double)ext / 255.0)) : -126;
            if(ex < -126) ex = -126;
            // make sure 255 steps really cover the extent in float arithmetic
            while(n.box.lo[a] + 255.0f * std::ldexp(1.0f, ex) < n.box.hi[a]) ex++;
            // (59: the traversal adds the exponent to that of 1/d, up to 2^67, in integer arithmetic: pt_cwbvh.cuh)
            if(ex > 59) { err = "box extent too large to quantise"; return 0; }
            e[a] = ex; scale[a] = std::ldexp(1.0f, ex);
        }
        uint8_t qlo[3][8] = {{0}}, qhi[3][8] = {{0}}, meta[8] = {0};
        uint32_t imask = 0, n_inner = 0;
        for(int s = 0; s < CW_WIDTH; ++s) if(child_in_slot[s] >= 0 && n.ch[child_in_slot[s]].inner >= 0) { imask |= 1u << s; n_inner++; }
        const uint32_t child_base = (uint32_t)(nodes.size() / 5);
        nodes.resize(nodes.size() + 5 * (size_t)n_inner);
        uint32_t tri_base = 0; bool have_tri_base = false; uint32_t tri_count = 0;
        // pass 1: boxes, meta bytes and the leaf payloads (kept contiguous per node)
        for(int s = 0; s < CW_WIDTH; ++s)
        {
            const int c = child_in_slot[s];
            if(c < 0) continue;
            const W8Child& ch = n.ch[c];
            for(int a = 0; a < 3; ++a)
            {
                int lo = (int)std::floor((ch.box.lo[a] - n.box.lo[a]) / scale[a]);
                int hi = (int)std::ceil((ch.box.hi[a] - n.box.lo[a]) / scale[a]);
                lo = std::min(std::max(lo, 0), 255); hi = std::min(std::max(hi, 0), 255);
                while(lo > 0 && n.box.lo[a] + (float)lo * scale[a] > ch.box.lo[a]) lo--;
                while(hi < 255 && n.box.lo[a] + (float)hi * scale[a] < ch.box.hi[a]) hi++;
                if(n.box.lo[a] + (float)lo * scale[a] > ch.box.lo[a] || n.box.lo[a] + (float)hi * scale[a] < ch.box.hi[a])
                { err = "quantised box does not enclose its child"; return 0; }
                qlo[a][s] = (uint8_t)lo; qhi[a][s] = (uint8_t)hi;
            }
            if(ch.inner >= 0) meta[s] = (uint8_t)((1u << 5) | (24u + (uint32_t)s));
            else
            {
                const uint32_t cnt = (uint32_t)ch.leaves.size();
                if(cnt == 0 || cnt > (uint32_t)CW_LEAF_MAX) { err = "leaf size out of range"; return 0; }
                const uint32_t first = emit_leaf(ch.leaves);
                if(!have_tri_base) { tri_base = first; have_tri_base = true; }
                const uint32_t off = first - tri_base;
                if(off != tri_count || off + cnt > 24u) { err = "leaf payloads of a node are not contiguous"; return 0; }
                tri_count += cnt;
                meta[s] = (uint8_t)((((1u << cnt) - 1u) << 5) | off);
            }
        }
        {   // empty slots: inverted box + the meta byte of a real child (pt_scene.cuh)
            static_assert(CW_LEAF_MAX == 1, "the maskless box test needs one triangle per leaf child");
            int first_used = -1;
            for(int s = 0; s < CW_WIDTH && first_used < 0; ++s) if(child_in_slot[s] >= 0) first_used = s;
            if(first_used < 0) { err = "node without children"; return 0; }
            for(int s = 0; s < CW_WIDTH; ++s)
            {
                if(child_in_slot[s] >= 0) continue;
                for(int a = 0; a < 3; ++a) { qlo[a][s] = 255; qhi[a][s] = 0; }
                meta[s] = meta[first_used];
            }
        }
        // pass 2: inner children, stored contiguously from child_base in slot order
        uint32_t deepest = 0, inner_rank = 0;
        for(int s = 0; s < CW_WIDTH; ++s)
        {
            const int c = child_in_slot[s];
            if(c < 0 || n.ch[c].inner < 0) continue;
            deepest = std::max(deepest, emit(n.ch[c].inner, child_base + inner_rank));
            if(!err.empty()) return 0;
            inner_rank++;
        }
        auto pack4 = [](const uint8_t* b) { return (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24); };
        float4* o = &nodes[5 * (size_t)cw];
        const uint32_t ew = (uint32_t)(uint8_t)(int8_t)e[0] | ((uint32_t)(uint8_t)(int8_t)e[1] << 8) | ((uint32_t)(uint8_t)(int8_t)e[2] << 16) | (imask << 24);
        o[0] = make_float4(n.box.lo[0], n.box.lo[1], n.box.lo[2], u2f(ew));
        o[1] = make_float4(u2f(child_base), u2f(tri_base), u2f(pack4(meta)), u2f(pack4(meta + 4)));
        o[2] = make_float4(u2f(pack4(qlo[0])), u2f(pack4(qlo[0] + 4)), u2f(pack4(qlo[1])), u2f(pack4(qlo[1] + 4)));
        o[3] = make_float4(u2f(pack4(qlo[2])), u2f(pack4(qlo[2] + 4)), u2f(pack4(qhi[0])), u2f(pack4(qhi[0] + 4)));
        o[4] = make_float4(u2f(pack4(qhi[1])), u2f(pack4(qhi[1] + 4)), u2f(pack4(qhi[2])), u2f(pack4(qhi[2] + 4)));
        return deepest + 1;
    }
};

// Builds the compressed BVH of one source tree; returns the root node index (or 0xFFFFFFFF).
uint32_t build_cw_tree(const std::vector<TNode>& tree, int leaf_max, std::vector<float4>& nodes,
                       const std::function<uint32_t(const std::vector<uint32_t>&)>& emit_leaf, uint32_t& depth, std::string& err)
{
    std::vector<W8Node> w8;
    CollapserW col{tree, CW_WIDTH, leaf_max, w8};
    int root;
    if(tree[0].count == 0)
    {   // single leaf: a node with one leaf child
        std::vector<GItem> items(1);
        items[0].box = tree[0].box; items[0].leaves.push_back(tree[0].payload);
        root = col.from_items(items, tree[0].box);
    }
    else if(leaf_max == 1 && CW_OPTIMAL_COLLAPSE)
    {
        OptimalCollapser opt(w8);
        root = opt.build(tree);
    }
    else root = col.build(0);
    CwEmitter em{w8, nodes, emit_leaf};
    const uint32_t cw_root = (uint32_t)(nodes.size() / 5);
    nodes.resize(nodes.size() + 5);
    depth = em.emit(root, cw_root);
    if(!em.err.empty()) { err = em.err; return 0xFFFFFFFFu; }
    return cw_root;
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

// Grows [lo, hi] by the eight corners of an object-space box taken through the instance transform.
void grow_by_transformed_box(const ptgpu_tlas_instance& inst, const float blo[3], const float bhi[3], float lo[3], float hi[3])
{
    const ptgpu_float4* c = inst.transform.r;
    for(int a = 0; a < 8; ++a)
    {
        const float x = (a & 1) ? bhi[0] : blo[0], y = (a & 2) ? bhi[1] : blo[1], z = (a & 4) ? bhi[2] : blo[2];
        const float v[3] = {
            c[0].x * x + c[1].x * y + c[2].x * z + c[3].x,
            c[0].y * x + c[1].y * y + c[2].y * z + c[3].y,
            c[0].z * x + c[1].z * y + c[2].z * z + c[3].z};
        for(int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], v[k]); hi[k] = std::max(hi[k], v[k]); }
    }
}

// The traversal tests triangles in object space, these boxes in world space: keep a margin of a few
// ulps of the coordinates between the two.
void pad_world_box(float lo[3], float hi[3])
{
    float extent = 0.0f, mag = 0.0f;
    for(int k = 0; k < 3; ++k)
    {
        extent = std::max(extent, hi[k] - lo[k]);
        mag = std::max(mag, std::max(std::fabs(lo[k]), std::fabs(hi[k])));
    }
    const float pad = 1e-5f * extent + 1e-6f * mag;
    for(int k = 0; k < 3; ++k) { lo[k] -= pad; hi[k] += pad; }
}


// ---- flat static scene: one compressed 8-wide BVH over the world-space triangles of all static instances ----
//
// The reference keeps every object as an instance of its mesh (bvh.cc:252-284) and so did the first
// version of this library. The static part of the scene is a forest: 884 trees and rocks whose boxes
// overlap each other and the terrain, so a ray entered 2.35 instances on average and 16 % of the entries
// hit nothing below the BLAS root. With 180 GB of HBM the instancing buys nothing: the 15.6 M instanced
// triangles are 750 MB as world-space vertices + 250 MB of nodes, and one SAH tree over all of them has
// no instance overlap, no ray transform and no second stack level. Per-frame instances (<= 7 hero
// objects, motion-blurred) stay instances.
//
// Build: primitives are split top-down (binned SAH) into ranges of at most FLAT_RANGE triangles; the ranges
// are built, collapsed and emitted in parallel with the BLAS code path; a small top tree over the ranges is
// collapsed on its own and its leaf slots are filled with copies of the subtree root nodes.

constexpr size_t FLAT_RANGE = 96 * 1024;

template<class F>
void parallel_for(size_t n, unsigned threads, F&& fn)
{
    if(threads <= 1 || n <= 1) { for(size_t i = 0; i < n; ++i) fn(i); return; }
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    const unsigned t = (unsigned)std::min<size_t>(threads, n);
    for(unsigned k = 0; k < t; ++k)
        pool.emplace_back([&]() { for(size_t i; (i = next.fetch_add(1)) < n;) fn(i); });
    for(auto& th : pool) th.join();
}

struct FlatRange { size_t begin, end; };

struct FlatTop
{
    std::deque<TNode> nodes;          // stable references while other threads append
    std::vector<FlatRange> ranges;
    std::mutex mu;
    uint32_t alloc2() { std::lock_guard<std::mutex> g(mu); nodes.emplace_back(); nodes.emplace_back(); return (uint32_t)nodes.size() - 2; }
    uint32_t add_range(size_t b, size_t e) { std::lock_guard<std::mutex> g(mu); ranges.push_back({b, e}); return (uint32_t)ranges.size() - 1; }
    TNode& at(uint32_t i) { std::lock_guard<std::mutex> g(mu); return nodes[i]; }
};

void split_top(std::vector<Prim>& prims, size_t begin, size_t end, FlatTop& top, uint32_t self, bool parallel)
{
    Box b; b.reset();
    for(size_t i = begin; i < end; ++i) b.grow(prims[i].box);
    TNode& me = top.at(self);
    me.box = b;
    if(end - begin <= FLAT_RANGE)
    {
        me.count = 0;
        me.payload = top.add_range(begin, end);
        return;
    }
    const size_t mid = binned_split(prims, begin, end);
    const uint32_t first = top.alloc2();
    me.first = first; me.count = 2;
    if(parallel && end - begin > 8 * FLAT_RANGE)
    {
        std::thread left([&]() { split_top(prims, begin, mid, top, first, true); });
        split_top(prims, mid, end, top, first + 1, true);
        left.join();
    }
    else
    {
        split_top(prims, begin, mid, top, first, parallel);
        split_top(prims, mid, end, top, first + 1, parallel);
    }
}

// Moves the nodes [0, count) of a compressed tree whose root is node 0 into breadth-first order, so that
// "index < K" selects its top levels (the traversal kernel can stage those in shared memory).
void cw_breadth_first(std::vector<float4>& nodes, uint32_t count, const std::vector<std::pair<uint32_t, int>>& stubs,
                      std::vector<std::pair<uint32_t, int>>& stubs_out)
{
    std::vector<uint8_t> is_stub(count, 0);
    for(const auto& sb : stubs) is_stub[sb.first] = 1;
    std::vector<float4> out(5 * (size_t)count);
    std::vector<uint32_t> new_of(count, 0xFFFFFFFFu);
    std::vector<uint32_t> queue{0u};
    new_of[0] = 0;
    uint32_t next = 1;
    for(size_t q = 0; q < queue.size(); ++q)
    {
        const uint32_t old = queue[q], nw = new_of[old];
        for(int k = 0; k < 5; ++k) out[5 * (size_t)nw + k] = nodes[5 * (size_t)old + k];
        if(is_stub[old]) continue;
        const uint32_t imask = f2u(nodes[5 * (size_t)old].w) >> 24, kids = (uint32_t)__builtin_popcount(imask);
        if(kids == 0) continue;
        const uint32_t base = f2u(nodes[5 * (size_t)old + 1].x);
        out[5 * (size_t)nw + 1].x = u2f(next);
        for(uint32_t j = 0; j < kids; ++j) { new_of[base + j] = next + j; queue.push_back(base + j); }
        next += kids;
    }
    stubs_out.clear();
    for(const auto& sb : stubs) stubs_out.push_back({new_of[sb.first], sb.second});
    for(size_t i = 0; i < out.size(); ++i) nodes[i] = out[i];
}

} // namespace

bool make_wide_instance(const WideScene& ws, const ptgpu_tlas_instance& inst, uint32_t ref_index, WideInstance& out)
{
    int found = -1;
    // an instance names its BLAS by the reference bvh handle, or (scene built from meshes) by its mesh
    for(size_t b = 0; b < ws.blas_info.size(); ++b)
        if(ws.from_meshes ? ws.blas_info[b].mesh.index_offset == inst.m.index_offset
                          : ws.blas_info[b].ref_node_offset == inst.blas.node_offset) { found = (int)b; break; }
    if(found < 0) return false;
    const WideBlasInfo& bi = ws.blas_info[found];
    if((!ws.from_meshes && bi.ref_node_count != inst.blas.node_count) || bi.mesh.index_offset != inst.m.index_offset ||
       bi.mesh.base_vertex_offset != inst.m.base_vertex_offset || bi.mesh.triangle_count != inst.m.triangle_count)
        return false;
    const ptgpu_float4* r = inst.inv_transform.r; // columns
    memset(&out, 0, sizeof(out));
    out.inv0 = make_float4(r[0].x, r[1].x, r[2].x, r[3].x);
    out.inv1 = make_float4(r[0].y, r[1].y, r[2].y, r[3].y);
    out.inv2 = make_float4(r[0].z, r[1].z, r[2].z, r[3].z);
    const WideBlas& wb = ws.blas[found];
    float lo[3] = {wb.lo.x, wb.lo.y, wb.lo.z}, hi[3] = {wb.hi.x, wb.hi.y, wb.hi.z}, wlo[3], whi[3];
    if(bi.sub_boxes.empty()) instance_world_box(inst, lo, hi, wlo, whi);
    else
    {
        for(int k = 0; k < 3; ++k) { wlo[k] = FLT_MAX; whi[k] = -FLT_MAX; }
        for(const WideSubBox& sb : bi.sub_boxes) grow_by_transformed_box(inst, sb.lo, sb.hi, wlo, whi);
        pad_world_box(wlo, whi);
    }
    out.lo = make_float4(wlo[0], wlo[1], wlo[2], 0.0f);
    out.hi = make_float4(whi[0], whi[1], whi[2], 0.0f);
    out.blas = (uint32_t)found;
    out.ref_instance = ref_index;
    out.cw_root = (size_t)found < ws.cw_blas_root.size() ? ws.cw_blas_root[found] : 0u;
    return true;
}

bool build_wide_scene(
    const ptgpu_bvh_node* nodes, size_t n_nodes, const ptgpu_bvh_link* links,
    const uint32_t* indices, size_t n_indices, const ptgpu_float3* pos, size_t n_verts,
    const ptgpu_tlas_instance* instances, size_t n_static,
    WideScene& out, std::string& err, const ptgpu_mesh* meshes, size_t n_meshes)
{
    out = WideScene();
    out.from_meshes = nodes == nullptr;
    // 1. every BLAS: recovered from the reference's static BVH region, in build order (= mesh load
    //    order, scene.cc:42-47), or built here from the triangles of each mesh of the table
    size_t off = 0, index_cursor = 0, vertex_cursor = 0, mesh_i = 0;
    std::vector<TNode> tree;
    uint32_t cw_blas_depth = 0;
    while(out.from_meshes ? mesh_i < n_meshes : off < n_nodes)
    {
        uint32_t count = 0, tri_count = 0;
        ptgpu_mesh m;
        if(out.from_meshes)
        {
            m = meshes[mesh_i++];
            if(m.triangle_count == 0) { err = "mesh " + std::to_string(mesh_i - 1) + " has no triangles"; return false; }
            if((size_t)m.index_offset + 3 * (size_t)m.triangle_count > n_indices) { err = "mesh " + std::to_string(mesh_i - 1) + ": indices beyond the index buffer"; return false; }
            uint32_t vmax = 0;
            for(size_t i = 0; i < 3 * (size_t)m.triangle_count; ++i) vmax = std::max(vmax, indices[m.index_offset + i]);
            if((size_t)m.base_vertex_offset + vmax >= n_verts) { err = "mesh " + std::to_string(mesh_i - 1) + ": vertex index beyond the vertex buffer"; return false; }
            m.vertex_count = vmax + 1;
            build_mesh_tree(indices, pos, m, tree);
            tri_count = m.triangle_count;
        }
        else
        {
            count = recover_tree(nodes, links, n_nodes, off, tree, err);
            if(count == 0) { err = "BVH at node " + std::to_string(off) + ": " + err; return false; }
            for(const TNode& t : tree) if(t.count == 0) tri_count++;
            // the i-th BLAS belongs to the i-th mesh: indices and vertices are appended in load order
            // (mesh.cc:118-119, 255-259), and every vertex is referenced by at least one index
            m.triangle_count = tri_count;
            m.index_offset = (uint32_t)index_cursor;
            m.base_vertex_offset = (uint32_t)vertex_cursor;
            if(index_cursor + 3 * (size_t)tri_count > n_indices) { err = "index buffer shorter than the BLAS leaves imply"; return false; }
            uint32_t vmax = 0;
            for(size_t i = 0; i < 3 * (size_t)tri_count; ++i) vmax = std::max(vmax, indices[index_cursor + i]);
            m.vertex_count = vmax + 1;
            if(vertex_cursor + m.vertex_count > n_verts) { err = "vertex buffer shorter than the indices imply"; return false; }
        }

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
        {   // the same tree in the compressed 8-wide layout
            auto emit_leaf_cw = [&](const std::vector<uint32_t>& prims) -> uint32_t {
                const uint32_t first = (uint32_t)(out.cw_tris.size() / 3);
                for(uint32_t p : prims)
                {
                    const uint32_t* ix = indices + m.index_offset + 3 * (size_t)p;
                    for(int k = 0; k < 3; ++k)
                    {
                        const ptgpu_float3& v = pos[m.base_vertex_offset + ix[k]];
                        float w = 0.0f;
                        if(k == 0) memcpy(&w, &p, 4);
                        out.cw_tris.push_back(make_float4(v.x, v.y, v.z, w));
                    }
                }
                return first;
            };
            uint32_t depth = 0;
            uint32_t root = build_cw_tree(tree, CW_LEAF_MAX, out.cw_nodes, emit_leaf_cw, depth, err);
            if(root == 0xFFFFFFFFu) { err = "compressed BVH of the BLAS at node " + std::to_string(off) + ": " + err; return false; }
            out.cw_blas_root.push_back(root);
            cw_blas_depth = std::max(cw_blas_depth, depth);
        }
        wb.node_count = (uint32_t)out.nodes.size() - wb.node_offset;
        wb.tri_count = (uint32_t)(out.tris.size() / 3) - wb.tri_offset;
        if(wb.tri_count != tri_count) { err = "triangle count mismatch after collapse"; return false; }
        if(wb.tri_count >= (1u << 27)) { err = "BLAS too large for the leaf encoding"; return false; }
        wb.lo = make_float4(tree[0].box.lo[0], tree[0].box.lo[1], tree[0].box.lo[2], 0.0f);
        wb.hi = make_float4(tree[0].box.hi[0], tree[0].box.hi[1], tree[0].box.hi[2], 0.0f);
        out.blas.push_back(wb);
        WideBlasInfo info{(uint32_t)off, count, m, stack, {}};
        {   // frontier of the source tree six levels below the root
            std::vector<std::pair<uint32_t, int>> todo{{0u, 0}};
            while(!todo.empty())
            {
                const auto [ni, depth] = todo.back(); todo.pop_back();
                const TNode& t = tree[ni];
                if(t.count == 0 || depth == 6)
                {
                    WideSubBox sb;
                    for(int k = 0; k < 3; ++k) { sb.lo[k] = t.box.lo[k]; sb.hi[k] = t.box.hi[k]; }
                    info.sub_boxes.push_back(sb);
                }
                else for(uint32_t c = 0; c < t.count; ++c) todo.push_back({t.first + c, depth + 1});
            }
        }
        out.blas_info.push_back(std::move(info));
        out.max_stack = std::max(out.max_stack, stack);
        off += count;
        index_cursor += 3 * (size_t)tri_count;
        vertex_cursor += m.vertex_count;
    }
    if(!out.from_meshes && (index_cursor != n_indices || vertex_cursor != n_verts))
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
    // Static instances get the exact world bounds of their transformed vertices (one pass at upload):
    // the reference bounds an instance by the transformed corners of its root box (bvh.cc:262-278),
    // which for a tree turned about its trunk is up to twice the footprint, and 28-36 % of all instance
    // entries of a frame met no child box of the BLAS root.
    for(size_t i = 0; i < n_static; ++i)
    {
        const ptgpu_tlas_instance& inst = instances[i];
        const ptgpu_float4* c = inst.transform.r;
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        const ptgpu_mesh& dm = out.blas_info[out.instances[i].blas].mesh;   // derived from the indices, checked above
        const ptgpu_float3* v = pos + dm.base_vertex_offset;
        for(uint32_t k = 0; k < dm.vertex_count; ++k)
        {
            const float w[3] = {
                c[0].x * v[k].x + c[1].x * v[k].y + c[2].x * v[k].z + c[3].x,
                c[0].y * v[k].x + c[1].y * v[k].y + c[2].y * v[k].z + c[3].y,
                c[0].z * v[k].x + c[1].z * v[k].y + c[2].z * v[k].z + c[3].z};
            for(int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], w[a]); hi[a] = std::max(hi[a], w[a]); }
        }
        if(dm.vertex_count == 0) continue;
        pad_world_box(lo, hi);
        out.instances[i].lo = make_float4(lo[0], lo[1], lo[2], 0.0f);
        out.instances[i].hi = make_float4(hi[0], hi[1], hi[2], 0.0f);
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
    {
        // compressed TLAS over all static instances
        std::vector<Prim> cprims;
        for(const Prim& pr : prims) cprims.push_back(pr);
        std::vector<TNode> ctree(1);
        ctree.reserve(2 * cprims.size());
        build_sah(cprims, 0, cprims.size(), ctree, 0);
        auto emit_inst_cw = [&](const std::vector<uint32_t>& ids) -> uint32_t {
            const uint32_t first = (uint32_t)out.cw_inst_index.size();
            for(uint32_t id : ids) out.cw_inst_index.push_back(id);
            return first;
        };
        uint32_t depth = 0;
        out.cw_tlas_root = build_cw_tree(ctree, 1, out.cw_nodes, emit_inst_cw, depth, err);
        if(out.cw_tlas_root == 0xFFFFFFFFu) { err = "compressed TLAS: " + err; return false; }
        // per node visit at most two pushes (rest of the node group, postponed leaf group); entering an
        // instance pushes the rest of its group and the exit marker; one entry for the dynamic instances
        // (+2: the world-space ray constants parked under the exit marker by the wavefront kernel)
        // (+4: the TLAS root group and the parked constants + marker of the world-space start instance)
        out.cw_max_stack = 2 * depth + 4 + 2 * cw_blas_depth + 1 + 4;
        if(out.cw_max_stack > (uint32_t)CW_STACK)
        {
            err = "compressed traversal stack bound " + std::to_string(out.cw_max_stack) + " exceeds CW_STACK";
            return false;
        }
        for(size_t i = 0; i < n_static; ++i) out.instances[i].cw_root = out.cw_blas_root[out.instances[i].blas];
    }
    out.max_stack += tstack + 1 /* exit marker */ + 16 /* dynamic instances */;
    if(out.max_stack > WIDE_STACK)
    {
        err = "traversal stack bound " + std::to_string(out.max_stack) + " exceeds WIDE_STACK " + std::to_string(WIDE_STACK);
        return false;
    }
    return true;
}

bool build_flat_scene(const WideScene& ws, const uint32_t* indices, const ptgpu_float3* pos,
                      const ptgpu_tlas_instance* instances, size_t n_static,
                      uint32_t node_base, uint32_t tri_base, FlatScene& out, std::string& err)
{
    const auto t_begin = std::chrono::steady_clock::now();
    out = FlatScene();
    unsigned threads = std::thread::hardware_concurrency();
    if(const char* e = getenv("PTGPU_BUILD_THREADS")) threads = (unsigned)atoi(e);
    threads = std::max(1u, std::min(threads, 64u));

    // 1. flat triangle numbering: instance i owns [first_tri[i], first_tri[i + 1])
    std::vector<size_t> first_tri(n_static + 1, 0);
    for(size_t i = 0; i < n_static; ++i)
        first_tri[i + 1] = first_tri[i] + ws.blas_info[ws.instances[i].blas].mesh.triangle_count;
    const size_t n_tris = first_tri[n_static];
    if(n_tris == 0) { err = "flat scene: no triangles"; return false; }
    if(n_tris >= 0x7FFFFFFFull || (size_t)tri_base + n_tris >= 0x7FFFFFFFull) { err = "flat scene: too many triangles"; return false; }
    // world-space vertex k of triangle t of instance i (the arithmetic of mul_m4v4 on a point, math.hh:230-240)
    auto world_vertex = [&](size_t i, uint32_t t, int k, float v[3]) {
        const ptgpu_mesh& m = ws.blas_info[ws.instances[i].blas].mesh;
        const ptgpu_float3& p = pos[m.base_vertex_offset + indices[m.index_offset + 3 * (size_t)t + k]];
        const ptgpu_float4* c = instances[i].transform.r;
        v[0] = c[0].x * p.x + c[1].x * p.y + c[2].x * p.z + c[3].x;
        v[1] = c[0].y * p.x + c[1].y * p.y + c[2].y * p.z + c[3].y;
        v[2] = c[0].z * p.x + c[1].z * p.y + c[2].z * p.z + c[3].z;
    };
    std::vector<uint8_t> mirrored(n_static, 0);   // a reflecting transform turns front faces into back faces
    for(size_t i = 0; i < n_static; ++i)
    {
        const ptgpu_float4* c = instances[i].transform.r;
        const double det = (double)c[0].x * ((double)c[1].y * c[2].z - (double)c[1].z * c[2].y)
                         - (double)c[1].x * ((double)c[0].y * c[2].z - (double)c[0].z * c[2].y)
                         + (double)c[2].x * ((double)c[0].y * c[1].z - (double)c[0].z * c[1].y);
        mirrored[i] = det < 0.0;
    }
    std::vector<Prim> prims(n_tris);
    parallel_for(n_static, threads, [&](size_t i) {
        const uint32_t count = ws.blas_info[ws.instances[i].blas].mesh.triangle_count;
        for(uint32_t t = 0; t < count; ++t)
        {
            Prim& pr = prims[first_tri[i] + t];
            pr.box.reset(); pr.id = (uint32_t)(first_tri[i] + t);
            for(int k = 0; k < 3; ++k)
            {
                float v[3]; world_vertex(i, t, k, v);
                for(int a = 0; a < 3; ++a) { pr.box.lo[a] = std::min(pr.box.lo[a], v[a]); pr.box.hi[a] = std::max(pr.box.hi[a], v[a]); }
            }
        }
    });

    // 2. top-down into ranges
    FlatTop top;
    top.nodes.emplace_back();
    split_top(prims, 0, n_tris, top, 0, threads > 1);
    const size_t n_ranges = top.ranges.size();

    // 3. every range: SAH tree -> optimal 8-wide collapse -> compressed nodes + world-space triangles
    struct Sub { std::vector<float4> nodes, tris; uint32_t depth = 0; std::string err; };
    std::vector<Sub> subs(n_ranges);
    parallel_for(n_ranges, threads, [&](size_t r) {
        Sub& sub = subs[r];
        const FlatRange fr = top.ranges[r];
        std::vector<TNode> tree(1);
        tree.reserve(2 * (fr.end - fr.begin));
        build_sah(prims, fr.begin, fr.end, tree, 0, 256);
        sub.tris.reserve(3 * (fr.end - fr.begin));
        auto emit_leaf = [&](const std::vector<uint32_t>& ids) -> uint32_t {
            const uint32_t first = (uint32_t)(sub.tris.size() / 3);
            for(uint32_t id : ids)
            {
                const size_t i = (size_t)(std::upper_bound(first_tri.begin(), first_tri.end(), (size_t)id) - first_tri.begin()) - 1;
                const uint32_t t = (uint32_t)(id - first_tri[i]);
                for(int k = 0; k < 3; ++k)
                {
                    float v[3]; world_vertex(i, t, k, v);
                    // .w: vertex 0 = primitive id in its mesh, vertex 1 = instance | mirrored << 31
                    const uint32_t w = k == 0 ? t : k == 1 ? ((uint32_t)i | (mirrored[i] ? 0x80000000u : 0u)) : 0u;
                    sub.tris.push_back(make_float4(v[0], v[1], v[2], u2f(w)));
                }
            }
            return first;
        };
        uint32_t root = build_cw_tree(tree, CW_LEAF_MAX, sub.nodes, emit_leaf, sub.depth, sub.err);
        if(root != 0u && sub.err.empty()) sub.err = "subtree root is not node 0";
    });
    for(const Sub& sub : subs) if(!sub.err.empty()) { err = "flat scene: " + sub.err; return false; }

    // 4. top tree over the ranges; its leaves are placeholders for the subtree roots
    std::vector<float4> top_nodes;
    std::vector<std::pair<uint32_t, int>> stubs;
    uint32_t top_depth = 0;
    if(n_ranges > 1)
    {
        std::vector<TNode> ttree(top.nodes.begin(), top.nodes.end());
        std::vector<W8Node> w8;
        OptimalCollapser opt(w8);
        const int root = opt.build(ttree);
        const size_t n_w8 = w8.size();
        for(size_t n = 0; n < n_w8; ++n)
            for(size_t c = 0; c < w8[n].ch.size(); ++c)
                if(w8[n].ch[c].inner < 0)
                {
                    W8Node stub; stub.box = w8[n].ch[c].box; stub.stub = (int)w8[n].ch[c].leaves[0];
                    w8[n].ch[c].leaves.clear();
                    w8[n].ch[c].inner = (int)w8.size();
                    w8.push_back(std::move(stub));
                }
        std::function<uint32_t(const std::vector<uint32_t>&)> no_leaf = [](const std::vector<uint32_t>&) -> uint32_t { return 0u; };
        CwEmitter em{w8, top_nodes, no_leaf};
        top_nodes.resize(5);
        top_depth = em.emit(root, 0u);
        if(!em.err.empty()) { err = "flat scene, top tree: " + em.err; return false; }
        cw_breadth_first(top_nodes, (uint32_t)(top_nodes.size() / 5), em.stubs, stubs);
        if(stubs.size() != n_ranges) { err = "flat scene: top tree lost a range"; return false; }
    }

    // 5. concatenate: [top nodes | subtree 0 | subtree 1 | ...], indices relocated to the device arrays
    const uint32_t n_top = (uint32_t)(top_nodes.size() / 5);
    std::vector<uint32_t> node_off(n_ranges), tri_off(n_ranges);
    size_t n_nodes = n_top, n_t = 0;
    for(size_t r = 0; r < n_ranges; ++r)
    {
        node_off[r] = (uint32_t)n_nodes; tri_off[r] = (uint32_t)n_t;
        n_nodes += subs[r].nodes.size() / 5; n_t += subs[r].tris.size() / 3;
        out.depth = std::max(out.depth, subs[r].depth);
    }
    if(n_t != n_tris) { err = "flat scene: triangle count mismatch"; return false; }
    if((size_t)node_base + n_nodes >= 0x7FFFFFFFull) { err = "flat scene: too many nodes"; return false; }
    out.depth += top_depth;
    out.n_top = n_top;
    out.nodes.resize(5 * n_nodes);
    out.tris.resize(3 * n_tris);
    parallel_for(n_ranges, threads, [&](size_t r) {
        Sub& sub = subs[r];
        const size_t cnt = sub.nodes.size() / 5;
        for(size_t n = 0; n < cnt; ++n)
        {
            float4* o = &out.nodes[5 * (node_off[r] + n)];
            for(int k = 0; k < 5; ++k) o[k] = sub.nodes[5 * n + k];
            o[1].x = u2f(f2u(o[1].x) + node_base + node_off[r]);
            o[1].y = u2f(f2u(o[1].y) + tri_base + tri_off[r]);
        }
        std::copy(sub.tris.begin(), sub.tris.end(), out.tris.begin() + 3 * (size_t)tri_off[r]);
        sub.nodes = std::vector<float4>(); sub.tris = std::vector<float4>();
    });
    for(uint32_t n = 0; n < n_top; ++n)
    {
        for(int k = 0; k < 5; ++k) out.nodes[5 * (size_t)n + k] = top_nodes[5 * (size_t)n + k];
        out.nodes[5 * (size_t)n + 1].x = u2f(f2u(top_nodes[5 * (size_t)n + 1].x) + node_base);
    }
    for(const auto& sb : stubs)   // the slot the top tree reserved for a subtree root gets a copy of that root
        for(int k = 0; k < 5; ++k) out.nodes[5 * (size_t)sb.first + k] = out.nodes[5 * (size_t)node_off[sb.second] + k];
    const TNode& rootn = top.nodes[0];
    for(int a = 0; a < 3; ++a) { out.lo[a] = rootn.box.lo[a]; out.hi[a] = rootn.box.hi[a]; }
    out.n_tris = n_tris;
    out.build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
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

    // compressed 8-wide layout: decode every node the way the kernel does and repeat the checks
    auto walk_cw = [&](uint32_t root, bool tlas, std::vector<uint32_t>& payload_seen, uint32_t payload_lo, uint32_t payload_hi) {
        // `box` = intersection of the slot boxes of all ancestors: what a ray must hit to get here
        struct Todo { uint32_t node; Box box; bool has_box; };
        std::vector<Todo> todo{{root, Box(), false}};
        while(!todo.empty())
        {
            Todo t = todo.back(); todo.pop_back();
            if(5 * (size_t)t.node + 4 >= ws.cw_nodes.size()) { fail("compressed child index outside the array"); continue; }
            const float4* n = &ws.cw_nodes[5 * (size_t)t.node];
            const uint32_t ew = f2u(n[0].w);
            const float sc[3] = {std::ldexp(1.0f, (int)(int8_t)(ew & 0xFF)), std::ldexp(1.0f, (int)(int8_t)((ew >> 8) & 0xFF)), std::ldexp(1.0f, (int)(int8_t)((ew >> 16) & 0xFF))};
            const float p[3] = {n[0].x, n[0].y, n[0].z};
            const uint32_t imask = ew >> 24, child_base = f2u(n[1].x), tri_base = f2u(n[1].y);
            const uint32_t metaw[2] = {f2u(n[1].z), f2u(n[1].w)};
            const uint32_t q[6][2] = {{f2u(n[2].x), f2u(n[2].y)}, {f2u(n[2].z), f2u(n[2].w)}, {f2u(n[3].x), f2u(n[3].y)},
                                      {f2u(n[3].z), f2u(n[3].w)}, {f2u(n[4].x), f2u(n[4].y)}, {f2u(n[4].z), f2u(n[4].w)}};
            uint32_t inner_rank = 0;
            for(int s = 0; s < 8; ++s)
            {
                const uint32_t meta = (metaw[s >> 2] >> (8 * (s & 3))) & 0xFF;
                if(meta == 0) { if(imask & (1u << s)) fail("imask set on an empty slot"); continue; }
                bool padding = false;   // an empty slot is an inverted box
                for(int a = 0; a < 3; ++a)
                    if(((q[a][s >> 2] >> (8 * (s & 3))) & 0xFF) > ((q[3 + a][s >> 2] >> (8 * (s & 3))) & 0xFF)) padding = true;
                if(padding) { if(imask & (1u << s)) fail("imask set on a padding slot"); continue; }
                Box cb;
                for(int a = 0; a < 3; ++a)
                {
                    cb.lo[a] = p[a] + (float)((q[a][s >> 2] >> (8 * (s & 3))) & 0xFF) * sc[a];
                    cb.hi[a] = p[a] + (float)((q[3 + a][s >> 2] >> (8 * (s & 3))) & 0xFF) * sc[a];
                }
                const bool inner = (meta & 0x18) == 0x18 && (meta >> 5) == 1;
                if(inner != ((imask >> s) & 1u)) fail("imask disagrees with meta");
                if(inner)
                {
                    if((meta & 0x1F) != 24u + (uint32_t)s) fail("inner meta does not encode its slot");
                    Box nb = cb;
                    if(t.has_box) for(int a = 0; a < 3; ++a) { nb.lo[a] = std::max(nb.lo[a], t.box.lo[a]); nb.hi[a] = std::min(nb.hi[a], t.box.hi[a]); }
                    todo.push_back({child_base + inner_rank, nb, true});
                    inner_rank++;
                    continue;
                }
                const uint32_t bits = meta >> 5, off = meta & 0x1F;
                const uint32_t cnt = bits == 1 ? 1 : bits == 3 ? 2 : bits == 7 ? 3 : 0;
                if(cnt == 0) { fail("bad unary leaf count"); continue; }
                for(uint32_t k = 0; k < cnt; ++k)
                {
                    const uint32_t idx = tri_base + off + k;
                    if(tlas)
                    {
                        if(idx >= ws.cw_inst_index.size()) { fail("instance slot outside the list"); continue; }
                        const uint32_t id = ws.cw_inst_index[idx];
                        if(id >= payload_hi) fail("instance id out of range"); else payload_seen[id]++;
                        const WideInstance& wi = ws.instances[id];
                        const float lo[3] = {wi.lo.x, wi.lo.y, wi.lo.z}, hi[3] = {wi.hi.x, wi.hi.y, wi.hi.z};
                        for(int a = 0; a < 3; ++a) if(lo[a] < cb.lo[a] || hi[a] > cb.hi[a]) fail("instance box outside its quantised leaf box");
                    }
                    else
                    {
                        if(idx < payload_lo || idx >= payload_hi) { fail("triangle slot outside the BLAS range"); continue; }
                        payload_seen[idx - payload_lo]++;
                        const float4* v = &ws.cw_tris[3 * (size_t)idx];
                        for(int c = 0; c < 3; ++c)
                        {
                            const float pp[3] = {v[c].x, v[c].y, v[c].z};
                            for(int a = 0; a < 3; ++a) if(pp[a] < cb.lo[a] || pp[a] > cb.hi[a]) fail("triangle vertex outside its quantised leaf box");
                            if(t.has_box)
                                for(int a = 0; a < 3; ++a) if(pp[a] < t.box.lo[a] || pp[a] > t.box.hi[a])
                                    fail("triangle " + std::to_string(idx) + " vertex outside an ancestor's slot box (node " + std::to_string(t.node) + ")");
                        }
                    }
                }
            }
        }
    };
    if(ws.cw_blas_root.size() != ws.blas.size()) fail("compressed BLAS table incomplete");
    else
    {
        uint32_t tri_cursor = 0;
        for(size_t b = 0; b < ws.blas.size(); ++b)
        {
            std::vector<uint32_t> seen(ws.blas[b].tri_count, 0);
            walk_cw(ws.cw_blas_root[b], false, seen, tri_cursor, tri_cursor + ws.blas[b].tri_count);
            for(uint32_t c : seen) if(c != 1) { fail("compressed BLAS " + std::to_string(b) + ": triangle missing or duplicated"); break; }
            tri_cursor += ws.blas[b].tri_count;
        }
        std::vector<uint32_t> seen(n_static, 0);
        walk_cw(ws.cw_tlas_root, true, seen, 0, (uint32_t)n_static);
        for(uint32_t c : seen) if(c != 1) { fail("compressed TLAS: instance missing or duplicated"); break; }
    }
    return bad;
}

// Structural check of the flat static scene (CPU tests): every triangle is a leaf exactly once, every
// vertex lies inside its quantised leaf box and inside the slot boxes of all its ancestors.
uint64_t verify_flat_scene(const FlatScene& fs, uint32_t node_base, uint32_t tri_base, std::string& err)
{
    uint64_t bad = 0;
    auto fail = [&](const std::string& m) { if(bad++ == 0) err = m; };
    const size_t n_nodes = fs.nodes.size() / 5, n_tris = fs.tris.size() / 3;
    std::vector<uint8_t> seen(n_tris, 0);
    struct Todo { uint32_t node; Box box; bool has_box; uint32_t depth; };
    std::vector<Todo> todo{{0u, Box(), false, 1u}};
    uint32_t deepest = 0;
    while(!todo.empty())
    {
        Todo t = todo.back(); todo.pop_back();
        if(t.node >= n_nodes) { fail("flat: child index outside the array"); continue; }
        deepest = std::max(deepest, t.depth);
        const float4* n = &fs.nodes[5 * (size_t)t.node];
        const uint32_t ew = f2u(n[0].w);
        const float sc[3] = {std::ldexp(1.0f, (int)(int8_t)(ew & 0xFF)), std::ldexp(1.0f, (int)(int8_t)((ew >> 8) & 0xFF)), std::ldexp(1.0f, (int)(int8_t)((ew >> 16) & 0xFF))};
        const float p[3] = {n[0].x, n[0].y, n[0].z};
        const uint32_t imask = ew >> 24, child_base = f2u(n[1].x) - node_base, tri_first = f2u(n[1].y) - tri_base;
        const uint32_t metaw[2] = {f2u(n[1].z), f2u(n[1].w)};
        const uint32_t q[6][2] = {{f2u(n[2].x), f2u(n[2].y)}, {f2u(n[2].z), f2u(n[2].w)}, {f2u(n[3].x), f2u(n[3].y)},
                                  {f2u(n[3].z), f2u(n[3].w)}, {f2u(n[4].x), f2u(n[4].y)}, {f2u(n[4].z), f2u(n[4].w)}};
        uint32_t inner_rank = 0;
        for(int s = 0; s < 8; ++s)
        {
            const uint32_t meta = (metaw[s >> 2] >> (8 * (s & 3))) & 0xFF;
            bool padding = meta == 0;
            for(int a = 0; a < 3; ++a)
                if(((q[a][s >> 2] >> (8 * (s & 3))) & 0xFF) > ((q[3 + a][s >> 2] >> (8 * (s & 3))) & 0xFF)) padding = true;
            if(padding) { if(imask & (1u << s)) fail("flat: imask set on an empty slot"); continue; }
            Box cb;
            for(int a = 0; a < 3; ++a)
            {
                cb.lo[a] = p[a] + (float)((q[a][s >> 2] >> (8 * (s & 3))) & 0xFF) * sc[a];
                cb.hi[a] = p[a] + (float)((q[3 + a][s >> 2] >> (8 * (s & 3))) & 0xFF) * sc[a];
            }
            if((imask >> s) & 1u)
            {
                Box nb = cb;
                if(t.has_box) for(int a = 0; a < 3; ++a) { nb.lo[a] = std::max(nb.lo[a], t.box.lo[a]); nb.hi[a] = std::min(nb.hi[a], t.box.hi[a]); }
                todo.push_back({child_base + inner_rank, nb, true, t.depth + 1});
                inner_rank++;
                continue;
            }
            const uint32_t idx = tri_first + (meta & 0x1F);
            if((meta >> 5) != 1u) { fail("flat: leaf child with more than one triangle"); continue; }
            if(idx >= n_tris) { fail("flat: triangle index outside the array"); continue; }
            if(seen[idx]++) fail("flat: triangle reachable twice");
            const float4* v = &fs.tris[3 * (size_t)idx];
            for(int c = 0; c < 3; ++c)
            {
                const float pp[3] = {v[c].x, v[c].y, v[c].z};
                for(int a = 0; a < 3; ++a)
                {
                    if(pp[a] < cb.lo[a] || pp[a] > cb.hi[a]) fail("flat: triangle vertex outside its quantised leaf box");
                    if(t.has_box && (pp[a] < t.box.lo[a] || pp[a] > t.box.hi[a])) fail("flat: triangle vertex outside an ancestor's slot box");
                }
            }
        }
    }
    size_t missing = 0;
    for(uint8_t c : seen) if(c != 1) missing++;
    if(missing) fail("flat: " + std::to_string(missing) + " triangles missing or duplicated");
    if(deepest > fs.depth) fail("flat: depth " + std::to_string(deepest) + " exceeds the recorded " + std::to_string(fs.depth));
    return bad;
}

} // namespace pt
