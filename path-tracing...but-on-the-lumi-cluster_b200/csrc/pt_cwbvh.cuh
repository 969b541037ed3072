// Traversal of the compressed 8-wide BVH (layout: bvh_wide.cu, after Ylitie/Karras/Laine 2017),
// two-level: static TLAS (leaves = instances) + per-BLAS trees, plus the subframe's few dynamic
// instances tested directly at ray start. Replaces ray_query.hh:111-290 on the fast path.
//
// One node = five 16-byte loads for eight children. Child boxes are 8-bit offsets on a per-node
// power-of-two grid and are supersets of the exact boxes, so the triangles a ray can hit are the
// same as with the reference's boxes; the triangle test and its acceptance rule (strict
// tmin < t < tmax) are the reference's (math.hh:340-401, ray_query.hh:245).
// The traversal stack holds GROUPS: (base index, bit mask) of node children or of leaf payloads that
// passed the box test, so a node visit pushes at most two entries and needs no distance sort —
// children are stored in octant order and visited in (slot XOR ray octant) order.
#pragma once
#include <type_traits>
#include "pt_scene.cuh"
#include "pt_trav_links.cuh"
#include "pt_wide.cuh"
#include "bvh_wide.hh"

namespace pt {

#define CW_MARK_X 0xFFFFFFFFu
#ifndef CW_DYN_UNION
#define CW_DYN_UNION 1
#endif

// Traversal stack storage. LocalStack: a per-thread array (local memory). HybridStack: the first
// CW_SM_STACK entries in shared memory laid out [entry][thread] (bank = thread, conflict-free for any
// mix of depths), the rest in local memory. ncu on the all-local version: the loads of the stack top
// and of the pending-triangle list were the top three stall sites of wf_trace_cw (25 % of samples).
#ifndef CW_SM_STACK_N
#define CW_SM_STACK_N 7
#endif
constexpr int CW_SM_STACK = CW_SM_STACK_N;
// Both also hold the query's closest-hit record (barycentrics, instance, primitive | back_face << 31) in two
// entries behind the stack: it is written a few times per query and read once, so it has no business in
// registers that the box test could use (wf_trace_cw_kernel: 56 per thread).
#define CW_NO_HIT 0xFFFFFFFFu
struct LocalStack
{
    uint2* p;       // CW_STACK + 2 entries
    PT_D void set(int i, uint2 v) { p[i] = v; }
    PT_D uint2 get(int i) const { return p[i]; }
    PT_D void hit_set(float u, float v, uint32_t inst, uint32_t prim_bf) { p[CW_STACK] = make_uint2(__float_as_uint(u), __float_as_uint(v)); p[CW_STACK + 1] = make_uint2(inst, prim_bf); }
    PT_D void hit_clear() { p[CW_STACK] = make_uint2(0u, 0u); p[CW_STACK + 1] = make_uint2(CW_NO_HIT, 0u); }
    PT_D uint2 hit_uv() const { return p[CW_STACK]; }
    PT_D uint2 hit_id() const { return p[CW_STACK + 1]; }
};
struct HybridStack
{
    uint2* sm;      // &shared[0][thread]; CW_SM_STACK + 2 rows
    uint2* lo;      // overflow (local memory)
    int stride;     // threads per block
    PT_D void set(int i, uint2 v) { if(i < CW_SM_STACK) sm[i * stride] = v; else lo[i - CW_SM_STACK] = v; }
    PT_D uint2 get(int i) const { return i < CW_SM_STACK ? sm[i * stride] : lo[i - CW_SM_STACK]; }
    PT_D void hit_set(float u, float v, uint32_t inst, uint32_t prim_bf) { sm[CW_SM_STACK * stride] = make_uint2(__float_as_uint(u), __float_as_uint(v)); sm[(CW_SM_STACK + 1) * stride] = make_uint2(inst, prim_bf); }
    PT_D void hit_clear() { sm[CW_SM_STACK * stride] = make_uint2(0u, 0u); sm[(CW_SM_STACK + 1) * stride] = make_uint2(CW_NO_HIT, 0u); }
    PT_D uint2 hit_uv() const { return sm[CW_SM_STACK * stride]; }
    PT_D uint2 hit_id() const { return sm[(CW_SM_STACK + 1) * stride]; }
};

// Per-query traversal state. Kept as small as the algorithm allows: wf_trace_cw_kernel runs at 56 registers
// (9 blocks per SM), and every value held here across the box test is one the test cannot use. So the state
// does NOT hold the world-space ray (o and idir are the world ray while outside an instance; on the rare exit
// from one — <= 7 per-frame instances — the ray is read again from where the caller keeps it, see the `World`
// argument of the functions below), nor the sign bits (they follow from the octant), nor the subframe, nor the
// closest-hit record (it lives with the stack; `hit` only says that there is one, and its distance IS tmax).
struct CwState
{
    v3 o, idir;         // current space (world or instance): origin and clamped 1/direction
    v3 S;               // triangle-test preprocess of the current-space direction (math.hh:340-356)
    int axis;
    uint32_t oct_inv4;  // ray octant (bit 2: d.x >= 0, bit 1: d.y >= 0, bit 0: d.z >= 0), replicated in 4 bytes
    float tmin, tmax;
    uint2 ngroup, tgroup;
    int sp;
    bool in_blas, any, hit;
    uint32_t cur_inst;
#ifdef WF_STATS
    uint32_t n_nodes = 0, n_tris = 0;   // census builds: node tests and triangle tests of this query
#endif
    template<class Stack>
    PT_D Hit result(const Stack& stack) const
    {
        const uint2 uv = stack.hit_uv(), id = stack.hit_id();
        Hit h;
        h.t = hit ? tmax : -1.0f; h.u = __uint_as_float(uv.x); h.v = __uint_as_float(uv.y);
        h.inst = id.x; h.prim = id.y & 0x7FFFFFFFu; h.back_face = (id.y >> 31) != 0u;
        return h;
    }
};

// Where the world-space ray of a query lives while the state does not hold it. The plain traversal keeps it in
// registers (it has no register budget to meet); wf_trace_cw_kernel reads it back from the path-state pool.
struct WorldRayRegs
{
    v3 o, d;
    uint32_t sub;
    PT_D v3 origin() const { return o; }
    PT_D v3 dir() const { return d; }
    PT_D uint32_t subframe() const { return sub; }
};

PT_D uint32_t sign_extend_s8x4(uint32_t x)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, 0x0, 0x0000BA98;" : "=r"(r) : "r"(x));
    return r;
}

PT_D float rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

PT_D void cw_set_space(CwState& st, v3 o, v3 d)
{
    st.o = o;
    // a zero component would give inf * 0 = NaN against the quantised grid: clamp to +-1e-20
    const float eps = 1e-20f;
    // MUFU.RCP alone (1 ulp, rcp.approx) is enough here: the quantised child boxes carry far more slack
    // than that, and the box test already runs against a 1e-5 longer ray. The IEEE-rounded __frcp_rn
    // this replaces compiled to ~30 instructions per component (Newton step + special cases) and made
    // cw_set_space 3.8 % of all instructions of the traversal kernel.
    st.idir = mk3(rcp_approx(fabsf(d.x) > eps ? d.x : copysignf(eps, d.x)),
                  rcp_approx(fabsf(d.y) > eps ? d.y : copysignf(eps, d.y)),
                  rcp_approx(fabsf(d.z) > eps ? d.z : copysignf(eps, d.z)));
    // slot bit 4 = +x side, 2 = +y, 1 = +z; a ray going +x meets the -x children first
    const uint32_t oct = (d.x < 0.0f ? 0u : 4u) | (d.y < 0.0f ? 0u : 2u) | (d.z < 0.0f ? 0u : 1u);
    st.oct_inv4 = oct * 0x01010101u;
}

// Byte j of w as a float: I2F.U8 on the XU pipe. Forms that were measured and lost (profiles/r01_ and
// r02_trace_kernel_history.md): PRMT building 2^23 + b with a FADD to remove the 2^23 (one more instruction per
// byte, +1.4-6 %); the 2^23 folded into the slab offset (the folded offset rounds to half a grid cell, so boxes
// were stored one cell larger: +4-6 %). What is used next to it, for one axis of three, is the 2^15 form below.
PT_D float u8f(uint32_t w, int j) { return (float)((w >> (8 * j)) & 0xFFu); }

// Byte j of w as the float 32768 + b, built by ONE PRMT on the ALU pipe (no XU instruction): the byte lands in
// bits 8..15 of 0x47000000 = 2^15, where one unit of the byte is 2^8 ulp, so the value is exact. The slab offset
// of such an axis carries -32768 * scale (cw_intersect_node); folding at 2^15 instead of the 2^23 of the earlier
// experiment costs 2^-9 of a grid cell in rounding instead of half a cell: no box padding. `magic` = 0x47000000
// comes from the constant bank (Scene::cw_magic) so that PRMT takes the byte selector as its immediate.
#ifndef CW_MAGIC
#define CW_MAGIC 1      // bit mask of the axes converted this way: 1 x, 2 y, 4 z. Measured at 36 warps/SM, where the XU
                        // queue was what warps waited for (XU 69.5 % busy): x -2.2 %, x+y -0.8 %, all three +0.1 %
#endif
template<int AXIS, int J>
PT_D float u8f_axis(uint32_t w, uint32_t magic)
{
    if(CW_MAGIC & (1 << AXIS))
    {
        uint32_t r;
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(magic), "n"(0x7604 | (J << 4)));
        return __uint_as_float(r);
    }
    return u8f(w, J);
}

// Top levels of the flat BVH staged in shared memory (north_star: "staging the top levels in shared memory
// and leaving the rest to L2"): nodes [base, base + n) of the node array are also at sm[5 * (index - base)].
struct TopLevels
{
    const float4* sm;
    uint32_t base, n;
};
constexpr uint32_t CW_TOP_NODES = 96;   // 7.5 KB per block

// Box test of the eight children of one node; fills the node group (inner children hit) and the
// leaf group (leaf payloads whose box was hit).
template<bool TOP>
PT_D void cw_intersect_node(const float4* __restrict__ nodes, uint32_t node_index, const CwState& st,
                            uint2& ngroup, uint2& tgroup, const TopLevels& top, uint32_t magic)
{
    float4 n0, n1, n2, n3, n4;
    if(TOP && node_index - top.base < top.n)
    {
        const float4* n = top.sm + 5 * (node_index - top.base);
        n0 = n[0]; n1 = n[1]; n2 = n[2]; n3 = n[3]; n4 = n[4];
    }
    else
    {
        const float4* n = nodes + 5 * (size_t)node_index;
        n0 = __ldg(n); n1 = __ldg(n + 1); n2 = __ldg(n + 2); n3 = __ldg(n + 3); n4 = __ldg(n + 4);
    }
    const uint32_t ew = __float_as_uint(n0.w);
    // scale x 1/d as an integer add on the exponent field (three instructions fewer per node than decoding the
    // scale and multiplying; same bits): |1/d| lies in [1, 1e20] (cw_set_space) and the builder keeps the
    // exponents in [-126, 59] (bvh_wide.cu), so the sum never leaves the normal range
    const float ax = __uint_as_float(__float_as_uint(st.idir.x) + (uint32_t)((((int)(ew << 24)) >> 1) & (int)0xFF800000));
    const float ay = __uint_as_float(__float_as_uint(st.idir.y) + (uint32_t)((((int)(ew << 16)) >> 1) & (int)0xFF800000));
    const float az = __uint_as_float(__float_as_uint(st.idir.z) + (uint32_t)((((int)(ew << 8)) >> 1) & (int)0xFF800000));
    float ox = (n0.x - st.o.x) * st.idir.x, oy = (n0.y - st.o.y) * st.idir.y, oz = (n0.z - st.o.z) * st.idir.z;
    if(CW_MAGIC & 1) ox = fmaf(-32768.0f, ax, ox);
    if(CW_MAGIC & 2) oy = fmaf(-32768.0f, ay, oy);
    if(CW_MAGIC & 4) oz = fmaf(-32768.0f, az, oz);
    const bool nx = !(st.oct_inv4 & 4u), ny = !(st.oct_inv4 & 2u), nz = !(st.oct_inv4 & 1u);
    // Boxes are culled against a slightly longer ray than the triangles: the box arithmetic rounds, so
    // with the exact tmax a node holding a hit a few ulp closer than the current one could be skipped or
    // not depending on the order in which the two were found (coincident leaf cards in the tree
    // meshes). With the slack the closest hit is found in any order: results stay deterministic.
    const float tmax_box = fmaf(st.tmax, 1.0e-5f, st.tmax);
    uint32_t hitmask = 0;
    #pragma unroll
    for(int half = 0; half < 2; ++half)
    {
        const uint32_t meta4 = __float_as_uint(half ? n1.w : n1.z);
        const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
        const uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
        const uint32_t bit_index4 = (meta4 ^ (st.oct_inv4 & inner_mask4)) & 0x1F1F1F1Fu;
        const uint32_t qlx = __float_as_uint(half ? n2.y : n2.x), qly = __float_as_uint(half ? n2.w : n2.z);
        const uint32_t qlz = __float_as_uint(half ? n3.y : n3.x), qhx = __float_as_uint(half ? n3.w : n3.z);
        const uint32_t qhy = __float_as_uint(half ? n4.y : n4.x), qhz = __float_as_uint(half ? n4.w : n4.z);
        const uint32_t x_near = nx ? qhx : qlx, x_far = nx ? qlx : qhx;
        const uint32_t y_near = ny ? qhy : qly, y_far = ny ? qly : qhy;
        const uint32_t z_near = nz ? qhz : qlz, z_far = nz ? qlz : qhz;
        auto child = [&](auto jc) {
            constexpr int j = decltype(jc)::value;
            const float t0x = fmaf(u8f_axis<0, j>(x_near, magic), ax, ox), t1x = fmaf(u8f_axis<0, j>(x_far, magic), ax, ox);
            const float t0y = fmaf(u8f_axis<1, j>(y_near, magic), ay, oy), t1y = fmaf(u8f_axis<1, j>(y_far, magic), ay, oy);
            const float t0z = fmaf(u8f_axis<2, j>(z_near, magic), az, oz), t1z = fmaf(u8f_axis<2, j>(z_far, magic), az, oz);
            const float cmin = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, st.tmin));
            const float cmax = fminf(fminf(t1x, t1y), fminf(t1z, tmax_box));
            // every slot holds a real child or an inverted box: no validity mask (the shift uses the low 5 bits)
            if(cmin <= cmax) hitmask |= 1u << ((bit_index4 >> (8 * j)) & 31u);
        };
        child(std::integral_constant<int, 0>{}); child(std::integral_constant<int, 1>{});
        child(std::integral_constant<int, 2>{}); child(std::integral_constant<int, 3>{});
    }
    ngroup.x = __float_as_uint(n1.x);
    ngroup.y = (hitmask & 0xFF000000u) | (ew >> 24);
    tgroup.x = __float_as_uint(n1.y);
    tgroup.y = hitmask & 0x00FFFFFFu;
}

// hit.inst of a world-space query phase: the instance comes from the triangle record
#define CW_FLAT_INST 0xFFFFFFFEu

// The world-space ray enters the flat static scene: its leaves are triangles (in_blas), stored in world space
PT_D void cw_enter_flat(CwState& st, v3 world_dir)
{
    tri_preprocess(world_dir, st.axis, st.S);
    st.in_blas = true;
    st.cur_inst = CW_FLAT_INST;
}

// Query start: the subframe's dynamic instances (world-box test) go on the stack as one instance
// group. Flat static scene: the query then starts IN the static world (its triangles are stored in world
// space: no transform, no entry step), the dynamic group waits under an exit marker. Otherwise the static
// TLAS root is the first node group.
// PARKED: the caller's exit protocol keeps the world-space ray constants under the exit marker
// (wf_trace_cw_kernel); otherwise the marker alone (cw_pop_phase recomputes them).
template<bool PARKED, class Stack>
PT_D void cw_begin(const Scene& sc, CwState& st, Stack& stack, uint32_t subframe, v3 ro, v3 rd,
                   float tmin, float tmax, bool any)
{
    st.tmin = tmin; st.tmax = tmax; st.any = any; st.hit = false;
    stack.hit_clear();
    st.sp = 0; st.in_blas = false; st.cur_inst = 0; st.axis = 2; st.S = mk3(0, 0, 1);
    cw_set_space(st, ro, rd);
    const uint2 r = __ldg(sc.dyn_range + subframe);
    const uint32_t p = r.x, a = r.y & 0xFFFFFu, len = r.y >> 20;
    uint32_t mask = 0;
#if CW_DYN_UNION
    // one box around the whole set first: most bounce and shadow rays miss it and skip the loop
    if(p + len > 0u && box_hit(__ldg(sc.dyn_union + 2 * subframe), __ldg(sc.dyn_union + 2 * subframe + 1), ro, st.idir, tmin, tmax))
#endif
    for(uint32_t k = 0; k < p + len; ++k)
    {
        const uint32_t id = sc.n_static + (k < p ? k : a + (k - p));
        const WideInstance* wi = sc.winst + id;
        if(box_hit(__ldg(&wi->lo), __ldg(&wi->hi), ro, st.idir, tmin, tmax)) mask |= 1u << k;
    }
    if(mask) stack.set(st.sp++, make_uint2(0x80000000u, mask));
    st.tgroup = make_uint2(0u, 0u);
    if(sc.flat_root == 0xFFFFFFFFu)
    {
        st.ngroup = make_uint2(sc.cw_tlas_root, 0x80000000u);
        return;
    }
    if(mask && sc.dyn_first)
    {   // the per-frame instances first (the hero objects stand in front of the camera: their hits shorten the
        // ray before the static world is walked); the static world waits on the stack as a node group
        st.sp = 0;
        stack.set(st.sp++, make_uint2(sc.flat_root, 0x80000000u));
        st.tgroup = make_uint2(0x80000000u, mask);
        st.ngroup = make_uint2(0u, 0u);
        return;
    }
    if(mask)
    {
        if(PARKED)
        {
            stack.set(st.sp++, make_uint2(__float_as_uint(st.idir.x), __float_as_uint(st.idir.y)));
            stack.set(st.sp++, make_uint2(__float_as_uint(st.idir.z), st.oct_inv4 & 0xFFu));
        }
        stack.set(st.sp++, make_uint2(CW_MARK_X, 0u));
    }
    cw_enter_flat(st, rd);
    st.ngroup = make_uint2(sc.flat_root, 0x80000000u);
}

template<class World>
PT_D uint32_t cw_decode_instance(const Scene& sc, const World& world, uint32_t base, uint32_t bit)
{
    if(base & 0x80000000u)
    {   // dynamic instance k of the subframe
        const uint2 r = __ldg(sc.dyn_range + world.subframe());
        const uint32_t p = r.x, a = r.y & 0xFFFFFu, k = (base & 0x7FFFFFFFu) + bit;
        return sc.n_static + (k < p ? k : a + (k - p));
    }
    return __ldg(sc.cw_inst_index + base + bit);
}

// ray_query_enter_blas (ray_query.hh:153-182) on the compressed layout
template<class Stack, class World>
PT_D void cw_enter_instance(const Scene& sc, CwState& st, Stack& stack, const World& world, uint32_t id)
{
    st.cur_inst = id;
    const WideInstance* wi = sc.winst + id;
    const float4 r0 = __ldg(&wi->inv0), r1 = __ldg(&wi->inv1), r2 = __ldg(&wi->inv2);
    const uint32_t root = __ldg(&wi->cw_root);
    const v3 ro = world.origin(), rd = world.dir();
    const v3 o = mk3(r0.x * ro.x + r0.y * ro.y + r0.z * ro.z + r0.w,
                     r1.x * ro.x + r1.y * ro.y + r1.z * ro.z + r1.w,
                     r2.x * ro.x + r2.y * ro.y + r2.z * ro.z + r2.w);
    const v3 d = mk3(r0.x * rd.x + r0.y * rd.y + r0.z * rd.z,
                     r1.x * rd.x + r1.y * rd.y + r1.z * rd.z,
                     r2.x * rd.x + r2.y * rd.y + r2.z * rd.z);
    cw_set_space(st, o, d);
    tri_preprocess(d, st.axis, st.S);
    st.in_blas = true;
    stack.set(st.sp++, make_uint2(CW_MARK_X, 0u));
    st.ngroup = make_uint2(root, 0x80000000u);
    st.tgroup = make_uint2(0u, 0u);
}

// One triangle of the leaf group (ray_query_test_triangle, ray_query.hh:225-246)
template<class Stack>
PT_D void cw_test_triangle(const Scene& sc, CwState& st, Stack& stack, uint32_t tri)
{
    const float4* tp = sc.cwtris + 3 * (size_t)tri;
    const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
    float u, v, t; bool bf;
#ifdef WF_STATS
    st.n_tris++;
#endif
    const bool ok = tri_intersect(st.o, st.axis, st.S, mk3(a), mk3(b), mk3(c), u, v, t, bf);
    // The reference accepts strictly closer candidates, so among exactly tied hits (a ray through a
    // shared edge) the first one ITS traversal order finds wins (ray_query.hh:245,289). This kernel
    // tests triangles in an order that depends on warp scheduling, so ties are broken by the smallest
    // (instance, primitive) instead: deterministic and independent of traversal order.
    const uint32_t prim = __float_as_uint(a.w);
    // world-space triangles of the flat static scene name their instance (and whether its transform mirrors,
    // which swaps front and back); BLAS triangles belong to the instance the ray is in (their b.w is 0)
    const uint32_t bw = __float_as_uint(b.w);
    const uint32_t inst = st.cur_inst == CW_FLAT_INST ? (bw & 0x7FFFFFFFu) : st.cur_inst;
    if(!(ok && t > st.tmin)) return;
    bool accept = t < st.tmax;
    if(!accept && st.hit && t == st.tmax)
    {   // exact tie with the recorded hit
        const uint2 id = stack.hit_id();
        accept = inst < id.x || (inst == id.x && prim < (id.y & 0x7FFFFFFFu));
    }
    if(accept)
    {
        stack.hit_set(u, v, inst, prim | ((bf != ((bw >> 31) != 0u)) ? 0x80000000u : 0u));
        st.hit = true;
        st.tmax = t;
        if(st.any) { st.sp = 0; st.ngroup.y = 0u; st.tgroup.y = 0u; }
    }
}

// node phase: take the next child of the node group (highest bit = nearest in octant order), test it
template<bool TOP = false, class Stack>
PT_D void cw_node_phase(const Scene& sc, CwState& st, Stack& stack, const TopLevels& top = TopLevels{nullptr, 0u, 0u})
{
    if(st.ngroup.y > 0x00FFFFFFu)
    {
        const uint32_t hits_imask = st.ngroup.y;
        const uint32_t child_bit = 31u - (uint32_t)__clz(hits_imask);
        const uint32_t base = st.ngroup.x;
        st.ngroup.y &= ~(1u << child_bit);
        if(st.ngroup.y > 0x00FFFFFFu) stack.set(st.sp++, st.ngroup);
        const uint32_t slot = (child_bit - 24u) ^ (st.oct_inv4 & 0xFFu);
        const uint32_t rel = (uint32_t)__popc(hits_imask & ~(0xFFFFFFFFu << slot));
#ifdef WF_STATS
        st.n_nodes++;
#endif
        cw_intersect_node<TOP>(sc.cwnodes, base + rel, st, st.ngroup, st.tgroup, top, sc.cw_magic);
    }
    else
    {   // the entry popped last was a leaf group
        st.tgroup = st.ngroup;
        st.ngroup = make_uint2(0u, 0u);
    }
}

// instance phase (TLAS context): enter one instance of the leaf group, park everything else
template<class Stack, class World>
PT_D void cw_instance_phase(const Scene& sc, CwState& st, Stack& stack, const World& world)
{
    const uint32_t bit = 31u - (uint32_t)__clz(st.tgroup.y);
    st.tgroup.y &= ~(1u << bit);
    if(st.ngroup.y > 0x00FFFFFFu) stack.set(st.sp++, st.ngroup);
    if(st.tgroup.y) stack.set(st.sp++, st.tgroup);
    cw_enter_instance(sc, st, stack, world, cw_decode_instance(sc, world, st.tgroup.x, bit));
}

// pop phase; returns false when the query is complete
template<class Stack, class World>
PT_D bool cw_pop_phase(const Scene& sc, CwState& st, Stack& stack, const World& world)
{
    if(st.ngroup.y <= 0x00FFFFFFu)
    {
        if(st.sp == 0) return false;
        const uint2 e = stack.get(--st.sp);
        if(e.y == 0u)
        {   // exit marker: BLAS finished, back to world space
            st.in_blas = false;
            cw_set_space(st, world.origin(), world.dir());
            st.ngroup = make_uint2(0u, 0u);
        }
        else
        {
            st.ngroup = e;
            // a node group met in world space with a flat static scene is that scene's root (dyn_first)
            if(!st.in_blas && e.y > 0x00FFFFFFu && sc.flat_root != 0xFFFFFFFFu) cw_enter_flat(st, world.dir());
        }
    }
    return true;
}

template<bool ANY>
PT_D bool trace_cw(const Scene& sc, uint32_t subframe, v3 origin, v3 dir, float tmin, float tmax, Hit& hit,
                   uint32_t* census = nullptr)
{
    uint2 stack_mem[CW_STACK + 2];
    LocalStack stack{stack_mem};
    CwState st;
    const WorldRayRegs world{origin, dir, subframe};
    cw_begin<false>(sc, st, stack, subframe, origin, dir, tmin, tmax, ANY);
    for(;;)
    {
        // (a leaf group may already wait in tgroup: the per-frame instances of a dyn_first query start there)
        if(st.ngroup.y > 0x00FFFFFFu || st.tgroup.y == 0u) cw_node_phase(sc, st, stack);
        if(st.in_blas)
        {
            while(st.tgroup.y)
            {
                const uint32_t bit = 31u - (uint32_t)__clz(st.tgroup.y);
                st.tgroup.y &= ~(1u << bit);
                cw_test_triangle(sc, st, stack, st.tgroup.x + bit);
            }
        }
        else if(st.tgroup.y) cw_instance_phase(sc, st, stack, world);
        if(!cw_pop_phase(sc, st, stack, world)) break;
    }
    hit = st.result(stack);
#ifdef WF_STATS
    if(census) { census[0] = st.n_nodes; census[1] = st.n_tris; }
#endif
    return hit.t >= 0.0f;
}

struct CwTrav
{
    template<bool ANY>
    static PT_D bool trace(const Scene& sc, const SubframeCtx& sf, v3 o, v3 d, float tmin, float tmax,
                           Hit& h, TravCounters&)
    {
        return trace_cw<ANY>(sc, sf.index, o, d, tmin, tmax, h);
    }
};

} // namespace pt
