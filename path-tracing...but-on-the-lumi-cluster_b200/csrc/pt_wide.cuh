// Two-level traversal of the GPU layout: 4-wide nodes (128 B, SoA child boxes, eight LDG.128),
// multi-triangle leaves with pre-gathered vertices (three LDG.128 per triangle), a static TLAS
// plus the subframe's few dynamic instances tested directly. Replaces ray_query.hh:111-290 on the
// fast path; keeps the reference's triangle test (math.hh:340-401) and its acceptance rules
// (strict tmin < t < tmax, node test near <= far && far > tmin && near < tmax) so that edge
// behaviour matches. Closest hits do not depend on BVH topology.
#pragma once
#include "pt_scene.cuh"
#include "pt_trav_links.cuh"
#include "bvh_wide.hh"

namespace pt {

#define PT_EXIT_MARK 0xFFFFFFFEu
#define PT_EMPTY 0xFFFFFFFFu

struct WideRay
{
    v3 o, d, inv;     // current space (world or object)
    int axis; v3 S;   // triangle-test preprocess of d (object space)
};

// Slab test of four children at once; returns sort keys: float bits of max(near, tmin) with the
// slot in the two low bits, PT_EMPTY for a miss. near >= 0 so the bits order like the floats.
PT_D void test4(const WideNode* __restrict__ n, v3 o, v3 inv, float tmin, float tmax, uint32_t key[4], uint4& child)
{
    const float4 lox = __ldg(&n->lox), loy = __ldg(&n->loy), loz = __ldg(&n->loz);
    const float4 hix = __ldg(&n->hix), hiy = __ldg(&n->hiy), hiz = __ldg(&n->hiz);
    child = __ldg(&n->child);
    const float lx[4] = {lox.x, lox.y, lox.z, lox.w}, ly[4] = {loy.x, loy.y, loy.z, loy.w}, lz[4] = {loz.x, loz.y, loz.z, loz.w};
    const float hx[4] = {hix.x, hix.y, hix.z, hix.w}, hy[4] = {hiy.x, hiy.y, hiy.z, hiy.w}, hz[4] = {hiz.x, hiz.y, hiz.z, hiz.w};
    const uint32_t ch[4] = {child.x, child.y, child.z, child.w};
    #pragma unroll
    for(int i = 0; i < 4; ++i)
    {
        float t0x = (lx[i] - o.x) * inv.x, t1x = (hx[i] - o.x) * inv.x;
        float t0y = (ly[i] - o.y) * inv.y, t1y = (hy[i] - o.y) * inv.y;
        float t0z = (lz[i] - o.z) * inv.z, t1z = (hz[i] - o.z) * inv.z;
        float near = fmaxf(fminf(t0x, t1x), fmaxf(fminf(t0y, t1y), fminf(t0z, t1z)));
        float far = fminf(fmaxf(t0x, t1x), fminf(fmaxf(t0y, t1y), fmaxf(t0z, t1z)));
        bool hit = near <= far && far > tmin && near < tmax && ch[i] != PT_EMPTY;
        key[i] = hit ? ((__float_as_uint(fmaxf(near, 0.0f)) & ~3u) | (uint32_t)i) : PT_EMPTY;
    }
}

PT_D void cswap(uint32_t& a, uint32_t& b) { uint32_t lo = min(a, b), hi = max(a, b); a = lo; b = hi; }

PT_D uint32_t pick_child(const uint4& c, uint32_t slot)
{
    return slot == 0 ? c.x : slot == 1 ? c.y : slot == 2 ? c.z : c.w;
}

// Slab test of one box (dynamic instances)
PT_D bool box_hit(float4 lo, float4 hi, v3 o, v3 inv, float tmin, float tmax)
{
    float t0x = (lo.x - o.x) * inv.x, t1x = (hi.x - o.x) * inv.x;
    float t0y = (lo.y - o.y) * inv.y, t1y = (hi.y - o.y) * inv.y;
    float t0z = (lo.z - o.z) * inv.z, t1z = (hi.z - o.z) * inv.z;
    float near = fmaxf(fminf(t0x, t1x), fmaxf(fminf(t0y, t1y), fminf(t0z, t1z)));
    float far = fminf(fmaxf(t0x, t1x), fminf(fmaxf(t0y, t1y), fmaxf(t0z, t1z)));
    return near <= far && far > tmin && near < tmax;
}

template<bool ANY>
PT_D bool trace_wide(const Scene& sc, uint32_t subframe, v3 origin, v3 dir, float tmin, float tmax, Hit& hit)
{
    hit.t = -1.0f; hit.u = 0.0f; hit.v = 0.0f; hit.inst = 0xFFFFFFFFu; hit.prim = 0; hit.back_face = false;

    uint32_t stack[WIDE_STACK];
    int sp = 0;
    const v3 winv = safe_inv_dir(dir);

    // the subframe's dynamic instances: ids n_static + [0,p) and n_static + [a, a+len)
    {
        const uint2 r = __ldg(sc.dyn_range + subframe);
        const uint32_t p = r.x, a = r.y & 0xFFFFFu, len = r.y >> 20;
        for(uint32_t k = 0; k < p + len; ++k)
        {
            const uint32_t id = sc.n_static + (k < p ? k : a + (k - p));
            const WideInstance* wi = sc.winst + id;
            if(box_hit(__ldg(&wi->lo), __ldg(&wi->hi), origin, winv, tmin, tmax))
                stack[sp++] = 0x80000000u | id;
        }
    }

    v3 o = origin, d = dir, inv = winv;
    int axis = 2; v3 S = mk3(0, 0, 1);
    bool in_blas = false;
    const WideNode* nodes = sc.wtlas;
    const float4* tris = sc.wtris;
    uint32_t cur_inst = 0;
    uint32_t cur = 0; // TLAS root

    for(;;)
    {
        if(cur == PT_EMPTY)
        {
            if(sp == 0) break;
            cur = stack[--sp];
        }
        if(cur == PT_EXIT_MARK)
        {   // BLAS finished: back to world space
            in_blas = false; nodes = sc.wtlas; o = origin; d = dir; inv = winv;
            cur = PT_EMPTY;
            continue;
        }
        if(!(cur & 0x80000000u))
        {   // inner node
            uint32_t key[4]; uint4 child;
            test4(nodes + cur, o, inv, tmin, tmax, key, child);
            // sort ascending by entry distance (5-comparator network)
            cswap(key[0], key[1]); cswap(key[2], key[3]); cswap(key[0], key[2]); cswap(key[1], key[3]); cswap(key[1], key[2]);
            if(key[3] != PT_EMPTY) stack[sp++] = pick_child(child, key[3] & 3u);
            if(key[2] != PT_EMPTY) stack[sp++] = pick_child(child, key[2] & 3u);
            if(key[1] != PT_EMPTY) stack[sp++] = pick_child(child, key[1] & 3u);
            cur = key[0] != PT_EMPTY ? pick_child(child, key[0] & 3u) : PT_EMPTY;
            continue;
        }
        if(!in_blas)
        {   // TLAS leaf: enter the instance (ray_query_enter_blas, ray_query.hh:153-182)
            cur_inst = cur & 0x7FFFFFFFu;
            const WideInstance* wi = sc.winst + cur_inst;
            const float4 r0 = __ldg(&wi->inv0), r1 = __ldg(&wi->inv1), r2 = __ldg(&wi->inv2);
            const uint32_t b = __ldg(&wi->blas);
            o = mk3(r0.x * origin.x + r0.y * origin.y + r0.z * origin.z + r0.w,
                    r1.x * origin.x + r1.y * origin.y + r1.z * origin.z + r1.w,
                    r2.x * origin.x + r2.y * origin.y + r2.z * origin.z + r2.w);
            d = mk3(r0.x * dir.x + r0.y * dir.y + r0.z * dir.z,
                    r1.x * dir.x + r1.y * dir.y + r1.z * dir.z,
                    r2.x * dir.x + r2.y * dir.y + r2.z * dir.z);
            inv = safe_inv_dir(d);
            tri_preprocess(d, axis, S);
            const uint2 bo = __ldg(reinterpret_cast<const uint2*>(sc.wblas + b)); // node_offset, tri_offset
            nodes = sc.wnodes + bo.x;
            tris = sc.wtris + 3 * (size_t)bo.y;
            in_blas = true;
            stack[sp++] = PT_EXIT_MARK;
            cur = 0; // BLAS root
            continue;
        }
        {   // BLAS leaf: up to WIDE_LEAF_MAX triangles (ray_query_test_triangle, ray_query.hh:225-246)
            const uint32_t first = cur & 0x07FFFFFFu, count = ((cur >> 27) & 0xFu) + 1u;
            const float4* tp = tris + 3 * (size_t)first;
            for(uint32_t k = 0; k < count; ++k, tp += 3)
            {
                const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
                float u, v, t; bool bf;
                bool ok = tri_intersect(o, axis, S, mk3(a), mk3(b), mk3(c), u, v, t, bf);
                if(ok && t < tmax && t > tmin)
                {
                    hit.t = t; hit.u = u; hit.v = v; hit.inst = cur_inst; hit.prim = __float_as_uint(a.w); hit.back_face = bf;
                    if(ANY) return true;
                    tmax = t;
                }
            }
            cur = PT_EMPTY;
        }
    }
    return !ANY && hit.t >= 0.0f;
}

struct WideTrav
{
    template<bool ANY>
    static PT_D bool trace(const Scene& sc, const SubframeCtx& sf, v3 o, v3 d, float tmin, float tmax,
                           Hit& h, TravCounters&)
    {
        return trace_wide<ANY>(sc, sf.index, o, d, tmin, tmax, h);
    }
};

} // namespace pt
