// Traversal of the REFERENCE acceleration layout as it is: stackless two-level walk over the
// direction-indexed link tables (device restatement of ray_query.hh:111-290). This is the
// cross-check / event-counting mode (ptgpu_set_option "traversal" = 1), not the fast path: it
// visits exactly the nodes the reference visits, so its counters are the N_* of SURVEY.md §8(d).
#pragma once
#include "pt_scene.cuh"

namespace pt {

struct Hit
{
    float t;            // < 0: miss
    float u, v;         // barycentrics (u, v, 1-u-v) (ray_query.hh:243)
    uint32_t inst;      // index into Scene::instances
    uint32_t prim;      // triangle index within the instance's mesh
    bool back_face;
};

struct TravCounters { uint32_t nodes, tris, blas; };

// ray_triangle_intersection_preprocess (math.hh:340-356)
PT_D void tri_preprocess(v3 dir, int& axis, v3& S)
{
    float ax = fabsf(dir.x), ay = fabsf(dir.y), az = fabsf(dir.z);
    v3 r = dir;
    axis = 2;
    if(ax > ay && ax > az) { axis = 0; r = mk3(dir.z, dir.y, dir.x); }
    else if(ay > az) { axis = 1; r = mk3(dir.x, dir.z, dir.y); }
    float inv = 1.0f / r.z;
    S = mk3(r.x * inv, r.y * inv, inv);
}

// ray_triangle_intersection (math.hh:358-401). Returns the reference's hit predicate; t,u,v are
// the components of *uvt.
// Every product and sum is written with an explicit-rounding intrinsic so that nvcc cannot contract
// them differently at different inlining sites: with free contraction a degenerate triangle
// (three identical vertices: the dragon mesh has some) gave cross products that were rounding
// residues instead of exact zeros, with a sign pattern that depended on which copy of this function
// ran — a spurious hit in one kernel and none in another, and run-to-run differences in the
// scheduled traversal kernel, which inlines the test at two sites. Written this way det is exactly
// zero for such triangles (rejected, as in the reference's strict-math semantics) and every copy
// of the function returns identical bits.
PT_D bool tri_intersect(v3 origin, int axis, v3 S, v3 p0, v3 p1, v3 p2,
                        float& u, float& v, float& t, bool& back_face)
{
    const float Ax0 = __fsub_rn(p0.x, origin.x), Ay0 = __fsub_rn(p0.y, origin.y), Az0 = __fsub_rn(p0.z, origin.z);
    const float Bx0 = __fsub_rn(p1.x, origin.x), By0 = __fsub_rn(p1.y, origin.y), Bz0 = __fsub_rn(p1.z, origin.z);
    const float Cx0 = __fsub_rn(p2.x, origin.x), Cy0 = __fsub_rn(p2.y, origin.y), Cz0 = __fsub_rn(p2.z, origin.z);
    // math.hh:376-385: axis 0 swaps x and z, axis 1 swaps y and z
    const bool a0 = axis == 0, a1 = axis == 1;
    const float Az = a0 ? Ax0 : a1 ? Ay0 : Az0, Bz = a0 ? Bx0 : a1 ? By0 : Bz0, Cz = a0 ? Cx0 : a1 ? Cy0 : Cz0;
    const float Ax = a0 ? Az0 : Ax0, Bx = a0 ? Bz0 : Bx0, Cx = a0 ? Cz0 : Cx0;
    const float Ay = a1 ? Az0 : Ay0, By = a1 ? Bz0 : By0, Cy = a1 ? Cz0 : Cy0;
    // x -= S.x * z; y -= S.y * z (math.hh:387-388)
    const float xa = __fmaf_rn(-S.x, Az, Ax), xb = __fmaf_rn(-S.x, Bz, Bx), xc = __fmaf_rn(-S.x, Cz, Cx);
    const float ya = __fmaf_rn(-S.y, Az, Ay), yb = __fmaf_rn(-S.y, Bz, By), yc = __fmaf_rn(-S.y, Cz, Cy);
    // uvw = cross(y, x) (math.hh:390): both products rounded, then subtracted
    const float U = __fsub_rn(__fmul_rn(yb, xc), __fmul_rn(yc, xb));
    const float V = __fsub_rn(__fmul_rn(yc, xa), __fmul_rn(ya, xc));
    const float W = __fsub_rn(__fmul_rn(ya, xb), __fmul_rn(yb, xa));
    const float det = __fadd_rn(__fadd_rn(U, V), W);
    const float inv = __fdiv_rn(1.0f, det);
    u = __fmul_rn(U, inv);
    v = __fmul_rn(V, inv);
    // dot(uvw, S.z * z) * (1 / det) (math.hh:392)
    const float dz = __fmaf_rn(W, __fmul_rn(S.z, Cz), __fmaf_rn(V, __fmul_rn(S.z, Bz), __fmul_rn(U, __fmul_rn(S.z, Az))));
    t = __fmul_rn(dz, inv);
    back_face = (det < 0.0f) != ((S.z < 0.0f) != (axis != 2)); // math.hh:393-395
    return det != 0.0f && t >= 0.0f &&
        ((U >= 0.0f && V >= 0.0f && W >= 0.0f) || (U <= 0.0f && V <= 0.0f && W <= 0.0f));
}

PT_D v3 safe_inv_dir(v3 d)
{   // ray_query.hh:130-133: 1/d, and +inf (1e40 narrowed to float) for a zero component
    return mk3(d.x == 0.0f ? __int_as_float(0x7f800000) : 1.0f / d.x,
               d.y == 0.0f ? __int_as_float(0x7f800000) : 1.0f / d.y,
               d.z == 0.0f ? __int_as_float(0x7f800000) : 1.0f / d.z);
}

PT_D uint32_t octant_of(v3 d)
{   // ray_query.hh:135-138
    return (d.x > 0.0f ? 1u : 0u) | (d.y > 0.0f ? 2u : 0u) | (d.z > 0.0f ? 4u : 0u);
}

// slab test of ray_query_traverse (ray_query.hh:195-207)
PT_D bool slab_hit(const float2* nodes, uint32_t idx, v3 origin, v3 inv_dir, float tmin, float tmax)
{
    const float2* n = nodes + 3u * idx;
    float2 a = __ldg(n), b = __ldg(n + 1), c = __ldg(n + 2); // (minx,miny) (minz,maxx) (maxy,maxz)
    float t0x = (a.x - origin.x) * inv_dir.x, t1x = (b.y - origin.x) * inv_dir.x;
    float t0y = (a.y - origin.y) * inv_dir.y, t1y = (c.x - origin.y) * inv_dir.y;
    float t0z = (b.x - origin.z) * inv_dir.z, t1z = (c.y - origin.z) * inv_dir.z;
    float near = fmaxf(fminf(t0x, t1x), fmaxf(fminf(t0y, t1y), fminf(t0z, t1z)));
    float far = fminf(fmaxf(t0x, t1x), fminf(fmaxf(t0y, t1y), fmaxf(t0z, t1z)));
    return near <= far && far > tmin && near < tmax;
}

// Closest hit (ANY = false: trace_ray's proceed/confirm loop, path_tracer.hh:342-349) or first
// candidate (ANY = true: trace_shadow_ray, path_tracer.hh:415-427).
template<bool ANY, bool COUNT>
PT_D bool trace_links(const Scene& sc, uint32_t tlas_count, uint32_t tlas_offset,
                      v3 origin, v3 dir, float tmin, float tmax, Hit& hit, TravCounters& cnt)
{
    hit.t = -1.0f; hit.u = 0.0f; hit.v = 0.0f; hit.inst = 0xFFFFFFFFu; hit.prim = 0; hit.back_face = false;
    const v3 inv_dir = safe_inv_dir(dir);
    const uint32_t tlink = tlas_offset * 8u + octant_of(dir) * tlas_count;
    uint32_t tnode = 0;
    while(tnode < tlas_count)
    {
        if(COUNT) cnt.nodes++;
        uint2 link = __ldg(sc.ref_links + tlink + tnode);
        if(!slab_hit(sc.ref_nodes, tlas_offset + tnode, origin, inv_dir, tmin, tmax)) { tnode = link.y; continue; }
        if(!(link.x & 0x80000000u)) { tnode = link.x; continue; }
        tnode = link.y;
        const uint32_t inst_id = link.x & 0x7FFFFFFFu;

        // ray_query_enter_blas (ray_query.hh:153-182)
        if(COUNT) cnt.blas++;
        const RefInstance* in = sc.instances + inst_id;
        const uint4 h0 = __ldg(reinterpret_cast<const uint4*>(in));       // blas{count,offset}, vertex_count, triangle_count
        const uint2 h1 = __ldg(reinterpret_cast<const uint2*>(in) + 2);   // index_offset, base_vertex
        const float4 i0 = __ldg(&in->inv_transform[0]), i1 = __ldg(&in->inv_transform[1]);
        const float4 i2 = __ldg(&in->inv_transform[2]), i3 = __ldg(&in->inv_transform[3]);
        // mul_m4v4(inv, (o,1)): columns i0..i3
        v3 o = mk3(i0.x * origin.x + i1.x * origin.y + i2.x * origin.z + i3.x,
                   i0.y * origin.x + i1.y * origin.y + i2.y * origin.z + i3.y,
                   i0.z * origin.x + i1.z * origin.y + i2.z * origin.z + i3.z);
        v3 d = mk3(i0.x * dir.x + i1.x * dir.y + i2.x * dir.z,
                   i0.y * dir.x + i1.y * dir.y + i2.y * dir.z,
                   i0.z * dir.x + i1.z * dir.y + i2.z * dir.z);
        const v3 binv = safe_inv_dir(d);
        const uint32_t bcount = h0.x, boffset = h0.y;
        const uint32_t blink = boffset * 8u + octant_of(d) * bcount;
        int axis; v3 S;
        tri_preprocess(d, axis, S);

        uint32_t bnode = 0;
        while(bnode < bcount)
        {
            if(COUNT) cnt.nodes++;
            uint2 bl = __ldg(sc.ref_links + blink + bnode);
            if(!slab_hit(sc.ref_nodes, boffset + bnode, o, binv, tmin, tmax)) { bnode = bl.y; continue; }
            if(!(bl.x & 0x80000000u)) { bnode = bl.x; continue; }
            bnode = bl.y;
            const uint32_t prim = bl.x & 0x7FFFFFFFu;
            // ray_query_test_triangle (ray_query.hh:225-246)
            if(COUNT) cnt.tris++;
            const uint32_t* ip = sc.indices + h1.x + prim * 3u;
            uint32_t k0 = __ldg(ip), k1 = __ldg(ip + 1), k2 = __ldg(ip + 2);
            v3 p0 = mk3(__ldg(sc.pos + h1.y + k0)), p1 = mk3(__ldg(sc.pos + h1.y + k1)), p2 = mk3(__ldg(sc.pos + h1.y + k2));
            float u, v, t; bool bf;
            bool ok = tri_intersect(o, axis, S, p0, p1, p2, u, v, t, bf);
            if(ok && t < tmax && t > tmin)
            {
                hit.t = t; hit.u = u; hit.v = v; hit.inst = inst_id; hit.prim = prim; hit.back_face = bf;
                if(ANY) return true;
                tmax = t; // ray_query_confirm (ray_query.hh:289)
            }
        }
    }
    return hit.t >= 0.0f && !ANY ? true : false;
}

} // namespace pt
