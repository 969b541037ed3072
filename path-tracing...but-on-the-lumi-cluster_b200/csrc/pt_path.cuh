// One path sample: device restatement of path_trace_pixel (path_tracer.hh:637-741) with its
// helpers trace_ray (:340-412), get_camera_ray (:429-450) and nee_branch (:594-620).
// The RNG draw order of SURVEY.md Appendix A is kept exactly; `Trav` supplies the ray queries.
#pragma once
#include "pt_scene.cuh"
#include "pt_trav_links.cuh"

namespace pt {

struct HitInfo
{
    float thit;
    v3 pos;
    m3 tbn;
    Surface s;
    float emission;
    float nee_pdf;
};

struct PathCounters
{
    uint32_t paths, rays, nodes, tris, blas, bounces, shadow, sky, att, hits, misses;
};

// Event sink: a no-op unless COUNT.
template<bool COUNT> struct Events;
template<> struct Events<false>
{
    PT_D void ray(const TravCounters&) {}
    PT_D void shadow(const TravCounters&) {}
    PT_D void bounce() {} PT_D void sky() {} PT_D void att() {} PT_D void hit() {} PT_D void miss() {}
    PT_D void path() {}
};
template<> struct Events<true>
{
    PathCounters c = {};
    PT_D void ray(const TravCounters& t) { c.rays++; c.nodes += t.nodes; c.tris += t.tris; c.blas += t.blas; }
    PT_D void shadow(const TravCounters& t) { ray(t); c.shadow++; }
    PT_D void bounce() { c.bounces++; } PT_D void sky() { c.sky++; } PT_D void att() { c.att++; }
    PT_D void hit() { c.hits++; } PT_D void miss() { c.misses++; }
    PT_D void path() { c.paths++; }
};

// get_camera_ray (path_tracer.hh:429-450)
PT_D void camera_ray(const Scene& sc, const RefSubframe* sf, float uz, float uw, float cx, float cy,
                     v3& dir, v3& origin)
{
    const float4 o0 = __ldg(&sf->orient[0]), o1 = __ldg(&sf->orient[1]), o2 = __ldg(&sf->orient[2]);
    const float4 cp = __ldg(&sf->position);
    const float4 cf = __ldg(reinterpret_cast<const float4*>(&sf->aspect_ratio)); // aspect, inv_focal, focal_dist, ap_angle
    const int32_t polygon = __ldg(&sf->aperture_polygon);
    const float radius = __ldg(&sf->aperture_radius);

    float uvx = cx / (float)sc.width * 2.0f - 1.0f;
    float uvy = cy / (float)sc.height * 2.0f - 1.0f;
    uvx *= cf.x;
    uvy = -uvy;
    float ax = 0.0f, ay = 0.0f;
    if(polygon > 3 && radius != 0.0f)
    {   // a zero radius multiplies the polygon sample to exactly zero in the reference (:437)
        v2 a = sample_regular_polygon(uz, uw, cf.w, (uint32_t)polygon);
        ax = a.x * radius; ay = a.y * radius;
    }
    v3 o = mk3(ax, ay, 0.0f);
    v3 d = mk3(uvx * cf.y, uvy * cf.y, -1.0f) * cf.z;
    d = normalize(d - o);
    m3 m; m.c0 = mk3(o0); m.c1 = mk3(o1); m.c2 = mk3(o2);
    dir = mul_m3v3(m, d);
    origin = mul_m3v3(m, o) + mk3(cp);
}

// The part of trace_ray after the query (path_tracer.hh:351-411)
PT_D void shade_hit(const Scene& sc, const Light& light, const Hit& h, v3 origin, v3 dir, HitInfo& hi)
{
    hi.thit = h.t;
    hi.nee_pdf = 0.0f;
    if(h.t < 0.0f)
    {   // sky: sun disk only, the atmosphere is a separate volumetric pass (:355-366)
        float visible = dot(light.dir, dir) > light.cos_solid_angle ? 1.0f : 0.0f;
        hi.nee_pdf = visible / (PT_TWO_PI * (1.0f - light.cos_solid_angle));
        hi.s.albedo = (visible * (hi.nee_pdf == 0.0f ? 1.0f : hi.nee_pdf)) * light.color;
        hi.emission = 1.0f;
        return;
    }
    hi.pos = origin + dir * h.t;
    const RefInstance* in = sc.instances + h.inst;
    const uint2 mo = __ldg(reinterpret_cast<const uint2*>(in) + 2); // index_offset, base_vertex
    const float4 t0 = __ldg(&in->transform[0]), t1 = __ldg(&in->transform[1]), t2 = __ldg(&in->transform[2]);
    const float bx = h.u, by = h.v, bz = 1.0f - h.u - h.v;
    // the reference gathers 3 indices, then 9 attribute vectors through them (path_tracer.hh:375-409): two
    // dependent levels of scattered 16-byte loads. shade_tris holds the same nine vectors per triangle,
    // contiguous (144 B), one level after the instance record.
    const float4* rec = sc.shade_tris + 9 * (size_t)(mo.x / 3u + h.prim);
    const float4 n0 = __ldg(rec), n1 = __ldg(rec + 1), n2 = __ldg(rec + 2);
    const float4 a0 = __ldg(rec + 3), a1 = __ldg(rec + 4), a2 = __ldg(rec + 5);
    const float4 m0 = __ldg(rec + 6), m1 = __ldg(rec + 7), m2 = __ldg(rec + 8);
    v3 n = mk3(n0) * bx + mk3(n1) * by + mk3(n2) * bz;
    m3 rot; rot.c0 = mk3(t0); rot.c1 = mk3(t1); rot.c2 = mk3(t2);
    n = normalize(mul_m3v3(rot, n)); // forward 3x3, not the inverse transpose (:371,392)
    if(h.back_face) { hi.s.eta = 1.5f; n = -n; }
    else hi.s.eta = 1.0f / 1.5f;
    hi.tbn = tangent_space(n);
    hi.s.albedo = mk3(a0) * bx + mk3(a1) * by + mk3(a2) * bz;
    float mx = m0.x * bx + m1.x * by + m2.x * bz;
    hi.s.roughness = mx * mx;
    hi.s.metallic = m0.y * bx + m1.y * by + m2.y * bz;
    hi.s.transmission = m0.z * bx + m1.z * by + m2.z * bz;
    hi.emission = m0.w * bx + m1.w * by + m2.w * bz;
}

// Traversal policy walking the reference link tables.
template<bool COUNT>
struct LinksTrav
{
    template<bool ANY>
    static PT_D bool trace(const Scene& sc, const SubframeCtx& sf, v3 o, v3 d, float tmin, float tmax,
                           Hit& h, TravCounters& tc)
    {
        return trace_links<ANY, COUNT>(sc, sf.tlas_count, sf.tlas_offset, o, d, tmin, tmax, h, tc);
    }
};

PT_D void load_subframe(const Scene& sc, int sample_index, SubframeCtx& sf, const RefSubframe*& rsf)
{
    // path_tracer.hh:655-657
    uint32_t si = sample_index < 0 ? 0u : (uint32_t)sample_index / (uint32_t)sc.samples_per_subframe;
    rsf = sc.subframes + si;
    sf.index = si;
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(rsf));
    sf.tlas_count = t.x; sf.tlas_offset = t.y;
    sf.light.dir = mk3(__ldg(&rsf->light_dir));
    sf.light.color = mk3(__ldg(&rsf->light_color));
    sf.light.cos_solid_angle = __ldg(&rsf->cos_solid_angle);
}

template<class Trav, bool COUNT>
PT_D v3 path_trace_sample(const Scene& sc, uint32_t px, uint32_t py, int sample_index, Events<COUNT>& ev)
{
    SubframeCtx sf;
    const RefSubframe* rsf;
    load_subframe(sc, sample_index, sf, rsf);
    ev.path();

    rng4 seed = {px, py, (uint32_t)sample_index, sc.student_id};
    pcg4d(seed); // :660

    float4 u = rand4(seed);
    v2 film = sample_gaussian_disk(u.x, u.y, 0.4f);
    v3 ray_d, ray_o;
    camera_ray(sc, rsf, u.z, u.w, (float)px + (film.x + 0.5f), (float)py + (film.y + 0.5f), ray_d, ray_o);

    Hit h;
    HitInfo info;
    TravCounters tc = {0, 0, 0};
    Trav::template trace<false>(sc, sf, ray_o, ray_d, 0.0f, PT_MAX_RAY_DIST, h, tc);
    ev.ray(tc);
    shade_hit(sc, sf.light, h, ray_o, ray_d, info);
    if(h.t < 0.0f) ev.miss(); else ev.hit();

    v3 attenuation = mk3(1, 1, 1);
    v3 in_scatter;
    if(sky_scattering(seed, sf.light, ray_o, ray_d, info.thit, attenuation, in_scatter)) ev.sky();
    v3 contribution = in_scatter + attenuation * info.s.albedo * info.emission;

    float regularization = 1.0f;
    for(int bounce = 0; bounce < sc.max_bounces && info.thit > 0.0f; ++bounce)
    {
        ev.bounce();
        v3 view = mul_v3m3(-ray_d, info.tbn);
        if(view.z < 1e-7f) view.z = fmaxf(view.z, 1e-7f);
        view = normalize(view);

        {   // nee_branch (:594-620)
            float4 un = rand4(seed);
            v3 light_dir = sample_cone(sf.light.dir, sf.light.cos_solid_angle, un.x, un.y);
            float nee_pdf = 1.0f / (PT_TWO_PI * (1.0f - sf.light.cos_solid_angle));
            float bsdf_pdf = 0.0f;
            v3 color = bsdf_eval(mul_v3m3(light_dir, info.tbn), view, info.s, bsdf_pdf) * nee_pdf * sf.light.color;
            bool lit = !(color.x == 0.0f && color.y == 0.0f && color.z == 0.0f);
            if(lit)
            {
                Hit sh;
                TravCounters sc_cnt = {0, 0, 0};
                bool occluded = Trav::template trace<true>(sc, sf, info.pos, light_dir, PT_MIN_RAY_DIST, PT_MAX_RAY_DIST, sh, sc_cnt);
                ev.shadow(sc_cnt);
                if(!occluded)
                {
                    float mis_pdf = 1.0f;
                    if(sf.light.cos_solid_angle < 1.0f)
                        mis_pdf = (nee_pdf * nee_pdf + bsdf_pdf * bsdf_pdf) / nee_pdf;
                    ev.att();
                    color *= sky_attenuation(un.w, info.pos, light_dir);
                    contribution += attenuation * (color * (1.0f / mis_pdf));
                }
            }
        }

        float4 ub = rand4(seed);
        v3 tdir, bsdf_att;
        float bsdf_pdf;
        bsdf_sample(ub.x, ub.y, ub.z, view, info.s, tdir, bsdf_att, bsdf_pdf);

        ray_d = normalize(mul_m3v3(info.tbn, tdir));
        ray_o = info.pos;
        tc.nodes = tc.tris = tc.blas = 0;
        Trav::template trace<false>(sc, sf, ray_o, ray_d, PT_MIN_RAY_DIST, PT_MAX_RAY_DIST, h, tc);
        ev.ray(tc);
        shade_hit(sc, sf.light, h, ray_o, ray_d, info);
        if(h.t < 0.0f) ev.miss(); else ev.hit();

        float mis_pdf = bsdf_pdf < 0.0f ? -bsdf_pdf :
            (info.nee_pdf * info.nee_pdf + bsdf_pdf * bsdf_pdf) / bsdf_pdf;
        attenuation *= bsdf_att;

        v3 atmo_att, scat;
        if(sky_scattering(seed, sf.light, ray_o, ray_d, info.thit, atmo_att, scat)) ev.sky();
        contribution += attenuation * (scat + atmo_att * info.s.albedo * info.emission) * (1.0f / mis_pdf);
        attenuation *= atmo_att * (1.0f / fabsf(bsdf_pdf));

        // path-space regularisation (:735-737), applied to the NEXT hit's roughness
        if(bsdf_pdf > 0.0f)
            regularization *= fmaxf(1.0f - PT_REG_GAMMA / sqrtf(sqrtf(bsdf_pdf)), 0.0f);
        info.s.roughness = 1.0f - (1.0f - info.s.roughness) * regularization;
    }
    return contribution;
}

} // namespace pt
