// Sub-function evaluator (ptgpu_debug_eval): ONE device function of the path per item, on caller-supplied
// inputs, so that each of them can be compared directly with the reference's own function of the same name
// (oracle/ref_harness.cc: ref_eval) instead of only through whole paths and images. Item i reads
// in[24*i ..] and writes out[32*i ..]; function codes and layouts: PTGPU_FN_* in include/ptgpu.h.
// uint32 values travel as the bits of a float.
#pragma once
#include "pt_kernels.cuh"
#include "pt_cwbvh.cuh"

namespace pt {

template<class Trav>
__global__ void debug_eval_kernel(Scene sc, int fn, const float* __restrict__ in_all, size_t n, float* __restrict__ out_all)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const float* in = in_all + 24 * i;
    float* out = out_all + 32 * i;
    for(int k = 0; k < 32; ++k) out[k] = 0.0f;
    switch(fn)
    {
    case PTGPU_FN_RAND4: {            // generate_uniform_random4 (math.hh:475-485)
        rng4 s = {__float_as_uint(in[0]), __float_as_uint(in[1]), __float_as_uint(in[2]), __float_as_uint(in[3])};
        const float4 f = rand4(s);
        out[0] = __uint_as_float(s.x); out[1] = __uint_as_float(s.y); out[2] = __uint_as_float(s.z); out[3] = __uint_as_float(s.w);
        out[4] = f.x; out[5] = f.y; out[6] = f.z; out[7] = f.w;
        break; }
    case PTGPU_FN_FILM_OFFSET: {      // sample_gaussian_weighted_disk(u, 0.4) (path_tracer.hh:19-25, :665)
        const v2 o = sample_gaussian_disk(in[0], in[1], 0.4f);
        out[0] = o.x; out[1] = o.y;
        break; }
    case PTGPU_FN_CAMERA_RAY: {       // get_camera_ray (path_tracer.hh:429-450)
        const uint32_t sub = (uint32_t)in[4];
        if(sub >= sc.n_subframes) break;
        v3 d, o;
        camera_ray(sc, sc.subframes + sub, in[0], in[1], in[2], in[3], d, o);
        out[0] = d.x; out[1] = d.y; out[2] = d.z; out[3] = o.x; out[4] = o.y; out[5] = o.z;
        break; }
    case PTGPU_FN_GGX_VNDF: {         // sample_ggx_vndf (path_tracer.hh:67-83)
        const v3 h = sample_ggx_vndf(mk3(in[0], in[1], in[2]), in[3], in[4], in[5]);
        out[0] = h.x; out[1] = h.y; out[2] = h.z;
        break; }
    case PTGPU_FN_BSDF: {             // bsdf (path_tracer.hh:184-222)
        Surface s; s.albedo = mk3(in[6], in[7], in[8]); s.roughness = in[9]; s.metallic = in[10]; s.transmission = in[11]; s.eta = in[12];
        float pdf = 0.0f;
        const v3 a = bsdf_eval(mk3(in[0], in[1], in[2]), mk3(in[3], in[4], in[5]), s, pdf);
        out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = pdf;
        break; }
    case PTGPU_FN_SAMPLE_BSDF: {      // sample_bsdf (path_tracer.hh:224-296)
        Surface s; s.albedo = mk3(in[6], in[7], in[8]); s.roughness = in[9]; s.metallic = in[10]; s.transmission = in[11]; s.eta = in[12];
        v3 dir, att; float pdf = 0.0f;
        bsdf_sample(in[0], in[1], in[2], mk3(in[3], in[4], in[5]), s, dir, att, pdf);
        out[0] = dir.x; out[1] = dir.y; out[2] = dir.z; out[3] = att.x; out[4] = att.y; out[5] = att.z; out[6] = pdf;
        break; }
    case PTGPU_FN_SKY_ATTENUATION: {  // nishita_atmosphere_attenuation as nee_branch calls it (path_tracer.hh:456-497, :615-617)
        const v3 a = sky_attenuation(in[0], mk3(in[1], in[2], in[3]), mk3(in[4], in[5], in[6]));
        out[0] = a.x; out[1] = a.y; out[2] = a.z;
        break; }
    case PTGPU_FN_SKY_SCATTERING: {   // nishita_atmosphere_scattering (path_tracer.hh:499-588)
        rng4 s = {__float_as_uint(in[0]), __float_as_uint(in[1]), __float_as_uint(in[2]), __float_as_uint(in[3])};
        Light l; l.dir = mk3(in[4], in[5], in[6]); l.color = mk3(in[7], in[8], in[9]); l.cos_solid_angle = in[10];
        v3 att, scat;
        sky_scattering(s, l, mk3(in[11], in[12], in[13]), mk3(in[14], in[15], in[16]), in[17], att, scat);
        out[0] = att.x; out[1] = att.y; out[2] = att.z; out[3] = scat.x; out[4] = scat.y; out[5] = scat.z;
        out[6] = __uint_as_float(s.x); out[7] = __uint_as_float(s.y); out[8] = __uint_as_float(s.z); out[9] = __uint_as_float(s.w);
        break; }
    case PTGPU_FN_SAMPLE_CONE: {      // sample_cone (path_tracer.hh:40-48)
        const v3 d = sample_cone(mk3(in[0], in[1], in[2]), in[3], in[4], in[5]);
        out[0] = d.x; out[1] = d.y; out[2] = d.z;
        break; }
    case PTGPU_FN_SHADOW_RAY: {       // trace_shadow_ray (path_tracer.hh:415-427)
        const uint32_t sub = (uint32_t)in[8];
        if(sub >= sc.n_subframes) break;
        SubframeCtx sf; const RefSubframe* rsf;
        load_subframe(sc, (int)(sub * sc.samples_per_subframe), sf, rsf);
        Hit h; TravCounters tc = {0, 0, 0};
        out[0] = Trav::template trace<true>(sc, sf, mk3(in[0], in[1], in[2]), mk3(in[3], in[4], in[5]), in[6], in[7], h, tc) ? 1.0f : 0.0f;
        break; }
    case PTGPU_FN_TRACE_RAY: {        // trace_ray -> hit_info (path_tracer.hh:340-412)
        const uint32_t sub = (uint32_t)in[7];
        if(sub >= sc.n_subframes) break;
        SubframeCtx sf; const RefSubframe* rsf;
        load_subframe(sc, (int)(sub * sc.samples_per_subframe), sf, rsf);
        Hit h; TravCounters tc = {0, 0, 0};
        const v3 o = mk3(in[0], in[1], in[2]), d = mk3(in[3], in[4], in[5]);
        Trav::template trace<false>(sc, sf, o, d, in[6], PT_MAX_RAY_DIST, h, tc);
        HitInfo hi;
        shade_hit(sc, sf.light, h, o, d, hi);
        out[0] = hi.thit;
        out[13] = hi.s.albedo.x; out[14] = hi.s.albedo.y; out[15] = hi.s.albedo.z;
        out[18] = hi.emission; out[21] = hi.nee_pdf;
        if(hi.thit >= 0.0f)
        {
            out[1] = hi.pos.x; out[2] = hi.pos.y; out[3] = hi.pos.z;
            out[4] = hi.tbn.c0.x; out[5] = hi.tbn.c0.y; out[6] = hi.tbn.c0.z;
            out[7] = hi.tbn.c1.x; out[8] = hi.tbn.c1.y; out[9] = hi.tbn.c1.z;
            out[10] = hi.tbn.c2.x; out[11] = hi.tbn.c2.y; out[12] = hi.tbn.c2.z;
            out[16] = hi.s.roughness; out[17] = hi.s.metallic; out[19] = hi.s.transmission; out[20] = hi.s.eta;
        }
        break; }
    default: break;
    }
}

} // namespace pt
