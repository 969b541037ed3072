// Material model, sky ray-march and tonemap of the path tracer (device restatement of
// path_tracer.hh:89-296, 456-588, 753-771). Pure functions, no memory traffic.
#pragma once
#include "pt_math.cuh"

namespace pt {

// config.hh:30-42
#define PT_MIN_RAY_DIST 1e-4f
#define PT_MAX_RAY_DIST 1e9f
#define PT_REG_GAMMA 0.15f
#define PT_EARTH_RADIUS 6.3781e6f
#define PT_ATMO_PRIMARY 8
#define PT_ATMO_SECONDARY 4
#define PT_ATMO_HEIGHT 1.0e5f
#define PT_RAYLEIGH_X 5.8e-6f
#define PT_RAYLEIGH_Y 13.6e-6f
#define PT_RAYLEIGH_Z 33.1e-6f
#define PT_RAYLEIGH_SCALE 7994.0f
#define PT_MIE_COEF 4.0e-6f
#define PT_MIE_G 0.80f
#define PT_MIE_SCALE 1200.0f

struct Light { v3 dir; v3 color; float cos_solid_angle; };

// ---- BSDF (path_tracer.hh:89-296) -------------------------------------------------------------

PT_D float pow5(float x) { float x2 = x * x; return x2 * x2 * x; }

// fresnel_schlick_bidir_attenuated (path_tracer.hh:89-98)
PT_D float fresnel_att(float v_dot_h, float f0, float eta, float roughness)
{
    if(eta > 1.0f)
    {
        float sin_theta2 = eta * eta * (1.0f - v_dot_h * v_dot_h);
        if(sin_theta2 >= 1.0f) return 1.0f;
        v_dot_h = sqrtf(1.0f - sin_theta2);
    }
    return f0 + (fmaxf(1.0f - roughness, f0) - f0) * pow5(fmaxf(1.0f - v_dot_h, 0.0f));
}

// trowbridge_reitz_distribution (path_tracer.hh:105-110)
PT_D float ggx_d(float hdotn, float a)
{
    float a2 = a * a;
    float denom = hdotn * hdotn * (a2 - 1.0f) + 1.0f;
    return a2 / fmaxf(PT_PI * denom * denom, 1e-10f);
}

// trowbridge_reitz_masking_shadowing (path_tracer.hh:112-123)
PT_D float ggx_g2(float ldotn, float ldoth, float vdotn, float vdoth, float a)
{
    if(vdotn * vdoth < 0.0f) return 0.0f;
    if(ldotn * ldoth < 0.0f) return 0.0f;
    float a2 = a * a;
    return 0.5f / (fabsf(vdotn) * sqrtf(ldotn * ldotn - a2 * ldotn * ldotn + a2) +
                   fabsf(ldotn) * sqrtf(vdotn * vdotn - a2 * vdotn * vdotn + a2));
}

// trowbridge_reitz_masking (path_tracer.hh:125-129)
PT_D float ggx_g1(float vdotn, float vdoth, float a)
{
    if(vdotn * vdoth < 0.0f) return 0.0f;
    return 2.0f * vdotn / (vdotn + sqrtf(vdotn * vdotn * (1.0f - a * a) + a * a));
}

struct Surface
{
    v3 albedo;
    float roughness, metallic, transmission, eta;
};

// bsdf_core (path_tracer.hh:131-181)
PT_D v3 bsdf_core(v3 light, v3 h, v3 view, const Surface& s, float f0, float distribution,
                  float& reflection_pdf, float& diffuse_pdf, float& transmission_pdf)
{
    const float ldotn = light.z, vdotn = view.z;
    const float vdoth = dot(view, h), ldoth = dot(light, h);
    float fresnel = fresnel_att(vdoth, f0, s.eta, 0.0f);
    float geometry = ggx_g2(ldotn, ldoth, vdotn, vdoth, s.roughness);
    float G1 = ggx_g1(vdotn, vdoth, s.roughness);
    v3 color;
    if(light.z > 0.0f)
    {   // BRDF
        float spec = geometry * distribution;
        float om = 1.0f - s.metallic;
        color = mk3((s.albedo.x * s.metallic + fresnel * om) * spec,
                    (s.albedo.y * s.metallic + fresnel * om) * spec,
                    (s.albedo.z * s.metallic + fresnel * om) * spec);
        float kd = (1.0f - fresnel) * om * (1.0f - s.transmission) / PT_PI;
        color += kd * s.albedo;
        reflection_pdf = G1 * distribution / (4.0f * view.z);
        diffuse_pdf = fmaxf(light.z * (1.0f / PT_PI), 0.0f);
        transmission_pdf = 0.0f;
    }
    else
    {   // BTDF
        float denom = s.eta * vdoth + ldoth;
        float d2 = denom * denom;
        color = s.albedo * (s.transmission * fabsf(vdoth * ldoth) * (1.0f - fresnel) * 4.0f * geometry * distribution / d2);
        reflection_pdf = 0.0f;
        diffuse_pdf = 0.0f;
        transmission_pdf = fabsf(vdoth * ldoth) * G1 * distribution / (fabsf(view.z) * d2);
    }
    return color * fabsf(ldotn);
}

PT_D void lobe_probs(v3 view, const Surface& s, float& f0, float& p_refl, float& p_trans, float& p_diff)
{
    f0 = (1.0f - s.eta) / (1.0f + s.eta);
    f0 *= f0;
    p_refl = mixf(1.0f, fresnel_att(view.z, f0, s.eta, s.roughness), luminance(s.albedo) * (1.0f - s.metallic));
    p_trans = (1.0f - p_refl) * s.transmission;
    p_diff = (1.0f - p_refl) * (1.0f - s.transmission);
}

// bsdf() — evaluation for NEE (path_tracer.hh:184-222)
PT_D v3 bsdf_eval(v3 light, v3 view, const Surface& s, float& out_pdf)
{
    v3 h;
    if(light.z > 0.0f) h = normalize(view + light);
    else h = signf(s.eta - 1.0f) * normalize(light + s.eta * view);
    float distribution = ggx_d(h.z, s.roughness);
    float f0, p_refl, p_trans, p_diff;
    lobe_probs(view, s, f0, p_refl, p_trans, p_diff);
    float rp, dp, tp;
    v3 att = bsdf_core(light, h, view, s, f0, s.roughness < 1e-3f ? 0.0f : distribution, rp, dp, tp);
    out_pdf = rp * p_refl + dp * p_diff + tp * p_trans;
    return att;
}

// sample_bsdf (path_tracer.hh:224-296). A negative pdf marks a delta lobe.
PT_D void bsdf_sample(float ux, float uy, float uz, v3 view, const Surface& s,
                      v3& out_dir, v3& out_att, float& out_pdf)
{
    v3 h = sample_ggx_vndf(view, s.roughness, ux, uy);
    float f0, p_refl, p_trans, p_diff;
    lobe_probs(view, s, f0, p_refl, p_trans, p_diff);
    bool diffuse = false, bad;
    if((uz -= p_refl) <= 0.0f)
    {
        out_dir = reflect(-view, h);
        bad = out_dir.z <= 0.0f;
    }
    else if((uz -= p_trans) <= 0.0f)
    {
        out_dir = refract(-view, h, s.eta);
        bad = out_dir.z >= 0.0f;
    }
    else
    {
        out_dir = sample_cosine_hemisphere(ux, uy);
        h = normalize(out_dir + view);
        diffuse = true;
        bad = out_dir.z == 0.0f;
    }
    if(bad)
    {
        out_dir = mk3(0, 0, 1);
        out_att = mk3(0, 0, 0);
        out_pdf = 1.0f;
        return;
    }
    float distribution = ggx_d(h.z, s.roughness);
    if(s.roughness < 1e-3f) distribution = diffuse ? 0.0f : fabsf(4.0f * out_dir.z * view.z);
    float rp, dp, tp;
    out_att = bsdf_core(out_dir, h, view, s, f0, distribution, rp, dp, tp);
    out_pdf = rp * p_refl + tp * p_trans;
    if(s.roughness < 1e-3f && !diffuse) out_pdf = -out_pdf;
    else out_pdf += dp * p_diff;
}

// ---- sky (path_tracer.hh:456-588) --------------------------------------------------------------

// ray_sphere_intersection against the atmosphere shell centred at (0,-R,0) (math.hh:404-417)
PT_D bool atmo_sphere(v3 origin, v3 dir, float& tmin, float& tmax)
{
    const float radius = PT_EARTH_RADIUS + PT_ATMO_HEIGHT;
    v3 oc = mk3(origin.x, origin.y + PT_EARTH_RADIUS, origin.z);
    float b = dot(oc, dir);
    float c = dot(oc, oc) - radius * radius;
    float disc = b * b - c;
    if(disc < 0.0f) return false;
    disc = sqrtf(disc);
    tmin = -b - disc;
    tmax = -b + disc;
    return true;
}

PT_D float atmo_height(v3 p)
{
    return length(mk3(p.x, p.y + PT_EARTH_RADIUS, p.z)) - PT_EARTH_RADIUS;
}

#ifdef PT_FAST_EXP
#define PT_EXP(x) __expf(x)
#else
#define PT_EXP(x) expf(x)
#endif

// nishita_atmosphere_attenuation (path_tracer.hh:456-497), iterations = 8, tmax = MAX_RAY_DIST
PT_D v3 sky_attenuation(float jitter, v3 pos, v3 view)
{
    float tmin, atmax;
    bool hit = atmo_sphere(pos, view, tmin, atmax);
    if(!hit) return mk3(1, 1, 1);
    tmin = fmaxf(tmin, 0.0f);
    float tmax = fminf(atmax, PT_MAX_RAY_DIST);
    float segment = (tmax - tmin) / (float)PT_ATMO_PRIMARY;
    float ray = 0.0f, mie = 0.0f;
    bool shadowed = false;
    #pragma unroll
    for(int i = 0; i < PT_ATMO_PRIMARY; ++i)
    {
        float t = segment * (jitter + (float)i);
        float height = atmo_height(pos + t * view);
        ray += PT_EXP(-height / PT_RAYLEIGH_SCALE);
        mie += PT_EXP(-height / PT_MIE_SCALE);
        if(height < 0.0f) shadowed = true;
    }
    if(shadowed) return mk3(0, 0, 0);
    float tx = (PT_RAYLEIGH_X * ray + PT_MIE_COEF * mie) * segment;
    float ty = (PT_RAYLEIGH_Y * ray + PT_MIE_COEF * mie) * segment;
    float tz = (PT_RAYLEIGH_Z * ray + PT_MIE_COEF * mie) * segment;
    return mk3(PT_EXP(-tx), PT_EXP(-ty), PT_EXP(-tz));
}

// nishita_atmosphere_scattering (path_tracer.hh:499-588). Draws one rand4 only after both
// early-outs (:513, :521). Returns true if the march ran (for the event counters).
PT_D bool sky_scattering(rng4& seed, const Light& light, v3 pos, v3 view, float tmax,
                         v3& attenuation, v3& in_scatter)
{
    attenuation = mk3(1, 1, 1);
    in_scatter = mk3(0, 0, 0);
    if(tmax > 0.0f && tmax < 1e3f) return false;
    float tmin, atmax;
    if(!atmo_sphere(pos, view, tmin, atmax)) return false;
    tmin = fmaxf(tmin, 0.0f);
    tmax = fminf(atmax, tmax < 0.0f ? PT_MAX_RAY_DIST : tmax);

    float segment = (tmax - tmin) / (float)PT_ATMO_PRIMARY;
    float4 jitter = rand4(seed);

    float mu = dot(view, light.dir);
    float rayleigh_phase = 3.0f / (16.0f * PT_PI) * (1.0f + mu * mu);
    const float g = PT_MIE_G;
    float mb = 1.0f + g * g - 2.0f * g * mu;
    float mie_phase = 3.0f / (8.0f * PT_PI) * (1.0f - g * g) * (1.0f + mu * mu) /
        ((2.0f + g * g) * (mb * sqrtf(mb)));

    float ray_od = 0.0f, mie_od = 0.0f;
    v3 ray_sum = mk3(0, 0, 0), mie_sum = mk3(0, 0, 0);
    float l0 = tmin, l1 = tmax;
    #pragma unroll 1   // (unrolled by 2: 104 registers, same time)
    for(int i = 0; i < PT_ATMO_PRIMARY; ++i)
    {
        float t = segment * (jitter.x + (float)i);
        v3 p = pos + t * view;
        // the reference reuses its outer tmin/tmax here (:542-543): on a miss (only possible through
        // rounding, p is inside the shell) they keep the values of the previous iteration
        atmo_sphere(p, light.dir, l0, l1);
        float light_segment = (l1 - l0) / (float)PT_ATMO_SECONDARY;
        float lray = 0.0f, lmie = 0.0f;
        bool shadowed = false;
        #pragma unroll
        for(int j = 0; j < PT_ATMO_SECONDARY; ++j)
        {
            float tl = light_segment * (jitter.y + (float)j);
            float height = atmo_height(p + tl * light.dir);
            lray += PT_EXP(-height / PT_RAYLEIGH_SCALE);
            lmie += PT_EXP(-height / PT_MIE_SCALE);
            if(height < 0.0f) shadowed = true;
        }
        float height = fmaxf(atmo_height(p), 0.0f);
        float ray_density = PT_EXP(-height / PT_RAYLEIGH_SCALE) * segment;
        float mie_density = PT_EXP(-height / PT_MIE_SCALE) * segment;
        ray_od += ray_density;
        mie_od += mie_density;
        float rr = lray * light_segment + ray_od;
        float mm = PT_MIE_COEF * (lmie * light_segment + mie_od);
        if(!shadowed)
        {
            v3 la = mk3(PT_EXP(-(PT_RAYLEIGH_X * rr + mm)), PT_EXP(-(PT_RAYLEIGH_Y * rr + mm)),
                        PT_EXP(-(PT_RAYLEIGH_Z * rr + mm)));
            ray_sum += la * ray_density;
            mie_sum += la * mie_density;
        }
    }
    float mt = PT_MIE_COEF * mie_od;
    attenuation = mk3(PT_EXP(-(PT_RAYLEIGH_X * ray_od + mt)), PT_EXP(-(PT_RAYLEIGH_Y * ray_od + mt)),
                      PT_EXP(-(PT_RAYLEIGH_Z * ray_od + mt)));
    v3 rs = mk3(ray_sum.x * PT_RAYLEIGH_X, ray_sum.y * PT_RAYLEIGH_Y, ray_sum.z * PT_RAYLEIGH_Z) * rayleigh_phase;
    v3 ms = mie_sum * (PT_MIE_COEF * mie_phase);
    in_scatter = (rs + ms) * light.color * 4.0f;
    return true;
}

// ---- tonemap_pixel (path_tracer.hh:753-771): returns B,G,R,A bytes -------------------------------
PT_D float srgb_oetf(float c)
{
    return c < 0.0031308f ? c * 12.92f : powf(c, 1.0f / 2.4f) * 1.055f - 0.055f;
}
PT_D float aces_fit(float c)
{
    return (c * (2.51f * c + 0.03f)) / (c * (2.43f * c + 0.59f) + 0.14f);
}
PT_D uchar4 tonemap(v3 color)
{
    float r = clampf(srgb_oetf(aces_fit(color.x)), 0.0f, 1.0f);
    float g = clampf(srgb_oetf(aces_fit(color.y)), 0.0f, 1.0f);
    float b = clampf(srgb_oetf(aces_fit(color.z)), 0.0f, 1.0f);
    // round() = half away from zero; the operands are non-negative
    return make_uchar4((unsigned char)roundf(b * 255.0f), (unsigned char)roundf(g * 255.0f),
                       (unsigned char)roundf(r * 255.0f), 255);
}

} // namespace pt
