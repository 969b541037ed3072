// Device math for the path tracer: small vector type, matrix conventions, RNG, sampling.
// Restates (does not include) the reference's math.hh — CUDA's float3 is 12 bytes and collides
// with the host's 16-byte float3 (math.hh:36), so device code uses its own `v3`.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PT_D __device__ __forceinline__
#define PT_HD __host__ __device__ __forceinline__

#define PT_PI 3.14159265358979323846f
#define PT_TWO_PI 6.28318530717958647692f

namespace pt {

struct v3 { float x, y, z; };
struct v2 { float x, y; };

PT_HD v3 mk3(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
PT_HD v3 mk3(float4 a) { return mk3(a.x, a.y, a.z); }
PT_HD v3 operator+(v3 a, v3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
PT_HD v3 operator-(v3 a, v3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
PT_HD v3 operator*(v3 a, v3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
PT_HD v3 operator*(v3 a, float b) { return mk3(a.x * b, a.y * b, a.z * b); }
PT_HD v3 operator*(float b, v3 a) { return mk3(a.x * b, a.y * b, a.z * b); }
PT_HD v3 operator/(v3 a, v3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
PT_HD v3 operator-(v3 a) { return mk3(-a.x, -a.y, -a.z); }
PT_HD v3& operator+=(v3& a, v3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
PT_HD v3& operator*=(v3& a, v3 b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; return a; }
PT_HD float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PT_HD v3 cross(v3 a, v3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

// normalize(a) = a / length(a) (math.hh:106,110). One IEEE sqrt and one IEEE reciprocal; the
// reference's three divisions differ from this by at most 1 ulp per component.
PT_D v3 normalize(v3 a)
{
    float inv = 1.0f / sqrtf(dot(a, a));
    return a * inv;
}
PT_D float length(v3 a) { return sqrtf(dot(a, a)); }

// mat3 with r[i] = COLUMNS for mul_m3v3 (math.hh:224-227: mul_m3v3 transposes, then dots rows).
struct m3 { v3 c0, c1, c2; };
// mul_m3v3(M, v): M.c0*v.x + M.c1*v.y + M.c2*v.z, summed in the reference's dot order
PT_HD v3 mul_m3v3(const m3& m, v3 v)
{
    return mk3(m.c0.x * v.x + m.c1.x * v.y + m.c2.x * v.z,
               m.c0.y * v.x + m.c1.y * v.y + m.c2.y * v.z,
               m.c0.z * v.x + m.c1.z * v.y + m.c2.z * v.z);
}
// mul_v3m3(v, M) = (dot(c0,v), dot(c1,v), dot(c2,v)) (math.hh:224) — world -> tangent for a TBN
PT_HD v3 mul_v3m3(v3 v, const m3& m) { return mk3(dot(m.c0, v), dot(m.c1, v), dot(m.c2, v)); }

// create_tangent_space (math.hh:419-435)
PT_D m3 tangent_space(v3 n)
{
    const float k = 0.57735026918962576451f;
    v3 major;
    if(fabsf(n.x) < k) major = mk3(1, 0, 0);
    else if(fabsf(n.y) < k) major = mk3(0, 1, 0);
    else major = mk3(0, 0, 1);
    m3 m;
    m.c0 = normalize(cross(n, major));
    m.c1 = cross(n, m.c0);
    m.c2 = n;
    return m;
}

PT_D float luminance(v3 c) { return c.x * 0.2126f + c.y * 0.7152f + c.z * 0.0722f; }
PT_D float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
PT_D float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
PT_D float signf(float v) { return v < 0.0f ? -1.0f : (v > 0.0f ? 1.0f : 0.0f); }

// reflect / refract (math.hh:442-453)
PT_D v3 reflect(v3 I, v3 N) { return I - (2.0f * dot(N, I)) * N; }
PT_D v3 refract(v3 I, v3 N, float eta)
{
    float ndoti = dot(N, I);
    float k = 1.0f - eta * eta * (1.0f - ndoti * ndoti);
    if(k < 0.0f) return mk3(0, 0, 0);
    return eta * I - (eta * ndoti + sqrtf(k)) * N;
}

// ---- RNG: pcg4d (math.hh:466-473), bit-exact uint32 arithmetic -------------------------------
struct rng4 { uint32_t x, y, z, w; };

PT_HD void pcg4d(rng4& s)
{
    s.x = s.x * 1664525u + 1013904223u;
    s.y = s.y * 1664525u + 1013904223u;
    s.z = s.z * 1664525u + 1013904223u;
    s.w = s.w * 1664525u + 1013904223u;
    // each "+=" is a whole-vector statement in the reference: all four products use the old values
    uint32_t ax = s.y * s.w, ay = s.z * s.x, az = s.x * s.y, aw = s.y * s.z;
    s.x += ax; s.y += ay; s.z += az; s.w += aw;
    s.x ^= s.x >> 16; s.y ^= s.y >> 16; s.z ^= s.z >> 16; s.w ^= s.w >> 16;
    ax = s.y * s.w; ay = s.z * s.x; az = s.x * s.y; aw = s.y * s.z;
    s.x += ax; s.y += ay; s.z += az; s.w += aw;
}

// generate_uniform_random4 (math.hh:475-485): (float)uint32 (round-to-nearest) * 2^-32
PT_D float4 rand4(rng4& s)
{
    pcg4d(s);
    const float k = 2.3283064365386963e-10f;
    return make_float4(__uint2float_rn(s.x) * k, __uint2float_rn(s.y) * k,
                       __uint2float_rn(s.z) * k, __uint2float_rn(s.w) * k);
}

// ---- sampling (path_tracer.hh:12-83) ---------------------------------------------------------

// inv_erf (math.hh:455-463)
PT_D float inv_erf(float x)
{
    float ln1x2 = logf(1.0f - x * x);
    const float a = 0.147f;
    const float p = 2.0f / (PT_PI * a);
    float k = p + ln1x2 * 0.5f;
    float k2 = k * k;
    return signf(x) * sqrtf(sqrtf(k2 - ln1x2 * (1.0f / a)) - k);
}

// sample_gaussian_weighted_disk(u, sigma) (path_tracer.hh:12-25)
PT_D v2 sample_gaussian_disk(float ux, float uy, float sigma)
{
    float r = sqrtf(ux);
    float theta = PT_TWO_PI * uy;
    float k = clampf(r * 2.0f - 1.0f, -(1.0f - 1e-6f), 1.0f - 1e-6f);
    r = sigma * 1.41421356f * inv_erf(k);
    float s, c;
    sincosf(theta, &s, &c);
    v2 o; o.x = r * c; o.y = r * s;
    return o;
}

// sample_cosine_hemisphere (path_tracer.hh:27-33)
PT_D v3 sample_cosine_hemisphere(float ux, float uy)
{
    float r = sqrtf(ux);
    float s, c;
    sincosf(PT_TWO_PI * uy, &s, &c);
    float dx = r * c, dy = r * s;
    return mk3(dx, dy, sqrtf(fmaxf(0.0f, 1.0f - (dx * dx + dy * dy))));
}

// sample_cone (path_tracer.hh:40-48)
PT_D v3 sample_cone(v3 dir, float cos_theta_min, float ux, float uy)
{
    float cos_theta = mixf(1.0f, cos_theta_min, ux);
    float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
    float s, c;
    sincosf(uy * PT_TWO_PI, &s, &c);
    m3 t = tangent_space(dir);
    return mul_m3v3(t, mk3(c * sin_theta, s * sin_theta, cos_theta));
}

// sample_regular_polygon (path_tracer.hh:50-62)
PT_D v2 sample_regular_polygon(float ux, float uy, float angle, uint32_t sides)
{
    float fs = (float)sides;
    float side = floorf(ux * fs);
    ux *= fs;
    ux = ux - floorf(ux);
    float side_radians = PT_TWO_PI / fs;
    float a1 = side_radians * side + angle;
    float a2 = side_radians * (side + 1.0f) + angle;
    float s1, c1, s2, c2;
    sincosf(a1, &s1, &c1);
    sincosf(a2, &s2, &c2);
    if(ux + uy > 1.0f) { ux = 1.0f - ux; uy = 1.0f - uy; }
    v2 o; o.x = s1 * ux + s2 * uy; o.y = c1 * ux + c2 * uy;
    return o;
}

// sample_ggx_vndf (path_tracer.hh:67-83)
PT_D v3 sample_ggx_vndf(v3 view, float roughness, float ux, float uy)
{
    if(roughness < 1e-3f) return mk3(0, 0, 1);
    v3 v = normalize(mk3(roughness * view.x, roughness * view.y, view.z));
    float phi = PT_TWO_PI * ux;
    float z = fmaf(1.0f - uy, 1.0f + v.z, -v.z);
    float sin_theta = sqrtf(clampf(1.0f - z * z, 0.0f, 1.0f));
    float s, c;
    sincosf(phi, &s, &c);
    v3 h = mk3(sin_theta * c, sin_theta * s, z) + v;
    return normalize(mk3(roughness * h.x, roughness * h.y, fmaxf(0.0f, h.z)));
}

} // namespace pt
