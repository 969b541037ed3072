// Host-only: per-frame scene state WITHOUT the reference's setup_animation_frame.
//
// setup_animation_frame (scene.cc:271-718) replays a keyframe table once per motion-blur subframe,
// appends up to five dynamic instances per subframe and then builds one SAH TLAS per subframe over
// all ~890 instances (32-128 builds per frame, 75 ms single-threaded: as long as a GPU frame). The
// kernels here never needed those TLASes — the static TLAS is built once and a subframe's few dynamic
// instances are tested directly — so the per-frame host work reduces to the keyframe replay and
// <= 7 matrix products per subframe: microseconds, re-entrant, one call per frame per GPU worker.
//
// The keyframe rows and the mesh handles are DATA handed in by the caller (extracted from the
// reference at build time, oracle/extract_animation.py); the player below restates scene.cc's
// arithmetic (same operations in float, double where the reference's unqualified sin/cos/tan/cos
// resolve to the double overloads).
#include "../../include/ptgpu.h"

#include <cmath>
#include <cstring>
#include <new>
#include <vector>

namespace {

struct f4 { float x, y, z, w; };
struct m4 { f4 r[4]; };   // columns, as math.hh:153 / 226-228
struct f3 { float x, y, z; };
struct m3 { f3 r[3]; };

inline float dot4(f4 a, f4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// mul_m3m3 / mul_m4m4 (math.hh:238-256): result.r[i] = (a.r[i] . rows of b)
m3 mul33(const m3& b, const m3& a)
{
    const f3 bt[3] = {{b.r[0].x, b.r[1].x, b.r[2].x}, {b.r[0].y, b.r[1].y, b.r[2].y}, {b.r[0].z, b.r[1].z, b.r[2].z}};
    m3 o;
    for(int i = 0; i < 3; ++i) o.r[i] = {dot3(a.r[i], bt[0]), dot3(a.r[i], bt[1]), dot3(a.r[i], bt[2])};
    return o;
}
m4 mul44(const m4& b, const m4& a)
{
    const f4 bt[4] = {{b.r[0].x, b.r[1].x, b.r[2].x, b.r[3].x}, {b.r[0].y, b.r[1].y, b.r[2].y, b.r[3].y},
                      {b.r[0].z, b.r[1].z, b.r[2].z, b.r[3].z}, {b.r[0].w, b.r[1].w, b.r[2].w, b.r[3].w}};
    m4 o;
    for(int i = 0; i < 4; ++i) o.r[i] = {dot4(a.r[i], bt[0]), dot4(a.r[i], bt[1]), dot4(a.r[i], bt[2]), dot4(a.r[i], bt[3])};
    return o;
}

// rotation_euler (math.hh:305-318): roll * yaw * pitch, angles in radians
m4 rotation_euler(f3 e)
{
    const float sp = (float)std::sin((double)e.x), cp = (float)std::cos((double)e.x);
    const float sy = (float)std::sin((double)e.y), cy = (float)std::cos((double)e.y);
    const float sr = (float)std::sin((double)e.z), cr = (float)std::cos((double)e.z);
    const m3 pitch = {{{1, 0, 0}, {0, cp, -sp}, {0, sp, cp}}};
    const m3 yaw = {{{cy, 0, sy}, {0, 1, 0}, {-sy, 0, cy}}};
    const m3 roll = {{{cr, -sr, 0}, {sr, cr, 0}, {0, 0, 1}}};
    const m3 r = mul33(roll, mul33(yaw, pitch));
    return {{{r.r[0].x, r.r[0].y, r.r[0].z, 0}, {r.r[1].x, r.r[1].y, r.r[1].z, 0}, {r.r[2].x, r.r[2].y, r.r[2].z, 0}, {0, 0, 0, 1}}};
}
m4 translation(f3 p) { return {{{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {p.x, p.y, p.z, 1}}}; }
m4 scaling(f3 s) { return {{{s.x, 0, 0, 0}, {0, s.y, 0, 0}, {0, 0, s.z, 0}, {0, 0, 0, 1}}}; }

// inverse4 (math.hh:179-221): cofactor expansion (after GLM)
m4 inverse4(const m4& a)
{
    const f4 *c = a.r;
    const float c00 = c[2].z * c[3].w - c[3].z * c[2].w, c02 = c[1].z * c[3].w - c[3].z * c[1].w, c03 = c[1].z * c[2].w - c[2].z * c[1].w;
    const float c04 = c[2].y * c[3].w - c[3].y * c[2].w, c06 = c[1].y * c[3].w - c[3].y * c[1].w, c07 = c[1].y * c[2].w - c[2].y * c[1].w;
    const float c08 = c[2].y * c[3].z - c[3].y * c[2].z, c10 = c[1].y * c[3].z - c[3].y * c[1].z, c11 = c[1].y * c[2].z - c[2].y * c[1].z;
    const float c12 = c[2].x * c[3].w - c[3].x * c[2].w, c14 = c[1].x * c[3].w - c[3].x * c[1].w, c15 = c[1].x * c[2].w - c[2].x * c[1].w;
    const float c16 = c[2].x * c[3].z - c[3].x * c[2].z, c18 = c[1].x * c[3].z - c[3].x * c[1].z, c19 = c[1].x * c[2].z - c[2].x * c[1].z;
    const float c20 = c[2].x * c[3].y - c[3].x * c[2].y, c22 = c[1].x * c[3].y - c[3].x * c[1].y, c23 = c[1].x * c[2].y - c[2].x * c[1].y;
    const f4 f0 = {c00, c00, c02, c03}, f1 = {c04, c04, c06, c07}, f2 = {c08, c08, c10, c11};
    const f4 f3_ = {c12, c12, c14, c15}, f4_ = {c16, c16, c18, c19}, f5 = {c20, c20, c22, c23};
    const f4 v0 = {c[1].x, c[0].x, c[0].x, c[0].x}, v1 = {c[1].y, c[0].y, c[0].y, c[0].y};
    const f4 v2 = {c[1].z, c[0].z, c[0].z, c[0].z}, v3 = {c[1].w, c[0].w, c[0].w, c[0].w};
    auto comb = [](f4 p, f4 fa, f4 q, f4 fb, f4 r, f4 fc, f4 sgn) {
        return f4{(p.x * fa.x - q.x * fb.x + r.x * fc.x) * sgn.x, (p.y * fa.y - q.y * fb.y + r.y * fc.y) * sgn.y,
                  (p.z * fa.z - q.z * fb.z + r.z * fc.z) * sgn.z, (p.w * fa.w - q.w * fb.w + r.w * fc.w) * sgn.w};
    };
    const f4 pos = {+1, -1, +1, -1}, neg = {-1, +1, -1, +1};
    m4 inv = {{comb(v1, f0, v2, f1, v3, f2, pos), comb(v0, f0, v2, f3_, v3, f4_, neg),
               comb(v0, f1, v1, f3_, v3, f5, pos), comb(v0, f2, v1, f4_, v2, f5, neg)}};
    const float det = dot4(c[0], f4{inv.r[0].x, inv.r[1].x, inv.r[2].x, inv.r[3].x});
    const float k = 1.0f / det;
    for(int i = 0; i < 4; ++i) inv.r[i] = {inv.r[i].x * k, inv.r[i].y * k, inv.r[i].z * k, inv.r[i].w * k};
    return inv;
}

inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; } // math.hh:145
inline float clamp01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }

void put_instance(ptgpu_tlas_instance& out, const ptgpu_mesh_handle& h, const m4& t)
{
    memset(&out, 0, sizeof(out));
    out.blas = h.blas;
    out.m = h.m;
    const m4 inv = inverse4(t);
    memcpy(&out.transform, &t, sizeof(m4));
    memcpy(&out.inv_transform, &inv, sizeof(m4));
}

// add_instance(s, name, pos, pitch_yaw_roll, scale = 1) (scene.cc:62-74)
m4 instance_transform(f3 pos, f3 pyr_degrees)
{
    const float pi = (float)M_PI;
    m4 t = scaling({1, 1, 1});
    t = mul44(rotation_euler({pyr_degrees.x * pi / 180.0f, pyr_degrees.y * pi / 180.0f, pyr_degrees.z * pi / 180.0f}), t);
    return mul44(translation(pos), t);
}

} // namespace

struct ptgpu_anim
{
    std::vector<ptgpu_anim_key> keys;
    ptgpu_mesh_handle meshes[PTGPU_MESH_COUNT];
    ptgpu_config cfg;
};

extern "C" {

int ptgpu_anim_create(ptgpu_anim** out, const ptgpu_anim_key* keys, size_t n_keys,
                      const ptgpu_mesh_handle* meshes, const ptgpu_config* cfg)
{
    if(!out) return 1;
    *out = nullptr;
    if(!keys || !meshes || !cfg || cfg->spp <= 0 || cfg->samples_per_subframe <= 0 || cfg->height <= 0) return 1;
    for(size_t i = 0; i < n_keys; ++i)
        if(keys[i].var < 0 || keys[i].var >= PTGPU_ANIM_VAR_COUNT) return 1;
    ptgpu_anim* a = new(std::nothrow) ptgpu_anim();
    if(!a) return 1;
    a->keys.assign(keys, keys + n_keys);
    memcpy(a->meshes, meshes, sizeof(a->meshes));
    a->cfg = *cfg;
    *out = a;
    return 0;
}

void ptgpu_anim_destroy(ptgpu_anim* a) { delete a; }

size_t ptgpu_anim_subframe_count(const ptgpu_anim* a)
{   // scene.cc:648-650
    return a ? (size_t)((a->cfg.spp + a->cfg.samples_per_subframe - 1) / a->cfg.samples_per_subframe) : 0;
}

size_t ptgpu_anim_max_instances(const ptgpu_anim* a)
{   // logo + buddha, then teapot/armadillo/dragon/bunny/end per subframe (scene.cc:634-674)
    return a ? 2 + 5 * ptgpu_anim_subframe_count(a) : 0;
}

uint32_t ptgpu_anim_frame_count(const ptgpu_anim*) { return 60 * 30; } // scene.cc:720-724

int ptgpu_anim_frame(const ptgpu_anim* a, uint32_t frame, ptgpu_subframe* subframes,
                     ptgpu_tlas_instance* dyn, size_t* n_dyn, uint32_t* dyn_begin, uint32_t* dyn_end)
{
    if(!a || !subframes || !dyn || !n_dyn || !dyn_begin || !dyn_end) return 1;
    const float framerate = 30.0f; // FRAMERATE, config.hh:17/24
    float var[PTGPU_ANIM_VAR_COUNT];
    // initial values (scene.cc:279-316)
    for(float& v : var) v = 0.0f;
    const f3 camera_start_pos = {-81.4f, 65.0f, -113.6f}, camera_start_ori = {30.6f, 146.6f, 0.0f};
    var[PTGPU_VAR_CAM_POS_X] = camera_start_pos.x; var[PTGPU_VAR_CAM_POS_Y] = camera_start_pos.y; var[PTGPU_VAR_CAM_POS_Z] = camera_start_pos.z;
    var[PTGPU_VAR_CAM_ORI_X] = camera_start_ori.x; var[PTGPU_VAR_CAM_ORI_Y] = camera_start_ori.y; var[PTGPU_VAR_CAM_ORI_Z] = camera_start_ori.z;
    var[PTGPU_VAR_FOV] = 80.0f;
    var[PTGPU_VAR_FOCAL_DISTANCE] = 2.0f;
    var[PTGPU_VAR_APERTURE_RADIUS] = 0.0f;
    var[PTGPU_VAR_TEAPOT_POS_X] = 40.1f; var[PTGPU_VAR_TEAPOT_POS_Y] = 13.95f; var[PTGPU_VAR_TEAPOT_POS_Z] = 13.611633f;

    // play_animation_track (scene.cc:33-42): every key whose start has passed sets its variable
    auto play = [&](float t) {
        for(const ptgpu_anim_key& k : a->keys)
        {
            if(!(k.start <= t)) break;
            const float lt = k.duration == 0.0f ? 1.0f : clamp01((t - k.start) / k.duration);
            var[k.var] = mixf(k.from, k.to, lt);
        }
    };
    const float anim_t0 = float(frame) / framerate * 30.0f;
    play(anim_t0);

    size_t n = 0;
    // frame-static instances (scene.cc:634-644)
    if(var[PTGPU_VAR_LOGO_VISIBLE] != 0.0f)
    {
        const float pi = (float)M_PI;
        m4 t = rotation_euler({camera_start_ori.x * pi / 180.0f, camera_start_ori.y * pi / 180.0f, camera_start_ori.z * pi / 180.0f});
        const f3 logo_pos = {camera_start_pos.x - (-1.3f), camera_start_pos.y - 2.0f, camera_start_pos.z - (-2.0f)};
        t = mul44(translation(logo_pos), t);
        put_instance(dyn[n++], a->meshes[PTGPU_MESH_LOGO], t);
    }
    put_instance(dyn[n++], a->meshes[PTGPU_MESH_BUDDHA], instance_transform({-39.255131f, 30.395447f, 40.472446f}, {0, 0, 0}));

    const uint32_t subframe_count = (uint32_t)ptgpu_anim_subframe_count(a);
    for(uint32_t i = 0; i < subframe_count; ++i)
    {
        const float anim_t = float(frame + float(i) / subframe_count) / framerate * 30.0f; // scene.cc:661
        play(anim_t);
        dyn_begin[i] = (uint32_t)n;
        put_instance(dyn[n++], a->meshes[PTGPU_MESH_TEAPOT],
                     instance_transform({var[PTGPU_VAR_TEAPOT_POS_X], var[PTGPU_VAR_TEAPOT_POS_Y], var[PTGPU_VAR_TEAPOT_POS_Z]},
                                        {var[PTGPU_VAR_TEAPOT_ORI_X], var[PTGPU_VAR_TEAPOT_ORI_Y], var[PTGPU_VAR_TEAPOT_ORI_Z]}));
        if(var[PTGPU_VAR_ARMADILLO_VISIBLE] != 0.0f)
            put_instance(dyn[n++], a->meshes[PTGPU_MESH_ARMADILLO],
                         instance_transform({var[PTGPU_VAR_ARMADILLO_POS_X], var[PTGPU_VAR_ARMADILLO_POS_Y], var[PTGPU_VAR_ARMADILLO_POS_Z]},
                                            {var[PTGPU_VAR_ARMADILLO_ORI_X], var[PTGPU_VAR_ARMADILLO_ORI_Y], var[PTGPU_VAR_ARMADILLO_ORI_Z]}));
        if(var[PTGPU_VAR_DRAGON_VISIBLE] != 0.0f)
            put_instance(dyn[n++], a->meshes[PTGPU_MESH_DRAGON],
                         instance_transform({var[PTGPU_VAR_DRAGON_POS_X], var[PTGPU_VAR_DRAGON_POS_Y], var[PTGPU_VAR_DRAGON_POS_Z]},
                                            {var[PTGPU_VAR_DRAGON_ORI_X], var[PTGPU_VAR_DRAGON_ORI_Y], var[PTGPU_VAR_DRAGON_ORI_Z]}));
        if(var[PTGPU_VAR_BUNNY_VISIBLE] != 0.0f)
            put_instance(dyn[n++], a->meshes[PTGPU_MESH_BUNNY],
                         instance_transform({var[PTGPU_VAR_BUNNY_POS_X], var[PTGPU_VAR_BUNNY_POS_Y], var[PTGPU_VAR_BUNNY_POS_Z]},
                                            {var[PTGPU_VAR_BUNNY_ORI_X], var[PTGPU_VAR_BUNNY_ORI_Y], var[PTGPU_VAR_BUNNY_ORI_Z]}));
        if(var[PTGPU_VAR_END_VISIBLE] != 0.0f)
            put_instance(dyn[n++], a->meshes[PTGPU_MESH_END],
                         instance_transform({var[PTGPU_VAR_END_POS_X], var[PTGPU_VAR_END_POS_Y], var[PTGPU_VAR_END_POS_Z]},
                                            {var[PTGPU_VAR_END_ORI_X], var[PTGPU_VAR_END_ORI_Y], var[PTGPU_VAR_END_ORI_Z]}));
        dyn_end[i] = (uint32_t)n;

        // camera and light of the subframe (scene.cc:682-695)
        ptgpu_subframe& sf = subframes[i];
        memset(&sf, 0, sizeof(sf));
        const float pi = (float)M_PI;
        const m4 ori = rotation_euler({var[PTGPU_VAR_CAM_ORI_X] * pi / 180.0f, var[PTGPU_VAR_CAM_ORI_Y] * pi / 180.0f, var[PTGPU_VAR_CAM_ORI_Z] * pi / 180.0f});
        for(int c = 0; c < 3; ++c) { sf.cam.orientation.r[c].x = ori.r[c].x; sf.cam.orientation.r[c].y = ori.r[c].y; sf.cam.orientation.r[c].z = ori.r[c].z; }
        sf.cam.position.x = var[PTGPU_VAR_CAM_POS_X]; sf.cam.position.y = var[PTGPU_VAR_CAM_POS_Y]; sf.cam.position.z = var[PTGPU_VAR_CAM_POS_Z];
        sf.cam.aspect_ratio = a->cfg.width / float(a->cfg.height);
        sf.cam.inv_focal_length = (float)std::tan((double)var[PTGPU_VAR_FOV] * M_PI / 360.0f);
        sf.cam.focal_distance = var[PTGPU_VAR_FOCAL_DISTANCE];
        sf.cam.aperture_angle = (float)(M_PI / 16.0f);
        sf.cam.aperture_polygon = 6;
        sf.cam.aperture_radius = var[PTGPU_VAR_APERTURE_RADIUS];
        const float sunset_t = anim_t / (30.0f * 60.0f) * 1.1f - 0.05f;
        sf.light.direction.x = 0.0f;
        sf.light.direction.y = sinf((float)(sunset_t * M_PI));
        sf.light.direction.z = cosf((float)(sunset_t * M_PI));
        sf.light.color.x = sf.light.color.y = sf.light.color.z = 4.0f;
        sf.light.cos_solid_angle = (float)std::cos(4.0f * M_PI / 180.0f);
    }
    *n_dyn = n;
    return 0;
}

} // extern "C"
