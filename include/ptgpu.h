/* ptgpu.h — C ABI of libptgpu.so, the B200 (sm_100a) implementation of the reference's rendering
 * hot path: baseline_render -> path_trace_pixel -> ray_query traversal -> tonemap_pixel.
 *
 * C99-callable, plain pointers and sizes only. Every entry point names the reference interface
 * it replaces (file:line relative to the reference tree).
 *
 * Seam (reference main.cc:82-101):
 *     setup_animation_frame(s, f);          // caller (reference scene.cc), unchanged
 *     baseline_render(s, image);            // <- replaced by ptgpu_render_frame()
 *     write_bmp(..., image);                // caller, unchanged (or ptgpu_render_frame_bmp)
 *
 * Conventions
 *   - return 0 on success, non-zero on error; text via ptgpu_last_error(). The reference has no
 *     error returns (it prints and exit(1)s, mesh.cc:25-41, bmp.cc:54-59).
 *   - the context copies everything it is given; host pointers are never retained
 *     (setup_animation_frame reallocates them every frame, scene.cc:274-277, 712-717).
 *   - one context per GPU, one host thread per context.
 *   - host `float3` is 16 bytes (alignas(16), math.hh:36): all vec3 arrays are passed as 16-byte
 *     elements.
 */
#ifndef PTGPU_H
#define PTGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- POD mirrors of the reference types that cross the seam (layout-identical) ------------- */

typedef struct { float x, y, z, pad; } ptgpu_float3;                 /* math.hh:36  (16 B) */
typedef struct { float x, y, z, w; } ptgpu_float4;                   /* math.hh:37  (16 B) */
typedef struct { ptgpu_float3 r[3]; } ptgpu_mat3;                    /* math.hh:152 (48 B) */
typedef struct { ptgpu_float4 r[4]; } ptgpu_mat4;                    /* math.hh:153 (64 B) */

typedef struct { uint32_t node_count, node_offset; } ptgpu_bvh;      /* bvh.hh:35-39  (8 B) */
typedef struct { float min_x, min_y, min_z, max_x, max_y, max_z; } ptgpu_bvh_node; /* bvh.hh:45-49 (24 B) */
typedef struct { uint32_t accept, cancel; } ptgpu_bvh_link;          /* bvh.hh:57-67  (8 B) */
typedef struct {                                                      /* mesh.hh:18-28 (16 B) */
    uint32_t vertex_count, triangle_count, index_offset, base_vertex_offset;
} ptgpu_mesh;
typedef struct {                                                      /* bvh.hh:73-79 (160 B) */
    ptgpu_bvh blas;          /* @0  */
    ptgpu_mesh m;            /* @8  */
    uint32_t pad_[2];        /* @24 (alignment of mat4) */
    ptgpu_mat4 transform;    /* @32 */
    ptgpu_mat4 inv_transform;/* @96 */
} ptgpu_tlas_instance;
typedef struct {                                                      /* scene.hh:7-18 (96 B) */
    ptgpu_mat3 orientation;  /* @0  */
    ptgpu_float3 position;   /* @48 */
    float aspect_ratio;      /* @64 */
    float inv_focal_length;  /* @68 */
    float focal_distance;    /* @72 */
    float aperture_angle;    /* @76 */
    int32_t aperture_polygon;/* @80 */
    float aperture_radius;   /* @84 */
    uint32_t pad_[2];
} ptgpu_camera;
typedef struct {                                                      /* scene.hh:20-25 (48 B) */
    ptgpu_float3 direction;
    ptgpu_float3 color;
    float cos_solid_angle;
    uint32_t pad_[3];
} ptgpu_directional_light;
typedef struct {                                                      /* scene.hh:27-35 (160 B) */
    ptgpu_bvh tlas;          /* @0   */
    uint32_t pad_[2];
    ptgpu_camera cam;        /* @16  */
    ptgpu_directional_light light; /* @112 */
} ptgpu_subframe;

/* The compile-time constants of config.hh, made run-time so one library serves the shipped
 * TESTING config (config.hh:14-18), production (config.hh:21-25) and test-sized renders. */
typedef struct {
    int32_t width;                 /* IMAGE_WIDTH   config.hh:14/21 */
    int32_t height;                /* IMAGE_HEIGHT  config.hh:15/22 */
    int32_t spp;                   /* SAMPLES_PER_PIXEL config.hh:16/23 */
    int32_t max_bounces;           /* MAX_BOUNCES   config.hh:18/25 */
    uint32_t student_id;           /* STUDENT_ID    config.hh:5 — 4th word of the RNG key */
    int32_t samples_per_subframe;  /* SAMPLES_PER_MOTION_BLUR_STEP config.hh:29 (8) */
} ptgpu_config;

typedef struct ptgpu_ctx ptgpu_ctx;

/* ---- life cycle ----------------------------------------------------------------------------- */

/* Fills cfg with the shipped config.hh values (TESTING: 640x360, 256 spp, 4 bounces,
 * STUDENT_ID 152121358, 8 samples per motion-blur step). */
void ptgpu_default_config(ptgpu_config* cfg);

/* Creates a context on CUDA device `device`. Fails (non-zero) if no sm_100 device is present:
 * there is no CPU fallback. On failure *out is NULL and ptgpu_last_error(NULL) has the text. */
int ptgpu_create(ptgpu_ctx** out, int device, const ptgpu_config* cfg);
/* Creates the CUDA context of `device` ahead of ptgpu_create (about a second per GPU on an 8-GPU box, serialised
 * by the driver): a multi-GPU host calls it from one thread per GPU at program start and loads its scene
 * meanwhile. Optional; 0 = ok. */
int ptgpu_warm_up(int device);
void ptgpu_destroy(ptgpu_ctx* ctx);
const char* ptgpu_last_error(const ptgpu_ctx* ctx);

/* ---- static scene: everything load_scene() produces (scene.cc:135-269) ---------------------- */

/* Replaces the by-pointer hand-over of main.cc:29-37 for the data that does not change after
 * load_scene(): all BLAS nodes/links (bvh.hh:88-92; links hold 8 direction tables per BVH at
 * links[8*node_offset + octant*node_count + i], bvh.cc:218-226), the mesh buffers
 * (mesh.hh:32-44) and instances[0 .. n_static) (scene.hh:55-60). Uploads once and builds the
 * GPU traversal layout (wide BVH per BLAS + one static TLAS). n_links must be 8*n_nodes. */
int ptgpu_upload_static(
    ptgpu_ctx* ctx,
    const ptgpu_bvh_node* nodes, size_t n_nodes,
    const ptgpu_bvh_link* links, size_t n_links,
    const uint32_t* indices, size_t n_indices,
    const ptgpu_float3* pos, const ptgpu_float3* normal,
    const ptgpu_float4* albedo, const ptgpu_float4* material, size_t n_verts,
    const ptgpu_tlas_instance* instances, size_t n_static);

/* the static scene from its meshes alone. The reference hands baseline_render BVHs that
 * build_blas made (bvh.cc:231-260: full-sweep SAH, eight link tables per BVH, 2.6 s of load_scene);
 * this entry point needs none of them: every BLAS is built here from the triangles of `meshes`
 * (binned SAH above 256 primitives, full sweep below, single-triangle leaves, then the cost-optimal
 * collapse into compressed 8-wide nodes), and instances name their BLAS by their `m` member
 * (`blas` is ignored). Afterwards use ptgpu_set_frame_ranges / ptgpu_set_animation_frame for the
 * per-frame state (ptgpu_set_frame wants the reference's TLAS arrays); "traversal" = 1 and the event
 * counters, which walk the reference's link tables, are unavailable. `meshes` must list every mesh
 * that a static or per-frame instance will use. Hits are the same as with ptgpu_upload_static. */
int ptgpu_upload_meshes(
    ptgpu_ctx* ctx,
    const uint32_t* indices, size_t n_indices,
    const ptgpu_float3* pos, const ptgpu_float3* normal,
    const ptgpu_float4* albedo, const ptgpu_float4* material, size_t n_verts,
    const ptgpu_mesh* meshes, size_t n_meshes,
    const ptgpu_tlas_instance* static_instances, size_t n_static);

/* ---- per frame: everything setup_animation_frame() produces (scene.cc:271-718) -------------- */

/* subframes[n_subframes]         s.subframes (scene.hh:65), one per 8 samples
 * dyn_instances[n_dyn]           s.instances[static_instance_count ..) — instance ids in TLAS
 *                                leaves are n_static + index into this array
 * tlas_nodes/links               s.bvh_buf.nodes/links from the first per-frame TLAS on, i.e.
 *                                nodes + tlas_node_base and links + 8*tlas_node_base;
 *                                subframe.tlas.node_offset is absolute (scene.cc:714)
 * The per-subframe TLAS link tables are used to recover which dynamic instances each subframe
 * sees (leaf payloads, scene.cc:84-88). */
int ptgpu_set_frame(
    ptgpu_ctx* ctx,
    const ptgpu_subframe* subframes, size_t n_subframes,
    const ptgpu_tlas_instance* dyn_instances, size_t n_dyn,
    const ptgpu_bvh_node* tlas_nodes, const ptgpu_bvh_link* tlas_links,
    size_t n_tlas_nodes, size_t tlas_node_base);

/* A subframe may see at most this many per-frame instances (shared prefix + its own range): the traversal
 * kernels keep them as one bit group / 16 stack entries. The reference animation uses at most 7
 * (scene.cc:634-674). More is an error, not a silent truncation. */
#define PTGPU_MAX_DYNAMIC_PER_SUBFRAME 16

/* Same without reference TLAS arrays: the caller states each subframe's dynamic range
 * [dyn_begin[i], dyn_end[i]) into dyn_instances (what scene.cc:651-678 keeps in `entries`).
 * subframes[i].tlas is ignored. */
int ptgpu_set_frame_ranges(
    ptgpu_ctx* ctx,
    const ptgpu_subframe* subframes, size_t n_subframes,
    const ptgpu_tlas_instance* dyn_instances, size_t n_dyn,
    const uint32_t* dyn_begin, const uint32_t* dyn_end);

/* ---- render --------------------------------------------------------------------------------- */

/* baseline_render(const scene&, uchar4* image) (main.cc:12-46): all cfg.spp samples of every
 * pixel, mean, tonemap_pixel. out_bgra: width*height*4 bytes, B,G,R,255, row 0 = top. */
int ptgpu_render(ptgpu_ctx* ctx, uint8_t* out_bgra);

/* baseline_render + write_bmp's packing (bmp.cc:15-52) fused: out_bmp receives the complete
 * file image (54-byte header, bottom-up BGR rows padded to 4 bytes); size from ptgpu_bmp_size. */
int ptgpu_render_bmp(ptgpu_ctx* ctx, uint8_t* out_bmp);
size_t ptgpu_bmp_size(const ptgpu_ctx* ctx);

/* ptgpu_set_frame + ptgpu_render in one call: the drop-in for main.cc:88. */
int ptgpu_render_frame(
    ptgpu_ctx* ctx,
    const ptgpu_subframe* subframes, size_t n_subframes,
    const ptgpu_tlas_instance* dyn_instances, size_t n_dyn,
    const ptgpu_bvh_node* tlas_nodes, const ptgpu_bvh_link* tlas_links,
    size_t n_tlas_nodes, size_t tlas_node_base,
    uint8_t* out_bgra);

/* Test hook: the loop nest of main.cc:16-43 over the pixel rectangle (x0,y0,w,h) and the sample
 * set {s_begin + k*s_stride : k < s_count}. out_rgb (may be NULL): w*h*3 floats, mean linear
 * radiance; out_bgra (may be NULL): w*h*4 bytes, tonemap_pixel of that mean. */
int ptgpu_render_rect(
    ptgpu_ctx* ctx, int32_t x0, int32_t y0, int32_t w, int32_t h,
    int32_t s_begin, int32_t s_count, int32_t s_stride,
    float* out_rgb, uint8_t* out_bgra);

/* path_trace_pixel(xy, sample_index, ...) (path_tracer.hh:637-741) for n independent
 * (x, y, sample_index) triples; out_rgb: n*3 floats. sample_index keeps the reference meaning:
 * RNG key (x, y, (uint)sample_index, STUDENT_ID) and subframe sample_index/8, negative -> 0. */
int ptgpu_trace_samples(
    ptgpu_ctx* ctx, const uint32_t* xy, const int32_t* sample_index, size_t n, float* out_rgb);

/* tonemap_pixel(float3) (path_tracer.hh:753-771) for n colours (3 floats each) -> n*4 bytes BGRA. */
int ptgpu_tonemap(ptgpu_ctx* ctx, const float* rgb, size_t n, uint8_t* out_bgra);

/* Closest-hit ray query (ray_query.hh:111-290 driven as in path_tracer.hh:342-349) for n rays
 * against subframe `subframe`: ray = {ox,oy,oz,tmin, dx,dy,dz,tmax} (8 floats);
 * out_f = {thit,u,v,w} (4 floats), out_u = {instance_id, primitive_id, back_face} per ray. */
int ptgpu_trace_closest(
    ptgpu_ctx* ctx, const float* rays, size_t n, uint32_t subframe, float* out_f, uint32_t* out_u);

/* pcg4d (math.hh:466-473) applied `steps` times to n 4-word states, in place. */
int ptgpu_pcg4d(ptgpu_ctx* ctx, uint32_t* states, size_t n, int32_t steps);

/* ---- device-resident rendering for pipelined drivers and the benchmark ---------------------- */

/* Launches the frame render on the context's stream and returns without waiting; the finished
 * frame stays in device memory. ptgpu_fetch_* waits and copies it out. */
int ptgpu_render_async(ptgpu_ctx* ctx);
int ptgpu_fetch_bgra(ptgpu_ctx* ctx, uint8_t* out_bgra);
int ptgpu_fetch_bmp(ptgpu_ctx* ctx, uint8_t* out_bmp);
int ptgpu_sync(ptgpu_ctx* ctx);

/* validator.py:41-52 on the device-resident frame of the last render: the frame is
 * box-downscaled by 2 (skimage.transform.downscale_local_mean, zero padded), truncated to 8 bits and
 * compared by PSNR (data range 255) with `ref_rgb_half`, the course's half-size reference image as
 * ceil(height/2) x ceil(width/2) x 3 bytes, RGB, row 0 = top. *good = 1 when PSNR >= 32 dB
 * (ACCEPT_MIN_PSNR). Only the half-size reference goes up and one 8-byte sum comes back, so a whole
 * animation can be validated without reading its frames back or a Python pass over 1800 files. */
int ptgpu_validate_frame(ptgpu_ctx* ctx, const uint8_t* ref_rgb_half, double* psnr, int32_t* good);
/* Device time (CUDA events on the context's stream) of the most recent finished render, in ms;
 * `launches` (may be NULL) receives the number of kernels it launched. */
int ptgpu_last_render_ms(ptgpu_ctx* ctx, float* ms, int32_t* launches);

/* ---- options and instrumentation ------------------------------------------------------------ */

/* "traversal": 0 = wide GPU BVH (default), 1 = walk the reference link tables as they are
 *              (stackless, ray_query.hh:184-223) — the counting / cross-check mode
 * "counters":  1 = count per-path events (rays, node visits, triangle tests, ...) on the
 *              reference link tables; only with traversal = 1
 * "kernel":    2 = wavefront (default), 0 = persistent megakernel, 1 = one-thread-per-path tile kernel
 * "bvh":       1 = compressed 8-wide BVH (default, wavefront only), 0 = 4-wide float BVH
 * "flat":      1 (default) = the static instances are traversed as ONE world-space BVH over all their
 *              triangles, built at upload (no instancing for the static part: 180 GB of HBM make the 15.6 M
 *              instanced triangles of the shipped scene a 1 GB array); 0 = static TLAS + instanced BLASes.
 *              Set it before the scene is uploaded. Per-frame instances are instances either way.
 * "sort":      1 (default) = the wavefront renderer sorts bounce and shadow rays by direction octant and
 *              origin cell before every traversal launch; 0 = queue order.
 * "dyn_first": 1 (default) = with the flat scene a query enters the per-frame instances (the hero objects in
 *              front of the camera) before the static world; 0 = after.
 * "l2_persist": percent (0..100, default 0) of the device's persisting-L2 maximum set aside for the compressed BVH nodes
 *              through an access-policy window on the render stream (hits persist, everything else streams).
 * "plain_trace": 1 = every round, 2 = the primary round of the wavefront renderer is traced by the plain single-ray
 *              loop, one thread per ray, instead of the warp-scheduled kernel (reference point for measurements; default 0).
 * "top_smem":  1 = the traversal kernel stages the top levels of the flat BVH in shared memory (default 0:
 *              measured 0-3 % slower, profiles/r02_trace_kernel_history.md).
 * "lanes", "pool_budget_mb": path slots per pixel of the wavefront pool (power of two) and its budget
 * "min_active", "node_threshold", "node_burst", "tri_threshold", "xform_threshold": warp scheduling
 *              of the traversal kernels (see csrc/pt_wave.cuh); results do not depend on them
 * "validate":  1 = cross-check every traversal query (debug, slow), see ptgpu_get_stat */
int ptgpu_set_option(ptgpu_ctx* ctx, const char* key, int64_t value);

enum {
    PTGPU_CNT_PATHS = 0, PTGPU_CNT_RAYS, PTGPU_CNT_NODE_VISITS, PTGPU_CNT_TRI_TESTS,
    PTGPU_CNT_BLAS_ENTERS, PTGPU_CNT_BOUNCES, PTGPU_CNT_SHADOW_RAYS, PTGPU_CNT_SKY_MARCHES,
    PTGPU_CNT_SKY_ATTENUATIONS, PTGPU_CNT_HITS, PTGPU_CNT_MISSES, PTGPU_CNT_COUNT = 16
};
/* Reads and clears the event counters accumulated since the last call. */
int ptgpu_read_counters(ptgpu_ctx* ctx, uint64_t out[PTGPU_CNT_COUNT]);

/* Sizes of the device-side scene, for reporting: out = {static bytes, per-frame bytes,
 * wide-BVH nodes, triangles, static instances}. */
int ptgpu_scene_stats(ptgpu_ctx* ctx, uint64_t out[8]);

/* ---- per-frame scene state without setup_animation_frame ------------------------ */

/* The reference's setup_animation_frame (scene.cc:271-718) replays a keyframe table per motion-blur
 * subframe, appends the dynamic instances and builds one SAH TLAS per subframe (75 ms per frame on one
 * core). This library never needs those TLASes, so the per-frame host work is only the replay: these
 * entry points restate it (csrc/frame_setup.cu) on top of DATA the caller hands in — the rows of the
 * reference's `animation_stop anim[]` table (scene.cc:24-31, 319-627; the float* target replaced by a
 * variable index) and the (mesh, bvh) handles of the seven dynamic meshes (scene.hh:49). */
enum ptgpu_anim_var {
    PTGPU_VAR_LOGO_VISIBLE, PTGPU_VAR_ARMADILLO_VISIBLE, PTGPU_VAR_DRAGON_VISIBLE, PTGPU_VAR_BUNNY_VISIBLE, PTGPU_VAR_END_VISIBLE,
    PTGPU_VAR_CAM_POS_X, PTGPU_VAR_CAM_POS_Y, PTGPU_VAR_CAM_POS_Z, PTGPU_VAR_CAM_ORI_X, PTGPU_VAR_CAM_ORI_Y, PTGPU_VAR_CAM_ORI_Z,
    PTGPU_VAR_FOV, PTGPU_VAR_FOCAL_DISTANCE, PTGPU_VAR_APERTURE_RADIUS,
    PTGPU_VAR_TEAPOT_POS_X, PTGPU_VAR_TEAPOT_POS_Y, PTGPU_VAR_TEAPOT_POS_Z, PTGPU_VAR_TEAPOT_ORI_X, PTGPU_VAR_TEAPOT_ORI_Y, PTGPU_VAR_TEAPOT_ORI_Z,
    PTGPU_VAR_ARMADILLO_POS_X, PTGPU_VAR_ARMADILLO_POS_Y, PTGPU_VAR_ARMADILLO_POS_Z, PTGPU_VAR_ARMADILLO_ORI_X, PTGPU_VAR_ARMADILLO_ORI_Y, PTGPU_VAR_ARMADILLO_ORI_Z,
    PTGPU_VAR_DRAGON_POS_X, PTGPU_VAR_DRAGON_POS_Y, PTGPU_VAR_DRAGON_POS_Z, PTGPU_VAR_DRAGON_ORI_X, PTGPU_VAR_DRAGON_ORI_Y, PTGPU_VAR_DRAGON_ORI_Z,
    PTGPU_VAR_BUNNY_POS_X, PTGPU_VAR_BUNNY_POS_Y, PTGPU_VAR_BUNNY_POS_Z, PTGPU_VAR_BUNNY_ORI_X, PTGPU_VAR_BUNNY_ORI_Y, PTGPU_VAR_BUNNY_ORI_Z,
    PTGPU_VAR_END_POS_X, PTGPU_VAR_END_POS_Y, PTGPU_VAR_END_POS_Z, PTGPU_VAR_END_ORI_X, PTGPU_VAR_END_ORI_Y, PTGPU_VAR_END_ORI_Z,
    PTGPU_ANIM_VAR_COUNT
};
enum ptgpu_anim_mesh {
    PTGPU_MESH_LOGO, PTGPU_MESH_BUDDHA, PTGPU_MESH_TEAPOT, PTGPU_MESH_ARMADILLO, PTGPU_MESH_DRAGON, PTGPU_MESH_BUNNY, PTGPU_MESH_END,
    PTGPU_MESH_COUNT
};
typedef struct { float start, duration, from, to; int32_t var; } ptgpu_anim_key;   /* animation_stop, scene.cc:24-31 */
typedef struct { ptgpu_mesh m; ptgpu_bvh blas; } ptgpu_mesh_handle;                /* scene::meshes entry, scene.hh:49 */
typedef struct ptgpu_anim ptgpu_anim;

int ptgpu_anim_create(ptgpu_anim** out, const ptgpu_anim_key* keys, size_t n_keys,
                      const ptgpu_mesh_handle* meshes /* [PTGPU_MESH_COUNT] */, const ptgpu_config* cfg);
void ptgpu_anim_destroy(ptgpu_anim* anim);
size_t ptgpu_anim_subframe_count(const ptgpu_anim* anim);   /* ceil(spp / 8), scene.cc:648-650 */
size_t ptgpu_anim_max_instances(const ptgpu_anim* anim);    /* capacity needed for `dyn` below */
uint32_t ptgpu_anim_frame_count(const ptgpu_anim* anim);    /* get_animation_frame_count, scene.cc:720-724 */
/* What setup_animation_frame(s, frame) leaves in s.subframes and s.instances[static_instance_count..),
 * minus the TLAS handles: subframes[subframe_count] (tlas zeroed), dyn[*n_dyn], and each subframe's
 * range [dyn_begin[i], dyn_end[i]) of per-subframe instances (scene.cc:651-678); the instances before
 * the first range are the frame-static extras (logo, buddha). Pure function of `frame`: re-entrant. */
int ptgpu_anim_frame(const ptgpu_anim* anim, uint32_t frame, ptgpu_subframe* subframes,
                     ptgpu_tlas_instance* dyn, size_t* n_dyn, uint32_t* dyn_begin, uint32_t* dyn_end);
/* ptgpu_anim_frame + ptgpu_set_frame_ranges */
int ptgpu_set_animation_frame(ptgpu_ctx* ctx, const ptgpu_anim* anim, uint32_t frame);

/* Facts about the most recent wavefront render: "wave_rounds", "wave_lanes" (path slots per pixel),
 * "pool_bytes" (path-state pool), "validate_mismatches" (with option "validate" = 1 every ray of
 * every round is re-traced with the plain single-ray traversal and compared with what the scheduled
 * traversal kernel stored; the number of disagreeing queries), "trace_us" / "shade_us" (device time
 * of the traversal launches, and of the classify + shade launches, of that frame, from CUDA events
 * recorded on the render stream around them; first 48 rounds), "trace_launches" (how many
 * traversal launches "trace_us" covers), "sort_us" (ray sort + camera-ray generation of the same rounds).
 * About the uploaded scene: "flat_tris", "flat_nodes", "flat_depth", "flat_build_ms" (0 without a flat scene). */
int ptgpu_get_stat(ptgpu_ctx* ctx, const char* key, uint64_t* out);

/* ---- sub-function evaluator (parity tests) ----------------------------------------------------- */

/* One device function of the path per item, on caller-supplied inputs: each can be compared directly with
 * the reference function it restates instead of only through whole paths. Item i reads in[24*i ..] and
 * writes out[32*i ..] (floats; uint32 values as the bits of a float). Functions that look at the scene
 * need an uploaded scene and frame. */
enum ptgpu_fn
{
    PTGPU_FN_RAND4 = 0,           /* generate_uniform_random4, math.hh:475-485. in: state[4] (bits) -> state[4] (bits), u[4] */
    PTGPU_FN_FILM_OFFSET = 1,     /* sample_gaussian_weighted_disk(u, 0.4), path_tracer.hh:19-25. in: u[2] -> offset[2] */
    PTGPU_FN_CAMERA_RAY = 2,      /* get_camera_ray, path_tracer.hh:429-450. in: u[2], coord[2], subframe -> dir[3], origin[3] */
    PTGPU_FN_GGX_VNDF = 3,        /* sample_ggx_vndf, path_tracer.hh:67-83. in: view[3], roughness, u[2] -> h[3] */
    PTGPU_FN_BSDF = 4,            /* bsdf, path_tracer.hh:184-222. in: light[3], view[3], albedo[3], roughness, metallic,
                                   * transmission, eta -> attenuation[3], pdf */
    PTGPU_FN_SAMPLE_BSDF = 5,     /* sample_bsdf, path_tracer.hh:224-296. in: u[3], view[3], albedo[3], roughness, metallic,
                                   * transmission, eta -> dir[3], attenuation[3], pdf (negative: delta lobe) */
    PTGPU_FN_SKY_ATTENUATION = 6, /* nishita_atmosphere_attenuation(jitter, 8, pos, view, 1e9), path_tracer.hh:456-497.
                                   * in: jitter, pos[3], view[3] -> attenuation[3] */
    PTGPU_FN_SKY_SCATTERING = 7,  /* nishita_atmosphere_scattering, path_tracer.hh:499-588. in: seed[4] (bits), light dir[3],
                                   * light color[3], cos_solid_angle, pos[3], view[3], tmax -> attenuation[3], in_scatter[3], seed[4] */
    PTGPU_FN_SAMPLE_CONE = 8,     /* sample_cone, path_tracer.hh:40-48. in: dir[3], cos_theta_min, u[2] -> dir[3] */
    PTGPU_FN_SHADOW_RAY = 9,      /* trace_shadow_ray, path_tracer.hh:415-427. in: origin[3], dir[3], tmin, tmax, subframe -> occluded */
    PTGPU_FN_TRACE_RAY = 10,      /* trace_ray, path_tracer.hh:340-412. in: origin[3], dir[3], tmin, subframe -> thit, pos[3],
                                   * tbn columns[9], albedo[3], roughness, metallic, emission, transmission, eta, nee_pdf */
    PTGPU_FN_COUNT = 11
};
int ptgpu_debug_eval(ptgpu_ctx* ctx, int32_t fn, const float* in, size_t n, float* out);

/* ---- OBJ/MTL loader: load_mesh, mesh.cc:104-265, without the reference's mesh.cc ----- */

/* A growing set of mesh buffers = `mesh_buffers` (mesh.hh:31-43): every ptgpu_meshes_load_obj appends one
 * mesh's indices and de-duplicated vertices and returns its handle (`mesh`, mesh.hh:18-28). Same parsing
 * rules, vertex numbering and attribute packing as the reference, so the arrays can be handed to
 * ptgpu_upload_meshes (or to the reference's own build_blas). Host-only, needs no GPU. Errors (missing OBJ
 * or MTL file) return non-zero with ptgpu_meshes_last_error(); the reference prints and exits (mesh.cc:25-29). */
typedef struct ptgpu_mesh_set ptgpu_mesh_set;
int ptgpu_meshes_create(ptgpu_mesh_set** out);
void ptgpu_meshes_destroy(ptgpu_mesh_set* set);
const char* ptgpu_meshes_last_error(const ptgpu_mesh_set* set);
int ptgpu_meshes_load_obj(ptgpu_mesh_set* set, const char* obj_path, ptgpu_mesh* out_mesh);
size_t ptgpu_meshes_index_count(const ptgpu_mesh_set* set);
size_t ptgpu_meshes_vertex_count(const ptgpu_mesh_set* set);
const uint32_t* ptgpu_meshes_indices(const ptgpu_mesh_set* set);      /* valid until the next load / destroy */
const ptgpu_float3* ptgpu_meshes_pos(const ptgpu_mesh_set* set);
const ptgpu_float3* ptgpu_meshes_normal(const ptgpu_mesh_set* set);
const ptgpu_float4* ptgpu_meshes_albedo(const ptgpu_mesh_set* set);
const ptgpu_float4* ptgpu_meshes_material(const ptgpu_mesh_set* set);

/* Host-only, needs no GPU: runs the BVH flattening that ptgpu_upload_static performs on the
 * reference arrays (bvh.cc:43-229 output) and checks the result structurally (every triangle
 * reachable exactly once, boxes nested, every static instance in the TLAS once).
 * out = {BLAS count, wide nodes, triangles, TLAS nodes, stack bound, violations, stack capacity, 0}. */
int ptgpu_host_flatten_check(
    const ptgpu_bvh_node* nodes, size_t n_nodes, const ptgpu_bvh_link* links, size_t n_links,
    const uint32_t* indices, size_t n_indices, const ptgpu_float3* pos, size_t n_verts,
    const ptgpu_tlas_instance* instances, size_t n_static, uint64_t out[8], char* err, size_t err_len);
/* Host-only: builds the flat static scene (all static instances as world-space triangles under ONE
 * compressed 8-wide BVH, no instancing; replaces the per-subframe build_tlas of bvh.cc:252-284 for the
 * static part) from the same arrays and checks it structurally.
 * out = {triangles, nodes, top-tree nodes, depth, violations, build milliseconds, 0, 0}. */
int ptgpu_host_flat_check(
    const ptgpu_bvh_node* nodes, size_t n_nodes, const ptgpu_bvh_link* links, size_t n_links,
    const uint32_t* indices, size_t n_indices, const ptgpu_float3* pos, size_t n_verts,
    const ptgpu_tlas_instance* instances, size_t n_static, uint64_t out[8], char* err, size_t err_len);
/* Host-only, needs no GPU and no context: does the CPU work of ptgpu_upload_static (BVH flattening and, with
 * flat != 0, the flat static scene: seconds for the shipped scene) and keeps the result in the process-wide
 * cache, where every later ptgpu_upload_static of the same arrays finds it. A multi-GPU driver calls it right
 * after load_scene(), while its worker threads are still creating their CUDA contexts. */
int ptgpu_host_prepare_static(
    const ptgpu_bvh_node* nodes, size_t n_nodes, const ptgpu_bvh_link* links, size_t n_links,
    const uint32_t* indices, size_t n_indices, const ptgpu_float3* pos, size_t n_verts,
    const ptgpu_tlas_instance* instances, size_t n_static, int32_t flat, char* err, size_t err_len);
/* Host-only: the check ptgpu_set_frame_ranges applies to the per-subframe dynamic sets (ranges inside
 * [0, n_dyn], at most PTGPU_MAX_DYNAMIC_PER_SUBFRAME instances per subframe). 0 = accepted. */
int ptgpu_host_check_dynamic_ranges(const uint32_t* dyn_begin, const uint32_t* dyn_end, size_t n_subframes, size_t n_dyn,
                                    char* err, size_t err_len);
/* The same for ptgpu_upload_meshes' own BLAS builder. */
int ptgpu_host_build_check(
    const uint32_t* indices, size_t n_indices, const ptgpu_float3* pos, size_t n_verts,
    const ptgpu_mesh* meshes, size_t n_meshes,
    const ptgpu_tlas_instance* instances, size_t n_static, uint64_t out[8], char* err, size_t err_len);

#ifdef __cplusplus
}
#endif
#endif /* PTGPU_H */
