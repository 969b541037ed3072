// C ABI of libptgpu.so (include/ptgpu.h): context, device memory, uploads, launches.
#include "../../include/ptgpu.h"
#include "pt_kernels.cuh"
#include "pt_wide.cuh"
#include "pt_mega.cuh"
#include "pt_wave.cuh"
#include "pt_cwbvh.cuh"
#include "pt_debug.cuh"
#include "bvh_wide.hh"

#include <algorithm>
#include <cfloat>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

static_assert(sizeof(ptgpu_tlas_instance) == 160, "tlas_instance must be 160 B (bvh.hh:73-79)");
static_assert(sizeof(ptgpu_subframe) == 160, "subframe must be 160 B (scene.hh:27-35)");
static_assert(sizeof(ptgpu_camera) == 96, "camera must be 96 B (scene.hh:7-18)");
static_assert(sizeof(ptgpu_directional_light) == 48, "directional_light must be 48 B (scene.hh:20-25)");
static_assert(sizeof(ptgpu_bvh_node) == 24 && sizeof(ptgpu_bvh_link) == 8 && sizeof(ptgpu_mesh) == 16, "bvh PODs");
static_assert(offsetof(ptgpu_tlas_instance, transform) == 32 && offsetof(ptgpu_tlas_instance, inv_transform) == 96, "instance offsets");
static_assert(offsetof(ptgpu_subframe, cam) == 16 && offsetof(ptgpu_subframe, light) == 112, "subframe offsets");

using namespace pt;

namespace {

thread_local std::string g_create_error;

template<class T>
struct DevBuf
{
    T* p = nullptr;
    size_t cap = 0; // elements
    cudaError_t reserve(size_t n)
    {
        if(n <= cap) return cudaSuccess;
        if(p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if(e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if(p) cudaFree(p); p = nullptr; cap = 0; }
};

// The flat static scene takes seconds to build (15.6 M triangles at the shipped scene) and depends only on
// the uploaded arrays: contexts of one process that upload the same scene (one per GPU in the driver, one
// per test in the test-suite) share one build.
uint64_t hash_bytes(uint64_t h, const void* p, size_t n)
{
    const uint64_t* w = (const uint64_t*)p;
    for(size_t i = 0; i < n / 8; ++i) { h ^= w[i]; h *= 0x100000001B3ull; h ^= h >> 29; }
    const uint8_t* b = (const uint8_t*)p + (n & ~size_t(7));
    for(size_t i = 0; i < (n & 7); ++i) { h ^= b[i]; h *= 0x100000001B3ull; }
    return h;
}
// Host side of an uploaded static scene: the compressed BVHs of the meshes (+ static TLAS) and, with option
// "flat", the flat static scene. Built once per process and set of input arrays.
struct HostScene
{
    WideScene wide;
    FlatScene flat;
    bool have_flat = false;
};
std::mutex g_scene_mutex;
std::map<uint64_t, std::shared_ptr<const HostScene>> g_scene_cache;

std::shared_ptr<const HostScene> get_host_scene(
    const ptgpu_bvh_node* nodes, size_t n_nodes, const ptgpu_bvh_link* links, size_t n_links,
    const uint32_t* indices, size_t n_indices, const ptgpu_float3* pos, size_t n_verts,
    const ptgpu_tlas_instance* instances, size_t n_static, const ptgpu_mesh* meshes, size_t n_meshes,
    bool want_flat, std::string& err)
{
    uint64_t h = 0xCBF29CE484222325ull;
    if(nodes) { h = hash_bytes(h, nodes, n_nodes * sizeof(ptgpu_bvh_node)); h = hash_bytes(h, links, n_links * sizeof(ptgpu_bvh_link)); }
    if(meshes) h = hash_bytes(h, meshes, n_meshes * sizeof(ptgpu_mesh));
    h = hash_bytes(h, instances, n_static * sizeof(ptgpu_tlas_instance));
    h = hash_bytes(h, indices, n_indices * 4);
    h = hash_bytes(h, pos, n_verts * sizeof(ptgpu_float3));
    const uint64_t flag = want_flat ? 1 : 0;
    h = hash_bytes(h, &flag, 8);
    std::lock_guard<std::mutex> lock(g_scene_mutex);   // a second context waits for the first one's build
    auto it = g_scene_cache.find(h);
    if(it != g_scene_cache.end()) return it->second;
    auto hs = std::make_shared<HostScene>();
    if(!build_wide_scene(nodes, n_nodes, links, indices, n_indices, pos, n_verts, instances, n_static, hs->wide, err, meshes, n_meshes))
    { err = "wide BVH build failed: " + err; return nullptr; }
    if(want_flat)
    {
        const uint32_t node_base = (uint32_t)(hs->wide.cw_nodes.size() / 5), tri_base = (uint32_t)(hs->wide.cw_tris.size() / 3);
        if(!build_flat_scene(hs->wide, indices, pos, instances, n_static, node_base, tri_base, hs->flat, err))
        { err = "flat scene build failed: " + err; return nullptr; }
        if(2 * hs->flat.depth + 8 > (uint32_t)CW_STACK)
        { err = "flat scene: BVH depth " + std::to_string(hs->flat.depth) + " needs a deeper traversal stack than CW_STACK"; return nullptr; }
        hs->have_flat = true;
    }
    if(g_scene_cache.size() >= 3) g_scene_cache.erase(g_scene_cache.begin());
    g_scene_cache[h] = hs;
    return hs;
}

} // namespace

struct ptgpu_ctx
{
    int device = 0;
    ptgpu_config cfg{};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    std::string error;
    int sm_count = 0;

    // options
    int traversal = 0; // 0 wide, 1 links
    int counters_on = 0;
    int kernel = 2;    // 0 megakernel, 1 simple tiles, 2 wavefront
    int min_active = -1; // -1: per-kernel default
    int bvh = 1;         // 0: 4-wide float BVH, 1: compressed 8-wide BVH
    int validate = 0;                          // debug: re-trace every ray with the plain traversal and compare
    int max_lanes = 256;                       // wavefront: slots per pixel (power of two)
    size_t pool_budget_bytes = 16ull << 30;    // wavefront: path-state pool budget
    // warp scheduling of wf_trace_cw (-1 = by scene kind: instanced 4 / 2, flat 1 / 3 — with the flat scene
    // instance entries are rare (the <= 7 per-frame objects), so a lane should not wait for company to enter one,
    // and longer node bursts pay because no ENTER step competes; profiles/r02_trace_kernel_history.md)
    int tri_threshold = 8, xform_threshold = -1, node_threshold = 16, node_burst = -1;
    int flat = 1;                              // static instances as one world-space BVH (built at upload)
    int sort = 1;                              // wavefront: bounce and shadow rays sorted by octant + origin cell
    int top_smem = 0;                          // wavefront: top levels of the flat BVH staged in shared memory
    int plain_trace = 0;                       // wavefront: 1 = every round, 2 = the primary round traced by the plain single-ray loop (reference point)
    int l2_persist = 0;                        // percent of the persisting-L2 maximum set aside for the BVH nodes (0 = no access-policy window)
    int l2_persist_applied = -1;               // what the render stream currently carries
    size_t stat_l2_set_aside = 0, stat_l2_window = 0;
    int dyn_first = 1;                         // flat scene: per-frame instances are entered before the static world

    // static scene, reference layout
    DevBuf<float2> ref_nodes;     // 3 per node; static region then per-frame TLAS region
    DevBuf<uint2> ref_links;
    DevBuf<uint32_t> indices;
    DevBuf<float4> pos, normal, albedo, material;
    DevBuf<float4> shade_tris;      // 9 per triangle of the index buffer (pt_scene.cuh)
    DevBuf<RefInstance> instances; // static + dynamic
    size_t n_static_nodes = 0, n_static = 0, n_verts = 0, n_indices = 0;
    bool have_static = false;
    bool have_ref_bvh = false;      // false after ptgpu_upload_meshes: no reference nodes/links on the device
    std::vector<ptgpu_tlas_instance> host_static; // kept for the wide instance records

    // wide layout
    std::shared_ptr<const HostScene> host;   // host-side build results (shared by the contexts of a process)
    DevBuf<WideNode> wnodes;
    DevBuf<float4> wtris;
    DevBuf<WideBlas> wblas;
    DevBuf<WideInstance> winst;     // static + dynamic
    DevBuf<WideNode> wtlas;
    size_t n_wtlas = 0;
    DevBuf<float4> cwnodes, cwtris;   // BLASes + static TLAS, then (flat scene) the world-space BVH of all static instances
    DevBuf<uint32_t> cw_inst_index;
    bool have_flat = false;
    uint32_t flat_root = 0xFFFFFFFFu, flat_top = 0, flat_depth = 0;
    size_t flat_tris = 0, flat_nodes = 0;
    double flat_build_seconds = 0.0;
    float key_lo[3] = {0, 0, 0}, key_scale[3] = {0, 0, 0};

    // per frame
    DevBuf<RefSubframe> subframes;
    DevBuf<uint2> dyn_range;
    DevBuf<float4> dyn_union;       // per subframe: world box (lo, hi) around all per-frame instances it sees
    size_t n_subframes = 0, n_dyn = 0;
    bool have_frame = false, frame_has_ref_tlas = false;
    std::vector<uint8_t> staging;   // pinned-size-agnostic host staging for per-frame uploads
    void* pinned = nullptr; size_t pinned_cap = 0;

    // outputs
    DevBuf<uchar4> out_bgra;
    DevBuf<uint8_t> out_bmp;
    DevBuf<float> out_rgb;
    DevBuf<Counters> counters;
    DevBuf<MegaState> mega_state;
    // wavefront pool
    DevBuf<uint8_t> wave_mem;
    DevBuf<uint32_t> wave_flag;
    uint32_t* wave_flag_host = nullptr;
    int last_wave_rounds = 0;
    uint32_t last_wave_lanes = 0;
    size_t last_pool_bytes = 0;
    // per-kernel device time of the last wavefront frame (CUDA events on the render stream)
    std::vector<cudaEvent_t> wave_events;
    double last_trace_us = 0.0, last_shade_us = 0.0, last_sort_us = 0.0;
    uint64_t last_trace_launches = 0;
    int wave_timed_rounds = 0;
    unsigned long long last_validate_mismatches = 0;
    uint32_t bmp_pitch = 0;
    bool bmp_header_done = false;
    bool render_pending = false;        // some render was launched (ptgpu_last_render_ms has something to time)
    bool frame_on_device = false;       // out_bgra / out_bmp hold a full frame (ptgpu_validate_frame, ptgpu_fetch_*)
    DevBuf<uchar4> rect_bgra;           // ptgpu_render_rect's own image: a rect never overwrites the frame
    int last_launches = 0;

    // scratch for the small utility entry points
    DevBuf<uint8_t> scratch_a, scratch_b, scratch_c;
};

namespace {

int fail(ptgpu_ctx* c, const char* fmt, ...)
{
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    if(c) c->error = buf; else g_create_error = buf;
    return 1;
}

#define CK(call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) \
    return fail(ctx, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while(0)

int use(ptgpu_ctx* ctx)
{
    CK(cudaSetDevice(ctx->device));
    return 0;
}

void* pinned_staging(ptgpu_ctx* ctx, size_t bytes)
{
    if(bytes > ctx->pinned_cap)
    {
        if(ctx->pinned) cudaFreeHost(ctx->pinned);
        ctx->pinned = nullptr; ctx->pinned_cap = 0;
        size_t cap = bytes + bytes / 2 + 4096;
        if(cudaMallocHost(&ctx->pinned, cap) != cudaSuccess) return nullptr;
        ctx->pinned_cap = cap;
    }
    return ctx->pinned;
}

Scene make_scene(ptgpu_ctx* ctx)
{
    Scene s{};
    s.ref_nodes = ctx->ref_nodes.p;
    s.ref_links = ctx->ref_links.p;
    s.indices = ctx->indices.p;
    s.pos = ctx->pos.p; s.normal = ctx->normal.p; s.albedo = ctx->albedo.p; s.material = ctx->material.p;
    s.shade_tris = ctx->shade_tris.p;
    s.instances = ctx->instances.p;
    s.subframes = ctx->subframes.p;
    s.wnodes = ctx->wnodes.p; s.wtris = ctx->wtris.p; s.wblas = ctx->wblas.p; s.winst = ctx->winst.p;
    s.wtlas = ctx->wtlas.p; s.dyn_range = ctx->dyn_range.p; s.dyn_union = ctx->dyn_union.p;
    s.cwnodes = ctx->cwnodes.p; s.cwtris = ctx->cwtris.p; s.cw_inst_index = ctx->cw_inst_index.p;
    s.cw_tlas_root = ctx->host ? ctx->host->wide.cw_tlas_root : 0u;
    s.flat_root = ctx->flat && ctx->have_flat ? ctx->flat_root : 0xFFFFFFFFu;
    s.flat_top = ctx->flat_top;
    s.dyn_first = (uint32_t)ctx->dyn_first;
    s.cw_magic = 0x47000000u;
    for(int a = 0; a < 3; ++a) { s.key_lo[a] = ctx->key_lo[a]; s.key_scale[a] = ctx->key_scale[a]; }
    s.n_static = (uint32_t)ctx->n_static;
    s.n_subframes = (uint32_t)ctx->n_subframes;
    s.width = ctx->cfg.width; s.height = ctx->cfg.height;
    s.max_bounces = ctx->cfg.max_bounces;
    s.samples_per_subframe = ctx->cfg.samples_per_subframe;
    s.student_id = ctx->cfg.student_id;
    return s;
}

int check_ready(ptgpu_ctx* ctx)
{
    if(!ctx) return 1;
    if(!ctx->have_static) return fail(ctx, "ptgpu_upload_static has not been called");
    if(!ctx->have_frame) return fail(ctx, "ptgpu_set_frame has not been called");
    if(ctx->traversal == 1 && !ctx->frame_has_ref_tlas)
        return fail(ctx, "traversal=links needs the reference TLAS arrays (use ptgpu_set_frame, not _ranges)");
    return 0;
}

// Sums the per-round events of the last wavefront frame (waits for the frame).
void wave_timing(ptgpu_ctx* ctx)
{
    if(ctx->wave_timed_rounds <= 0) return;
    cudaEventSynchronize(ctx->wave_events[4 * (ctx->wave_timed_rounds - 1) + 3]);
    ctx->last_trace_us = ctx->last_shade_us = ctx->last_sort_us = 0.0;
    for(int i = 0; i < ctx->wave_timed_rounds; ++i)
    {
        float a = 0.f, b = 0.f, c = 0.f;
        cudaEventElapsedTime(&c, ctx->wave_events[4 * i], ctx->wave_events[4 * i + 1]);
        cudaEventElapsedTime(&a, ctx->wave_events[4 * i + 1], ctx->wave_events[4 * i + 2]);
        cudaEventElapsedTime(&b, ctx->wave_events[4 * i + 2], ctx->wave_events[4 * i + 3]);
        ctx->last_trace_us += 1e3 * a; ctx->last_shade_us += 1e3 * b; ctx->last_sort_us += 1e3 * c;
    }
    ctx->last_trace_launches = (uint64_t)ctx->wave_timed_rounds;
    ctx->wave_timed_rounds = 0;
}

// Carves the wavefront pool out of one allocation and runs rounds of generate / trace / shade until
// no slot has a ray or a sample left. Returns kernels launched, or -1.
#ifndef WF_SHADE_GRID
#define WF_SHADE_GRID 8   // blocks per SM of the shade kernels (4 resident at 106 registers: two waves)
#endif
// L2 access-policy window over the compressed BVH nodes on the render stream: the traversal kernel's node loads
// keep their lines as "persisting" while the 11.9 GB path-state pool streams through the same 126 MB L2
// (option "l2_persist" = percent of cudaDevAttrMaxPersistingL2CacheSize to set aside; 0 = off).
static void apply_l2_policy(ptgpu_ctx* ctx)
{
    if(ctx->l2_persist_applied == ctx->l2_persist) return;
    ctx->l2_persist_applied = ctx->l2_persist;
    int dev = 0, max_persist = 0, max_window = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    const size_t node_bytes = ctx->cwnodes.cap * sizeof(float4);
    if(ctx->l2_persist <= 0 || max_persist <= 0 || max_window <= 0 || node_bytes == 0)
    {
        attr.accessPolicyWindow.num_bytes = 0;   // disables the window
        cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
        cudaCtxResetPersistingL2Cache();
        cudaGetLastError();
        return;
    }
    const size_t set_aside = (size_t)max_persist / 100 * (size_t)ctx->l2_persist;
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, set_aside);
    // the flat scene's nodes are the tail of the array: the window covers the end if the array is larger than a window
    const size_t win = node_bytes < (size_t)max_window ? node_bytes : (size_t)max_window;
    attr.accessPolicyWindow.base_ptr = (char*)ctx->cwnodes.p + (node_bytes - win);
    attr.accessPolicyWindow.num_bytes = win;
    attr.accessPolicyWindow.hitRatio = set_aside >= win ? 1.0f : (float)((double)set_aside / (double)win);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
    ctx->stat_l2_set_aside = set_aside; ctx->stat_l2_window = win;
    cudaGetLastError();
}

int launch_wave(ptgpu_ctx* ctx, const Scene& sc, RenderJob job)
{
    apply_l2_policy(ctx);
    const int tiles_x = (job.w + WF_TILE - 1) / WF_TILE, tiles_y = (job.h + WF_TILE - 1) / WF_TILE;
    // slots per pixel: the largest power of two <= the sample count that keeps the pool in budget
    const size_t n_pix_padded = (size_t)tiles_x * tiles_y * WF_TILE * WF_TILE;
    const size_t bytes_per_slot = 16 * 9 + 4 + 8 + 4 + 10 * 4 + 1;
    uint32_t lanes = 1, lane_shift = 0;
    while(lanes * 2 <= (uint32_t)job.s_count && lanes * 2 <= (uint32_t)ctx->max_lanes &&
          n_pix_padded * (lanes * 2) * bytes_per_slot <= ctx->pool_budget_bytes && n_pix_padded * (lanes * 2) < 0x7FFF0000ull)
    { lanes *= 2; lane_shift++; }
    const uint32_t n_slots = (uint32_t)(n_pix_padded * lanes);
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
    const size_t o_rng = carve(16ull * n_slots), o_ro = carve(16ull * n_slots), o_rd = carve(16ull * n_slots),
        o_sd = carve(16ull * n_slots), o_hit = carve(16ull * n_slots), o_prim = carve(4ull * n_slots),
        o_att = carve(16ull * n_slots), o_con = carve(16ull * n_slots), o_nee = carve(16ull * n_slots),
        o_sum = carve(16ull * n_slots), o_cur = carve(8ull * n_slots), o_vis = carve(4ull * n_slots),
        o_qt = carve(4ull * 3ull * (n_slots + 64)), o_qk = carve(4ull * 3ull * (n_slots + 64)), o_qs = carve(4ull * 2ull * (n_slots + 64)),
        o_sh = carve(4ull * 2ull * WF_SORT_BINS), o_st = carve(n_slots), o_qf = carve(4ull * (n_slots + 64)), o_qn = carve(4ull * (n_slots + 64)), o_cnt = carve(sizeof(WaveCounters)), o_stats = carve(40 * 8);
    if(ctx->wave_mem.reserve(off) != cudaSuccess)
    {
        size_t free_b = 0, total_b = 0;
        cudaGetLastError();
        cudaMemGetInfo(&free_b, &total_b);
        fail(ctx, "the path-state pool of %zu MB (%u slots per pixel) does not fit: %zu MB of device memory free; "
                  "lower option pool_budget_mb or lanes", off >> 20, lanes, free_b >> 20);
        return -2;
    }
    if(!ctx->wave_flag_host)
    {
        if(cudaMallocHost(&ctx->wave_flag_host, 64) != cudaSuccess) return -1;
        if(ctx->wave_flag.reserve(16) != cudaSuccess) return -1;
    }
    uint8_t* m = ctx->wave_mem.p;
    WaveBuffers wb{};
    wb.rng = (uint4*)(m + o_rng); wb.ray_o = (float4*)(m + o_ro); wb.ray_d = (float4*)(m + o_rd);
    wb.shadow_d = (float4*)(m + o_sd); wb.hit = (float4*)(m + o_hit); wb.hit_prim = (uint32_t*)(m + o_prim);
    wb.atten = (float4*)(m + o_att); wb.contrib = (float4*)(m + o_con); wb.nee = (float4*)(m + o_nee);
    wb.sum = (float4*)(m + o_sum); wb.cursor = (int2*)(m + o_cur); wb.visible = (uint32_t*)(m + o_vis);
    wb.q_trace = (uint32_t*)(m + o_qt); wb.status = m + o_st;
    wb.q_key = (uint32_t*)(m + o_qk); wb.q_sorted = (uint32_t*)(m + o_qs); wb.sort_hist = (uint32_t*)(m + o_sh);
    wb.sort = ctx->sort && ctx->bvh == 1 ? 1 : 0;
    wb.q_far = (uint32_t*)(m + o_qf); wb.q_near = (uint32_t*)(m + o_qn);
    wb.cnt = (WaveCounters*)(m + o_cnt);
    wb.stats = (unsigned long long*)(m + o_stats);
    wb.n_slots = n_slots; wb.seg_cap = n_slots + 64; wb.tiles_x = tiles_x;
    wb.lanes = lanes; wb.lane_shift = lane_shift;
    if(job.min_active < 1) job.min_active = 1;
    const bool flat_scene = sc.flat_root != 0xFFFFFFFFu;
    job.tri_threshold = ctx->tri_threshold; job.xform_threshold = ctx->xform_threshold > 0 ? ctx->xform_threshold : (flat_scene ? 1 : 4);
    job.node_threshold = ctx->node_threshold; job.node_burst = ctx->node_burst > 0 ? ctx->node_burst : (flat_scene ? 3 : 2);

    cudaStream_t st = ctx->stream;
    int launches = 0;
    cudaMemsetAsync(wb.stats, 0, 40 * 8, st);
    wf_init_kernel<<<(n_slots + 255) / 256, 256, 0, st>>>(wb, job); launches++;
    const int sms = ctx->sm_count;
    // worst case: every sample of a slot takes all bounces
    const int samples_per_slot = (job.s_count + (int)lanes - 1) / (int)lanes;
    const int max_rounds = samples_per_slot * (sc.max_bounces + 1) + 2;
    const int check_every = 8;
    int rounds = 0;
    // four events per round (round begin, trace begin, trace end, shade end) for the first WAVE_TIMED_ROUNDS rounds
    const int WAVE_TIMED_ROUNDS = 48;
    if(ctx->wave_events.empty())
    {
        ctx->wave_events.resize(4 * WAVE_TIMED_ROUNDS);
        for(auto& e : ctx->wave_events) if(cudaEventCreate(&e) != cudaSuccess) return -1;
    }
    for(;;)
    {
        for(int b = 0; b < check_every && rounds < max_rounds; ++b, ++rounds)
        {
            const bool timed = rounds < WAVE_TIMED_ROUNDS;
            if(timed) cudaEventRecord(ctx->wave_events[4 * rounds], st);
            if(wb.sort && rounds > 0)
            {   // the bounce and shadow rays the previous round's shading appended (round 0 has primary rays only)
                cudaMemsetAsync(wb.sort_hist, 0, 4ull * 2ull * WF_SORT_BINS, st);
                wf_sort_count_kernel<<<dim3(sms * 2, 2), WF_SORT_THREADS, 0, st>>>(wb);
                wf_sort_scan_kernel<<<2, 1024, 0, st>>>(wb);
                wf_sort_scatter_kernel<<<dim3(sms * 2, 2), WF_SORT_THREADS, 0, st>>>(wb);
                launches += 3;
            }
            wf_generate_kernel<<<sms * 4, 256, 0, st>>>(sc, job, wb);
            if(timed) cudaEventRecord(ctx->wave_events[4 * rounds + 1], st);
            if(ctx->plain_trace == 1 || (ctx->plain_trace == 2 && rounds == 0)) wf_trace_plain_kernel<<<sms * 16, 128, 0, st>>>(sc, job, wb);
            else if(ctx->top_smem) wf_trace_cw_kernel<true><<<sms * WF_CW_BLOCKS, WF_TRACE_THREADS, 0, st>>>(sc, job, wb);
            else wf_trace_cw_kernel<false><<<sms * WF_CW_BLOCKS, WF_TRACE_THREADS, 0, st>>>(sc, job, wb);
            if(timed) cudaEventRecord(ctx->wave_events[4 * rounds + 2], st);
            if(ctx->validate) wf_validate_kernel<<<sms * 8, 128, 0, st>>>(sc, job, wb, wb.stats + 39);
            wf_phase_kernel<<<1, 32, 0, st>>>(wb, 1, nullptr);
            wf_classify_kernel<<<sms * 8, 256, 0, st>>>(wb);
            wf_shade_kernel<true><<<sms * WF_SHADE_GRID, 128, 0, st>>>(sc, job, wb);
            wf_shade_kernel<false><<<sms * WF_SHADE_GRID, 128, 0, st>>>(sc, job, wb);
            if(timed) cudaEventRecord(ctx->wave_events[4 * rounds + 3], st);
            wf_phase_kernel<<<1, 32, 0, st>>>(wb, 2, ctx->wave_flag.p);
            launches += 7;
        }
        // every possible round is in flight: no need to ask the device whether paths are left, and the
        // call stays asynchronous (the shipped 256-spp config: one sample per slot, bounces + 3 rounds)
        if(rounds >= max_rounds && !ctx->validate) break;
        if(cudaMemcpyAsync(ctx->wave_flag_host, ctx->wave_flag.p, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
        if(cudaStreamSynchronize(st) != cudaSuccess) return -1;
        if(*ctx->wave_flag_host == 0u || rounds >= max_rounds) break;
    }
    ctx->last_wave_rounds = rounds;
    ctx->wave_timed_rounds = rounds < WAVE_TIMED_ROUNDS ? rounds : WAVE_TIMED_ROUNDS;   // read back lazily by wave_timing()
    if(ctx->validate)
    {
        unsigned long long h = 0;
        cudaMemcpyAsync(&h, wb.stats + 39, sizeof(h), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        ctx->last_validate_mismatches = h;
    }
    ctx->last_wave_lanes = lanes;
    ctx->last_pool_bytes = off;
#ifdef WF_STATS
    {
        unsigned long long h[40];
        cudaMemcpyAsync(h, wb.stats, sizeof(h), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        fprintf(stderr, "WF_STATS node-iters %llu (executed %llu): lanes node %.2f idle %.2f enter-wait %.2f pend-full %.2f tri-wait %.2f | tri steps %llu lanes %.2f | enter steps %llu lanes %.2f | forced %llu (n %.2f t %.2f e %.2f)\n",
                h[0], h[6], (double)h[1] / h[0], (double)h[2] / h[0], (double)h[3] / h[0], (double)h[4] / h[0], (double)h[5] / h[0],
                h[8], h[8] ? (double)h[9] / h[8] : 0.0, h[10], h[10] ? (double)h[11] / h[10] : 0.0,
                h[12], h[12] ? (double)h[13] / h[12] : 0.0, h[12] ? (double)h[14] / h[12] : 0.0, h[12] ? (double)h[15] / h[12] : 0.0);
        if(h[22]) fprintf(stderr, "WF_STATS plain traversal of the same %llu rays (validate): %.2f node tests, %.2f triangle tests per ray\n",
                          h[22], (double)h[20] / h[22], (double)h[21] / h[22]);
        fprintf(stderr, "WF_STATS scheduled kernel per ray-lane: node tests %llu, triangle tests %llu\n", h[1], h[9]);
        fprintf(stderr, "WF_STATS entries by BLAS:");
        for(int i = 0; i < 16; ++i) if(h[24 + i]) fprintf(stderr, " [%d] %llu", i, h[24 + i]);
        fprintf(stderr, "\n");
        fprintf(stderr, "WF_STATS instance root tests %llu, of which no child box hit %llu (%.1f %%)\n", h[7], h[16], h[7] ? 100.0 * h[16] / h[7] : 0.0);
    }
#endif
    wf_finalize_kernel<<<(unsigned)((n_pix_padded + 255) / 256), 256, 0, st>>>(job, wb); launches++;
    return launches;
}

// Launch one render job on the context's stream. Returns kernels launched, or -1.
int launch_job(ptgpu_ctx* ctx, const RenderJob& job)
{
    Scene sc = make_scene(ctx);
    int launches = 0;
    if(ctx->traversal == 1)
    {
        const int tiles = ((job.w + TILE_W - 1) / TILE_W) * ((job.h + TILE_H - 1) / TILE_H);
        if(ctx->counters_on)
            render_tiles_kernel<LinksTrav<true>, true><<<tiles, TILE_THREADS, 0, ctx->stream>>>(sc, job, ctx->counters.p);
        else
            render_tiles_kernel<LinksTrav<false>, false><<<tiles, TILE_THREADS, 0, ctx->stream>>>(sc, job, nullptr);
        launches = 1;
    }
    else if(ctx->kernel == 1)
    {
        const int tiles = ((job.w + TILE_W - 1) / TILE_W) * ((job.h + TILE_H - 1) / TILE_H);
        // the straight-line kernel over the plain single-ray traversal of whichever BVH is selected
        if(ctx->bvh == 1) render_tiles_kernel<CwTrav, false><<<tiles, TILE_THREADS, 0, ctx->stream>>>(sc, job, nullptr);
        else render_tiles_kernel<WideTrav, false><<<tiles, TILE_THREADS, 0, ctx->stream>>>(sc, job, nullptr);
        launches = 1;
    }
    else if(ctx->kernel == 0)
    {
        launches = launch_mega(sc, job, ctx->mega_state.p, ctx->sm_count, ctx->stream);
    }
    else
    {
        launches = launch_wave(ctx, sc, job);
    }
    if(launches < 0) return launches;
    if(cudaGetLastError() != cudaSuccess) return -1;
    return launches;
}

// Per-subframe dynamic instance sets -> the kernels' encoding {prefix p, a | len << 20}: ids [n_static,
// n_static + p) are seen by every subframe (frame-static extras: logo, buddha), ids n_static + [a, a + len) by
// this one. The kernels keep a subframe's set as one 24-bit group / 16 stack entries: more is refused here.
bool encode_dynamic_ranges(const uint32_t* dyn_begin, const uint32_t* dyn_end, size_t n_subframes, size_t n_dyn,
                           std::vector<uint2>& ranges, std::string& why)
{
    // shared prefix = instances before the first subframe's range
    uint32_t prefix = n_dyn ? 0xFFFFFFFFu : 0;
    for(size_t i = 0; i < n_subframes; ++i)
    {
        if(dyn_begin[i] > dyn_end[i] || dyn_end[i] > n_dyn) { why = "bad range " + std::to_string(i); return false; }
        if(dyn_begin[i] < prefix) prefix = dyn_begin[i];
    }
    ranges.resize(n_subframes);
    for(size_t i = 0; i < n_subframes; ++i)
    {
        const uint32_t a = dyn_begin[i], b = dyn_end[i];
        if(a >= (1u << 20) || prefix + (b - a) > (uint32_t)PTGPU_MAX_DYNAMIC_PER_SUBFRAME)
        {
            why = "subframe " + std::to_string(i) + " sees " + std::to_string(prefix + (b - a)) + " dynamic instances (limit " +
                std::to_string(PTGPU_MAX_DYNAMIC_PER_SUBFRAME) + ")";
            return false;
        }
        ranges[i] = make_uint2(prefix, a | ((b - a) << 20));
    }
    return true;
}

int upload_frame_common(ptgpu_ctx* ctx, const ptgpu_subframe* subframes, size_t n_subframes,
                        const ptgpu_tlas_instance* dyn, size_t n_dyn, const std::vector<uint2>& ranges)
{
    // one pinned staging block, three async copies, all on the render stream
    const size_t sz_sub = n_subframes * sizeof(RefSubframe);
    const size_t sz_dyn = n_dyn * sizeof(RefInstance);
    const size_t sz_rng = n_subframes * sizeof(uint2);
    const size_t sz_win = n_dyn * sizeof(WideInstance);
    const size_t sz_uni = n_subframes * 2 * sizeof(float4);
    // the previous frame's copies must have left the staging block
    CK(cudaStreamSynchronize(ctx->stream));
    uint8_t* st = (uint8_t*)pinned_staging(ctx, sz_sub + sz_dyn + sz_rng + sz_win + sz_uni + 128);
    if(!st) return fail(ctx, "pinned staging allocation failed");
    CK(ctx->subframes.reserve(n_subframes));
    CK(ctx->dyn_range.reserve(n_subframes));
    CK(ctx->dyn_union.reserve(2 * n_subframes));
    if(ctx->instances.cap < ctx->n_static + n_dyn)
    {   // grow, keeping the static part
        DevBuf<RefInstance> nb; CK(nb.reserve(ctx->n_static + n_dyn + 64));
        CK(cudaMemcpyAsync(nb.p, ctx->instances.p, ctx->n_static * sizeof(RefInstance), cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->instances.release(); ctx->instances = nb;
        DevBuf<WideInstance> wb; CK(wb.reserve(ctx->n_static + n_dyn + 64));
        CK(cudaMemcpyAsync(wb.p, ctx->winst.p, ctx->n_static * sizeof(WideInstance), cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->winst.release(); ctx->winst = wb;
    }
    size_t off = 0;
    const WideInstance* wi_host = nullptr;
    memcpy(st + off, subframes, sz_sub);
    CK(cudaMemcpyAsync(ctx->subframes.p, st + off, sz_sub, cudaMemcpyHostToDevice, ctx->stream));
    off += sz_sub;
    if(n_dyn)
    {
        memcpy(st + off, dyn, sz_dyn);
        CK(cudaMemcpyAsync(ctx->instances.p + ctx->n_static, st + off, sz_dyn, cudaMemcpyHostToDevice, ctx->stream));
        off += sz_dyn;
        WideInstance* wi = (WideInstance*)(st + ((off + 15) & ~size_t(15)));
        for(size_t i = 0; i < n_dyn; ++i)
            if(!make_wide_instance(ctx->host->wide, dyn[i], (uint32_t)(ctx->n_static + i), wi[i]))
                return fail(ctx, "dynamic instance %zu references an unknown BLAS (node_offset %u)", i, dyn[i].blas.node_offset);
        CK(cudaMemcpyAsync(ctx->winst.p + ctx->n_static, wi, sz_win, cudaMemcpyHostToDevice, ctx->stream));
        wi_host = wi;
        off = ((off + 15) & ~size_t(15)) + sz_win;
    }
    memcpy(st + off, ranges.data(), sz_rng);
    CK(cudaMemcpyAsync(ctx->dyn_range.p, st + off, sz_rng, cudaMemcpyHostToDevice, ctx->stream));
    off = (off + sz_rng + 15) & ~size_t(15);
    {   // one box per subframe around everything its per-frame instances can cover: a query that misses it
        // (most bounce and shadow rays) skips the per-instance box tests at its start (cw_begin)
        float4* un = (float4*)(st + off);
        for(size_t i = 0; i < n_subframes; ++i)
        {
            float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
            const uint32_t p = ranges[i].x, a = ranges[i].y & 0xFFFFFu, len = ranges[i].y >> 20;
            for(uint32_t k = 0; k < p + len; ++k)
            {
                const WideInstance& w = wi_host[k < p ? k : a + (k - p)];
                lo[0] = std::min(lo[0], w.lo.x); lo[1] = std::min(lo[1], w.lo.y); lo[2] = std::min(lo[2], w.lo.z);
                hi[0] = std::max(hi[0], w.hi.x); hi[1] = std::max(hi[1], w.hi.y); hi[2] = std::max(hi[2], w.hi.z);
            }
            un[2 * i] = make_float4(lo[0], lo[1], lo[2], 0.0f);
            un[2 * i + 1] = make_float4(hi[0], hi[1], hi[2], 0.0f);
        }
        CK(cudaMemcpyAsync(ctx->dyn_union.p, un, sz_uni, cudaMemcpyHostToDevice, ctx->stream));
    }
    ctx->n_subframes = n_subframes;
    ctx->n_dyn = n_dyn;
    ctx->have_frame = true;
    return 0;
}

} // namespace

extern "C" {

void ptgpu_default_config(ptgpu_config* cfg)
{
    cfg->width = 640; cfg->height = 360; cfg->spp = 256; cfg->max_bounces = 4;
    cfg->student_id = 152121358u; cfg->samples_per_subframe = 8;
}

int ptgpu_create(ptgpu_ctx** out, int device, const ptgpu_config* cfg)
{
    if(!out) return 1;
    *out = nullptr;
    ptgpu_ctx* ctx = nullptr; // for CK's fail()
    if(!cfg || cfg->width <= 0 || cfg->height <= 0 || cfg->spp <= 0 || cfg->max_bounces < 0 || cfg->samples_per_subframe <= 0)
        return fail(nullptr, "invalid ptgpu_config");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if(e != cudaSuccess || n == 0)
        return fail(nullptr, "no CUDA device: %s (libptgpu has no CPU fallback)", cudaGetErrorString(e));
    if(device < 0 || device >= n) return fail(nullptr, "device %d out of range (have %d)", device, n);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if(prop.major != 10)
        return fail(nullptr, "device %d is sm_%d%d; libptgpu is built for sm_100a only", device, prop.major, prop.minor);
    CK(cudaSetDevice(device));
    ctx = new ptgpu_ctx();
    ctx->device = device;
    ctx->cfg = *cfg;
    ctx->sm_count = prop.multiProcessorCount;
    cudaError_t e1 = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    cudaError_t e2 = cudaEventCreate(&ctx->ev_begin);
    cudaError_t e3 = cudaEventCreate(&ctx->ev_end);
    if(e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
    {
        delete ctx;
        return fail(nullptr, "stream/event creation failed");
    }
    ctx->bmp_pitch = ((uint32_t)cfg->width * 3u + 3u) / 4u * 4u;
    const size_t npix = (size_t)cfg->width * cfg->height;
    if(ctx->out_bgra.reserve(npix) != cudaSuccess || ctx->out_bmp.reserve(54 + (size_t)ctx->bmp_pitch * cfg->height) != cudaSuccess ||
       ctx->counters.reserve(1) != cudaSuccess || ctx->mega_state.reserve(1) != cudaSuccess)
    {
        delete ctx;
        return fail(nullptr, "output allocation failed");
    }
    cudaMemset(ctx->out_bmp.p, 0, 54 + (size_t)ctx->bmp_pitch * cfg->height);
    cudaMemset(ctx->counters.p, 0, sizeof(Counters));
    cudaMemset(ctx->mega_state.p, 0, sizeof(MegaState));
    *out = ctx;
    return 0;
}

int ptgpu_warm_up(int device)
{
    // creates the device's primary CUDA context (what the first CUDA call on a device pays for: ~1 s per GPU,
    // serialised by the driver), so that a host can overlap it with its own start-up work
    if(cudaSetDevice(device) != cudaSuccess) return 1;
    return cudaFree(nullptr) == cudaSuccess ? 0 : 1;
}

void ptgpu_destroy(ptgpu_ctx* ctx)
{
    if(!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->ref_nodes.release(); ctx->ref_links.release(); ctx->indices.release();
    ctx->pos.release(); ctx->normal.release(); ctx->albedo.release(); ctx->material.release();
    ctx->instances.release(); ctx->wnodes.release(); ctx->wtris.release(); ctx->wblas.release();
    ctx->winst.release(); ctx->wtlas.release(); ctx->cwnodes.release(); ctx->cwtris.release(); ctx->shade_tris.release(); ctx->cw_inst_index.release(); ctx->subframes.release(); ctx->dyn_range.release(); ctx->dyn_union.release();
    ctx->out_bgra.release(); ctx->out_bmp.release(); ctx->out_rgb.release(); ctx->counters.release(); ctx->rect_bgra.release();
    ctx->mega_state.release(); ctx->wave_mem.release(); ctx->wave_flag.release();
    if(ctx->wave_flag_host) cudaFreeHost(ctx->wave_flag_host);
    for(auto& e : ctx->wave_events) cudaEventDestroy(e);
    ctx->scratch_a.release(); ctx->scratch_b.release(); ctx->scratch_c.release();
    if(ctx->pinned) cudaFreeHost(ctx->pinned);
    cudaEventDestroy(ctx->ev_begin); cudaEventDestroy(ctx->ev_end);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* ptgpu_last_error(const ptgpu_ctx* ctx)
{
    return ctx ? ctx->error.c_str() : g_create_error.c_str();
}

// Static scene upload. nodes/links == nullptr: no reference BVH, every BLAS is built from `meshes`.
static int upload_static_common(
    ptgpu_ctx* ctx,
    const ptgpu_bvh_node* nodes, size_t n_nodes,
    const ptgpu_bvh_link* links, size_t n_links,
    const uint32_t* indices, size_t n_indices,
    const ptgpu_float3* pos, const ptgpu_float3* normal,
    const ptgpu_float4* albedo, const ptgpu_float4* material, size_t n_verts,
    const ptgpu_tlas_instance* instances, size_t n_static,
    const ptgpu_mesh* meshes, size_t n_meshes)
{
    if(use(ctx)) return 1;

    // reference layout, with head-room after the static region for the per-frame TLAS (links mode)
    CK(ctx->ref_nodes.reserve(3 * n_nodes + 16));
    CK(ctx->ref_links.reserve(n_links + 16));
    CK(ctx->indices.reserve(n_indices));
    CK(ctx->pos.reserve(n_verts)); CK(ctx->normal.reserve(n_verts));
    CK(ctx->albedo.reserve(n_verts)); CK(ctx->material.reserve(n_verts));
    CK(ctx->instances.reserve(n_static + 64));
    if(nodes)
    {
        CK(cudaMemcpy(ctx->ref_nodes.p, nodes, n_nodes * sizeof(ptgpu_bvh_node), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(ctx->ref_links.p, links, n_links * sizeof(ptgpu_bvh_link), cudaMemcpyHostToDevice));
    }
    ctx->have_ref_bvh = nodes != nullptr;
    if(!ctx->have_ref_bvh) ctx->traversal = 0;
    CK(cudaMemcpy(ctx->indices.p, indices, n_indices * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->pos.p, pos, n_verts * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->normal.p, normal, n_verts * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->albedo.p, albedo, n_verts * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->material.p, material, n_verts * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->instances.p, instances, n_static * sizeof(RefInstance), cudaMemcpyHostToDevice));
    ctx->n_static_nodes = n_nodes; ctx->n_static = n_static; ctx->n_verts = n_verts; ctx->n_indices = n_indices;
    ctx->host_static.assign(instances, instances + n_static);

    // GPU traversal layout: a compressed BVH per distinct BLAS + one static TLAS (+ the flat static scene)
    std::string err;
    ctx->host = get_host_scene(nodes, n_nodes, links, n_links, indices, n_indices, pos, n_verts, instances, n_static,
                               meshes, n_meshes, ctx->flat != 0, err);
    if(!ctx->host) return fail(ctx, "%s", err.c_str());
    const WideScene& w = ctx->host->wide;
    {   // shading records: the nine attribute vectors of every triangle, contiguous (pt_scene.cuh)
        std::vector<float4> rec(9 * (n_indices / 3), make_float4(0.f, 0.f, 0.f, 0.f));
        for(const WideBlasInfo& bi : w.blas_info)
        {
            if(bi.mesh.index_offset % 3u) return fail(ctx, "mesh index_offset %u is not a multiple of 3", bi.mesh.index_offset);
            for(uint32_t t = 0; t < bi.mesh.triangle_count; ++t)
            {
                const size_t k = bi.mesh.index_offset / 3u + t;
                for(int j = 0; j < 3; ++j)
                {
                    const size_t v = (size_t)bi.mesh.base_vertex_offset + indices[3 * k + j];
                    const ptgpu_float3& n = normal[v];
                    rec[9 * k + j] = make_float4(n.x, n.y, n.z, 0.f);
                    rec[9 * k + 3 + j] = make_float4(albedo[v].x, albedo[v].y, albedo[v].z, albedo[v].w);
                    rec[9 * k + 6 + j] = make_float4(material[v].x, material[v].y, material[v].z, material[v].w);
                }
            }
        }
        CK(ctx->shade_tris.reserve(rec.size() + 1));
        CK(cudaMemcpy(ctx->shade_tris.p, rec.data(), rec.size() * sizeof(float4), cudaMemcpyHostToDevice));
    }
    CK(ctx->wnodes.reserve(w.nodes.size()));
    CK(ctx->wtris.reserve(w.tris.size()));
    CK(ctx->wblas.reserve(w.blas.size()));
    CK(ctx->winst.reserve(n_static + 64));
    CK(ctx->wtlas.reserve(w.tlas.size()));
    CK(cudaMemcpy(ctx->wnodes.p, w.nodes.data(), w.nodes.size() * sizeof(WideNode), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->wtris.p, w.tris.data(), w.tris.size() * sizeof(float4), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->wblas.p, w.blas.data(), w.blas.size() * sizeof(WideBlas), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->winst.p, w.instances.data(), n_static * sizeof(WideInstance), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->wtlas.p, w.tlas.data(), w.tlas.size() * sizeof(WideNode), cudaMemcpyHostToDevice));
    ctx->n_wtlas = w.tlas.size();
    // flat static scene: appended to the compressed node / triangle arrays (its indices are absolute)
    const FlatScene* flat = ctx->host->have_flat ? &ctx->host->flat : nullptr;
    ctx->have_flat = flat != nullptr; ctx->flat_root = 0xFFFFFFFFu;
    if(flat)
    {
        ctx->flat_root = (uint32_t)(w.cw_nodes.size() / 5); ctx->flat_top = flat->n_top; ctx->flat_depth = flat->depth;
        ctx->flat_tris = flat->n_tris; ctx->flat_nodes = flat->nodes.size() / 5;
        ctx->flat_build_seconds = flat->build_seconds;
    }
    {   // ray-sort grid over the static scene
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for(const WideInstance& wi : w.instances)
        {
            lo[0] = std::min(lo[0], wi.lo.x); lo[1] = std::min(lo[1], wi.lo.y); lo[2] = std::min(lo[2], wi.lo.z);
            hi[0] = std::max(hi[0], wi.hi.x); hi[1] = std::max(hi[1], wi.hi.y); hi[2] = std::max(hi[2], wi.hi.z);
        }
        const float cells[3] = {32.0f, 8.0f, 32.0f};
        for(int a = 0; a < 3; ++a)
        {
            ctx->key_lo[a] = lo[a];
            ctx->key_scale[a] = hi[a] > lo[a] ? cells[a] / (hi[a] - lo[a]) : 0.0f;
        }
    }
    const size_t n_flat_nodes4 = flat ? flat->nodes.size() : 0, n_flat_tris4 = flat ? flat->tris.size() : 0;
    CK(ctx->cwnodes.reserve(w.cw_nodes.size() + n_flat_nodes4));
    CK(ctx->cwtris.reserve(w.cw_tris.size() + n_flat_tris4));
    CK(ctx->cw_inst_index.reserve(w.cw_inst_index.size()));
    CK(cudaMemcpy(ctx->cwnodes.p, w.cw_nodes.data(), w.cw_nodes.size() * sizeof(float4), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->cwtris.p, w.cw_tris.data(), w.cw_tris.size() * sizeof(float4), cudaMemcpyHostToDevice));
    if(flat)
    {
        CK(cudaMemcpy(ctx->cwnodes.p + w.cw_nodes.size(), flat->nodes.data(), n_flat_nodes4 * sizeof(float4), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(ctx->cwtris.p + w.cw_tris.size(), flat->tris.data(), n_flat_tris4 * sizeof(float4), cudaMemcpyHostToDevice));
    }
    CK(cudaMemcpy(ctx->cw_inst_index.p, w.cw_inst_index.data(), w.cw_inst_index.size() * 4, cudaMemcpyHostToDevice));
    ctx->have_static = true;
    ctx->have_frame = false;
    ctx->frame_on_device = false;
    ctx->l2_persist_applied = -1;
    return 0;
}

int ptgpu_upload_static(
    ptgpu_ctx* ctx,
    const ptgpu_bvh_node* nodes, size_t n_nodes,
    const ptgpu_bvh_link* links, size_t n_links,
    const uint32_t* indices, size_t n_indices,
    const ptgpu_float3* pos, const ptgpu_float3* normal,
    const ptgpu_float4* albedo, const ptgpu_float4* material, size_t n_verts,
    const ptgpu_tlas_instance* instances, size_t n_static)
{
    if(!ctx) return 1;
    if(!nodes || !links || !indices || !pos || !normal || !albedo || !material || !instances)
        return fail(ctx, "ptgpu_upload_static: null array");
    if(n_links != 8 * n_nodes) return fail(ctx, "ptgpu_upload_static: n_links (%zu) must be 8*n_nodes (%zu)", n_links, n_nodes);
    if(n_static == 0 || n_nodes == 0) return fail(ctx, "ptgpu_upload_static: empty scene");
    return upload_static_common(ctx, nodes, n_nodes, links, n_links, indices, n_indices, pos, normal, albedo, material, n_verts,
                                instances, n_static, nullptr, 0);
}

int ptgpu_upload_meshes(
    ptgpu_ctx* ctx,
    const uint32_t* indices, size_t n_indices,
    const ptgpu_float3* pos, const ptgpu_float3* normal,
    const ptgpu_float4* albedo, const ptgpu_float4* material, size_t n_verts,
    const ptgpu_mesh* meshes, size_t n_meshes,
    const ptgpu_tlas_instance* instances, size_t n_static)
{
    if(!ctx) return 1;
    if(!indices || !pos || !normal || !albedo || !material || !meshes || !instances)
        return fail(ctx, "ptgpu_upload_meshes: null array");
    if(n_static == 0 || n_meshes == 0) return fail(ctx, "ptgpu_upload_meshes: empty scene");
    return upload_static_common(ctx, nullptr, 0, nullptr, 0, indices, n_indices, pos, normal, albedo, material, n_verts,
                                instances, n_static, meshes, n_meshes);
}

int ptgpu_set_frame(
    ptgpu_ctx* ctx,
    const ptgpu_subframe* subframes, size_t n_subframes,
    const ptgpu_tlas_instance* dyn_instances, size_t n_dyn,
    const ptgpu_bvh_node* tlas_nodes, const ptgpu_bvh_link* tlas_links,
    size_t n_tlas_nodes, size_t tlas_node_base)
{
    if(!ctx) return 1;
    if(!ctx->have_static) return fail(ctx, "ptgpu_set_frame before ptgpu_upload_static");
    if(!ctx->have_ref_bvh) return fail(ctx, "ptgpu_set_frame: the scene was built from meshes (no reference BVH); use ptgpu_set_frame_ranges");
    if(!subframes || n_subframes == 0) return fail(ctx, "ptgpu_set_frame: no subframes");
    if(!tlas_nodes || !tlas_links) return fail(ctx, "ptgpu_set_frame: null TLAS arrays");
    if(tlas_node_base != ctx->n_static_nodes)
        return fail(ctx, "ptgpu_set_frame: tlas_node_base %zu != static node count %zu", tlas_node_base, ctx->n_static_nodes);
    if(use(ctx)) return 1;

    // Which dynamic instances does each subframe see? Read the leaf payloads of its TLAS
    // (octant-0 link table; all eight tables hold the same leaves, bvh.cc:170-193). The set is
    // [n_static, static_end) U [dyn_begin, dyn_end) (scene.cc:75-91, 634-674): a prefix shared by
    // all subframes (logo, buddha) plus one contiguous per-subframe range.
    std::vector<uint8_t> seen;
    std::vector<uint2> final_ranges(n_subframes);
    for(size_t i = 0; i < n_subframes; ++i)
    {
        const ptgpu_bvh& t = subframes[i].tlas;
        if(t.node_offset < tlas_node_base || (size_t)t.node_offset + t.node_count > tlas_node_base + n_tlas_nodes)
            return fail(ctx, "ptgpu_set_frame: subframe %zu TLAS [%u,+%u) outside the passed arrays", i, t.node_offset, t.node_count);
        const ptgpu_bvh_link* l = tlas_links + 8 * (size_t)(t.node_offset - tlas_node_base);
        seen.assign(n_dyn, 0);
        size_t n_static_seen = 0;
        for(uint32_t k = 0; k < t.node_count; ++k)
        {
            if(!(l[k].accept & 0x80000000u)) continue;
            const uint32_t id = l[k].accept & 0x7FFFFFFFu;
            if(id < ctx->n_static) { n_static_seen++; continue; }
            if(id >= ctx->n_static + n_dyn) return fail(ctx, "ptgpu_set_frame: TLAS leaf %u beyond the instance array", id);
            seen[id - ctx->n_static] = 1;
        }
        if(n_static_seen != ctx->n_static)
            return fail(ctx, "ptgpu_set_frame: subframe %zu TLAS holds %zu of %zu static instances", i, n_static_seen, ctx->n_static);
        uint32_t p = 0; while(p < n_dyn && seen[p]) ++p;
        uint32_t a = p; while(a < n_dyn && !seen[a]) ++a;
        uint32_t b = a; while(b < n_dyn && seen[b]) ++b;
        for(uint32_t k = b; k < n_dyn; ++k)
            if(seen[k]) return fail(ctx, "ptgpu_set_frame: subframe %zu dynamic set is not prefix + one range", i);
        if(a == n_dyn) { a = b = p; }
        // kernel encoding: x = prefix length p (ids n_static .. n_static+p), y = a | (b-a) << 20
        if(a >= (1u << 20) || p + (b - a) > (uint32_t)PTGPU_MAX_DYNAMIC_PER_SUBFRAME)
            return fail(ctx, "ptgpu_set_frame: subframe %zu sees %u dynamic instances (limit %d)", i, p + (b - a), PTGPU_MAX_DYNAMIC_PER_SUBFRAME);
        final_ranges[i] = make_uint2(p, a | ((b - a) << 20));
    }
    if(upload_frame_common(ctx, subframes, n_subframes, dyn_instances, n_dyn, final_ranges)) return 1;

    if(ctx->traversal == 1)
    {   // links mode also needs the reference TLAS behind the static region
        const size_t need_nodes = ctx->n_static_nodes + n_tlas_nodes;
        if(ctx->ref_nodes.cap < 3 * need_nodes)
        {
            DevBuf<float2> nb; CK(nb.reserve(3 * (need_nodes + need_nodes / 64)));
            CK(cudaMemcpy(nb.p, ctx->ref_nodes.p, ctx->n_static_nodes * 24, cudaMemcpyDeviceToDevice));
            ctx->ref_nodes.release(); ctx->ref_nodes = nb;
            DevBuf<uint2> lb; CK(lb.reserve(8 * (need_nodes + need_nodes / 64)));
            CK(cudaMemcpy(lb.p, ctx->ref_links.p, ctx->n_static_nodes * 64, cudaMemcpyDeviceToDevice));
            ctx->ref_links.release(); ctx->ref_links = lb;
        }
        CK(cudaMemcpyAsync(ctx->ref_nodes.p + 3 * ctx->n_static_nodes, tlas_nodes, n_tlas_nodes * 24, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->ref_links.p + 8 * ctx->n_static_nodes, tlas_links, n_tlas_nodes * 64, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->frame_has_ref_tlas = true;
    }
    else ctx->frame_has_ref_tlas = false;
    return 0;
}

int ptgpu_set_frame_ranges(
    ptgpu_ctx* ctx,
    const ptgpu_subframe* subframes, size_t n_subframes,
    const ptgpu_tlas_instance* dyn_instances, size_t n_dyn,
    const uint32_t* dyn_begin, const uint32_t* dyn_end)
{
    if(!ctx) return 1;
    if(!ctx->have_static) return fail(ctx, "ptgpu_set_frame_ranges before ptgpu_upload_static");
    if(!subframes || n_subframes == 0 || !dyn_begin || !dyn_end) return fail(ctx, "ptgpu_set_frame_ranges: null argument");
    if(use(ctx)) return 1;
    std::vector<uint2> ranges;
    std::string why;
    if(!encode_dynamic_ranges(dyn_begin, dyn_end, n_subframes, n_dyn, ranges, why))
        return fail(ctx, "ptgpu_set_frame_ranges: %s", why.c_str());
    ctx->frame_has_ref_tlas = false;
    return upload_frame_common(ctx, subframes, n_subframes, dyn_instances, n_dyn, ranges);
}

static int render_full(ptgpu_ctx* ctx, bool bgra, bool bmp)
{
    if(check_ready(ctx)) return 1;
    if(use(ctx)) return 1;
    if((size_t)ctx->cfg.spp > ctx->n_subframes * (size_t)ctx->cfg.samples_per_subframe)
        return fail(ctx, "frame has %zu subframes, %d spp needs %d", ctx->n_subframes, ctx->cfg.spp,
                    (ctx->cfg.spp + ctx->cfg.samples_per_subframe - 1) / ctx->cfg.samples_per_subframe);
    RenderJob job{};
    job.x0 = 0; job.y0 = 0; job.w = ctx->cfg.width; job.h = ctx->cfg.height;
    job.s_begin = 0; job.s_count = ctx->cfg.spp; job.s_stride = 1;
    job.out_rgb = nullptr;
    job.out_bgra = bgra ? ctx->out_bgra.p : nullptr;
    job.out_bmp = bmp ? ctx->out_bmp.p : nullptr;
    job.bmp_pitch = ctx->bmp_pitch;
    job.min_active = ctx->min_active < 0 ? (ctx->kernel == 2 ? 6 : 0) : ctx->min_active;
    int launches = 0;
    CK(cudaEventRecord(ctx->ev_begin, ctx->stream));
    if(bmp && !ctx->bmp_header_done)
    {
        bmp_header_kernel<<<1, 32, 0, ctx->stream>>>(ctx->out_bmp.p, (uint32_t)job.w, (uint32_t)job.h, ctx->bmp_pitch);
        ctx->bmp_header_done = true;
        launches++;
    }
    int l = launch_job(ctx, job);
    if(l == -2) return 1;   // launch_wave said why
    if(l < 0) return fail(ctx, "kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    launches += l;
    CK(cudaEventRecord(ctx->ev_end, ctx->stream));
    ctx->last_launches = launches;
    ctx->render_pending = true;
    ctx->frame_on_device = true;
    return 0;
}

int ptgpu_render_async(ptgpu_ctx* ctx) { return render_full(ctx, true, true); }

int ptgpu_sync(ptgpu_ctx* ctx)
{
    if(!ctx) return 1;
    if(use(ctx)) return 1;
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int ptgpu_fetch_bgra(ptgpu_ctx* ctx, uint8_t* out_bgra)
{
    if(!ctx || !out_bgra) return 1;
    if(!ctx->frame_on_device) return fail(ctx, "ptgpu_fetch_bgra: no rendered frame on the device");
    if(use(ctx)) return 1;
    CK(cudaMemcpyAsync(out_bgra, ctx->out_bgra.p, (size_t)ctx->cfg.width * ctx->cfg.height * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

size_t ptgpu_bmp_size(const ptgpu_ctx* ctx)
{
    return ctx ? 54 + (size_t)ctx->bmp_pitch * ctx->cfg.height : 0;
}

int ptgpu_fetch_bmp(ptgpu_ctx* ctx, uint8_t* out_bmp)
{
    if(!ctx || !out_bmp) return 1;
    if(!ctx->frame_on_device) return fail(ctx, "ptgpu_fetch_bmp: no rendered frame on the device");
    if(use(ctx)) return 1;
    CK(cudaMemcpyAsync(out_bmp, ctx->out_bmp.p, ptgpu_bmp_size(ctx), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int ptgpu_validate_frame(ptgpu_ctx* ctx, const uint8_t* ref_rgb_half, double* psnr, int32_t* good)
{
    if(!ctx || !ref_rgb_half || !psnr) return 1;
    if(use(ctx)) return 1;
    if(!ctx->frame_on_device) return fail(ctx, "ptgpu_validate_frame: no rendered frame on the device");
    const uint32_t w = (uint32_t)ctx->cfg.width, h = (uint32_t)ctx->cfg.height, hw = (w + 1) / 2, hh = (h + 1) / 2;
    const size_t ref_bytes = (size_t)hw * hh * 3;
    CK(ctx->scratch_a.reserve(ref_bytes + 16));
    CK(ctx->scratch_b.reserve(16));
    CK(cudaMemcpyAsync(ctx->scratch_a.p, ref_rgb_half, ref_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->scratch_b.p, 0, 8, ctx->stream));
    validate_psnr_kernel<<<ctx->sm_count * 2, 256, 0, ctx->stream>>>(ctx->out_bgra.p, w, h, (const uint8_t*)ctx->scratch_a.p, hw, hh,
                                                                   (unsigned long long*)ctx->scratch_b.p);
    unsigned long long sse = 0;
    CK(cudaMemcpyAsync(&sse, ctx->scratch_b.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    // skimage.metrics.peak_signal_noise_ratio for uint8 images: data range 255, mean over all samples
    const double mse = (double)sse / (double)ref_bytes;
    *psnr = sse == 0 ? INFINITY : 10.0 * log10(255.0 * 255.0 / mse);
    if(good) *good = *psnr >= 32.0 ? 1 : 0; // ACCEPT_MIN_PSNR, validator.py:11
    return 0;
}

int ptgpu_render(ptgpu_ctx* ctx, uint8_t* out_bgra)
{
    if(!out_bgra) return fail(ctx, "ptgpu_render: null output");
    if(render_full(ctx, true, false)) return 1;
    return ptgpu_fetch_bgra(ctx, out_bgra);
}

int ptgpu_render_bmp(ptgpu_ctx* ctx, uint8_t* out_bmp)
{
    if(!out_bmp) return fail(ctx, "ptgpu_render_bmp: null output");
    if(render_full(ctx, true, true)) return 1; // the BGRA frame stays on the device for ptgpu_validate_frame
    return ptgpu_fetch_bmp(ctx, out_bmp);
}

int ptgpu_render_frame(
    ptgpu_ctx* ctx,
    const ptgpu_subframe* subframes, size_t n_subframes,
    const ptgpu_tlas_instance* dyn_instances, size_t n_dyn,
    const ptgpu_bvh_node* tlas_nodes, const ptgpu_bvh_link* tlas_links,
    size_t n_tlas_nodes, size_t tlas_node_base,
    uint8_t* out_bgra)
{
    if(ptgpu_set_frame(ctx, subframes, n_subframes, dyn_instances, n_dyn, tlas_nodes, tlas_links, n_tlas_nodes, tlas_node_base)) return 1;
    return ptgpu_render(ctx, out_bgra);
}

int ptgpu_last_render_ms(ptgpu_ctx* ctx, float* ms, int32_t* launches)
{
    if(!ctx || !ms) return 1;
    if(use(ctx)) return 1;
    if(!ctx->render_pending) return fail(ctx, "no render to time");
    CK(cudaEventSynchronize(ctx->ev_end));
    CK(cudaEventElapsedTime(ms, ctx->ev_begin, ctx->ev_end));
    if(launches) *launches = ctx->last_launches;
    return 0;
}

int ptgpu_render_rect(
    ptgpu_ctx* ctx, int32_t x0, int32_t y0, int32_t w, int32_t h,
    int32_t s_begin, int32_t s_count, int32_t s_stride,
    float* out_rgb, uint8_t* out_bgra)
{
    if(check_ready(ctx)) return 1;
    if(w <= 0 || h <= 0 || s_count <= 0) return fail(ctx, "ptgpu_render_rect: empty job");
    if(x0 < 0 || y0 < 0) return fail(ctx, "ptgpu_render_rect: negative origin");
    const long last = (long)s_begin + (long)(s_count - 1) * s_stride;
    const long lo = s_stride >= 0 ? s_begin : last, hi = s_stride >= 0 ? last : s_begin;
    if(hi >= 0 && (size_t)(hi / ctx->cfg.samples_per_subframe) >= ctx->n_subframes)
        return fail(ctx, "ptgpu_render_rect: sample %ld needs subframe %ld of %zu", hi, hi / ctx->cfg.samples_per_subframe, ctx->n_subframes);
    (void)lo;
    if(use(ctx)) return 1;
    const size_t npix = (size_t)w * h;
    CK(ctx->out_rgb.reserve(npix * 3));
    uchar4* d_bgra = nullptr;
    if(out_bgra)
    {   // a buffer of its own: the full frame of the last ptgpu_render stays valid for ptgpu_validate_frame / fetch
        CK(ctx->rect_bgra.reserve(npix));
        d_bgra = ctx->rect_bgra.p;
    }
    RenderJob job{};
    job.x0 = x0; job.y0 = y0; job.w = w; job.h = h;
    job.s_begin = s_begin; job.s_count = s_count; job.s_stride = s_stride;
    job.out_rgb = ctx->out_rgb.p; job.out_bgra = d_bgra; job.out_bmp = nullptr; job.bmp_pitch = 0;
    job.min_active = ctx->min_active < 0 ? (ctx->kernel == 2 ? 6 : 0) : ctx->min_active;
    CK(cudaEventRecord(ctx->ev_begin, ctx->stream));
    int l = launch_job(ctx, job);
    if(l == -2) return 1;   // launch_wave said why
    if(l < 0) return fail(ctx, "kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    CK(cudaEventRecord(ctx->ev_end, ctx->stream));
    ctx->last_launches = l; ctx->render_pending = true;
    if(out_rgb) CK(cudaMemcpyAsync(out_rgb, ctx->out_rgb.p, npix * 12, cudaMemcpyDeviceToHost, ctx->stream));
    if(out_bgra) CK(cudaMemcpyAsync(out_bgra, d_bgra, npix * 4, cudaMemcpyDeviceToHost, ctx->stream));
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if(e != cudaSuccess) return fail(ctx, "render failed: %s", cudaGetErrorString(e));
    return 0;
}

int ptgpu_trace_samples(ptgpu_ctx* ctx, const uint32_t* xy, const int32_t* sample_index, size_t n, float* out_rgb)
{
    if(check_ready(ctx)) return 1;
    if(n == 0) return 0;
    if(!xy || !sample_index || !out_rgb) return fail(ctx, "ptgpu_trace_samples: null argument");
    for(size_t i = 0; i < n; ++i)
        if(sample_index[i] >= 0 && (size_t)(sample_index[i] / ctx->cfg.samples_per_subframe) >= ctx->n_subframes)
            return fail(ctx, "ptgpu_trace_samples: sample %d beyond the frame's subframes", sample_index[i]);
    if(use(ctx)) return 1;
    CK(ctx->scratch_a.reserve(n * 8)); CK(ctx->scratch_b.reserve(n * 4)); CK(ctx->scratch_c.reserve(n * 12));
    CK(cudaMemcpyAsync(ctx->scratch_a.p, xy, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->scratch_b.p, sample_index, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    Scene sc = make_scene(ctx);
    const int threads = 128; const int blocks = (int)((n + threads - 1) / threads);
    if(ctx->traversal == 1)
        trace_samples_kernel<LinksTrav<false>><<<blocks, threads, 0, ctx->stream>>>(sc, (uint32_t*)ctx->scratch_a.p, (int32_t*)ctx->scratch_b.p, n, (float*)ctx->scratch_c.p);
    else if(ctx->bvh == 1)
        trace_samples_kernel<CwTrav><<<blocks, threads, 0, ctx->stream>>>(sc, (uint32_t*)ctx->scratch_a.p, (int32_t*)ctx->scratch_b.p, n, (float*)ctx->scratch_c.p);
    else
        trace_samples_kernel<WideTrav><<<blocks, threads, 0, ctx->stream>>>(sc, (uint32_t*)ctx->scratch_a.p, (int32_t*)ctx->scratch_b.p, n, (float*)ctx->scratch_c.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_rgb, ctx->scratch_c.p, n * 12, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int ptgpu_tonemap(ptgpu_ctx* ctx, const float* rgb, size_t n, uint8_t* out_bgra)
{
    if(!ctx) return 1;
    if(n == 0) return 0;
    if(!rgb || !out_bgra) return fail(ctx, "ptgpu_tonemap: null argument");
    if(use(ctx)) return 1;
    CK(ctx->scratch_a.reserve(n * 12)); CK(ctx->scratch_b.reserve(n * 4));
    CK(cudaMemcpyAsync(ctx->scratch_a.p, rgb, n * 12, cudaMemcpyHostToDevice, ctx->stream));
    tonemap_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((float*)ctx->scratch_a.p, n, (uchar4*)ctx->scratch_b.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_bgra, ctx->scratch_b.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int ptgpu_trace_closest(ptgpu_ctx* ctx, const float* rays, size_t n, uint32_t subframe, float* out_f, uint32_t* out_u)
{
    if(check_ready(ctx)) return 1;
    if(n == 0) return 0;
    if(!rays || !out_f || !out_u) return fail(ctx, "ptgpu_trace_closest: null argument");
    if(subframe >= ctx->n_subframes) return fail(ctx, "ptgpu_trace_closest: subframe %u of %zu", subframe, ctx->n_subframes);
    if(use(ctx)) return 1;
    CK(ctx->scratch_a.reserve(n * 32)); CK(ctx->scratch_b.reserve(n * 16)); CK(ctx->scratch_c.reserve(n * 12));
    CK(cudaMemcpyAsync(ctx->scratch_a.p, rays, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    Scene sc = make_scene(ctx);
    const int threads = 128; const int blocks = (int)((n + threads - 1) / threads);
    if(ctx->traversal == 1)
        trace_closest_kernel<LinksTrav<false>><<<blocks, threads, 0, ctx->stream>>>(sc, (float*)ctx->scratch_a.p, n, subframe, (float*)ctx->scratch_b.p, (uint32_t*)ctx->scratch_c.p);
    else if(ctx->bvh == 1)
        trace_closest_kernel<CwTrav><<<blocks, threads, 0, ctx->stream>>>(sc, (float*)ctx->scratch_a.p, n, subframe, (float*)ctx->scratch_b.p, (uint32_t*)ctx->scratch_c.p);
    else
        trace_closest_kernel<WideTrav><<<blocks, threads, 0, ctx->stream>>>(sc, (float*)ctx->scratch_a.p, n, subframe, (float*)ctx->scratch_b.p, (uint32_t*)ctx->scratch_c.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_f, ctx->scratch_b.p, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(out_u, ctx->scratch_c.p, n * 12, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int ptgpu_pcg4d(ptgpu_ctx* ctx, uint32_t* states, size_t n, int32_t steps)
{
    if(!ctx) return 1;
    if(n == 0) return 0;
    if(!states) return fail(ctx, "ptgpu_pcg4d: null argument");
    if(use(ctx)) return 1;
    CK(ctx->scratch_a.reserve(n * 16));
    CK(cudaMemcpyAsync(ctx->scratch_a.p, states, n * 16, cudaMemcpyHostToDevice, ctx->stream));
    pcg4d_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((uint32_t*)ctx->scratch_a.p, n, steps);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(states, ctx->scratch_a.p, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int ptgpu_debug_eval(ptgpu_ctx* ctx, int32_t fn, const float* in, size_t n, float* out)
{
    if(!ctx) return 1;
    if(fn < 0 || fn >= PTGPU_FN_COUNT) return fail(ctx, "ptgpu_debug_eval: unknown function %d", fn);
    if(n == 0) return 0;
    if(!in || !out) return fail(ctx, "ptgpu_debug_eval: null argument");
    const bool needs_scene = fn == PTGPU_FN_CAMERA_RAY || fn == PTGPU_FN_SHADOW_RAY || fn == PTGPU_FN_TRACE_RAY;
    if(needs_scene && check_ready(ctx)) return 1;
    if(use(ctx)) return 1;
    CK(ctx->scratch_a.reserve(n * 24 * 4)); CK(ctx->scratch_b.reserve(n * 32 * 4));
    CK(cudaMemcpyAsync(ctx->scratch_a.p, in, n * 24 * 4, cudaMemcpyHostToDevice, ctx->stream));
    Scene sc = make_scene(ctx);
    const int threads = 64; const int blocks = (int)((n + threads - 1) / threads);
    if(ctx->traversal == 1)
        debug_eval_kernel<LinksTrav<false>><<<blocks, threads, 0, ctx->stream>>>(sc, fn, (const float*)ctx->scratch_a.p, n, (float*)ctx->scratch_b.p);
    else
        debug_eval_kernel<CwTrav><<<blocks, threads, 0, ctx->stream>>>(sc, fn, (const float*)ctx->scratch_a.p, n, (float*)ctx->scratch_b.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, ctx->scratch_b.p, n * 32 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int ptgpu_set_option(ptgpu_ctx* ctx, const char* key, int64_t value)
{
    if(!ctx || !key) return 1;
    if(!strcmp(key, "traversal")) { if(value != 0 && value != 1) return fail(ctx, "traversal must be 0 or 1");
        if(value == 1 && ctx->have_static && !ctx->have_ref_bvh) return fail(ctx, "traversal 1 walks the reference's link tables: the scene was built from meshes");
        ctx->traversal = (int)value; ctx->have_frame = false; return 0; }
    if(!strcmp(key, "counters")) { ctx->counters_on = value != 0; return 0; }
    if(!strcmp(key, "kernel")) { if(value < 0 || value > 2) return fail(ctx, "kernel must be 0, 1 or 2"); ctx->kernel = (int)value; return 0; }
    if(!strcmp(key, "bvh")) { if(value != 0 && value != 1) return fail(ctx, "bvh must be 0 or 1"); ctx->bvh = (int)value; return 0; }
    if(!strcmp(key, "tri_threshold")) { if(value < 1 || value > 32) return fail(ctx, "tri_threshold must be 1..32"); ctx->tri_threshold = (int)value; return 0; }
    if(!strcmp(key, "xform_threshold")) { if(value != -1 && (value < 1 || value > 32)) return fail(ctx, "xform_threshold must be 1..32 or -1"); ctx->xform_threshold = (int)value; return 0; }
    if(!strcmp(key, "node_threshold")) { if(value < 1 || value > 32) return fail(ctx, "node_threshold must be 1..32"); ctx->node_threshold = (int)value; return 0; }
    if(!strcmp(key, "node_burst")) { if(value != -1 && (value < 1 || value > 64)) return fail(ctx, "node_burst must be 1..64 or -1"); ctx->node_burst = (int)value; return 0; }
    if(!strcmp(key, "validate")) { ctx->validate = value != 0; return 0; }
    if(!strcmp(key, "flat")) { if(value != 0 && value != 1) return fail(ctx, "flat must be 0 or 1");
        if(value == 1 && ctx->have_static && !ctx->have_flat) return fail(ctx, "flat = 1 must be set before the scene is uploaded (the flat BVH is built at upload)");
        ctx->flat = (int)value; return 0; }
    if(!strcmp(key, "sort")) { if(value != 0 && value != 1) return fail(ctx, "sort must be 0 or 1"); ctx->sort = (int)value; return 0; }
    if(!strcmp(key, "dyn_first")) { if(value != 0 && value != 1) return fail(ctx, "dyn_first must be 0 or 1"); ctx->dyn_first = (int)value; return 0; }
    if(!strcmp(key, "l2_persist")) { if(value < 0 || value > 100) return fail(ctx, "l2_persist is a percentage"); ctx->l2_persist = (int)value; return 0; }
    if(!strcmp(key, "plain_trace")) { if(value < 0 || value > 2) return fail(ctx, "plain_trace must be 0, 1 or 2"); ctx->plain_trace = (int)value; return 0; }
    if(!strcmp(key, "top_smem")) { if(value != 0 && value != 1) return fail(ctx, "top_smem must be 0 or 1"); ctx->top_smem = (int)value; return 0; }
    if(!strcmp(key, "lanes")) { if(value < 1 || value > 4096 || (value & (value - 1))) return fail(ctx, "lanes must be a power of two in 1..4096"); ctx->max_lanes = (int)value; return 0; }
    if(!strcmp(key, "pool_budget_mb")) { if(value < 1) return fail(ctx, "pool_budget_mb must be positive"); ctx->pool_budget_bytes = (size_t)value << 20; return 0; }
    if(!strcmp(key, "min_active")) { if(value < -1 || value > 32) return fail(ctx, "min_active must be -1..32"); ctx->min_active = (int)value; return 0; }
    return fail(ctx, "unknown option '%s'", key);
}

int ptgpu_read_counters(ptgpu_ctx* ctx, uint64_t out[PTGPU_CNT_COUNT])
{
    if(!ctx || !out) return 1;
    if(use(ctx)) return 1;
    CK(cudaStreamSynchronize(ctx->stream));
    Counters h;
    CK(cudaMemcpy(&h, ctx->counters.p, sizeof(h), cudaMemcpyDeviceToHost));
    CK(cudaMemset(ctx->counters.p, 0, sizeof(Counters)));
    for(int i = 0; i < PTGPU_CNT_COUNT; ++i) out[i] = h.v[i];
    return 0;
}

int ptgpu_scene_stats(ptgpu_ctx* ctx, uint64_t out[8])
{
    if(!ctx || !out) return 1;
    static const WideScene empty_scene;
    const WideScene& w = ctx->host ? ctx->host->wide : empty_scene;
    uint64_t ref_bytes = ctx->n_static_nodes * (24 + 64) + ctx->n_indices * 4 + ctx->n_verts * 64 + ctx->n_static * 160;
    uint64_t wide_bytes = w.nodes.size() * sizeof(WideNode) + w.tris.size() * 16 + w.tlas.size() * sizeof(WideNode) +
        ctx->n_static * sizeof(WideInstance) + ctx->n_indices * 4 + ctx->n_verts * 48 + ctx->n_static * 160;
    out[0] = wide_bytes;
    out[1] = ctx->n_subframes * 160 + ctx->n_dyn * (160 + sizeof(WideInstance)) + ctx->n_subframes * 8;
    out[2] = w.nodes.size();
    out[3] = w.tris.size() / 3;
    out[4] = ctx->n_static;
    out[5] = w.tlas.size();
    out[6] = ref_bytes;
    out[7] = w.blas.size();
    return 0;
}

int ptgpu_host_flatten_check(
    const ptgpu_bvh_node* nodes, size_t n_nodes, const ptgpu_bvh_link* links, size_t n_links,
    const uint32_t* indices, size_t n_indices, const ptgpu_float3* pos, size_t n_verts,
    const ptgpu_tlas_instance* instances, size_t n_static, uint64_t out[8], char* err, size_t err_len)
{
    std::string e;
    WideScene ws;
    if(err && err_len) err[0] = 0;
    if(!nodes || !links || !indices || !pos || !instances || !out || n_links != 8 * n_nodes) e = "bad arguments";
    else if(build_wide_scene(nodes, n_nodes, links, indices, n_indices, pos, n_verts, instances, n_static, ws, e))
    {
        uint64_t bad = verify_wide_scene(ws, n_static, e);
        out[0] = ws.blas.size(); out[1] = ws.nodes.size(); out[2] = ws.tris.size() / 3; out[3] = ws.tlas.size();
        out[4] = ws.max_stack; out[5] = bad; out[6] = WIDE_STACK; out[7] = ws.cw_nodes.size() / 5;
        if(bad == 0) return 0;
    }
    if(err && err_len) { strncpy(err, e.c_str(), err_len - 1); err[err_len - 1] = 0; }
    return 1;
}

int ptgpu_host_flat_check(
    const ptgpu_bvh_node* nodes, size_t n_nodes, const ptgpu_bvh_link* links, size_t n_links,
    const uint32_t* indices, size_t n_indices, const ptgpu_float3* pos, size_t n_verts,
    const ptgpu_tlas_instance* instances, size_t n_static, uint64_t out[8], char* err, size_t err_len)
{
    std::string e;
    WideScene ws;
    FlatScene fs;
    if(err && err_len) err[0] = 0;
    if(!nodes || !links || !indices || !pos || !instances || !out || n_links != 8 * n_nodes) e = "bad arguments";
    else if(build_wide_scene(nodes, n_nodes, links, indices, n_indices, pos, n_verts, instances, n_static, ws, e))
    {
        const uint32_t node_base = (uint32_t)(ws.cw_nodes.size() / 5), tri_base = (uint32_t)(ws.cw_tris.size() / 3);
        if(build_flat_scene(ws, indices, pos, instances, n_static, node_base, tri_base, fs, e))
        {
            uint64_t bad = verify_flat_scene(fs, node_base, tri_base, e);
            out[0] = fs.n_tris; out[1] = fs.nodes.size() / 5; out[2] = fs.n_top; out[3] = fs.depth; out[4] = bad;
            out[5] = (uint64_t)(1e3 * fs.build_seconds); out[6] = 0; out[7] = 0;
            if(bad == 0) return 0;
        }
    }
    if(err && err_len) { strncpy(err, e.c_str(), err_len - 1); err[err_len - 1] = 0; }
    return 1;
}

int ptgpu_host_prepare_static(
    const ptgpu_bvh_node* nodes, size_t n_nodes, const ptgpu_bvh_link* links, size_t n_links,
    const uint32_t* indices, size_t n_indices, const ptgpu_float3* pos, size_t n_verts,
    const ptgpu_tlas_instance* instances, size_t n_static, int32_t flat, char* err, size_t err_len)
{
    std::string e;
    if(err && err_len) err[0] = 0;
    if(!nodes || !links || !indices || !pos || !instances || n_links != 8 * n_nodes || n_static == 0) e = "bad arguments";
    else if(get_host_scene(nodes, n_nodes, links, n_links, indices, n_indices, pos, n_verts, instances, n_static, nullptr, 0, flat != 0, e)) return 0;
    if(err && err_len) { strncpy(err, e.c_str(), err_len - 1); err[err_len - 1] = 0; }
    return 1;
}

int ptgpu_host_check_dynamic_ranges(const uint32_t* dyn_begin, const uint32_t* dyn_end, size_t n_subframes, size_t n_dyn,
                                    char* err, size_t err_len)
{
    std::vector<uint2> ranges;
    std::string why;
    if(err && err_len) err[0] = 0;
    if(!dyn_begin || !dyn_end || n_subframes == 0) why = "bad arguments";
    else if(encode_dynamic_ranges(dyn_begin, dyn_end, n_subframes, n_dyn, ranges, why)) return 0;
    if(err && err_len) { strncpy(err, why.c_str(), err_len - 1); err[err_len - 1] = 0; }
    return 1;
}

int ptgpu_host_build_check(
    const uint32_t* indices, size_t n_indices, const ptgpu_float3* pos, size_t n_verts,
    const ptgpu_mesh* meshes, size_t n_meshes,
    const ptgpu_tlas_instance* instances, size_t n_static, uint64_t out[8], char* err, size_t err_len)
{
    std::string e;
    WideScene ws;
    if(err && err_len) err[0] = 0;
    if(!indices || !pos || !meshes || !instances || !out || n_meshes == 0) e = "bad arguments";
    else if(build_wide_scene(nullptr, 0, nullptr, indices, n_indices, pos, n_verts, instances, n_static, ws, e, meshes, n_meshes))
    {
        uint64_t bad = verify_wide_scene(ws, n_static, e);
        out[0] = ws.blas.size(); out[1] = ws.nodes.size(); out[2] = ws.tris.size() / 3; out[3] = ws.tlas.size();
        out[4] = ws.max_stack; out[5] = bad; out[6] = WIDE_STACK; out[7] = ws.cw_nodes.size() / 5;
        if(bad == 0) return 0;
    }
    if(err && err_len) { strncpy(err, e.c_str(), err_len - 1); err[err_len - 1] = 0; }
    return 1;
}

int ptgpu_get_stat(ptgpu_ctx* ctx, const char* key, uint64_t* out)
{
    if(!ctx || !key || !out) return 1;
    if(!strcmp(key, "validate_mismatches")) { *out = ctx->last_validate_mismatches; return 0; }
    if(!strcmp(key, "wave_rounds")) { *out = (uint64_t)ctx->last_wave_rounds; return 0; }
    if(!strcmp(key, "wave_lanes")) { *out = ctx->last_wave_lanes; return 0; }
    if(!strcmp(key, "pool_bytes")) { *out = ctx->last_pool_bytes; return 0; }
    if(!strcmp(key, "trace_us")) { wave_timing(ctx); *out = (uint64_t)(ctx->last_trace_us + 0.5); return 0; }
    if(!strcmp(key, "shade_us")) { wave_timing(ctx); *out = (uint64_t)(ctx->last_shade_us + 0.5); return 0; }
    if(!strcmp(key, "trace_launches")) { wave_timing(ctx); *out = ctx->last_trace_launches; return 0; }
    if(!strcmp(key, "sort_us")) { wave_timing(ctx); *out = (uint64_t)(ctx->last_sort_us + 0.5); return 0; }
    if(!strcmp(key, "flat_tris")) { *out = ctx->have_flat ? ctx->flat_tris : 0; return 0; }
    if(!strcmp(key, "l2_set_aside")) { *out = (double)ctx->stat_l2_set_aside; return 0; }
    if(!strcmp(key, "l2_window")) { *out = (double)ctx->stat_l2_window; return 0; }
    if(!strcmp(key, "flat_nodes")) { *out = ctx->have_flat ? ctx->flat_nodes : 0; return 0; }
    if(!strcmp(key, "flat_depth")) { *out = ctx->have_flat ? ctx->flat_depth : 0; return 0; }
    if(!strcmp(key, "flat_build_ms")) { *out = ctx->have_flat ? (uint64_t)(1e3 * ctx->flat_build_seconds) : 0; return 0; }
    return fail(ctx, "unknown stat '%s'", key);
}

int ptgpu_set_animation_frame(ptgpu_ctx* ctx, const ptgpu_anim* anim, uint32_t frame)
{
    if(!ctx) return 1;
    if(!anim) return fail(ctx, "ptgpu_set_animation_frame: null animation");
    const size_t n_sub = ptgpu_anim_subframe_count(anim);
    std::vector<ptgpu_subframe> sub(n_sub);
    std::vector<ptgpu_tlas_instance> dyn(ptgpu_anim_max_instances(anim));
    std::vector<uint32_t> b(n_sub), e(n_sub);
    size_t n_dyn = 0;
    if(ptgpu_anim_frame(anim, frame, sub.data(), dyn.data(), &n_dyn, b.data(), e.data()) != 0)
        return fail(ctx, "ptgpu_anim_frame failed");
    return ptgpu_set_frame_ranges(ctx, sub.data(), n_sub, dyn.data(), n_dyn, b.data(), e.data());
}

} // extern "C"
