// Wavefront renderer: the same per-sample algorithm as pt_mega.cuh (path_trace_pixel,
// path_tracer.hh:637-741) split into homogeneous kernels over a pool of path slots.
//
// Why: in the megakernel ncu measured 6.9 of 32 lanes active per instruction — traversal lengths
// vary so much that a warp mostly waits for its slowest ray, and shading runs for the few lanes that
// happen to be done. Here
//   * wf_trace is a persistent kernel in which a lane whose ray ends immediately takes the next
//     ray from a global queue (closest-hit and shadow rays alike), so traversal lanes stay busy;
//   * wf_shade<FAR> processes finished closest-hit queries sorted into two queues: FAR (miss or
//     t >= 1e3: needs the sky ray-march, path_tracer.hh:513) and NEAR (skips it);
//   * wf_generate starts the next sample of slots whose path ended (camera rays, :655-671).
// A slot is (pixel, sample lane l): it traces samples k = l, l+L, ... of the job's sample set one
// after another and keeps their running sum. L (slots per pixel) is as large as the pool budget
// allows — 256 at the shipped config, i.e. one sample per slot and five rounds per frame: with L = 8
// (160 rounds) 23 % of all lane-iterations of wf_trace were idle in the tails of the launches.
#pragma once
#include "pt_kernels.cuh"
#include "pt_wide.cuh"
#include "pt_cwbvh.cuh"

namespace pt {

#define WF_INVALID 0xFFFFFFFFu
#define WF_SHADOW_BIT 0x80000000u
#define WF_SEG_BOUNCE 0
#define WF_SEG_PRIMARY 1
#define WF_SEG_SHADOW 2
#define WF_NEED_SAMPLE (-1)   // cursor.y: the slot waits for wf_generate
#define WF_FINISHED (-2)      // cursor.y: no samples left
#define WF_ST_NONE 0          // status: nothing to shade
#define WF_ST_NEAR 2          // status: closest-hit query done, 0 < t < 1e3 (no sky march, path_tracer.hh:513)
#define WF_ST_FAR 3           // status: closest-hit query done, miss or t >= 1e3

struct WaveCounters
{
    uint32_t n_seg[3];                        // ray-queue segment extents: 0 bounce, 1 primary, 2 shadow
    uint32_t n_far, n_near, n_new;            // shade-queue extents; slots waiting for their next sample
    uint32_t cur_trace;                       // fetch cursor of wf_trace over the concatenated segments
    uint32_t pad;
};

struct WaveBuffers
{
    // per slot
    uint4* rng;
    float4* ray_o;        // xyz origin, w tmin
    float4* ray_d;        // xyz extension-ray direction, w bsdf_pdf of the sample that made it
    float4* shadow_d;     // xyz NEE direction, w jitter for the attenuation march
    float4* hit;          // t, u, v, instance (bits)
    uint32_t* hit_prim;   // primitive | back_face << 31
    float4* atten;        // xyz path attenuation, w regularization
    float4* contrib;      // xyz contribution of the current sample, w unused
    float4* nee;          // xyz pending NEE radiance (before visibility/attenuation), w 1 = pending
    float4* sum;          // xyz sum over finished samples
    int2* cursor;         // x = k (index into the sample set), y = bounce
    uint32_t* visible;    // shadow query result: 1 = unoccluded
    // queues
    // The ray queue has one segment per ray kind so that a warp's 64-entry fetch is homogeneous:
    // primary rays of neighbouring pixels stay together (coherent), sun-ward shadow rays share a
    // direction, bounce rays are incoherent anyway. Entry = slot | WF_SHADOW_BIT.
    uint32_t* q_trace;    // 3 segments of seg_cap entries
    // Ray sort (wf_sort_*): shade writes a 13-bit key beside every bounce / shadow entry (direction octant
    // + Morton cell of the origin); one counting-sort pass per segment reorders the entries into q_sorted,
    // which is what wf_trace_cw then fetches from (the primary segment is already in pixel-tile order).
    uint32_t* q_key;      // 3 segments of seg_cap keys, parallel to q_trace (the primary segment's are unused)
    uint32_t* q_sorted;   // 2 segments of seg_cap entries: sorted bounce rays, sorted shadow rays
    uint32_t* sort_hist;  // 2 x WF_SORT_BINS: per-bin counts, then the running output cursors
    int32_t sort;         // 0: wf_trace_cw reads q_trace as it was appended
    uint8_t* status;      // per slot: WF_ST_*, written by wf_trace when a closest-hit query ends
    uint32_t* q_far;      // shade queues, filled by wf_classify in slot order
    uint32_t* q_near;
    WaveCounters* cnt;
    unsigned long long* stats;   // WF_STATS builds only
    uint32_t n_slots, seg_cap;
    uint32_t lanes, lane_shift;  // slots per pixel (power of two): slot = pixel << lane_shift | lane
    int32_t tiles_x;
};

constexpr int WF_TILE = 8;   // pixels per tile side in the slot order

PT_D bool slot_pixel(const WaveBuffers& wb, const RenderJob& job, uint32_t slot, int& lx, int& ly)
{
    const uint32_t pi = slot >> wb.lane_shift;
    const uint32_t tile = pi / (WF_TILE * WF_TILE), in = pi % (WF_TILE * WF_TILE);
    lx = (int)(tile % wb.tiles_x) * WF_TILE + (int)(in % WF_TILE);
    ly = (int)(tile / wb.tiles_x) * WF_TILE + (int)(in / WF_TILE);
    return lx < job.w && ly < job.h;
}

#ifndef WF_SHADE_PREFETCH
#define WF_SHADE_PREFETCH 1
#endif
PT_D void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

// Warp-aggregated append; must be reached by all 32 lanes of the warp.
PT_D void wf_append(uint32_t* q, uint32_t* count, bool pred, uint32_t val)
{
    const unsigned m = __ballot_sync(0xFFFFFFFFu, pred);
    if(m == 0u) return;
    const unsigned lane = threadIdx.x & 31u;
    uint32_t base = 0;
    if(lane == (unsigned)(__ffs(m) - 1)) base = atomicAdd(count, (uint32_t)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, __ffs(m) - 1);
    if(pred) q[base + __popc(m & ((1u << lane) - 1u))] = val;
}

// Block-aggregated append: one atomic per block and call instead of one per warp (1.8 M per round on one
// address with per-warp appends). Must be reached by every thread of the block the same number of times;
// `s_tmp` is WARPS + 1 words of shared memory per call site. Block order = warp order = slot order.
template<int WARPS>
PT_D void wf_append_block(uint32_t* q, uint32_t* count, bool pred, uint32_t val, uint32_t* s_tmp,
                          uint32_t* q_key = nullptr, uint32_t key = 0u)
{
    const unsigned m = __ballot_sync(0xFFFFFFFFu, pred);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if(lane == 0u) s_tmp[warp] = (uint32_t)__popc(m);
    __syncthreads();
    if(threadIdx.x == 0)
    {
        uint32_t total = 0;
        #pragma unroll
        for(int k = 0; k < WARPS; ++k) total += s_tmp[k];
        s_tmp[WARPS] = total ? atomicAdd(count, total) : 0u;
    }
    __syncthreads();
    if(pred)
    {
        uint32_t base = s_tmp[WARPS];
        for(unsigned k = 0; k < warp; ++k) base += s_tmp[k];
        const uint32_t pos = base + __popc(m & ((1u << lane) - 1u));
        q[pos] = val;
        if(q_key) q_key[pos] = key;
    }
    __syncthreads();
}

// Two block-aggregated appends behind one set of barriers (the shade kernels append a bounce and a shadow ray
// per slot: ncu put 10 % of wf_shade<NEAR>'s stall samples on the six barriers of two separate appends).
// `s_tmp` is 2 * (WARPS + 1) words.
template<int WARPS>
PT_D void wf_append2_block(uint32_t* q0, uint32_t* count0, bool pred0, uint32_t val0, uint32_t* key0, uint32_t k0,
                           uint32_t* q1, uint32_t* count1, bool pred1, uint32_t val1, uint32_t* key1, uint32_t k1,
                           uint32_t* s_tmp)
{
    const unsigned m0 = __ballot_sync(0xFFFFFFFFu, pred0), m1 = __ballot_sync(0xFFFFFFFFu, pred1);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if(lane == 0u) { s_tmp[warp] = (uint32_t)__popc(m0); s_tmp[WARPS + 1 + warp] = (uint32_t)__popc(m1); }
    __syncthreads();
    if(threadIdx.x < 2)
    {
        uint32_t* t = s_tmp + threadIdx.x * (WARPS + 1);
        uint32_t total = 0;
        #pragma unroll
        for(int k = 0; k < WARPS; ++k) total += t[k];
        t[WARPS] = total ? atomicAdd(threadIdx.x ? count1 : count0, total) : 0u;
    }
    __syncthreads();
    if(pred0)
    {
        uint32_t base = s_tmp[WARPS];
        for(unsigned k = 0; k < warp; ++k) base += s_tmp[k];
        const uint32_t pos = base + __popc(m0 & ((1u << lane) - 1u));
        q0[pos] = val0;
        if(key0) key0[pos] = k0;
    }
    if(pred1)
    {
        uint32_t base = s_tmp[2 * WARPS + 1];
        for(unsigned k = 0; k < warp; ++k) base += s_tmp[WARPS + 1 + k];
        const uint32_t pos = base + __popc(m1 & ((1u << lane) - 1u));
        q1[pos] = val1;
        if(key1) key1[pos] = k1;
    }
    __syncthreads();
}

// ---- ray sort ---------------------------------------------------------------------------------------------
// Bounce rays leave a surface in a random direction of the hemisphere and, from the second bounce on,
// from anywhere in the scene: a warp of 32 consecutive queue entries walked 32 unrelated parts of the BVH
// (19 of 32 lanes per instruction in wf_trace_cw, 11 in its triangle step). Key: the direction octant — the
// order the 8-wide nodes store their children in, and what the reference's eight link tables are indexed by
// (ray_query.hh:135-140) — and the Morton code of the origin's cell in a grid over the static
// scene. Shadow rays all point at the sun (4 degree cone), so their key spends all 13 bits on the origin.
// 13-bit keys: 32 KB of shared-memory counters per block. 15-bit keys (64 x 8 x 64 cells for shadow rays, 32 x 4 x 32
// + octant for bounce rays) measured the same traversal time and 1 ms more sorting per frame.
constexpr uint32_t WF_SORT_BINS = 8192;
constexpr uint32_t WF_SORT_CHUNK = 32768;    // entries per block iteration
constexpr int WF_SORT_THREADS = 512;

PT_D uint32_t spread_bits(uint32_t v)
{   // bit i of v -> bit 2i
    v = (v | (v << 4)) & 0x0F0F0F0Fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}
PT_D uint32_t sort_key(const Scene& sc, v3 o, v3 d, bool shadow)
{
    // cell of the origin in a 32 x 8 x 32 grid over the static scene
    const uint32_t cx = (uint32_t)fminf(fmaxf((o.x - sc.key_lo[0]) * sc.key_scale[0], 0.0f), 31.0f);
    const uint32_t cy = (uint32_t)fminf(fmaxf((o.y - sc.key_lo[1]) * sc.key_scale[1], 0.0f), 7.0f);
    const uint32_t cz = (uint32_t)fminf(fmaxf((o.z - sc.key_lo[2]) * sc.key_scale[2], 0.0f), 31.0f);
    // shadow: x5 z5 (Morton) y3; bounce: x4 z4 (Morton) y2, octant 3
    if(shadow) return ((spread_bits(cx) | (spread_bits(cz) << 1)) << 3) | cy;
    const uint32_t oct = (d.x < 0.0f ? 0u : 4u) | (d.y < 0.0f ? 0u : 2u) | (d.z < 0.0f ? 0u : 1u);
    return ((((spread_bits(cx >> 1) | (spread_bits(cz >> 1) << 1)) << 2) | (cy >> 1)) << 3) | oct;
}

// pass 1: per-bin counts of both segments (blockIdx.y: 0 bounce, 1 shadow), one shared-memory histogram
// per block, one global atomic per non-empty bin and block
__global__ void __launch_bounds__(WF_SORT_THREADS)
wf_sort_count_kernel(WaveBuffers wb)
{
    __shared__ uint32_t h[WF_SORT_BINS];
    const int seg = blockIdx.y ? WF_SEG_SHADOW : WF_SEG_BOUNCE;
    const uint32_t n = wb.cnt->n_seg[seg];
    if((size_t)blockIdx.x * WF_SORT_CHUNK >= n) return;
    const uint32_t* keys = wb.q_key + seg * (size_t)wb.seg_cap;
    for(uint32_t b = threadIdx.x; b < WF_SORT_BINS; b += WF_SORT_THREADS) h[b] = 0u;
    __syncthreads();
    for(size_t c = blockIdx.x; c * WF_SORT_CHUNK < n; c += gridDim.x)
    {
        const uint32_t begin = (uint32_t)(c * WF_SORT_CHUNK), end = min(n, begin + WF_SORT_CHUNK);
        for(uint32_t i = begin + threadIdx.x; i < end; i += WF_SORT_THREADS) atomicAdd(&h[keys[i] & (WF_SORT_BINS - 1u)], 1u);
    }
    __syncthreads();
    uint32_t* hist = wb.sort_hist + blockIdx.y * WF_SORT_BINS;
    for(uint32_t b = threadIdx.x; b < WF_SORT_BINS; b += WF_SORT_THREADS) if(h[b]) atomicAdd(&hist[b], h[b]);
}

// pass 2: exclusive scan of the bin counts, in place (one block per segment)
__global__ void __launch_bounds__(1024)
wf_sort_scan_kernel(WaveBuffers wb)
{
    __shared__ uint32_t s_warp[32];
    uint32_t* hist = wb.sort_hist + blockIdx.x * WF_SORT_BINS;
    constexpr uint32_t PER = WF_SORT_BINS / 1024;
    uint32_t v[PER], sum = 0;
    #pragma unroll
    for(uint32_t k = 0; k < PER; ++k) { v[k] = hist[threadIdx.x * PER + k]; sum += v[k]; }
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t incl = sum;
    #pragma unroll
    for(int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if(lane >= (unsigned)o) incl += t; }
    if(lane == 31u) s_warp[warp] = incl;
    __syncthreads();
    if(warp == 0)
    {
        uint32_t w = s_warp[lane], wi = w;
        #pragma unroll
        for(int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, wi, o); if(lane >= (unsigned)o) wi += t; }
        s_warp[lane] = wi - w;
    }
    __syncthreads();
    uint32_t run = s_warp[warp] + incl - sum;
    #pragma unroll
    for(uint32_t k = 0; k < PER; ++k) { hist[threadIdx.x * PER + k] = run; run += v[k]; }
}

// pass 3: scatter. A block counts its chunk again, claims a run in every bin it touches with one global
// atomic, and places its entries with shared-memory atomics. The order inside a bin is the order in which
// blocks claimed their runs — not reproducible, and it need not be: a query's result does not depend on
// the queue order (ties are broken by ids, pt_cwbvh.cuh), and the shade queues are rebuilt in slot order.
__global__ void __launch_bounds__(WF_SORT_THREADS)
wf_sort_scatter_kernel(WaveBuffers wb)
{
    __shared__ uint32_t h[WF_SORT_BINS];
    const int seg = blockIdx.y ? WF_SEG_SHADOW : WF_SEG_BOUNCE;
    const uint32_t n = wb.cnt->n_seg[seg];
    const uint32_t* keys = wb.q_key + seg * (size_t)wb.seg_cap;
    const uint32_t* in = wb.q_trace + seg * (size_t)wb.seg_cap;
    uint32_t* out = wb.q_sorted + blockIdx.y * (size_t)wb.seg_cap;
    uint32_t* cursor = wb.sort_hist + blockIdx.y * WF_SORT_BINS;
    for(size_t c = blockIdx.x; c * WF_SORT_CHUNK < n; c += gridDim.x)
    {
        const uint32_t begin = (uint32_t)(c * WF_SORT_CHUNK), end = min(n, begin + WF_SORT_CHUNK);
        for(uint32_t b = threadIdx.x; b < WF_SORT_BINS; b += WF_SORT_THREADS) h[b] = 0u;
        __syncthreads();
        for(uint32_t i = begin + threadIdx.x; i < end; i += WF_SORT_THREADS) atomicAdd(&h[keys[i] & (WF_SORT_BINS - 1u)], 1u);
        __syncthreads();
        for(uint32_t b = threadIdx.x; b < WF_SORT_BINS; b += WF_SORT_THREADS) if(h[b]) h[b] = atomicAdd(&cursor[b], h[b]);
        __syncthreads();
        for(uint32_t i = begin + threadIdx.x; i < end; i += WF_SORT_THREADS)
            out[atomicAdd(&h[keys[i] & (WF_SORT_BINS - 1u)], 1u)] = in[i];
        __syncthreads();
    }
}

// ---- init: every slot starts idle; valid ones are queued for their first sample ---------------
__global__ void wf_init_kernel(WaveBuffers wb, RenderJob job)
{
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if(slot >= wb.n_slots) return;
    int lx, ly;
    const bool valid = slot_pixel(wb, job, slot, lx, ly) && (int)(slot & (wb.lanes - 1u)) < job.s_count;
    wb.sum[slot] = make_float4(0, 0, 0, 0);
    wb.status[slot] = WF_ST_NONE;
    wb.cursor[slot] = make_int2((int)(slot & (wb.lanes - 1u)), valid ? WF_NEED_SAMPLE : WF_FINISHED);
    if(slot == 0)
    {
        wb.cnt->n_new = 1; wb.cnt->n_seg[0] = 0; wb.cnt->n_seg[1] = 0; wb.cnt->n_seg[2] = 0;
        wb.cnt->n_far = 0; wb.cnt->n_near = 0; wb.cnt->cur_trace = 0;
    }
}

// ---- generate: camera ray of the slot's next sample (path_tracer.hh:655-671) --------------------
__global__ void __launch_bounds__(256)
wf_generate_kernel(Scene sc, RenderJob job, WaveBuffers wb)
{
    // Scans all slots in slot order (pixel tiles), so the primary rays of one 2x2-pixel block and one
    // motion-blur subframe land next to each other in the primary segment.
    if(wb.cnt->n_new == 0u) return;
    __shared__ uint32_t s_tmp[9];
    const uint32_t rounded = (wb.n_slots + 255u) & ~255u;   // whole blocks: every thread of a block loops equally often
    for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < rounded; i += gridDim.x * blockDim.x)
    {
        const uint32_t slot = i;
        const bool ok = i < wb.n_slots && wb.cursor[i].y == WF_NEED_SAMPLE;
        if(ok)
        {
            int lx, ly;
            slot_pixel(wb, job, slot, lx, ly);
            const int k = wb.cursor[slot].x;
            const int sample = job.s_begin + k * job.s_stride;
            const uint32_t px = (uint32_t)(job.x0 + lx), py = (uint32_t)(job.y0 + ly);
            const uint32_t subframe = sample < 0 ? 0u : (uint32_t)sample / (uint32_t)sc.samples_per_subframe;
            rng4 seed = {px, py, (uint32_t)sample, sc.student_id};
            pcg4d(seed);
            float4 u = rand4(seed);
            v2 film = sample_gaussian_disk(u.x, u.y, 0.4f);
            v3 d, o;
            camera_ray(sc, sc.subframes + subframe, u.z, u.w, (float)px + (film.x + 0.5f), (float)py + (film.y + 0.5f), d, o);
            wb.rng[slot] = make_uint4(seed.x, seed.y, seed.z, seed.w);
            wb.ray_o[slot] = make_float4(o.x, o.y, o.z, 0.0f);
            wb.ray_d[slot] = make_float4(d.x, d.y, d.z, -1.0f); // bsdf_pdf -1: MIS weight 1, |pdf| 1, no regularisation
            // attenuation (1,1,1 | regularization 1), contribution 0 and "no NEE pending" are implied by bounce 0:
            // the shade kernels do not read them for a camera ray (48 B written + 48 B read less per path)
            wb.cursor[slot] = make_int2(k, 0);
        }
        wf_append_block<8>(wb.q_trace + WF_SEG_PRIMARY * (size_t)wb.seg_cap, &wb.cnt->n_seg[WF_SEG_PRIMARY], ok, slot, s_tmp);
    }
}

// ---- trace: persistent traversal kernel -----------------------------------------------------------
#ifndef WF_TRACE_THREADS_N
#define WF_TRACE_THREADS_N 128     // (64 x 18 and 96 x 12 blocks per SM measured the same)
#endif
constexpr int WF_TRACE_THREADS = WF_TRACE_THREADS_N;
#ifndef WF_FETCH_CHUNK_N
#define WF_FETCH_CHUNK_N 64
#endif
constexpr int WF_FETCH_CHUNK = WF_FETCH_CHUNK_N;

// ---- trace on the compressed 8-wide BVH (pt_cwbvh.cuh) ---------------------------------------------
//
// Scheduling inside a warp (each decision backed by an ncu source-level profile, profiles/):
//   * one code block per iteration, elected by ballot: NODE (box tests of one 8-wide node), TRI (one
//     triangle test per lane) or ENTER (transform the ray into an instance);
//   * leaf groups found by NODE go to a small per-lane pending list, so a lane keeps descending while
//     its triangles wait for the TRI block to be worth running (>= tri_threshold lanes): unvoted, the
//     triangle test ran with 2.5 of 32 lanes and was 37 % of all issued instructions;
//   * leaving an instance is free: the world-space ray constants are parked under the exit marker;
//   * a lane whose query ends takes the next ray from the global queue (refill when >= min_active
//     lanes are idle), closest-hit and shadow rays alike.
#ifndef CW_PEND_N
#define CW_PEND_N 4
#endif
constexpr int CW_PEND = CW_PEND_N;

// Resident blocks per SM. With the instanced scene 7 (72 registers, no spills) beat 8 (64 registers, 110 B of
// spills) by 7 %; with the flat scene the kernel waits on memory more and 8 blocks won 3.5 % in spite of the
// spills. With the lean traversal state (pt_cwbvh.cuh: no world ray, sign bits, subframe or hit record in
// registers) the kernel needs 56 registers without spills: 9 blocks = 36 warps, -8 % against the spilling
// 64-register kernel. 10 blocks (48 registers) spill again: +8 %.
#ifndef WF_CW_BLOCKS
#define WF_CW_BLOCKS 9
#endif
#ifndef WF_TRI_BURST
#define WF_TRI_BURST 4     // triangle steps per TRI block (2: +4.9 %, 3: +0.5 %, 6 and 16: +0.3 %; profiles/r02_trace_kernel_history.md)
#endif

// The world-space ray of the lane's query, read back from the path-state pool on the rare occasions the
// traversal needs it again (entering or leaving one of the <= 7 per-frame instances: ~5 M times per frame
// against 220 M queries): six registers and the subframe index the traversal state does not hold (pt_cwbvh.cuh).
struct SlotRay
{
    const Scene& sc;
    const RenderJob& job;
    const WaveBuffers& wb;
    const uint32_t& slot;
    const bool& shadow;
    PT_D v3 origin() const { const float4 f = wb.ray_o[slot]; return mk3(f.x, f.y, f.z); }
    PT_D v3 dir() const { const float4 f = shadow ? wb.shadow_d[slot] : wb.ray_d[slot]; return mk3(f.x, f.y, f.z); }
    PT_D uint32_t subframe() const
    {
        const int sample = job.s_begin + wb.cursor[slot].x * job.s_stride;
        return sample < 0 ? 0u : (uint32_t)sample / (uint32_t)sc.samples_per_subframe;
    }
};
template<bool TOP>
__global__ void __launch_bounds__(WF_TRACE_THREADS, WF_CW_BLOCKS)
wf_trace_cw_kernel(Scene sc, RenderJob job, WaveBuffers wb)
{
    // TOP: the first CW_TOP_NODES nodes of the flat BVH (its top levels, breadth-first) in shared memory
    __shared__ float4 s_top[TOP ? 5 * CW_TOP_NODES : 1];
    TopLevels top{s_top, sc.flat_root, 0u};
    if(TOP && sc.flat_root != 0xFFFFFFFFu)
    {
        top.n = min(sc.flat_top, CW_TOP_NODES);
        for(uint32_t i = threadIdx.x; i < 5u * top.n; i += WF_TRACE_THREADS) s_top[i] = __ldg(sc.cwnodes + 5 * (size_t)sc.flat_root + i);
        __syncthreads();
    }
    const unsigned lane = threadIdx.x & 31u;
    uint32_t res_base = 0, res_n = 0;
    bool exhausted = false;

    bool active = false;
    uint32_t slot = 0;
    __shared__ uint2 s_stack[CW_SM_STACK + 2][WF_TRACE_THREADS];   // + the closest-hit record (pt_cwbvh.cuh)
    __shared__ uint2 s_pend[CW_PEND][WF_TRACE_THREADS];
    uint2 stack_overflow[CW_STACK - CW_SM_STACK];
    HybridStack stack{&s_stack[0][threadIdx.x], stack_overflow, WF_TRACE_THREADS};
    uint2* pend = &s_pend[0][threadIdx.x];   // entry i at pend[i * WF_TRACE_THREADS]
    int np = 0;
    CwState st;
    st.ngroup = make_uint2(0u, 0u); st.tgroup = make_uint2(0u, 0u); st.sp = 0; st.in_blas = false; st.any = false;
    const SlotRay world{sc, job, wb, slot, st.any};

    for(;;)
    {
        const unsigned act = __ballot_sync(0xFFFFFFFFu, active);
        if(32 - __popc(act) >= job.min_active || act == 0u)
        {
            // -- refill idle lanes from the ray queue ------------------------------------------------
            // (the segment extents are read here, not held in registers across the traversal)
            const uint32_t n_s0 = wb.cnt->n_seg[0], n_s1 = wb.cnt->n_seg[1], n_s2 = wb.cnt->n_seg[2];
            const uint32_t n_entries = n_s0 + n_s1 + n_s2;
            const unsigned idle = ~act;
            const uint32_t want = (uint32_t)__popc(idle);
            const uint32_t my_rank = (uint32_t)__popc(idle & ((1u << lane) - 1u));
            uint32_t entry_index = WF_INVALID;
            uint32_t served = 0;
            while(want > served && !(exhausted && res_n == 0u))
            {
                if(res_n == 0u)
                {
                    uint32_t b = 0;
                    if(lane == 0) b = atomicAdd(&wb.cnt->cur_trace, (uint32_t)WF_FETCH_CHUNK);
                    b = __shfl_sync(0xFFFFFFFFu, b, 0);
                    if(b >= n_entries) { exhausted = true; break; }
                    res_base = b;
                    res_n = min((uint32_t)WF_FETCH_CHUNK, n_entries - b);
                }
                const uint32_t take = min(res_n, want - served);
                if(!active && my_rank >= served && my_rank < served + take) entry_index = res_base + (my_rank - served);
                res_base += take; res_n -= take; served += take;
            }
            if(entry_index != WF_INVALID)
            {
                // logical index over [bounce | primary | shadow] -> segment position (sorted copies of the
                // bounce and shadow segments when the ray sort is on)
                const uint32_t* seg_bounce = wb.sort ? wb.q_sorted : wb.q_trace;
                const uint32_t* seg_shadow = wb.sort ? wb.q_sorted + wb.seg_cap : wb.q_trace + 2 * (size_t)wb.seg_cap;
                const uint32_t e = entry_index < n_s0 ? __ldg(seg_bounce + entry_index) :
                    entry_index - n_s0 < n_s1 ? __ldg(wb.q_trace + wb.seg_cap + (entry_index - n_s0)) :
                    __ldg(seg_shadow + (entry_index - n_s0 - n_s1));
                if(e != WF_INVALID)
                {
                    slot = e & ~WF_SHADOW_BIT;
                    const bool shadow = (e & WF_SHADOW_BIT) != 0u;
                    const float4 fo = wb.ray_o[slot];
                    const float4 fd = shadow ? wb.shadow_d[slot] : wb.ray_d[slot];
                    const int sample = job.s_begin + wb.cursor[slot].x * job.s_stride;
                    const uint32_t subframe = sample < 0 ? 0u : (uint32_t)sample / (uint32_t)sc.samples_per_subframe;
                    cw_begin<true>(sc, st, stack, subframe, mk3(fo.x, fo.y, fo.z), mk3(fd.x, fd.y, fd.z), fo.w, PT_MAX_RAY_DIST, shadow);
                    np = 0;
                    active = true;
                }
            }
            if(__ballot_sync(0xFFFFFFFFu, active) == 0u)
            {
                if(exhausted && res_n == 0u) break;
                continue;
            }
        }

        // The three code blocks as lambdas; each is executed by the lanes that want it, in bursts.
        auto advance = [&]() {
            // a lane with neither node children nor a leaf group pops (cheap, unvoted)
            if(active && st.ngroup.y <= 0x00FFFFFFu && st.tgroup.y == 0u)
            {
                if(st.sp == 0)
                {
                    if(np == 0)
                    {   // query complete: write the result
                        active = false;
                        if(st.any) wb.visible[slot] = st.hit ? 0u : 1u;
                        else
                        {
                            const float t = st.hit ? st.tmax : -1.0f;
                            const uint2 uv = stack.hit_uv(), id = stack.hit_id();
                            wb.hit[slot] = make_float4(t, __uint_as_float(uv.x), __uint_as_float(uv.y), __uint_as_float(id.x));
                            wb.hit_prim[slot] = id.y;
                            wb.status[slot] = (t > 0.0f && t < 1e3f) ? WF_ST_NEAR : WF_ST_FAR;
                        }
                    }
                }
                else
                {
                    const uint2 e = stack.get(st.sp - 1);
                    if(e.y == 0u)
                    {   // exit marker: leave the instance once its pending triangles are done
                        if(np == 0)
                        {
                            st.sp -= 3;
                            const uint2 a = stack.get(st.sp), b = stack.get(st.sp + 1);
                            st.o = world.origin();
                            st.idir = mk3(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(b.x));
                            st.oct_inv4 = (b.y & 0xFFu) * 0x01010101u;
                            st.in_blas = false;
                        }
                    }
                    else
                    {
                        st.sp--;
                        if(e.y > 0x00FFFFFFu)
                        {
                            st.ngroup = e;
                            // a node group met in world space with a flat static scene is that scene's root (dyn_first)
                            if(!st.in_blas && sc.flat_root != 0xFFFFFFFFu) cw_enter_flat(st, world.dir());
                        }
                        else st.tgroup = e;
                    }
                }
            }
        };
        auto wants_node = [&]() { return active && st.ngroup.y > 0x00FFFFFFu && np < CW_PEND && !(!st.in_blas && st.tgroup.y != 0u); };
        auto wants_tri = [&]() { return active && np > 0; };
        auto wants_enter = [&]() { return active && !st.in_blas && st.tgroup.y != 0u; };
        auto node_step = [&]() {
#ifdef WF_STATS
            const bool at_root = st.in_blas && st.ngroup.y == 0x80000000u;
#endif
            cw_node_phase<TOP>(sc, st, stack, top);
#ifdef WF_STATS
            if(at_root) { atomicAdd(&wb.stats[7], 1ull); if(st.ngroup.y <= 0x00FFFFFFu && st.tgroup.y == 0u) atomicAdd(&wb.stats[16], 1ull); }
#endif
            if(st.tgroup.y != 0u)
            {
                if(st.in_blas) { pend[(np++) * WF_TRACE_THREADS] = st.tgroup; st.tgroup.y = 0u; }              // triangles wait for the TRI block
                // instances wait below the TLAS nodes (entering them first measured 3-4 % slower on forest frames:
                // entries spread out in time and fewer lanes share an ENTER step)
                else if(st.ngroup.y > 0x00FFFFFFu) { stack.set(st.sp++, st.tgroup); st.tgroup.y = 0u; }
            }
        };
        auto tri_step = [&]() {
            // one triangle of the newest pending leaf group
            uint2 g = pend[(np - 1) * WF_TRACE_THREADS];
            const uint32_t bit = 31u - (uint32_t)__clz(g.y);
            g.y &= ~(1u << bit);
            if(g.y) pend[(np - 1) * WF_TRACE_THREADS] = g; else np--;
            cw_test_triangle(sc, st, stack, g.x + bit);
            if(st.any && st.hit) np = 0; // any hit ends a shadow query (sp and groups are cleared)
        };
        auto enter_step = [&]() {
            // enter one instance of the group; the rest of the group and the world-space ray constants are
            // parked under the exit marker, which makes leaving the instance a few loads
            const uint32_t bit = 31u - (uint32_t)__clz(st.tgroup.y);
            st.tgroup.y &= ~(1u << bit);
            if(st.tgroup.y) stack.set(st.sp++, st.tgroup);
            stack.set(st.sp++, make_uint2(__float_as_uint(st.idir.x), __float_as_uint(st.idir.y)));
            stack.set(st.sp++, make_uint2(__float_as_uint(st.idir.z), st.oct_inv4 & 0xFFu));
            const uint32_t inst_id = cw_decode_instance(sc, world, st.tgroup.x, bit);
#ifdef WF_STATS
            atomicAdd(&wb.stats[24 + min(sc.winst[inst_id].blas, 15u)], 1ull);
#endif
            cw_enter_instance(sc, st, stack, world, inst_id);
        };

        bool progress = false;
        // -- NODE burst: while enough lanes have node children, nothing else is looked at --------------
        #pragma unroll 1
        for(int b = 0; b < job.node_burst; ++b)
        {
            advance();
            const bool w = wants_node();
#ifdef WF_STATS
            {
                const int c_node = __popc(__ballot_sync(0xFFFFFFFFu, w));
                const int c_idle = __popc(__ballot_sync(0xFFFFFFFFu, !active));
                const int c_enter = __popc(__ballot_sync(0xFFFFFFFFu, wants_enter()));
                const int c_pendfull = __popc(__ballot_sync(0xFFFFFFFFu, active && np >= CW_PEND));
                const int c_exitwait = __popc(__ballot_sync(0xFFFFFFFFu, active && !w && !wants_enter() && np > 0 && np < CW_PEND));
                if(lane == 0)
                {
                    atomicAdd(&wb.stats[0], 1ull); atomicAdd(&wb.stats[1], (unsigned long long)c_node);
                    atomicAdd(&wb.stats[2], (unsigned long long)c_idle); atomicAdd(&wb.stats[3], (unsigned long long)c_enter);
                    atomicAdd(&wb.stats[4], (unsigned long long)c_pendfull); atomicAdd(&wb.stats[5], (unsigned long long)c_exitwait);
                    if(c_node >= job.node_threshold) atomicAdd(&wb.stats[6], 1ull);
                }
            }
#endif
            if(__popc(__ballot_sync(0xFFFFFFFFu, w)) < job.node_threshold) break;
            if(w) node_step();
            progress = true;
        }
        // -- TRI block: worth running once enough lanes hold pending triangles ---------------------------
        {
            bool w = wants_tri();
            int n = __popc(__ballot_sync(0xFFFFFFFFu, w));
            #pragma unroll 1
            for(int b = 0; b < WF_TRI_BURST && n >= job.tri_threshold; ++b)
            {
#ifdef WF_STATS
                if(lane == 0) { atomicAdd(&wb.stats[8], 1ull); atomicAdd(&wb.stats[9], (unsigned long long)n); }
#endif
                if(w) tri_step();
                progress = true;
                w = wants_tri();
                n = __popc(__ballot_sync(0xFFFFFFFFu, w));
            }
        }
        // -- ENTER block -----------------------------------------------------------------------------------
        {
            advance();   // (measured: without this second pop opportunity per iteration the frame is 4.6 % slower)
            const bool w = wants_enter();
            if(__popc(__ballot_sync(0xFFFFFFFFu, w)) >= job.xform_threshold)
            {
#ifdef WF_STATS
                { const int n_e = __popc(__ballot_sync(0xFFFFFFFFu, w)); if(lane == 0) { atomicAdd(&wb.stats[10], 1ull); atomicAdd(&wb.stats[11], (unsigned long long)n_e); } }
#endif
                if(w) enter_step();
                progress = true;
            }
        }
        if(!progress)
        {   // no block met its threshold: run the fullest one once so that every lane eventually advances
            const bool wn = wants_node(), wt = wants_tri(), we = wants_enter();
            const int nn = __popc(__ballot_sync(0xFFFFFFFFu, wn));
            const int nt = __popc(__ballot_sync(0xFFFFFFFFu, wt));
            const int ne = __popc(__ballot_sync(0xFFFFFFFFu, we));
#ifdef WF_STATS
            if(lane == 0) { atomicAdd(&wb.stats[12], 1ull); atomicAdd(&wb.stats[13], (unsigned long long)nn); atomicAdd(&wb.stats[14], (unsigned long long)nt); atomicAdd(&wb.stats[15], (unsigned long long)ne); }
#endif
            if(nn >= nt && nn >= ne) { if(wn) node_step(); }
            else if(nt >= ne) { if(wt) tri_step(); }
            else { if(we) enter_step(); }
        }
    }
}

// ---- reference point: the same queue traced by the plain single-ray loop, one thread per ray -------------
// (ptgpu_set_option "plain_trace" = 1 for all rounds, 2 for the primary round only; what the warp scheduling
// of wf_trace_cw_kernel is measured against, profiles/r02_trace_kernel_history.md)
__global__ void __launch_bounds__(128)
wf_trace_plain_kernel(Scene sc, RenderJob job, WaveBuffers wb)
{
    const uint32_t n_s0 = wb.cnt->n_seg[0], n_s1 = wb.cnt->n_seg[1], n_s2 = wb.cnt->n_seg[2];
    const uint32_t n = n_s0 + n_s1 + n_s2;
    const uint32_t* seg_bounce = wb.sort ? wb.q_sorted : wb.q_trace;
    const uint32_t* seg_shadow = wb.sort ? wb.q_sorted + wb.seg_cap : wb.q_trace + 2 * (size_t)wb.seg_cap;
    for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const uint32_t e = i < n_s0 ? seg_bounce[i] : i - n_s0 < n_s1 ? wb.q_trace[wb.seg_cap + (i - n_s0)] : seg_shadow[i - n_s0 - n_s1];
        if(e == WF_INVALID) continue;
        const uint32_t slot = e & ~WF_SHADOW_BIT;
        const bool shadow = (e & WF_SHADOW_BIT) != 0u;
        const float4 fo = wb.ray_o[slot];
        const float4 fd = shadow ? wb.shadow_d[slot] : wb.ray_d[slot];
        const int sample = job.s_begin + wb.cursor[slot].x * job.s_stride;
        const uint32_t subframe = sample < 0 ? 0u : (uint32_t)sample / (uint32_t)sc.samples_per_subframe;
        Hit h;
        if(shadow)
        {
            trace_cw<true>(sc, subframe, mk3(fo.x, fo.y, fo.z), mk3(fd.x, fd.y, fd.z), fo.w, PT_MAX_RAY_DIST, h);
            wb.visible[slot] = h.t < 0.0f ? 1u : 0u;
        }
        else
        {
            trace_cw<false>(sc, subframe, mk3(fo.x, fo.y, fo.z), mk3(fd.x, fd.y, fd.z), fo.w, PT_MAX_RAY_DIST, h);
            wb.hit[slot] = make_float4(h.t, h.u, h.v, __uint_as_float(h.inst));
            wb.hit_prim[slot] = h.prim | (h.back_face ? 0x80000000u : 0u);
            wb.status[slot] = (h.t > 0.0f && h.t < 1e3f) ? WF_ST_NEAR : WF_ST_FAR;
        }
    }
}

// ---- debug: re-trace every queue entry with the plain single-ray traversal and compare ---------------
// (ptgpu_set_option "validate" = 1; counts and prints mismatches of the scheduled kernel)
__global__ void wf_validate_kernel(Scene sc, RenderJob job, WaveBuffers wb, unsigned long long* mismatch)
{
    const uint32_t n_s0 = wb.cnt->n_seg[0], n_s1 = wb.cnt->n_seg[1], n_s2 = wb.cnt->n_seg[2];
    const uint32_t n = n_s0 + n_s1 + n_s2;
    for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const uint32_t e = i < n_s0 ? wb.q_trace[i] : i - n_s0 < n_s1 ? wb.q_trace[wb.seg_cap + (i - n_s0)] : wb.q_trace[2 * (size_t)wb.seg_cap + (i - n_s0 - n_s1)];
        if(e == WF_INVALID) continue;
        const uint32_t slot = e & ~WF_SHADOW_BIT;
        const bool shadow = (e & WF_SHADOW_BIT) != 0u;
        const float4 fo = wb.ray_o[slot];
        const float4 fd = shadow ? wb.shadow_d[slot] : wb.ray_d[slot];
        const int sample = job.s_begin + wb.cursor[slot].x * job.s_stride;
        const uint32_t subframe = sample < 0 ? 0u : (uint32_t)sample / (uint32_t)sc.samples_per_subframe;
        Hit h;
        uint32_t census[2] = {0u, 0u};
#ifdef WF_STATS
        struct Flush { unsigned long long* s; uint32_t* c; __device__ ~Flush() {
            atomicAdd(&s[20], (unsigned long long)c[0]); atomicAdd(&s[21], (unsigned long long)c[1]); atomicAdd(&s[22], 1ull); } } flush{wb.stats, census};
#endif
        if(shadow)
        {
            trace_cw<true>(sc, subframe, mk3(fo.x, fo.y, fo.z), mk3(fd.x, fd.y, fd.z), fo.w, PT_MAX_RAY_DIST, h, census);
            const uint32_t vis = h.t < 0.0f ? 1u : 0u;
            if(vis != wb.visible[slot])
            {
                if(atomicAdd(mismatch, 1ull) < 8ull) printf("validate: shadow slot %u vis %u vs %u\n", slot, wb.visible[slot], vis);
            }
        }
        else
        {
            trace_cw<false>(sc, subframe, mk3(fo.x, fo.y, fo.z), mk3(fd.x, fd.y, fd.z), fo.w, PT_MAX_RAY_DIST, h, census);
            const float4 fh = wb.hit[slot];
            const uint32_t hp = wb.hit_prim[slot];
            if(fh.x != h.t || (h.t >= 0.0f && (__float_as_uint(fh.w) != h.inst || (hp & 0x7FFFFFFFu) != h.prim)))
            {
                if(atomicAdd(mismatch, 1ull) < 8ull)
                    printf("validate: slot %u sub %u o %.7g %.7g %.7g d %.7g %.7g %.7g tmin %g: kernel t %.9g inst %u prim %u | plain t %.9g inst %u prim %u\n",
                           slot, subframe, fo.x, fo.y, fo.z, fd.x, fd.y, fd.z, fo.w, fh.x, __float_as_uint(fh.w), hp & 0x7FFFFFFFu, h.t, h.inst, h.prim);
            }
        }
    }
}

// ---- classify: slots whose closest-hit query finished -> the FAR / NEAR shade queues ------------------
// wf_trace only writes a status byte per finished query; this scan appends the slots to the two
// shade queues in slot order (= pixel order, the samples of a pixel adjacent), so the shade kernels
// read the path state mostly coalesced and neighbouring lanes shade the same surface.
// One chunk of 4096 slots per block iteration, 16 status bytes per thread (one 16-byte load), one
// block-wide scan and ONE atomic per queue per chunk (the per-warp version issued 3.7 M atomics per
// round on two addresses and ran at 45 G slots/s). Order inside a chunk is slot order.
constexpr uint32_t WF_CLASSIFY_PER_THREAD = 16;
__global__ void __launch_bounds__(256)
wf_classify_kernel(WaveBuffers wb)
{
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_base[2];
    const uint32_t per_block = 256u * WF_CLASSIFY_PER_THREAD;
    const uint32_t n_chunks = (wb.n_slots + per_block - 1u) / per_block;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for(uint32_t c = blockIdx.x; c < n_chunks; c += gridDim.x)
    {
        const uint32_t first = c * per_block + threadIdx.x * WF_CLASSIFY_PER_THREAD;
        uint4 w = make_uint4(0u, 0u, 0u, 0u);
        if(first < wb.n_slots) w = *reinterpret_cast<const uint4*>(wb.status + first); // n_slots is a multiple of 64
        if((w.x | w.y | w.z | w.w) != 0u) *reinterpret_cast<uint4*>(wb.status + first) = make_uint4(0u, 0u, 0u, 0u);
        const uint32_t words[4] = {w.x, w.y, w.z, w.w};
        uint32_t far16 = 0u, near16 = 0u;
        #pragma unroll
        for(int k = 0; k < 4; ++k)
        {
            // status bytes are 0, 2 (NEAR) or 3 (FAR): bit 1 = finished, bit 0 = far
            const uint32_t done = (words[k] >> 1) & 0x01010101u, farb = words[k] & done & 0x01010101u;
            const uint32_t nearb = done & ~farb;
            // gather bit 0 of each byte into 4 adjacent bits
            far16 |= (((farb * 0x01020408u) >> 24) & 0xFu) << (4 * k);
            near16 |= (((nearb * 0x01020408u) >> 24) & 0xFu) << (4 * k);
        }
        // block-wide exclusive scan of (far count | near count << 16); a chunk holds at most 4096 of either
        const uint32_t mine = (uint32_t)__popc(far16) | ((uint32_t)__popc(near16) << 16);
        uint32_t incl = mine;
        #pragma unroll
        for(int o = 1; o < 32; o <<= 1)
        {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if(lane >= (unsigned)o) incl += v;
        }
        if(lane == 31u) s_warp[warp] = incl;
        __syncthreads();
        uint32_t before = 0u, total = 0u;
        #pragma unroll
        for(int k = 0; k < 8; ++k) { const uint32_t v = s_warp[k]; if(k < (int)warp) before += v; total += v; }
        if(threadIdx.x == 0)
        {
            s_base[0] = (total & 0xFFFFu) ? atomicAdd(&wb.cnt->n_far, total & 0xFFFFu) : 0u;
            s_base[1] = (total >> 16) ? atomicAdd(&wb.cnt->n_near, total >> 16) : 0u;
        }
        __syncthreads();
        const uint32_t excl = before + incl - mine;
        uint32_t pf = s_base[0] + (excl & 0xFFFFu), pn = s_base[1] + (excl >> 16);
        for(uint32_t m = far16; m; m &= m - 1u) wb.q_far[pf++] = first + (uint32_t)__ffs(m) - 1u;
        for(uint32_t m = near16; m; m &= m - 1u) wb.q_near[pn++] = first + (uint32_t)__ffs(m) - 1u;
        __syncthreads();
    }
}

// ---- shade: a closest-hit query finished (trace_ray tail + bounce loop body) -----------------------
template<bool FAR>
__global__ void __launch_bounds__(128)
wf_shade_kernel(Scene sc, RenderJob job, WaveBuffers wb)
{
    const uint32_t n = FAR ? wb.cnt->n_far : wb.cnt->n_near;
    const uint32_t* q = FAR ? wb.q_far : wb.q_near;
    __shared__ uint32_t s_tmp[2 * 5];
    const uint32_t rounded = (n + 127u) & ~127u;   // whole blocks (block-aggregated appends below)
    const uint32_t step = gridDim.x * blockDim.x;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t slot_ahead = i < n ? q[i] : WF_INVALID;
    for(; i < rounded; i += step)
    {
        const uint32_t slot = slot_ahead;
        // The slot of the NEXT iteration is fetched now and its path state pulled into L2 half-way through this
        // one: the kernel is bound by the latency of its dependent loads (queue -> state -> hit -> shading
        // record: issue slots 37 % busy, long_scoreboard the top stall), not by bandwidth (2.2 of 6.5 TB/s).
        slot_ahead = (i + step < n) ? q[i + step] : WF_INVALID;
        bool push_ext = false, push_shadow = false, push_new = false;
        uint32_t key_ext = 0u, key_shadow = 0u;
        if(slot != WF_INVALID)
        {
            int2 cursor = wb.cursor[slot];
            const int sample = job.s_begin + cursor.x * job.s_stride;
            const uint32_t subframe = sample < 0 ? 0u : (uint32_t)sample / (uint32_t)sc.samples_per_subframe;
            const RefSubframe* rsf = sc.subframes + subframe;
            Light light;
            light.dir = mk3(__ldg(&rsf->light_dir));
            light.color = mk3(__ldg(&rsf->light_color));
            light.cos_solid_angle = __ldg(&rsf->cos_solid_angle);

            const float4 fo = wb.ray_o[slot], fd = wb.ray_d[slot];
            v3 ray_o = mk3(fo.x, fo.y, fo.z), ray_d = mk3(fd.x, fd.y, fd.z);
            float bsdf_pdf = fd.w;
            int bounce = cursor.y;
            // a camera ray (bounce 0) carries the initial path state implicitly (wf_generate does not write it)
            float4 fa = make_float4(1, 1, 1, 1), fc = make_float4(0, 0, 0, 0), fn = make_float4(0, 0, 0, 0);
            if(bounce > 0) { fa = wb.atten[slot]; fc = wb.contrib[slot]; fn = wb.nee[slot]; }
            v3 attenuation = mk3(fa.x, fa.y, fa.z), contribution = mk3(fc.x, fc.y, fc.z);
            float regularization = fa.w;
            const uint4 fs = wb.rng[slot];
            rng4 seed = {fs.x, fs.y, fs.z, fs.w};

            // NEE of the previous bounce (nee_branch tail, path_tracer.hh:611-619)
            if(fn.w != 0.0f && wb.visible[slot] != 0u)
            {
                const float4 sd = wb.shadow_d[slot];
                contribution += mk3(fn.x, fn.y, fn.z) * sky_attenuation(sd.w, ray_o, mk3(sd.x, sd.y, sd.z));
            }

            const float4 fh = wb.hit[slot];
            const uint32_t hp = wb.hit_prim[slot];
            Hit hit; hit.t = fh.x; hit.u = fh.y; hit.v = fh.z; hit.inst = __float_as_uint(fh.w);
            hit.prim = hp & 0x7FFFFFFFu; hit.back_face = (hp & 0x80000000u) != 0u;
            HitInfo info;
            shade_hit(sc, light, hit, ray_o, ray_d, info);
#if WF_SHADE_PREFETCH
            if(slot_ahead != WF_INVALID)
            {
                prefetch_l2(wb.cursor + slot_ahead); prefetch_l2(wb.ray_o + slot_ahead); prefetch_l2(wb.ray_d + slot_ahead);
                prefetch_l2(wb.rng + slot_ahead); prefetch_l2(wb.hit + slot_ahead); prefetch_l2(wb.hit_prim + slot_ahead);
                if(bounce > 0)
                {   // (camera rays do not read these; the next slot is almost always at this slot's bounce)
                    prefetch_l2(wb.atten + slot_ahead); prefetch_l2(wb.contrib + slot_ahead); prefetch_l2(wb.nee + slot_ahead);
                    prefetch_l2(wb.visible + slot_ahead); prefetch_l2(wb.shadow_d + slot_ahead);
                }
            }
#endif

            const float mis_pdf = bsdf_pdf < 0.0f ? -bsdf_pdf :
                (info.nee_pdf * info.nee_pdf + bsdf_pdf * bsdf_pdf) / bsdf_pdf;
            v3 atmo_att = mk3(1, 1, 1), scat = mk3(0, 0, 0);
            if(FAR) sky_scattering(seed, light, ray_o, ray_d, info.thit, atmo_att, scat);
            contribution += attenuation * (scat + atmo_att * info.s.albedo * info.emission) * (1.0f / mis_pdf);
            attenuation *= atmo_att * (1.0f / fabsf(bsdf_pdf));
            if(bsdf_pdf > 0.0f)
                regularization *= fmaxf(1.0f - PT_REG_GAMMA / sqrtf(sqrtf(bsdf_pdf)), 0.0f);
            if(bounce > 0) info.s.roughness = 1.0f - (1.0f - info.s.roughness) * regularization;

            if(bounce >= sc.max_bounces || !(info.thit > 0.0f))
            {   // path complete
                float4 s = wb.sum[slot];
                s.x += contribution.x; s.y += contribution.y; s.z += contribution.z;
                wb.sum[slot] = s;
                cursor.x += (int)wb.lanes;
                push_new = cursor.x < job.s_count;
                cursor.y = push_new ? WF_NEED_SAMPLE : WF_FINISHED;
                wb.cursor[slot] = cursor;
            }
            else
            {
                bounce++;
                v3 view = mul_v3m3(-ray_d, info.tbn);
                if(view.z < 1e-7f) view.z = fmaxf(view.z, 1e-7f);
                view = normalize(view);

                float4 un = rand4(seed);
                v3 light_dir = sample_cone(light.dir, light.cos_solid_angle, un.x, un.y);
                const float nee_pdf = 1.0f / (PT_TWO_PI * (1.0f - light.cos_solid_angle));
                float eval_pdf = 0.0f;
                v3 color = bsdf_eval(mul_v3m3(light_dir, info.tbn), view, info.s, eval_pdf) * nee_pdf * light.color;
                const bool lit = !(color.x == 0.0f && color.y == 0.0f && color.z == 0.0f);
                float nee_mis = 1.0f;
                if(light.cos_solid_angle < 1.0f) nee_mis = (nee_pdf * nee_pdf + eval_pdf * eval_pdf) / nee_pdf;
                v3 nee_pending = attenuation * (color * (1.0f / nee_mis));

                float4 ub = rand4(seed);
                v3 tdir, bsdf_att;
                bsdf_sample(ub.x, ub.y, ub.z, view, info.s, tdir, bsdf_att, bsdf_pdf);
                v3 bounce_d = normalize(mul_m3v3(info.tbn, tdir));
                attenuation *= bsdf_att;

                wb.ray_o[slot] = make_float4(info.pos.x, info.pos.y, info.pos.z, PT_MIN_RAY_DIST);
                wb.ray_d[slot] = make_float4(bounce_d.x, bounce_d.y, bounce_d.z, bsdf_pdf);
                wb.shadow_d[slot] = make_float4(light_dir.x, light_dir.y, light_dir.z, un.w);
                wb.nee[slot] = make_float4(nee_pending.x, nee_pending.y, nee_pending.z, lit ? 1.0f : 0.0f);
                wb.atten[slot] = make_float4(attenuation.x, attenuation.y, attenuation.z, regularization);
                wb.contrib[slot] = make_float4(contribution.x, contribution.y, contribution.z, 0.0f);
                wb.rng[slot] = make_uint4(seed.x, seed.y, seed.z, seed.w);
                cursor.y = bounce;
                wb.cursor[slot] = cursor;
                push_ext = true;
                push_shadow = lit;
                if(wb.sort)
                {
                    key_ext = sort_key(sc, info.pos, bounce_d, false);
                    key_shadow = sort_key(sc, info.pos, light_dir, true);
                }
            }
        }
        uint32_t* keys = wb.sort ? wb.q_key : nullptr;
        wf_append2_block<4>(wb.q_trace + WF_SEG_SHADOW * (size_t)wb.seg_cap, &wb.cnt->n_seg[WF_SEG_SHADOW], push_shadow, slot | WF_SHADOW_BIT,
                            keys ? keys + WF_SEG_SHADOW * (size_t)wb.seg_cap : nullptr, key_shadow,
                            wb.q_trace + WF_SEG_BOUNCE * (size_t)wb.seg_cap, &wb.cnt->n_seg[WF_SEG_BOUNCE], push_ext, slot,
                            keys ? keys + WF_SEG_BOUNCE * (size_t)wb.seg_cap : nullptr, key_ext, s_tmp);
        // n_new is a flag (wf_generate only asks whether it is non-zero): a plain store, not an atomic per warp
        if(__any_sync(0xFFFFFFFFu, push_new) && (threadIdx.x & 31u) == 0u) wb.cnt->n_new = 1u;
    }
}

// ---- bookkeeping between phases ----------------------------------------------------------------------
// phase 1 (after trace): the ray queue is consumed -> reset it, the fetch cursor and the new-sample count (shade recounts)
// phase 2 (after shade): result queues consumed -> reset; report whether any ray is left
__global__ void wf_phase_kernel(WaveBuffers wb, int phase, uint32_t* host_visible_remaining)
{
    if(threadIdx.x != 0 || blockIdx.x != 0) return;
    WaveCounters* c = wb.cnt;
    if(phase == 1) { c->n_seg[0] = 0; c->n_seg[1] = 0; c->n_seg[2] = 0; c->cur_trace = 0; c->n_new = 0; }
    else if(phase == 2) { c->n_far = 0; c->n_near = 0; if(host_visible_remaining) *host_visible_remaining = c->n_seg[0] + c->n_seg[1] + c->n_seg[2] + c->n_new; }
}

// ---- finalize: fixed-order sum of the slots of each pixel, mean, tonemap, pack -------------------------
// Lane sums are added in ascending lane order. With one sample per slot (lanes == samples per pixel,
// the default when the pool fits) that is exactly the reference's order, j = 0 .. SPP-1 (main.cc:24-39).
__global__ void wf_finalize_kernel(RenderJob job, WaveBuffers wb)
{
    const uint32_t pi = blockIdx.x * blockDim.x + threadIdx.x;
    if(((size_t)pi << wb.lane_shift) >= wb.n_slots) return;
    int lx, ly;
    if(!slot_pixel(wb, job, pi << wb.lane_shift, lx, ly)) return;
    const float4* p = wb.sum + ((size_t)pi << wb.lane_shift);
    v3 s = mk3(0, 0, 0);
    for(uint32_t l = 0; l < wb.lanes; ++l) { const float4 f = p[l]; s += mk3(f.x, f.y, f.z); }
    store_pixel(job, lx, ly, s);
}

} // namespace pt
