// Render kernels: tile kernel (one thread per pixel x sample-lane) and utility kernels.
#pragma once
#include "pt_path.cuh"

namespace pt {

// What to render: the loop nest of baseline_render (main.cc:16-43) over a pixel rectangle and the
// sample set {s_begin + k*s_stride}. Outputs are optional.
struct RenderJob
{
    int32_t x0, y0, w, h;
    int32_t s_begin, s_count, s_stride;
    float* out_rgb;        // w*h*3 mean linear radiance, or null
    uchar4* out_bgra;      // w*h BGRA (tonemap_pixel), row 0 = top, or null
    uint8_t* out_bmp;      // full BMP file image (bmp.cc:15-52) for a full-frame job, or null
    uint32_t bmp_pitch;
    int32_t min_active;    // megakernel: leave the traversal loop when fewer lanes than this still traverse;
                           // wavefront: idle lanes needed before a warp refills from the ray queue
    int32_t tri_threshold, xform_threshold; // wavefront/compressed: lanes needed to elect the triangle / space-change block
    int32_t node_threshold, node_burst;     // wavefront/compressed: lanes needed to continue a node burst, and its length
};

constexpr int TILE_W = 8, TILE_H = 4, SAMPLE_LANES = 8;
constexpr int TILE_THREADS = TILE_W * TILE_H * SAMPLE_LANES; // 256

// Sum over the 8 sample lanes of one pixel in a fixed order (deterministic), lanes are the low 3
// bits of the thread index so the partners sit in one warp.
PT_D v3 reduce_lanes(v3 c)
{
    #pragma unroll
    for(int m = 1; m < SAMPLE_LANES; m <<= 1)
    {
        c.x += __shfl_xor_sync(0xFFFFFFFFu, c.x, m);
        c.y += __shfl_xor_sync(0xFFFFFFFFu, c.y, m);
        c.z += __shfl_xor_sync(0xFFFFFFFFu, c.z, m);
    }
    return c;
}

// Epilogue fused after accumulation: mean, tonemap_pixel (path_tracer.hh:753), BGRA store
// (main.cc:43) and BMP packing (bmp.cc:49-52: B,G,R of row h-1-y at pitch (3w+3)/4*4).
PT_D void store_pixel(const RenderJob& job, int lx, int ly, v3 sum)
{
    // colors[i] /= SAMPLES_PER_PIXEL (main.cc:42)
    v3 mean = mk3(sum.x / (float)job.s_count, sum.y / (float)job.s_count, sum.z / (float)job.s_count);
    const size_t i = (size_t)ly * job.w + lx;
    if(job.out_rgb)
    {
        job.out_rgb[i * 3 + 0] = mean.x;
        job.out_rgb[i * 3 + 1] = mean.y;
        job.out_rgb[i * 3 + 2] = mean.z;
    }
    if(job.out_bgra || job.out_bmp)
    {
        uchar4 p = tonemap(mean);
        if(job.out_bgra) job.out_bgra[i] = p;
        if(job.out_bmp)
        {
            uint8_t* row = job.out_bmp + 54 + (size_t)(job.h - 1 - ly) * job.bmp_pitch + (size_t)lx * 3;
            row[0] = p.x; row[1] = p.y; row[2] = p.z;
        }
    }
}

template<bool COUNT>
PT_D void flush_counters(Events<COUNT>&, Counters*) {}
template<>
PT_D void flush_counters<true>(Events<true>& ev, Counters* out)
{
    // warp-reduce then one atomic per warp per counter
    uint32_t vals[11] = {ev.c.paths, ev.c.rays, ev.c.nodes, ev.c.tris, ev.c.blas, ev.c.bounces, ev.c.shadow,
                         ev.c.sky, ev.c.att, ev.c.hits, ev.c.misses};
    #pragma unroll
    for(int k = 0; k < 11; ++k)
    {
        uint32_t v = vals[k];
        v = __reduce_add_sync(0xFFFFFFFFu, v);
        if((threadIdx.x & 31) == 0) atomicAdd(&out->v[k], (unsigned long long)v);
    }
}

// Simple tile kernel: block = 8x4 pixels x 8 sample lanes. Lane l of a pixel takes samples
// k = l, l+8, ... of the job's sample set, so with the default set (0..SPP-1) all 32 threads of a
// warp are in the same motion-blur subframe (sample/8) at the same time.
template<class Trav, bool COUNT>
__global__ void __launch_bounds__(TILE_THREADS)
render_tiles_kernel(Scene sc, RenderJob job, Counters* counters)
{
    const int tiles_x = (job.w + TILE_W - 1) / TILE_W;
    const int tile = blockIdx.x;
    const int lane = threadIdx.x & (SAMPLE_LANES - 1);
    const int p = threadIdx.x / SAMPLE_LANES;
    const int lx = (tile % tiles_x) * TILE_W + (p % TILE_W);
    const int ly = (tile / tiles_x) * TILE_H + (p / TILE_W);
    const bool inside = lx < job.w && ly < job.h;

    v3 sum = mk3(0, 0, 0);
    Events<COUNT> ev;
    if(inside)
    {
        for(int k = lane; k < job.s_count; k += SAMPLE_LANES)
        {
            const int sample = job.s_begin + k * job.s_stride;
            sum += path_trace_sample<Trav, COUNT>(sc, (uint32_t)(job.x0 + lx), (uint32_t)(job.y0 + ly), sample, ev);
        }
    }
    sum = reduce_lanes(sum);
    if(inside && lane == 0) store_pixel(job, lx, ly, sum);
    if(COUNT) flush_counters<COUNT>(ev, counters);
}

// path_trace_pixel for a list of (x, y, sample) triples (ptgpu_trace_samples)
template<class Trav>
__global__ void trace_samples_kernel(Scene sc, const uint32_t* xy, const int32_t* sample, size_t n, float* out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    Events<false> ev;
    v3 c = path_trace_sample<Trav, false>(sc, xy[2 * i], xy[2 * i + 1], sample[i], ev);
    out[3 * i + 0] = c.x; out[3 * i + 1] = c.y; out[3 * i + 2] = c.z;
}

template<class Trav>
__global__ void trace_closest_kernel(Scene sc, const float* rays, size_t n, uint32_t subframe, float* out_f, uint32_t* out_u)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    SubframeCtx sf;
    const RefSubframe* rsf;
    load_subframe(sc, (int)(subframe * sc.samples_per_subframe), sf, rsf);
    const float* r = rays + 8 * i;
    Hit h; TravCounters tc = {0, 0, 0};
    Trav::template trace<false>(sc, sf, mk3(r[0], r[1], r[2]), mk3(r[4], r[5], r[6]), r[3], r[7], h, tc);
    out_f[4 * i + 0] = h.t; out_f[4 * i + 1] = h.u; out_f[4 * i + 2] = h.v; out_f[4 * i + 3] = 1.0f - h.u - h.v;
    out_u[3 * i + 0] = h.inst; out_u[3 * i + 1] = h.prim; out_u[3 * i + 2] = h.back_face ? 1u : 0u;
}

__global__ void tonemap_kernel(const float* rgb, size_t n, uchar4* out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    out[i] = tonemap(mk3(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]));
}

__global__ void pcg4d_kernel(uint32_t* states, size_t n, int steps)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    rng4 s = {states[4 * i], states[4 * i + 1], states[4 * i + 2], states[4 * i + 3]};
    for(int k = 0; k < steps; ++k) pcg4d(s);
    states[4 * i] = s.x; states[4 * i + 1] = s.y; states[4 * i + 2] = s.z; states[4 * i + 3] = s.w;
}

// 54-byte BMP header (bmp.cc:21-44), written once per output buffer by thread 0
__global__ void bmp_header_kernel(uint8_t* bmp, uint32_t w, uint32_t h, uint32_t pitch)
{
    if(threadIdx.x != 0 || blockIdx.x != 0) return;
    for(int i = 0; i < 54; ++i) bmp[i] = 0;
    auto put32 = [&](int off, uint32_t v) { for(int k = 0; k < 4; ++k) bmp[off + k] = (uint8_t)(v >> (8 * k)); };
    auto put16 = [&](int off, uint32_t v) { bmp[off] = (uint8_t)v; bmp[off + 1] = (uint8_t)(v >> 8); };
    bmp[0] = 'B'; bmp[1] = 'M';
    put32(0x02, 54 + pitch * h);
    put32(0x0A, 54);
    put32(0x0E, 40);
    put32(0x12, w);
    put32(0x16, h);
    put16(0x1A, 1);
    put16(0x1C, 24);
    put32(0x1E, 0);
    put32(0x22, pitch * h);
    put32(0x26, 2835);
    put32(0x2A, 2835);
    put32(0x2E, 0);
    put32(0x32, 0);
}

// validator.py:41-52 on the finished device frame: 2x2 block mean of the own frame (zero padded,
// skimage.transform.downscale_local_mean), truncated to 8 bits (astype), squared difference against
// the half-size reference image (RGB, row 0 = top). Integer arithmetic throughout: sum[0] is exact.
__global__ void validate_psnr_kernel(const uchar4* __restrict__ bgra, uint32_t w, uint32_t h,
                                     const uint8_t* __restrict__ ref_rgb, uint32_t hw, uint32_t hh,
                                     unsigned long long* sum)
{
    unsigned long long sse = 0;
    for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < hw * hh; i += gridDim.x * blockDim.x)
    {
        const uint32_t x = i % hw, y = i / hw;
        uint32_t r = 0, g = 0, b = 0;
        #pragma unroll
        for(int k = 0; k < 4; ++k)
        {
            const uint32_t px = 2 * x + (k & 1), py = 2 * y + (k >> 1);
            if(px < w && py < h)
            {
                const uchar4 p = bgra[(size_t)py * w + px]; // x = blue, y = green, z = red (main.cc:39-43)
                r += p.z; g += p.y; b += p.x;
            }
        }
        const int dr = (int)(r >> 2) - (int)ref_rgb[3 * (size_t)i], dg = (int)(g >> 2) - (int)ref_rgb[3 * (size_t)i + 1],
                  db = (int)(b >> 2) - (int)ref_rgb[3 * (size_t)i + 2];
        sse += (unsigned long long)(dr * dr + dg * dg + db * db);
    }
    for(int o = 16; o > 0; o >>= 1) sse += __shfl_down_sync(0xFFFFFFFFu, sse, o);
    if((threadIdx.x & 31) == 0 && sse) atomicAdd(sum, sse);
}

} // namespace pt
