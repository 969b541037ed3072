// Persistent-thread megakernel: baseline_render's loop nest (main.cc:16-43) + path_trace_pixel
// (path_tracer.hh:637-741) as one per-thread state machine with ONE copy of every phase.
//
// Why a state machine: the straight-line tile kernel inlines the traversal three times and the sky
// march twice; ncu showed 7.9 of 32 threads active per instruction and `no_instruction`
// (instruction-cache) as the top stall. Here every lane is in one of four states and all lanes
// that are traversing — primary, bounce or shadow ray alike — execute the same loop.
//
// Work distribution: a warp owns a 2x2 pixel quad: lane = pixel*8 + sample_lane, and sample lane l
// traces samples k = l, l+8, ... of the job's sample set one after another (path regeneration:
// a lane starts its next sample as soon as its path ends, it never waits for the warp). With
// the default sample set all lanes of a warp start in the same motion-blur subframe. Each lane sums
// its own samples in ascending order and the 8 lanes of a pixel are reduced in a fixed shuffle
// order, so the result is deterministic. Warps fetch quads from a global counter (persistent
// threads: grid = SMs x resident warps), quads are ordered in 16x16-pixel tiles for L1/L2 locality.
#pragma once
#include "pt_kernels.cuh"
#include "pt_wide.cuh"

namespace pt {

struct MegaState { unsigned int next_quad; unsigned int pad[3]; };

constexpr int MEGA_THREADS = 128;
constexpr int MEGA_BLOCKS_PER_SM = 4;
constexpr int QUAD_TILE = 8; // quads per tile side (16 px)

enum : int { ST_NEW = 0, ST_TRAVERSE = 1, ST_SHADE = 2, ST_SHADOW_DONE = 3, ST_IDLE = 4 };

__global__ void __launch_bounds__(MEGA_THREADS, MEGA_BLOCKS_PER_SM)
mega_kernel(Scene sc, RenderJob job, MegaState* ms)
{
    const unsigned lane = threadIdx.x & 31u;
    const int sample_lane = lane & (SAMPLE_LANES - 1);
    const int qp = lane >> 3;                     // pixel within the quad
    const int quads_x = (job.w + 1) >> 1, quads_y = (job.h + 1) >> 1;
    const int tiles_x = (quads_x + QUAD_TILE - 1) / QUAD_TILE, tiles_y = (quads_y + QUAD_TILE - 1) / QUAD_TILE;
    const unsigned total_slots = (unsigned)(tiles_x * tiles_y * QUAD_TILE * QUAD_TILE);

    // ---- per-lane path state ----------------------------------------------------------------
    int state = ST_IDLE;
    int lx = 0, ly = 0;                 // pixel (job-local)
    bool pixel_valid = false;
    int k = 0;                          // index into the job's sample set
    v3 sum = mk3(0, 0, 0);
    rng4 seed = {0, 0, 0, 0};
    uint32_t subframe = 0;
    v3 ray_o = mk3(0, 0, 0), ray_d = mk3(0, 0, 1);
    v3 attenuation = mk3(1, 1, 1), contribution = mk3(0, 0, 0);
    v3 bounce_d = mk3(0, 0, 1), nee_pending = mk3(0, 0, 0);
    float bsdf_pdf = -1.0f, nee_jitter = 0.0f, regularization = 1.0f;
    int bounce = 0;
    bool shadow_ray = false;

    // ---- per-lane traversal state -------------------------------------------------------------
    uint32_t stack[WIDE_STACK];
    int sp = 0;
    uint32_t cur = PT_EMPTY;
    v3 o = mk3(0, 0, 0), d = mk3(0, 0, 1), inv = mk3(0, 0, 0), S = mk3(0, 0, 1);
    int axis = 2;
    bool in_blas = false;
    const WideNode* nodes = sc.wtlas;
    const float4* tris = sc.wtris;
    uint32_t cur_inst = 0;
    float tmin = 0.0f, tmax = 0.0f;
    Hit hit; hit.t = -1.0f; hit.u = hit.v = 0.0f; hit.inst = 0; hit.prim = 0; hit.back_face = false;

    bool warp_has_quad = false;

    for(;;)
    {
        // ---- fetch a quad when every lane of the warp is idle ------------------------------------
        if(__all_sync(0xFFFFFFFFu, state == ST_IDLE))
        {
            if(warp_has_quad)
            {   // epilogue of the finished quad: fixed-order reduction, mean, tonemap, pack
                v3 r = reduce_lanes(sum);
                if(pixel_valid && sample_lane == 0) store_pixel(job, lx, ly, r);
                warp_has_quad = false;
            }
            unsigned slot = 0;
            if(lane == 0) slot = atomicAdd(&ms->next_quad, 1u);
            slot = __shfl_sync(0xFFFFFFFFu, slot, 0);
            if(slot >= total_slots) break;
            const int tile = slot / (QUAD_TILE * QUAD_TILE), in_tile = slot % (QUAD_TILE * QUAD_TILE);
            const int qx = (tile % tiles_x) * QUAD_TILE + (in_tile % QUAD_TILE);
            const int qy = (tile / tiles_x) * QUAD_TILE + (in_tile / QUAD_TILE);
            lx = qx * 2 + (qp & 1);
            ly = qy * 2 + (qp >> 1);
            pixel_valid = lx < job.w && ly < job.h;
            warp_has_quad = true;
            sum = mk3(0, 0, 0);
            k = sample_lane;
            state = (pixel_valid && k < job.s_count) ? ST_NEW : ST_IDLE;
            if(!__any_sync(0xFFFFFFFFu, state != ST_IDLE)) continue; // quad entirely outside the image
        }

        // ---- ST_SHADOW_DONE: resolve the NEE sample (nee_branch tail, path_tracer.hh:611-619) ------
        if(state == ST_SHADOW_DONE)
        {
            if(hit.t < 0.0f) // unoccluded
                contribution += nee_pending * sky_attenuation(nee_jitter, ray_o, ray_d);
            ray_d = bounce_d;
            shadow_ray = false;
            tmin = PT_MIN_RAY_DIST; tmax = PT_MAX_RAY_DIST;
            state = ST_TRAVERSE; cur = PT_EMPTY - 2u; // (re)start marker, see below
        }

        // ---- ST_SHADE: a closest-hit query finished (trace_ray tail + the bounce loop body) ---------
        if(state == ST_SHADE)
        {
            const RefSubframe* rsf = sc.subframes + subframe;
            Light light;
            light.dir = mk3(__ldg(&rsf->light_dir));
            light.color = mk3(__ldg(&rsf->light_color));
            light.cos_solid_angle = __ldg(&rsf->cos_solid_angle);
            HitInfo info;
            shade_hit(sc, light, hit, ray_o, ray_d, info);

            // path_tracer.hh:722-737 (and :691-693 for the primary ray, where bsdf_pdf = -1 and
            // attenuation = 1 make the same expressions reduce to the primary-ray ones exactly)
            const float mis_pdf = bsdf_pdf < 0.0f ? -bsdf_pdf :
                (info.nee_pdf * info.nee_pdf + bsdf_pdf * bsdf_pdf) / bsdf_pdf;
            v3 atmo_att, scat;
            sky_scattering(seed, light, ray_o, ray_d, info.thit, atmo_att, scat);
            contribution += attenuation * (scat + atmo_att * info.s.albedo * info.emission) * (1.0f / mis_pdf);
            attenuation *= atmo_att * (1.0f / fabsf(bsdf_pdf));
            if(bsdf_pdf > 0.0f)
                regularization *= fmaxf(1.0f - PT_REG_GAMMA / sqrtf(sqrtf(bsdf_pdf)), 0.0f);
            if(bounce > 0) info.s.roughness = 1.0f - (1.0f - info.s.roughness) * regularization;

            if(bounce >= sc.max_bounces || !(info.thit > 0.0f))
            {   // path complete
                sum += contribution;
                k += SAMPLE_LANES;
                state = k < job.s_count ? ST_NEW : ST_IDLE;
            }
            else
            {   // next bounce: NEE sample + BSDF sample (path_tracer.hh:699-719, 594-609)
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
                nee_pending = attenuation * (color * (1.0f / nee_mis));
                nee_jitter = un.w;

                float4 ub = rand4(seed);
                v3 tdir, bsdf_att;
                bsdf_sample(ub.x, ub.y, ub.z, view, info.s, tdir, bsdf_att, bsdf_pdf);
                bounce_d = normalize(mul_m3v3(info.tbn, tdir));
                attenuation *= bsdf_att;

                ray_o = info.pos;
                ray_d = lit ? light_dir : bounce_d;
                shadow_ray = lit;
                tmin = PT_MIN_RAY_DIST; tmax = PT_MAX_RAY_DIST;
                state = ST_TRAVERSE; cur = PT_EMPTY - 2u;
            }
        }

        // ---- ST_NEW: start the next sample (path_tracer.hh:655-671) ---------------------------------
        if(state == ST_NEW)
        {
            const int sample = job.s_begin + k * job.s_stride;
            const uint32_t px = (uint32_t)(job.x0 + lx), py = (uint32_t)(job.y0 + ly);
            subframe = sample < 0 ? 0u : (uint32_t)sample / (uint32_t)sc.samples_per_subframe;
            seed.x = px; seed.y = py; seed.z = (uint32_t)sample; seed.w = sc.student_id;
            pcg4d(seed);
            float4 u = rand4(seed);
            v2 film = sample_gaussian_disk(u.x, u.y, 0.4f);
            camera_ray(sc, sc.subframes + subframe, u.z, u.w, (float)px + (film.x + 0.5f), (float)py + (film.y + 0.5f), ray_d, ray_o);
            attenuation = mk3(1, 1, 1); contribution = mk3(0, 0, 0);
            regularization = 1.0f; bounce = 0; bsdf_pdf = -1.0f; // -1: MIS weight 1, |pdf| 1, no regularisation
            shadow_ray = false;
            tmin = 0.0f; tmax = PT_MAX_RAY_DIST;
            state = ST_TRAVERSE; cur = PT_EMPTY - 2u;
        }

        // ---- (re)start of a query: push the subframe's dynamic instances and the TLAS root ----------
        if(state == ST_TRAVERSE && cur == PT_EMPTY - 2u)
        {
            hit.t = -1.0f; hit.u = 0.0f; hit.v = 0.0f; hit.inst = 0xFFFFFFFFu; hit.prim = 0; hit.back_face = false;
            sp = 0;
            o = ray_o; d = ray_d; inv = safe_inv_dir(ray_d);
            in_blas = false; nodes = sc.wtlas;
            const uint2 r = __ldg(sc.dyn_range + subframe);
            const uint32_t p = r.x, a = r.y & 0xFFFFFu, len = r.y >> 20;
            for(uint32_t i = 0; i < p + len; ++i)
            {
                const uint32_t id = sc.n_static + (i < p ? i : a + (i - p));
                const WideInstance* wi = sc.winst + id;
                if(box_hit(__ldg(&wi->lo), __ldg(&wi->hi), o, inv, tmin, tmax))
                    stack[sp++] = 0x80000000u | id;
            }
            cur = 0;
        }

        // ---- traversal: all traversing lanes, whatever their ray kind, run this one loop ------------
        // The loop is warp-uniform: it ends when no lane is traversing, or when fewer than
        // job.min_active lanes still are while others wait with shading work (those then shade and
        // come back with fresh rays instead of idling until the slowest query ends).
        for(;;)
        {
            const bool trav = state == ST_TRAVERSE;
            const unsigned tm = __ballot_sync(0xFFFFFFFFu, trav);
            if(tm == 0u) break;
            if(__popc(tm) < job.min_active)
            {
                const unsigned waiting = __ballot_sync(0xFFFFFFFFu, state == ST_SHADE || state == ST_SHADOW_DONE);
                if(waiting != 0u) break;
            }
            if(!trav) continue;
            if(cur == PT_EMPTY)
            {
                if(sp == 0)
                {   // query complete
                    state = shadow_ray ? ST_SHADOW_DONE : ST_SHADE;
                    continue;
                }
                cur = stack[--sp];
            }
            if(cur == PT_EXIT_MARK)
            {
                in_blas = false; nodes = sc.wtlas; o = ray_o; d = ray_d; inv = safe_inv_dir(ray_d);
                cur = PT_EMPTY;
            }
            else if(!(cur & 0x80000000u))
            {
                uint32_t key[4]; uint4 child;
                test4(nodes + cur, o, inv, tmin, tmax, key, child);
                cswap(key[0], key[1]); cswap(key[2], key[3]); cswap(key[0], key[2]); cswap(key[1], key[3]); cswap(key[1], key[2]);
                if(key[3] != PT_EMPTY) stack[sp++] = pick_child(child, key[3] & 3u);
                if(key[2] != PT_EMPTY) stack[sp++] = pick_child(child, key[2] & 3u);
                if(key[1] != PT_EMPTY) stack[sp++] = pick_child(child, key[1] & 3u);
                cur = key[0] != PT_EMPTY ? pick_child(child, key[0] & 3u) : PT_EMPTY;
            }
            else if(!in_blas)
            {
                cur_inst = cur & 0x7FFFFFFFu;
                const WideInstance* wi = sc.winst + cur_inst;
                const float4 r0 = __ldg(&wi->inv0), r1 = __ldg(&wi->inv1), r2 = __ldg(&wi->inv2);
                const uint32_t b = __ldg(&wi->blas);
                o = mk3(r0.x * ray_o.x + r0.y * ray_o.y + r0.z * ray_o.z + r0.w,
                        r1.x * ray_o.x + r1.y * ray_o.y + r1.z * ray_o.z + r1.w,
                        r2.x * ray_o.x + r2.y * ray_o.y + r2.z * ray_o.z + r2.w);
                d = mk3(r0.x * ray_d.x + r0.y * ray_d.y + r0.z * ray_d.z,
                        r1.x * ray_d.x + r1.y * ray_d.y + r1.z * ray_d.z,
                        r2.x * ray_d.x + r2.y * ray_d.y + r2.z * ray_d.z);
                inv = safe_inv_dir(d);
                tri_preprocess(d, axis, S);
                const uint2 bo = __ldg(reinterpret_cast<const uint2*>(sc.wblas + b));
                nodes = sc.wnodes + bo.x;
                tris = sc.wtris + 3 * (size_t)bo.y;
                in_blas = true;
                stack[sp++] = PT_EXIT_MARK;
                cur = 0;
            }
            else
            {
                const uint32_t first = cur & 0x07FFFFFFu, count = ((cur >> 27) & 0xFu) + 1u;
                const float4* tp = tris + 3 * (size_t)first;
                cur = PT_EMPTY;
                for(uint32_t i = 0; i < count; ++i, tp += 3)
                {
                    const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
                    float u, v, t; bool bf;
                    bool ok = tri_intersect(o, axis, S, mk3(a), mk3(b), mk3(c), u, v, t, bf);
                    if(ok && t < tmax && t > tmin)
                    {
                        hit.t = t; hit.u = u; hit.v = v; hit.inst = cur_inst; hit.prim = __float_as_uint(a.w); hit.back_face = bf;
                        tmax = t;
                        if(shadow_ray) { sp = 0; break; } // any hit ends a shadow query
                    }
                }
            }
        }
    }
}

inline int launch_mega(const Scene& sc, const RenderJob& job, MegaState* ms, int sm_count, cudaStream_t stream)
{
    cudaMemsetAsync(ms, 0, sizeof(MegaState), stream);
    const int quads = ((job.w + 1) / 2) * ((job.h + 1) / 2);
    const int warps_per_block = MEGA_THREADS / 32;
    int blocks = sm_count * MEGA_BLOCKS_PER_SM;
    const int needed = (quads + warps_per_block - 1) / warps_per_block;
    if(blocks > needed) blocks = needed;
    if(blocks < 1) blocks = 1;
    mega_kernel<<<blocks, MEGA_THREADS, 0, stream>>>(sc, job, ms);
    return 1;
}

} // namespace pt
