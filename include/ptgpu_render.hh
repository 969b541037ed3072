// ptgpu_render.hh — C++17 shim with baseline_render's signature (reference main.cc:12), so that
// main.cc:88 becomes a one-token swap:
//
//     #include "ptgpu_render.hh"
//     ...
//     gpu_render(s, image.get());          // was: baseline_render(s, image.get());
//
// Header-only and templated on the reference's own `scene` / `uchar4` types, so it needs no
// reference header at this repo's build time. Link with -lptgpu. Errors follow the reference's
// convention: message on stderr, exit(1) (mesh.cc:25-41, bmp.cc:54-59).
#ifndef PTGPU_RENDER_HH
#define PTGPU_RENDER_HH

#include "ptgpu.h"

#include <cstdio>
#include <cstdlib>

namespace ptgpu_detail {

inline void die(ptgpu_ctx* ctx, const char* what)
{
    std::fprintf(stderr, "%s: %s\n", what, ptgpu_last_error(ctx));
    std::exit(1);
}

// The CPU half of the static-scene upload (BVH flattening, flat static scene), done once for all workers of the
// process: call after load_scene(); every renderer<Scene>::set_frame then finds the result cached.
template<class Scene>
inline void prepare_static(const Scene& s)
{
    const size_t n_static_nodes = s.bvh_buf.nodes.size();   // before any setup_animation_frame: BLAS nodes only
    char err[512];
    if(ptgpu_host_prepare_static(
           reinterpret_cast<const ptgpu_bvh_node*>(s.bvh_buf.nodes.data()), n_static_nodes,
           reinterpret_cast<const ptgpu_bvh_link*>(s.bvh_buf.links.data()), 8 * n_static_nodes,
           s.mesh_buf.indices.data(), s.mesh_buf.indices.size(),
           reinterpret_cast<const ptgpu_float3*>(s.mesh_buf.pos.data()), s.mesh_buf.pos.size(),
           reinterpret_cast<const ptgpu_tlas_instance*>(s.instances.data()), s.static_instance_count, 1, err, sizeof(err)) != 0)
    {
        std::fprintf(stderr, "ptgpu_host_prepare_static: %s\n", err);
        std::exit(1);
    }
}

// One GPU worker bound to one `scene` object (scene.hh:40-65). The static part is uploaded on the
// first frame; it must not change afterwards (load_scene() runs once, main.cc:67).
template<class Scene>
class renderer
{
public:
    renderer(int device, const ptgpu_config& cfg)
    {
        if(ptgpu_create(&ctx_, device, &cfg) != 0) die(nullptr, "ptgpu_create");
    }
    ~renderer() { ptgpu_destroy(ctx_); }
    renderer(const renderer&) = delete;
    renderer& operator=(const renderer&) = delete;

    void set_frame(const Scene& s)
    {
        const auto* nodes = reinterpret_cast<const ptgpu_bvh_node*>(s.bvh_buf.nodes.data());
        const auto* links = reinterpret_cast<const ptgpu_bvh_link*>(s.bvh_buf.links.data());
        const auto* inst = reinterpret_cast<const ptgpu_tlas_instance*>(s.instances.data());
        // everything before the first per-frame TLAS is static (scene.cc:712-717)
        const size_t n_static_nodes = s.subframes.empty() ? s.bvh_buf.nodes.size() : s.subframes[0].tlas.node_offset;
        if(!uploaded_)
        {
            if(ptgpu_upload_static(
                   ctx_, nodes, n_static_nodes, links, 8 * n_static_nodes,
                   s.mesh_buf.indices.data(), s.mesh_buf.indices.size(),
                   reinterpret_cast<const ptgpu_float3*>(s.mesh_buf.pos.data()),
                   reinterpret_cast<const ptgpu_float3*>(s.mesh_buf.normal.data()),
                   reinterpret_cast<const ptgpu_float4*>(s.mesh_buf.albedo.data()),
                   reinterpret_cast<const ptgpu_float4*>(s.mesh_buf.material.data()), s.mesh_buf.pos.size(),
                   inst, s.static_instance_count) != 0)
                die(ctx_, "ptgpu_upload_static");
            uploaded_ = true;
        }
        if(ptgpu_set_frame(
               ctx_, reinterpret_cast<const ptgpu_subframe*>(s.subframes.data()), s.subframes.size(),
               inst + s.static_instance_count, s.instances.size() - s.static_instance_count,
               nodes + n_static_nodes, links + 8 * n_static_nodes,
               s.bvh_buf.nodes.size() - n_static_nodes, n_static_nodes) != 0)
            die(ctx_, "ptgpu_set_frame");
    }

    // baseline_render(s, image): BGRA, row 0 = top (main.cc:12-46)
    template<class Pixel>
    void render(const Scene& s, Pixel* image)
    {
        static_assert(sizeof(Pixel) == 4, "image must be uchar4 (math.hh:17)");
        set_frame(s);
        if(ptgpu_render(ctx_, reinterpret_cast<uint8_t*>(image)) != 0) die(ctx_, "ptgpu_render");
    }

    // baseline_render + write_bmp's packing in one go: `bmp` receives ptgpu_bmp_size() file bytes
    void render_bmp(const Scene& s, uint8_t* bmp)
    {
        set_frame(s);
        if(ptgpu_render_bmp(ctx_, bmp) != 0) die(ctx_, "ptgpu_render_bmp");
    }

    size_t bmp_size() const { return ptgpu_bmp_size(ctx_); }
    ptgpu_ctx* ctx() { return ctx_; }

private:
    ptgpu_ctx* ctx_ = nullptr;
    bool uploaded_ = false;
};

} // namespace ptgpu_detail

// Drop-in for baseline_render(const scene&, uchar4*): device 0, config.hh constants taken from the
// macros of the translation unit that includes this header after config.hh.
template<class Scene, class Pixel>
inline void gpu_render(const Scene& s, Pixel* image)
{
    static ptgpu_detail::renderer<Scene>* r = nullptr;
    if(!r)
    {
        ptgpu_config cfg;
        ptgpu_default_config(&cfg);
#if defined(IMAGE_WIDTH) && defined(IMAGE_HEIGHT) && defined(SAMPLES_PER_PIXEL) && defined(MAX_BOUNCES)
        cfg.width = IMAGE_WIDTH; cfg.height = IMAGE_HEIGHT; cfg.spp = SAMPLES_PER_PIXEL; cfg.max_bounces = MAX_BOUNCES;
        cfg.student_id = STUDENT_ID; cfg.samples_per_subframe = SAMPLES_PER_MOTION_BLUR_STEP;
#endif
        r = new ptgpu_detail::renderer<Scene>(0, cfg);
    }
    r->render(s, image);
}

#endif
