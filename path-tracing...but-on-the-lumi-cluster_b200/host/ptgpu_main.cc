// ptgpu_main.cc — multi-GPU replacement of the reference's main.cc:60-108.
//
// Same outputs (output/frame_NNNN.bmp, timing prints), but baseline_render is replaced by the C ABI
// of libptgpu.so and whole animation frames are sharded over the GPUs of one box: a shared atomic
// frame counter (frame cost varies ~7x over the animation, so the queue is dynamic, not blocked), no
// inter-GPU traffic; only the finished 8-bit BMP image returns to the host.
//
// Per GPU the three stages of main.cc's loop body run as a pipeline:
//     setup thread   setup_animation_frame(frame k+1)      host, 75 ms per frame on one core
//     render thread  upload + render frame k               device, 70-400 ms
//     writer thread  fwrite(frame k-1)                     host
// so the serial host work of the reference loop (scene.cc:271 + bmp.cc:54) hides behind the render.
// `--serial` keeps the reference's order (setup, render, write, one after the other) for comparison.
//
// This file is compiled TOGETHER with the reference's own scene.cc/bvh.cc/mesh.cc (it includes
// scene.hh / config.hh from the reference tree, -I$REF) — it is the integration a maintainer of
// the reference would build, see INTEGRATION.md. setup_animation_frame() mutates its `scene`
// (scene.cc:274-277), so every worker owns its copies of the loaded scene.
#include "scene.hh"
#include "config.hh"
#include "ptgpu_render.hh"

#include <atomic>
#include <chrono>
#include <clocale>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

static double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Bounded hand-over between two pipeline stages: indices into a ring of buffers owned by the worker.
class slot_queue
{
public:
    void push(int v) { { std::lock_guard<std::mutex> l(m_); q_.push_back(v); } cv_.notify_one(); }
    int pop() // blocks; -1 = closed and drained
    {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [&] { return !q_.empty() || closed_; });
        if(q_.empty()) return -1;
        int v = q_.front(); q_.pop_front();
        return v;
    }
    void close() { { std::lock_guard<std::mutex> l(m_); closed_ = true; } cv_.notify_all(); }
private:
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<int> q_;
    bool closed_ = false;
};

static void write_file(const std::string& out_dir, uint frame, const std::vector<uint8_t>& bmp)
{
    char name[512];
    snprintf(name, sizeof(name), "%s/frame_%04u.bmp", out_dir.c_str(), frame); // main.cc:93-101
    FILE* f = fopen(name, "w");
    if(!f) { fprintf(stderr, "Failed to write %s\n", name); exit(1); }
    // a short write (full disk) must not end in exit code 0 with a truncated frame
    const bool ok = fwrite(bmp.data(), 1, bmp.size(), f) == bmp.size();
    if(fclose(f) != 0 || !ok) { fprintf(stderr, "Failed to write %s (short write)\n", name); exit(1); }
}

// strict non-negative integer argument (atoi would turn "-3" into a huge uint and "x" into 0)
static bool parse_uint(const char* s, long lo, long hi, long& out)
{
    char* end = nullptr;
    const long v = strtol(s, &end, 10);
    if(end == s || *end != 0 || v < lo || v > hi) return false;
    out = v;
    return true;
}

int main(int argc, char** argv)
{
    setlocale(LC_ALL, "C"); // main.cc:63
    int gpus = 1;
    uint frame_begin = 0, frame_end = 0, frame_step = 1;
    std::string out_dir = "output", dump_dir;
    bool write_files = true, serial = false;
    for(int i = 1; i < argc; ++i)
    {
        long v = 0, w = 0;
        if(!strcmp(argv[i], "--gpus") && i + 1 < argc && parse_uint(argv[i + 1], 1, 1024, v)) { gpus = (int)v; ++i; }
        else if(!strcmp(argv[i], "--frames") && i + 2 < argc && parse_uint(argv[i + 1], 0, 1 << 30, v) && parse_uint(argv[i + 2], v + 1, 1 << 30, w))
        { frame_begin = (uint)v; frame_end = (uint)w; i += 2; }
        else if(!strcmp(argv[i], "--step") && i + 1 < argc && parse_uint(argv[i + 1], 1, 1 << 30, v)) { frame_step = (uint)v; ++i; }
        else if(!strcmp(argv[i], "--out") && i + 1 < argc) out_dir = argv[++i];
        else if(!strcmp(argv[i], "--dump") && i + 1 < argc) dump_dir = argv[++i];
        else if(!strcmp(argv[i], "--no-write")) write_files = false;
        else if(!strcmp(argv[i], "--serial")) serial = true;
        else { fprintf(stderr, "usage: %s [--gpus N>=1] [--frames BEGIN END (BEGIN < END)] [--step S>=1] [--out DIR] [--dump DIR] [--no-write] [--serial]\n", argv[0]); return 2; }
    }

    double t0 = now_s();
    // Creating eight CUDA contexts takes the driver ~7 s on an 8-GPU box: start that now, one thread per GPU, and load
    // the scene (1.4 s) and do the library's host-side BVH builds (3 s, once for all GPUs) meanwhile.
    std::vector<std::thread> warm;
    for(int g = 0; g < gpus; ++g) warm.emplace_back([g]() { ptgpu_warm_up(g); });
    scene loaded = load_scene();
    printf("EXECUTION TIME OF load_scene() : %.0fms\n", 1e3 * (now_s() - t0));
    if(frame_end == 0) frame_end = get_animation_frame_count(loaded); // main.cc:74 as intended
    if(loaded.subframes.empty()) ptgpu_detail::prepare_static(loaded);
    for(auto& w : warm) w.join();
    printf("START-UP (contexts, scene, static BVHs) : %.0fms\n", 1e3 * (now_s() - t0));

    if(!dump_dir.empty())
    {   // test hook: the arrays this program hands to the C ABI (main.cc:29-37), so that another host can
        // push exactly the same inputs through the library and compare the frames byte for byte
        auto put = [&](const std::string& name, const void* p, size_t bytes) {
            FILE* f = fopen((dump_dir + "/" + name).c_str(), "w");
            if(!f || fwrite(p, 1, bytes, f) != bytes || fclose(f) != 0) { fprintf(stderr, "Failed to write %s\n", name.c_str()); exit(1); }
        };
        scene s = loaded;
        for(uint frame = frame_begin; frame < frame_end; frame += frame_step)
        {
            setup_animation_frame(s, frame);
            const size_t n_static_nodes = s.subframes[0].tlas.node_offset, n_static = s.static_instance_count;
            if(frame == frame_begin)
            {
                put("nodes.bin", s.bvh_buf.nodes.data(), n_static_nodes * sizeof(bvh_node));
                put("links.bin", s.bvh_buf.links.data(), 8 * n_static_nodes * sizeof(bvh_link));
                put("indices.bin", s.mesh_buf.indices.data(), s.mesh_buf.indices.size() * 4);
                put("pos.bin", s.mesh_buf.pos.data(), s.mesh_buf.pos.size() * sizeof(float3));
                put("normal.bin", s.mesh_buf.normal.data(), s.mesh_buf.normal.size() * sizeof(float3));
                put("albedo.bin", s.mesh_buf.albedo.data(), s.mesh_buf.albedo.size() * sizeof(float4));
                put("material.bin", s.mesh_buf.material.data(), s.mesh_buf.material.size() * sizeof(float4));
                put("instances.bin", s.instances.data(), n_static * sizeof(tlas_instance));
            }
            char pre[64];
            snprintf(pre, sizeof(pre), "frame_%04u_", frame);
            put(std::string(pre) + "subframes.bin", s.subframes.data(), s.subframes.size() * sizeof(subframe));
            put(std::string(pre) + "dyn_instances.bin", s.instances.data() + n_static, (s.instances.size() - n_static) * sizeof(tlas_instance));
            put(std::string(pre) + "tlas_nodes.bin", s.bvh_buf.nodes.data() + n_static_nodes, (s.bvh_buf.nodes.size() - n_static_nodes) * sizeof(bvh_node));
            put(std::string(pre) + "tlas_links.bin", s.bvh_buf.links.data() + 8 * n_static_nodes, (s.bvh_buf.links.size() - 8 * n_static_nodes) * sizeof(bvh_link));
        }
    }

    ptgpu_config cfg;
    cfg.width = IMAGE_WIDTH; cfg.height = IMAGE_HEIGHT; cfg.spp = SAMPLES_PER_PIXEL; cfg.max_bounces = MAX_BOUNCES;
    cfg.student_id = STUDENT_ID; cfg.samples_per_subframe = SAMPLES_PER_MOTION_BLUR_STEP;

    std::atomic<uint> next_frame{frame_begin};
    std::atomic<uint> frames_done{0};
    std::atomic<long long> us_setup{0}, us_render{0}, us_write{0}; // busy time per stage, summed over workers
    auto timed = [](std::atomic<long long>& acc, auto&& fn) {
        const double a = now_s();
        fn();
        acc += (long long)(1e6 * (now_s() - a));
    };
    double render_t0 = now_s();
    std::vector<std::thread> workers;
    for(int g = 0; g < gpus; ++g)
    {
        if(serial)
        {
            workers.emplace_back([&, g]() {
                scene s = loaded; // private copy: setup_animation_frame is not re-entrant
                ptgpu_detail::renderer<scene> r(g, cfg);
                std::vector<uint8_t> bmp(r.bmp_size());
                for(;;)
                {
                    uint frame = next_frame.fetch_add(frame_step);
                    if(frame >= frame_end) break;
                    timed(us_setup, [&] { setup_animation_frame(s, frame); });
                    timed(us_render, [&] { r.render_bmp(s, bmp.data()); });
                    if(write_files) timed(us_write, [&] { write_file(out_dir, frame, bmp); });
                    frames_done++;
                }
            });
            continue;
        }
        workers.emplace_back([&, g]() {
            // rings: 2 scene copies between setup and render, 3 images between render and writer
            constexpr int N_SCENES = 2, N_IMAGES = 3;
            std::vector<scene> scenes(N_SCENES, loaded);
            uint scene_frame[N_SCENES] = {};
            ptgpu_detail::renderer<scene> r(g, cfg);
            const double t_ctx = now_s() - render_t0;
            bool first = true;
            std::vector<std::vector<uint8_t>> images(N_IMAGES, std::vector<uint8_t>(r.bmp_size()));
            uint image_frame[N_IMAGES] = {};
            slot_queue scenes_free, scenes_ready, images_free, images_ready;
            for(int i = 0; i < N_SCENES; ++i) scenes_free.push(i);
            for(int i = 0; i < N_IMAGES; ++i) images_free.push(i);

            std::thread setup([&]() {
                for(;;)
                {
                    int si = scenes_free.pop();
                    if(si < 0) break;
                    uint frame = next_frame.fetch_add(frame_step);
                    if(frame >= frame_end) break;
                    timed(us_setup, [&] { setup_animation_frame(scenes[si], frame); });
                    scene_frame[si] = frame;
                    scenes_ready.push(si);
                }
                scenes_ready.close();
            });
            std::thread writer([&]() {
                for(;;)
                {
                    int ii = images_ready.pop();
                    if(ii < 0) break;
                    if(write_files) timed(us_write, [&] { write_file(out_dir, image_frame[ii], images[ii]); });
                    frames_done++;
                    images_free.push(ii);
                }
            });
            for(;;)
            {
                int si = scenes_ready.pop();
                if(si < 0) break;
                int ii = images_free.pop();
                timed(us_render, [&] { r.render_bmp(scenes[si], images[ii].data()); });
                if(first)
                {   // start-up cost of this worker: CUDA context, then scene upload + BVH builds inside the first frame
                    fprintf(stderr, "GPU %d: context after %.2fs, first frame (with the scene upload) done after %.2fs\n", g, t_ctx, now_s() - render_t0);
                    first = false;
                }
                image_frame[ii] = scene_frame[si];
                scenes_free.push(si);
                images_ready.push(ii);
            }
            scenes_free.close();
            images_ready.close();
            setup.join();
            writer.join();
        });
    }
    for(auto& w : workers) w.join();
    double dt = now_s() - render_t0;
    uint n = frames_done.load();
    printf("\n\nRENDERED %u FRAMES ON %d GPU(S) IN %.3fs (%.1f Mpaths/s); TOTAL %.0fms\n", n, gpus, dt,
           n * (double)IMAGE_WIDTH * IMAGE_HEIGHT * SAMPLES_PER_PIXEL / dt / 1e6, 1e3 * (now_s() - t0));
    printf("STAGE BUSY TIME (all workers): setup_animation_frame %.3fs, upload+render+readback %.3fs, write %.3fs; %s\n",
           1e-6 * us_setup.load(), 1e-6 * us_render.load(), 1e-6 * us_write.load(), serial ? "serial" : "pipelined");
    return 0;
}
