// ptgpu_main.cc — multi-GPU replacement of the reference's main.cc:60-108.
//
// Same outputs (output/frame_NNNN.bmp, timing prints), but baseline_render is replaced by the C ABI
// of libptgpu.so and whole animation frames are sharded over the GPUs of one box: one host thread
// per GPU, a shared atomic frame counter (frame cost varies ~7x over the animation, so the queue is
// dynamic, not blocked), no inter-GPU traffic; only the finished 8-bit BMP image returns to the host.
//
// This file is compiled TOGETHER with the reference's own scene.cc/bvh.cc/mesh.cc (it includes
// scene.hh / config.hh from the reference tree, -I$REF) — it is the integration a maintainer of
// the reference would build, see INTEGRATION.md. setup_animation_frame() mutates its `scene`
// (scene.cc:274-277), so every worker owns a copy of the loaded scene.
#include "scene.hh"
#include "config.hh"
#include "ptgpu_render.hh"

#include <atomic>
#include <chrono>
#include <clocale>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

static double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv)
{
    setlocale(LC_ALL, "C"); // main.cc:63
    int gpus = 1;
    uint frame_begin = 0, frame_end = 0, frame_step = 1;
    std::string out_dir = "output";
    bool write_files = true;
    for(int i = 1; i < argc; ++i)
    {
        if(!strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = atoi(argv[++i]);
        else if(!strcmp(argv[i], "--frames") && i + 2 < argc) { frame_begin = atoi(argv[++i]); frame_end = atoi(argv[++i]); }
        else if(!strcmp(argv[i], "--step") && i + 1 < argc) frame_step = atoi(argv[++i]);
        else if(!strcmp(argv[i], "--out") && i + 1 < argc) out_dir = argv[++i];
        else if(!strcmp(argv[i], "--no-write")) write_files = false;
        else { fprintf(stderr, "usage: %s [--gpus N] [--frames BEGIN END] [--step S] [--out DIR] [--no-write]\n", argv[0]); return 2; }
    }

    double t0 = now_s();
    scene loaded = load_scene();
    printf("EXECUTION TIME OF load_scene() : %.0fms\n", 1e3 * (now_s() - t0));
    if(frame_end == 0) frame_end = get_animation_frame_count(loaded); // main.cc:74 as intended
    if(frame_step == 0) frame_step = 1;

    ptgpu_config cfg;
    cfg.width = IMAGE_WIDTH; cfg.height = IMAGE_HEIGHT; cfg.spp = SAMPLES_PER_PIXEL; cfg.max_bounces = MAX_BOUNCES;
    cfg.student_id = STUDENT_ID; cfg.samples_per_subframe = SAMPLES_PER_MOTION_BLUR_STEP;

    std::atomic<uint> next_frame{frame_begin};
    std::atomic<uint> frames_done{0};
    double render_t0 = now_s();
    std::vector<std::thread> workers;
    for(int g = 0; g < gpus; ++g)
    {
        workers.emplace_back([&, g]() {
            scene s = loaded; // private copy: setup_animation_frame is not re-entrant
            ptgpu_detail::renderer<scene> r(g, cfg);
            std::vector<uint8_t> bmp(r.bmp_size());
            for(;;)
            {
                uint frame = next_frame.fetch_add(frame_step);
                if(frame >= frame_end) break;
                setup_animation_frame(s, frame);
                r.render_bmp(s, bmp.data());
                if(write_files)
                {
                    char name[512];
                    snprintf(name, sizeof(name), "%s/frame_%04u.bmp", out_dir.c_str(), frame); // main.cc:93-101
                    FILE* f = fopen(name, "w");
                    if(!f) { fprintf(stderr, "Failed to write %s\n", name); exit(1); }
                    fwrite(bmp.data(), 1, bmp.size(), f);
                    fclose(f);
                }
                frames_done++;
            }
        });
    }
    for(auto& w : workers) w.join();
    double dt = now_s() - render_t0;
    uint n = frames_done.load();
    printf("\n\nRENDERED %u FRAMES ON %d GPU(S) IN %.3fs (%.1f Mpaths/s); TOTAL %.0fms\n", n, gpus, dt,
           n * (double)IMAGE_WIDTH * IMAGE_HEIGHT * SAMPLES_PER_PIXEL / dt / 1e6, 1e3 * (now_s() - t0));
    return 0;
}
