#!/usr/bin/env python3
"""Benchmark of the rendering hot path (baseline_render -> path_trace_pixel -> tonemap_pixel).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle/_ref)

A step renders ONE animation frame at the shipped config.hh size (640x360, 256 spp, 4 bounces =
58,982,400 paths). The frames are scene snapshots of the reference's own animation (the arrays its
setup_animation_frame() hands to baseline_render, dumped by oracle/make_snapshots.py where the
reference is mounted; stand-in terrain/pine/bunny geometry). The K steps of a run walk
the 14 snapshot frames in order (step j -> snapshot j mod 14, frames spread over the 1800-frame animation:
the "full default animation" workload of BASELINE.json configs[2]); rank r renders the same K-frame list
rotated by r (step i -> list[(i + r) mod K]), so at every N every rank renders exactly the same multiset
of frames and v_N / (N v_1) measures scaling, not the frame mix (frame 0 is 3.5x cheaper than the mean).
Frames are sharded over ranks with no data-path collective (weak scaling: K frames per rank). After the
timed loops every 10th frame of the real animation (180 frames, sharded over the ranks) is rendered through
the frame-setup module and read back: `animation_seconds_measured` = that wall time x 10.

Printed JSON (rank 0, one line): metric Mpaths/s; `value` = device-resident throughput (static scene
in HBM, per-frame input = ~40 KB of transforms), `e2e` = the same through ptgpu_render_frame() with
host buffers in and the BGRA frame out; `roofline` = FP32-issue roofline of the whole frame
(algorithmic flops of SURVEY.md 8(d), counted on the reference's own BVH); `cpu_baseline` = the
reference's baseline_render loop on one host core, on a bounded sample of the same frames.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT, SPP = 640, 360, 256
PATHS_PER_FRAME = WIDTH * HEIGHT * SPP
ANIMATION_FRAMES = 1800
# bounded CPU sample per frame: every 16th row, one sample per motion-blur subframe
CPU_ROW_STRIDE, CPU_SAMPLE_STRIDE = 16, 8
CPU_SAMPLE_FRAMES = [0, 330, 520, 1400]


def load_json(path, default=None):
    try:
        with open(path) as f:
            return json.load(f)
    except (OSError, ValueError):
        return default


def flops_table():
    """Frozen algorithmic flops per path, per snapshot frame (profiles/flops_per_path.json)."""
    t = load_json(os.path.join(ROOT, "profiles", "flops_per_path.json"), {}) or {}
    return {int(k): v for k, v in t.get("frames", {}).items()}


def traversal_flops_table():
    """The traversal terms of the same formula (25 N_node + 56 N_tri + 61 N_blas + 9 N_ray): the
    algorithmic flops of the dominant kernel, wf_trace_cw_kernel, per path."""
    t = load_json(os.path.join(ROOT, "profiles", "flops_per_path.json"), {}) or {}
    return {int(k): 25.0 * e["node_visits"] + 56.0 * e["tri_tests"] + 61.0 * e["blas_enters"] + 9.0 * e["rays"]
            for k, e in t.get("events_per_path", {}).items()}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons while the timed region runs: through NVML in-process
    (nvidia_ml_py), falling back to one nvidia-smi process per sample. (Forking nvidia-smi five times a
    second from a process that maps a CUDA context cost the timed loop ~5 ms per step.)"""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index, uuid=None):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.source = "nvidia-smi"
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
                except Exception:
                    h = None
            self.handle = h if h is not None else pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.source = "nvml"
        except Exception:
            self.nvml = self.handle = None

    def sample_nvml(self):
        n = self.nvml
        self.samples.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        mask = int(get(self.handle))
        for name, bit in self.REASONS:
            if mask & bit:
                self.reasons.add(name)

    def sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        f = [x.strip() for x in out.strip().split(",")]
        self.samples.append(float(f[0]))
        self.max_mhz = float(f[1])
        for n, v in zip(names, f[2:]):
            if v.lower().startswith("active"):
                self.reasons.add(n)

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self.nvml is not None:
                    self.sample_nvml()
                else:
                    self.sample_smi()
            except Exception:
                pass
            self.stop_flag.wait(float(os.environ.get("BENCH_SAMPLER_PERIOD", "0.1" if self.nvml is not None else "0.5")))

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s), "source": self.source}


def cpu_sample_rows():
    return list(range(0, HEIGHT, CPU_ROW_STRIDE))


def cpu_render_sample(oracle, frame, nthreads):
    """Times the reference loop nest on the bounded sample of one frame; returns (paths, seconds)."""
    oracle.setup_frame(frame)
    paths, t0 = 0, time.perf_counter()
    n_samples = SPP // CPU_SAMPLE_STRIDE
    for y in cpu_sample_rows():
        oracle.render_rect(0, y, WIDTH, 1, 0, n_samples, CPU_SAMPLE_STRIDE, nthreads=nthreads, tonemap=False)
        paths += WIDTH * n_samples
    return paths, time.perf_counter() - t0


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(args):
    """--impl reference: the reference's own CPU implementation (oracle/_ref), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import refbind
    if not refbind.available("fast"):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (needs /root/reference at build time)"}))
        return 0
    o = refbind.get("fast")
    o.load_scene()
    cores = os.cpu_count() or 1
    frames = frame_list()
    for i in range(min(args.warmup, 1)):   # one warm step pages the scene in; CPU steps are seconds long
        cpu_render_sample(o, frames[i % len(frames)], cores)
    paths, secs = 0, 0.0
    for i in range(args.steps):
        p, s = cpu_render_sample(o, frames[i % len(frames)], cores)
        paths += p
        secs += s
    value = paths / secs / 1e6
    sample = "per step: rows 0,%d,..,%d of one frame x samples 0,%d,..,%d (%d paths) of the 58,982,400-path frame" % (
        CPU_ROW_STRIDE, cpu_sample_rows()[-1], CPU_SAMPLE_STRIDE, SPP - CPU_SAMPLE_STRIDE, len(cpu_sample_rows()) * WIDTH * (SPP // CPU_SAMPLE_STRIDE))
    print(json.dumps({
        "impl": "reference", "metric": "Mpaths/s", "value": round(value, 4), "unit": "Mpaths/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1e3 * secs / max(args.steps, 1), 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (reference animation, stand-in terrain/pine/bunny assets)",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": round(value, 4), "unit": "Mpaths/s", "cores": cores, "kind": "reference",
                         "sample": sample, "cpu": cpu_model(), "build": "oracle/_ref/libptref.so (-O3 -ffast-math -fopenmp -march=x86-64-v3)"},
        "e2e": {"value": round(value, 4), "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))
    return 0


def frame_list():
    import __graft_entry__ as ge
    sio = ge.load_package().scene_io
    frames = sio.available_frames("testing")
    if not frames:
        raise SystemExit("bench.py: no scene snapshots under scenes/_cache (run __graft_entry__.build() where the reference is mounted)")
    return frames


def workload_config(n_gpus, animation=False):
    return {"workload": "full default animation, config.hh TESTING size: 640x360, 256 spp, 4 bounces, 32 motion-blur subframes; "
                        "one step = one frame (58,982,400 paths); " + (
                            "steps walk the 1800-frame animation at a uniform stride" if animation else
                            "steps walk 14 frames spread over the 1800-frame animation (0,100,200,330,420,520,660,800,1000,1100,1250,1400,1600,1750); "
                            "every rank renders the same K-frame list, rotated by its rank"),
            "frames_per_step": 1, "paths_per_step": PATHS_PER_FRAME, "parallelism": "frames sharded over %d GPU(s), no collective" % n_gpus,
            "l2_policy": "every step renders a different frame and streams the 11.9 GB path-state pool (one slot per path) and the 1 GB flat scene through HBM, far larger than the 126 MB L2"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=14)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel", type=int, default=None, help="0 megakernel, 1 tiles, 2 wavefront (default)")
    ap.add_argument("--animation-stride", type=int, default=10,
                    help="after the timed loops render every S-th frame of the 1800-frame animation end to end (0 = skip)")
    ap.add_argument("--animation", action="store_true",
                    help="walk the animation itself: step i renders frame (i*N + rank) * 1800 / (steps*N) through the "
                         "frame-setup module (--steps 1800 --gpus 1 = every frame); no roofline (flops are frozen for the 14 snapshot frames)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import numpy as np
    import torch
    import __graft_entry__ as ge
    pkg = ge.load_package()
    sio = pkg.scene_io

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the render path has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on stdout when the first communicator is created; the contract is
        # ONE JSON line there, so stdout points at stderr until the communicator exists
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    frames = frame_list()
    cfg = pkg.Config.testing()
    r = pkg.Renderer(cfg, device=local_rank)
    if args.kernel is not None:
        r.set_option("kernel", args.kernel)
    r.upload_static(**sio.load_static(sio.static_path("testing")))
    snaps = {f: sio.load_frame(sio.frame_path(f, "testing")) for f in frames}

    # per-frame scene state from the frame-setup module (csrc/frame_setup.cu): keyframe replay on the
    # host, ~30 KB up, no reference TLAS arrays
    anim = pkg.Animation(cfg) if os.path.exists(pkg.animation.default_path()) else None
    if args.animation and anim is None:
        raise SystemExit("bench.py: --animation needs scenes/_cache/animation.json (run __graft_entry__.build() where the reference is mounted)")

    def frame_of(step):
        if args.animation:
            return ((step * world + rank) * ANIMATION_FRAMES) // (args.steps * world)
        # the same K-frame list on every rank, rotated by the rank: identical multisets at every N
        return pkg.sharding.bench_frame(step, rank, max(args.steps, 1), frames)

    def set_frame(f):
        if anim is not None:
            anim.set_frame(r, f)
        else:
            s = snaps[f]
            r.set_frame(s["subframes"], s["dyn_instances"], s["tlas_nodes"], s["tlas_links"])

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        r.sync()

    # pinned output buffer for the end-to-end loop
    out_pinned = torch.empty((HEIGHT, WIDTH, 4), dtype=torch.uint8).pin_memory()
    out_np = out_pinned.numpy()

    for i in range(args.warmup):
        set_frame(frame_of(i))
        r.render_async()
        r.sync()

    sampler = ClockSampler(local_rank, getattr(torch.cuda.get_device_properties(local_rank), "uuid", None))
    sampler.start()

    # ---- loop A: device-resident. Static scene in HBM; per step ~40 KB of per-frame transforms go up,
    #      nothing comes back. Timed by the bracket (wall) and by CUDA events on the render stream.
    barrier()
    t0 = time.perf_counter()
    dev_ms, launches, flops = 0.0, 0, 0.0
    trace_us, trace_launches, trace_flops = 0.0, 0, 0.0
    ftab, ttab = flops_table(), traversal_flops_table()
    trace_steps = os.environ.get("BENCH_TRACE_STEPS") and rank == 0
    step_wall = []
    for i in range(args.steps):
        ts = time.perf_counter()
        f = frame_of(i)
        set_frame(f)
        r.render_async()
        ms, n = r.last_render_ms()   # waits for this frame's end event
        if trace_steps:
            print("step %d frame %d: wall %.2f ms, device %.2f ms" % (i, f, 1e3 * (time.perf_counter() - ts), ms), file=sys.stderr)
        dev_ms += ms
        launches += n
        flops += ftab.get(f, 0.0) * PATHS_PER_FRAME
        # the traversal launches of this frame, CUDA events on the render stream around each launch
        trace_us += r.get_stat("trace_us")
        trace_launches += r.get_stat("trace_launches")
        trace_flops += ttab.get(f, 0.0) * PATHS_PER_FRAME
        step_wall.append(1e3 * (time.perf_counter() - ts))
    t_loop = time.perf_counter() - t0
    barrier()
    wall_a = time.perf_counter() - t0
    if trace_steps:
        print("steps %.2f ms, loop %.2f ms, with closing barrier %.2f ms" % (sum(step_wall), 1e3 * t_loop, 1e3 * wall_a), file=sys.stderr)

    # ---- loop B: end to end through the C ABI call a drop-in user makes (ptgpu_render_frame): host
    #      arrays in (subframes, dynamic instances, reference TLAS arrays), pinned BGRA frame out.
    barrier()
    t0 = time.perf_counter()
    h2d = 0
    for i in range(args.steps):
        f = frame_of(i)
        if args.animation:
            sub, dyn, db, de = anim.frame(f)                       # host: keyframe replay
            r.set_frame_ranges(sub, dyn, db, de)                   # host arrays -> device
            r.render(out=out_np)                                   # render + BGRA frame back to pinned host memory
            h2d += sub.nbytes + dyn.nbytes + dyn.shape[0] * 128 + sub.shape[0] * 8
            continue
        s = snaps[f]
        r.render_frame(s["subframes"], s["dyn_instances"], s["tlas_nodes"], s["tlas_links"], out=out_np)
        h2d += s["subframes"].nbytes + s["dyn_instances"].nbytes + (s["dyn_instances"].shape[0] * 128) + s["subframes"].shape[0] * 8
    barrier()
    wall_b = time.perf_counter() - t0
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    # ---- loop C: the animation itself, every S-th frame, end to end (keyframe replay on the host, ~30 KB up,
    #      render, BGRA frame back to pinned host memory), frames dealt round-robin to the ranks. This is the
    #      loop of main.cc:78-102 without the file write; S = 1 and one rank is the whole 1800-frame animation.
    anim_wall, anim_frames = 0.0, 0
    if anim is not None and args.animation_stride > 0 and not args.animation:
        todo = list(range(0, ANIMATION_FRAMES, args.animation_stride))
        mine = todo[rank::world]
        barrier()
        t0 = time.perf_counter()
        for f in mine:
            sub, dyn, db, de = anim.frame(f)
            r.set_frame_ranges(sub, dyn, db, de)
            r.render(out=out_np)
        barrier()
        anim_wall, anim_frames = time.perf_counter() - t0, len(todo)

    clocks_mine = sampler.result()
    mine_stats = torch.tensor([wall_a, wall_b, dev_ms / 1e3, trace_us / 1e6, float(clocks_mine["sm_mhz"] or 0.0), anim_wall],
                              dtype=torch.float64, device="cuda")
    per_rank = [mine_stats.clone() for _ in range(world)]
    if dist is not None:
        dist.all_gather(per_rank, mine_stats)
    per_rank = [[float(x) for x in t.tolist()] for t in per_rank]
    anim_wall = max(p[5] for p in per_rank)

    times = torch.tensor([wall_a, wall_b, dev_ms / 1e3], dtype=torch.float64, device="cuda")
    sums = torch.tensor([float(launches), flops], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    wall_a, wall_b, dev_s = [float(x) for x in times.tolist()]
    launches, flops = [float(x) for x in sums.tolist()]

    if rank == 0:
        total_paths = args.steps * world * PATHS_PER_FRAME
        value = total_paths / wall_a / 1e6
        e2e = total_paths / wall_b / 1e6
        peaks = load_json(os.path.join(ROOT, "MEASURED_PEAKS.json"), {}) or {}
        sm_max = float(peaks.get("sm_max_mhz", 1965.0))
        peak_tflops = 148 * 128 * 2 * sm_max * 1e6 / 1e12   # FP32 FMA issue peak
        # roofline on the device time of ONE rank's frames (max over ranks), flops of all ranks / world
        achieved = (flops / world) / dev_s / 1e12 if dev_s > 0 and flops > 0 and not args.animation else None
        prof = load_json(os.path.join(ROOT, "profiles", "roofline_inputs.json"), {}) or {}
        # dominant kernel (rank 0's launches): algorithmic traversal flops / its own launch durations
        k_ok = trace_us > 0 and trace_flops > 0 and trace_launches > 0 and not args.animation
        k_achieved = trace_flops / (trace_us * 1e-6) / 1e12 if k_ok else None
        line = {
            "metric": "Mpaths/s", "value": round(value, 2), "unit": "Mpaths/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(1e3 * wall_a / args.steps, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (reference animation via scene snapshots; stand-in terrain/pine/bunny assets)",
            "config": workload_config(world, args.animation),
            "animation_seconds_estimate": round(ANIMATION_FRAMES * (wall_a / args.steps) / world, 1),
            # measured: every S-th frame of the animation rendered end to end on these N GPUs, wall x S
            "animation_seconds_measured": round(anim_wall * args.animation_stride, 1) if anim_frames else None,
            "animation_sample": ("frames 0,%d,..,%d (%d frames) in %.2f s on %d GPU(s), ptgpu_anim_frame + ptgpu_set_frame_ranges + ptgpu_render"
                                 % (args.animation_stride, ANIMATION_FRAMES - args.animation_stride, anim_frames, anim_wall, world)) if anim_frames else None,
            # every rank's own clock: which GPU is the slow one, and by how much (value uses the max)
            "per_rank": {"wall_ms_per_step": [round(1e3 * p[0] / args.steps, 3) for p in per_rank],
                         "e2e_ms_per_step": [round(1e3 * p[1] / args.steps, 3) for p in per_rank],
                         "device_ms_per_step": [round(1e3 * p[2] / args.steps, 3) for p in per_rank],
                         "trace_ms_per_step": [round(1e3 * p[3] / args.steps, 3) for p in per_rank],
                         "sm_mhz": [p[4] for p in per_rank]},
            "device_ms_per_step": round(1e3 * dev_s / args.steps, 3),
            "timing_breakdown_ms": {"sum_of_steps": round(sum(step_wall), 2), "loop": round(1e3 * t_loop, 2), "with_closing_barrier": round(1e3 * wall_a, 2)},
            "e2e": {"value": round(e2e, 2), "unit": "Mpaths/s", "h2d_bytes_per_step": int(h2d / args.steps),
                    "d2h_bytes_per_step": WIDTH * HEIGHT * 4, "ms_per_step": round(1e3 * wall_b / args.steps, 3),
                    "call": "ptgpu_anim_frame + ptgpu_set_frame_ranges + ptgpu_render" if args.animation else "ptgpu_render_frame (include/ptgpu.h)"},
            "frame_setup": "ptgpu_set_animation_frame (csrc/frame_setup.cu)" if anim is not None else "snapshot arrays via ptgpu_set_frame",
            "gpu_launches": int(launches),
            "clocks": clocks_mine,
            # the dominant kernel (wf_trace_cw_kernel), as the contract asks; the whole frame beside it
            "roofline": {"bound": "fp32", "kernel": "wf_trace_cw_kernel",
                         "achieved": round(k_achieved, 3) if k_ok else None, "peak": round(peak_tflops, 2),
                         "unit": "TFLOP/s", "frac": round(k_achieved / peak_tflops, 4) if k_ok else None,
                         "traffic": prof.get("wf_trace_cw_dram_bytes_per_launch"),
                         "launches": int(trace_launches),
                         "avg_launch_ms": round(trace_us / trace_launches / 1e3, 3) if k_ok else None,
                         "flops_per_launch": round(trace_flops / trace_launches, 0) if k_ok else None,
                         "share_of_device_time": round(trace_us / 1e3 / dev_ms, 4) if k_ok and dev_ms > 0 else None,
                         "whole_frame": {"achieved": round(achieved, 3) if achieved else None,
                                         "frac": round(achieved / peak_tflops, 4) if achieved else None,
                                         "flops": "all terms of the formula x paths / CUDA-event device time of the frames (all kernels)"},
                         "note": "FP32-issue roofline (SURVEY.md 8(d)): algorithmic flops counted on the REFERENCE's BVH "
                                 "(profiles/flops_per_path.json); kernel line = traversal terms 25 N_node + 56 N_tri + 61 N_blas + 9 N_ray "
                                 "/ CUDA-event time of the traversal launches (render stream, rank 0); peak = 148 SM x 128 lanes x 2 x %.0f MHz "
                                 "(sm_max_mhz of MEASURED_PEAKS.json); traffic = dram read+write bytes per launch from the ncu --set full capture "
                                 "(profiles/roofline_inputs.json); tensor cores unused; the path is not HBM-bound" % sm_max},
        }
        if not args.no_cpu_baseline:
            try:
                from oracle import refbind
                if refbind.available("fast"):
                    o = refbind.get("fast")
                    o.load_scene()
                    paths, secs = 0, 0.0
                    for f in CPU_SAMPLE_FRAMES:
                        p, s = cpu_render_sample(o, f, 1)
                        paths += p
                        secs += s
                    line["cpu_baseline"] = {
                        "value": round(paths / secs / 1e6, 4), "unit": "Mpaths/s", "cores": 1, "kind": "reference",
                        "sample": "frames %s: rows 0,%d,.. x samples 0,%d,.. (%d paths, %.1f s) of the 58,982,400-path frames" % (
                            CPU_SAMPLE_FRAMES, CPU_ROW_STRIDE, CPU_SAMPLE_STRIDE, paths, secs),
                        "cpu": cpu_model(), "host_cores": os.cpu_count()}
                else:
                    line["cpu_baseline"] = {"value": None, "unit": "Mpaths/s", "cores": 1, "kind": "reference", "sample": "oracle/_ref not present"}
            except Exception as e:  # the GPU numbers stand on their own
                line["cpu_baseline"] = {"value": None, "unit": "Mpaths/s", "cores": 1, "kind": "reference", "sample": "failed: %s" % e}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    r.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
