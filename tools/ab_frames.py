#!/usr/bin/env python3
"""Developer tool: frame times of traversal variants side by side (A/B), with their own correctness checks.

usage: ab_frames.py [--frames 0 520 1400] [--configs "flat=0,sort=0;flat=1,sort=1"] [--reps 2] [--check]

Every config is a set of ptgpu_set_option pairs; `flat` is an upload-time option, so the tool keeps one
context per flat value. Per config and frame: best device time of the frame (CUDA events), and the device
time of its traversal launches, ray sort + generation and shading. --check adds, per config: the ray-by-ray
validation of the scheduled kernel against the plain traversal on a window (option "validate"), closest hits
of the flat scene against the instanced one (same instance / primitive, t), and a window against the oracle.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


# options a config does not name go back to these (options are sticky in a context)
DEFAULTS = {"sort": 1, "dyn_first": 1, "top_smem": 0, "plain_trace": 0, "l2_persist": 0, "node_threshold": 16, "node_burst": -1, "tri_threshold": 8,
            "xform_threshold": -1, "min_active": -1}


def parse_configs(text):
    out = []
    for part in text.split(";"):
        part = part.strip()
        if part:
            out.append({k.strip(): int(v) for k, v in (kv.split("=") for kv in part.split(","))})
    return out


def rays_in_scene(n, seed):
    rng = np.random.RandomState(seed)
    o = np.stack([rng.uniform(-95, 95, n), rng.uniform(0, 60, n), rng.uniform(-95, 95, n)], 1)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3], rays[:, 3], rays[:, 4:7], rays[:, 7] = o, 0.0, d, 1e9
    return rays


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, nargs="*", default=[0, 520, 1400])
    ap.add_argument("--configs", default="flat=0,sort=0;flat=0,sort=1;flat=1,sort=0;flat=1,sort=1")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--spp", type=int, default=256)
    args = ap.parse_args()
    pkg = ge.load_package()
    sio = pkg.scene_io
    cfg = pkg.Config.testing()
    cfg.spp = args.spp
    configs = parse_configs(args.configs)
    static = sio.load_static(sio.static_path())
    ctx = {}
    for c in configs:
        fl = c.get("flat", 1)
        if fl not in ctx:
            r = pkg.Renderer(cfg, 0)
            r.set_option("flat", fl)
            r.upload_static(**static)
            ctx[fl] = r
            if fl:
                print("flat scene: %d triangles, %d nodes, depth %d, built in %.1f s" % (
                    r.get_stat("flat_tris"), r.get_stat("flat_nodes"), r.get_stat("flat_depth"), r.get_stat("flat_build_ms") / 1e3), flush=True)
    oracle = None
    if args.check and not os.environ.get("PTGPU_NO_ORACLE"):
        try:
            from oracle import refbind
            oracle = refbind.get("fast")
            oracle.load_scene()
        except Exception as e:  # noqa
            print("oracle unavailable: %s" % e)
    paths = cfg.width * cfg.height * cfg.spp
    for f in args.frames:
        fr = sio.load_frame(sio.frame_path(f))
        for r in ctx.values():
            r.set_frame(fr["subframes"], fr["dyn_instances"], fr["tlas_nodes"], fr["tlas_links"])
        ref = None
        if oracle is not None:
            oracle.setup_frame(f)
            ref = oracle.render_rect(160, 90, 320, 180, 0, 8, 32)
        if args.check and 0 in ctx and 1 in ctx:
            rays = rays_in_scene(20000, f)
            fa, ua = ctx[0].trace_closest(rays, 0)
            fb, ub = ctx[1].trace_closest(rays, 0)
            hit_a, hit_b = fa[:, 0] > 0, fb[:, 0] > 0
            both = hit_a & hit_b
            same = both & (ua[:, 0] == ub[:, 0]) & (ua[:, 1] == ub[:, 1])
            trel = np.abs(fa[same, 0] - fb[same, 0]) / np.maximum(fa[same, 0], 1e-6)
            print("frame %4d closest hits, flat vs instanced: hit/miss agree %.5f, same (inst, prim) %.5f of %d hits, "
                  "t rel max %.2e, back_face agree %.5f" % (
                      f, (hit_a == hit_b).mean(), same.sum() / max(both.sum(), 1), both.sum(),
                      trel.max() if trel.size else 0.0, (ua[same, 2] == ub[same, 2]).mean()), flush=True)
        for c in configs:
            r = ctx[c.get("flat", 1)]
            for k, v in {**DEFAULTS, **c}.items():
                if k != "flat":
                    r.set_option(k, v)
            name = ",".join("%s=%d" % kv for kv in c.items())
            r.render_async(); r.sync()
            best = (1e9, 0, 0, 0)
            for _ in range(args.reps):
                r.render_async(); r.sync()
                ms = r.last_render_ms()[0]
                if ms < best[0]:
                    best = (ms, r.get_stat("trace_us") / 1e3, r.get_stat("sort_us") / 1e3, r.get_stat("shade_us") / 1e3)
            msg = "frame %4d %-28s %8.2f ms %7.1f Mpaths/s | trace %7.2f sort+gen %6.2f shade %6.2f" % (
                (f, name, best[0], paths / best[0] / 1e3) + best[1:])
            if c.get("l2_persist"):
                msg += " | L2 set aside %.0f MB, window %.0f MB" % (r.get_stat("l2_set_aside") / 1e6, r.get_stat("l2_window") / 1e6)
            if args.check:
                r.set_option("validate", 1)
                r.render_rect(256, 148, 128, 64, 0, 16, 16)
                msg += " | validate mismatches %d" % r.get_stat("validate_mismatches")
                r.set_option("validate", 0)
            if ref is not None:
                g = r.render_rect(160, 90, 320, 180, 0, 8, 32)
                mae = np.abs(g[1][..., :3].astype(float) - ref[1][..., :3].astype(float)).mean()
                rel = abs(g[0].mean() - ref[0].mean()) / ref[0].mean()
                msg += " | oracle window MAE %.4f mean-rel %.2e" % (mae, rel)
            print(msg, flush=True)


if __name__ == "__main__":
    main()
