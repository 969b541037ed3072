#!/usr/bin/env python3
"""Render snapshot frames through the C ABI (no oracle): the command ncu profiles.
usage: prof_frame.py [--frames 520 ...] [--reps 2] [--kernel 0|1] [--traversal 0|1] [--spp N] [--opt flat=0 --opt sort=0 ...]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, nargs="*", default=[520])
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--kernel", type=int, default=2)
    ap.add_argument("--traversal", type=int, default=0)
    ap.add_argument("--spp", type=int, default=256)
    ap.add_argument("--opt", action="append", default=[], help="ptgpu_set_option pair key=value (applied before the upload)")
    args = ap.parse_args()
    pkg = ge.load_package()
    sio = pkg.scene_io
    cfg = pkg.Config.testing()
    cfg.spp = args.spp
    r = pkg.Renderer(cfg, 0)
    for kv in args.opt:
        k, v = kv.split("=")
        r.set_option(k, int(v))
    r.upload_static(**sio.load_static(sio.static_path()))
    r.set_option("kernel", args.kernel)
    r.set_option("traversal", args.traversal)
    for f in args.frames:
        fr = sio.load_frame(sio.frame_path(f))
        r.set_frame(fr["subframes"], fr["dyn_instances"], fr["tlas_nodes"], fr["tlas_links"])
        for _ in range(args.reps):
            r.render_async()
            r.sync()
            ms, n = r.last_render_ms()
            print("frame %d: %.2f ms, %.1f Mpaths/s (%d launches)" % (f, ms, cfg.width * cfg.height * cfg.spp / ms / 1e3, n), flush=True)


if __name__ == "__main__":
    main()
