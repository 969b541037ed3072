#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — dump scene snapshots (inputs of the render seam) with the oracle.

Runs the reference's own load_scene()/setup_animation_frame() (through oracle/_ref/libptref*.so)
and stores the arrays they hand to baseline_render (main.cc:29-37) under scenes/_cache/, so that
bench.py and the tools can drive the C ABI on the GPU box without touching oracle/.
"""
import argparse
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refbind  # noqa: E402

# frames the benchmark cycles through: spread over the animation's content map
DEFAULT_FRAMES = [0, 100, 200, 330, 420, 520, 660, 800, 1000, 1100, 1250, 1400, 1600, 1750]


def load_scene_io():
    p = os.path.join(ROOT, "path-tracing...but-on-the-lumi-cluster_b200", "scene_io.py")
    spec = importlib.util.spec_from_file_location("ptb200_scene_io", p)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="fast", help="oracle build: fast (TESTING), prod, mb")
    ap.add_argument("--tag", default=None)
    ap.add_argument("--frames", type=int, nargs="*", default=DEFAULT_FRAMES)
    ap.add_argument("--force", action="store_true")
    args = ap.parse_args()
    tag = args.tag or {"fast": "testing", "strict": "testing", "prod": "production", "mb": "motionblur"}[args.variant]
    sio = load_scene_io()
    if not refbind.available(args.variant):
        print("make_snapshots: oracle variant %s not built" % args.variant)
        return 1
    todo = [f for f in args.frames if args.force or not os.path.exists(sio.frame_path(f, tag))]
    need_static = args.force or not os.path.exists(sio.static_path(tag))
    if not todo and not need_static:
        return 0
    o = refbind.Oracle(args.variant)
    o.load_scene()
    if need_static:
        v = o.setup_frame(0)
        sio.save_static(sio.static_path(tag), v)
    for f in todo:
        v = o.setup_frame(f)
        sio.save_frame(sio.frame_path(f, tag), v, f)
    print("make_snapshots: %s: static%s + %d frames under %s" % (tag, "" if need_static else " (kept)", len(todo), sio.CACHE))
    return 0


if __name__ == "__main__":
    sys.exit(main())
