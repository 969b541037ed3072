#!/usr/bin/env python3
"""Scan a directory of output/frame_NNNN.bmp files (what main.cc:93-101 / pt_gpu write): every frame of the
range present, right size and header, not black, no frame that jumps away from both neighbours (a NaN or a
lost kernel shows as a black or flat frame: tonemap_pixel maps NaN to 0). Prints a summary; exit code 1 on
any finding.   usage: scan_frames.py DIR [--frames 1800] [--width 640 --height 360]"""
import argparse
import os
import sys

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dir")
    ap.add_argument("--frames", type=int, default=1800)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=360)
    args = ap.parse_args()
    pitch = (args.width * 3 + 3) // 4 * 4
    size = 54 + pitch * args.height
    missing, bad_size, black, means = [], [], [], np.full(args.frames, np.nan)
    for f in range(args.frames):
        p = os.path.join(args.dir, "frame_%04d.bmp" % f)
        if not os.path.exists(p):
            missing.append(f)
            continue
        raw = np.fromfile(p, dtype=np.uint8)
        if raw.size != size or raw[0] != ord("B") or raw[1] != ord("M"):
            bad_size.append(f)
            continue
        px = raw[54:].reshape(args.height, pitch)[:, :args.width * 3]
        means[f] = px.mean()
        if px.max() == 0:
            black.append(f)
    ok = np.isfinite(means)
    # a frame whose mean brightness is far from BOTH neighbours while they agree with each other (hard cuts
    # of the animation change one side only)
    jumps = []
    for f in range(1, args.frames - 1):
        if ok[f - 1] and ok[f] and ok[f + 1]:
            a, b, c = means[f - 1], means[f], means[f + 1]
            if abs(a - c) < 2.0 and abs(b - a) > 10.0 and abs(b - c) > 10.0:
                jumps.append(f)
    print("frames present %d of %d; missing %s; wrong size/header %s; black %s; isolated brightness jumps %s" % (
        int(ok.sum()), args.frames, missing[:10], bad_size[:10], black[:10], jumps[:10]))
    print("mean brightness over the animation: min %.2f (frame %d), max %.2f (frame %d)" % (
        np.nanmin(means), int(np.nanargmin(means)), np.nanmax(means), int(np.nanargmax(means))))
    return 1 if (missing or bad_size or black or jumps) else 0


if __name__ == "__main__":
    sys.exit(main())
