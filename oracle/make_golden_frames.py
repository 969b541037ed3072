#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — render whole animation frames with the oracle (the unmodified reference,
oracle/_ref/libptref.so: baseline_render's loop nest, main.cc:12-46, all host threads) and commit them as
golden fixtures: tests/golden/oracle_frames/frame_NNNN.png (the tonemapped 8-bit frame, 640x360, RGB) and
oracle_frames.json (image-mean linear radiance per channel, seconds, sha256 of the stand-in assets).

The GPU box has no /root/reference and a full frame is 20-150 s of CPU, so the frames are rendered once where
the reference is mounted:   python oracle/make_golden_frames.py [--frames 0 100 ... 1700] [--spp 256]
tests/test_animation_gpu.py validates the CUDA path against them (validator.py's rule, MAE, image mean).
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refbind  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "oracle_frames")


def main():
    from PIL import Image
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, nargs="*", default=list(range(0, 1800, 100)))
    ap.add_argument("--spp", type=int, default=256)
    ap.add_argument("--force", action="store_true")
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    meta_path = os.path.join(OUT, "oracle_frames.json")
    meta = {"frames": {}}
    if os.path.exists(meta_path):
        with open(meta_path) as f:
            meta = json.load(f)
    o = refbind.get("fast")
    o.load_scene()
    c = o.config
    sha = os.path.join(refbind.REF_DIR, "data", "STANDINS.sha256")
    meta.update({"width": c.width, "height": c.height, "spp": args.spp, "oracle": "oracle/_ref/libptref.so (-O3 -ffast-math -fopenmp, x86-64-v3)",
                 "standins_sha256": hashlib.sha256(open(sha, "rb").read()).hexdigest() if os.path.exists(sha) else None})
    for f in args.frames:
        png = os.path.join(OUT, "frame_%04d.png" % f)
        if os.path.exists(png) and str(f) in meta["frames"] and not args.force:
            continue
        o.setup_frame(f)
        t0 = time.time()
        rgb, bgra = o.render_rect(0, 0, c.width, c.height, 0, args.spp, 1)
        dt = time.time() - t0
        Image.fromarray(np.ascontiguousarray(bgra[..., 2::-1])).save(png, optimize=True)
        # the reference itself returns NaN for a few paths of some frames (-ffast-math build): those pixels are
        # listed and left out of the image mean on both sides
        flat = rgb.reshape(-1, 3)
        bad = np.flatnonzero(~np.isfinite(flat).all(axis=1))
        meta["frames"][str(f)] = {"mean_linear_rgb": [float(x) for x in np.delete(flat, bad, axis=0).mean(axis=0, dtype=np.float64)],
                                  "seconds": round(dt, 1), "nonfinite_pixels": [int(i) for i in bad]}
        with open(meta_path, "w") as fh:
            json.dump(meta, fh, indent=1, sort_keys=True)
        print("frame %4d: %.1f s, mean %s" % (f, dt, meta["frames"][str(f)]["mean_linear_rgb"]), flush=True)


if __name__ == "__main__":
    main()
