#!/usr/bin/env python3
"""Developer tool: verbose GPU-vs-oracle comparison on a few frames (run under gpurun)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
from oracle import refbind  # noqa: E402


def cmp(name, g_rgb, g_bgra, r_rgb, r_bgra):
    mae = np.abs(g_bgra[..., :3].astype(float) - r_bgra[..., :3].astype(float)).mean()
    rel = abs(g_rgb.mean() - r_rgb.mean()) / max(r_rgb.mean(), 1e-12)
    d = np.abs(g_rgb - r_rgb) / np.maximum(np.abs(r_rgb), 1e-3)
    print("  %-28s MAE %.4f/255  mean-rel %.2e  frac(px rel>1e-3) %.3f  max8bit %d" % (
        name, mae, rel, (d.max(axis=-1) > 1e-3).mean(),
        np.abs(g_bgra[..., :3].astype(int) - r_bgra[..., :3].astype(int)).max()), flush=True)


def main():
    pkg = ge.load_package()
    frames = [int(a) for a in sys.argv[1:]] or [0, 330, 520, 1400]
    o = refbind.get("fast")
    o.load_scene()
    r = pkg.Renderer(pkg.Config.testing(), 0)
    v = o.setup_frame(frames[0])
    t = time.time()
    r.upload_static(**pkg.scene_io.static_from_view(v))
    print("upload_static %.2fs" % (time.time() - t), r.scene_stats())
    print("pcg4d", r.pcg4d([[0, 0, 0, 152121358]]), o.pcg4d((0, 0, 0, 152121358)))
    print("tonemap", r.tonemap([[0.18, 0.09, 0.045], [1, 0.5, 0.25], [16, 8, 4]]).tolist())
    x0, y0, w, h = 160, 90, 320, 180
    for f in frames:
        v = o.setup_frame(f)
        fr = pkg.scene_io.frame_from_view(v)
        print("frame", f)
        t = time.time()
        ref = o.render_rect(x0, y0, w, h, 0, 8, 32)
        print("  oracle %.2fs" % (time.time() - t))
        for mode, opts in (("links", {"traversal": 1}), ("wide/tiles", {"traversal": 0, "kernel": 1}),
                           ("wide/mega", {"traversal": 0, "kernel": 0})):
            for k, val in opts.items():
                r.set_option(k, val)
            r.set_frame(**fr)
            g = r.render_rect(x0, y0, w, h, 0, 8, 32)
            ms, _ = r.last_render_ms()
            cmp("%s (%.2f ms)" % (mode, ms), g[0], g[1], ref[0], ref[1])
        # single-sample agreement
        rng = np.random.RandomState(f)
        xy = np.stack([rng.randint(0, 640, 512), rng.randint(0, 360, 512)], 1).astype(np.uint32)
        si = rng.randint(0, 256, 512).astype(np.int32)
        ref_s = np.stack([o.trace_sample(int(a), int(b), int(c)) for (a, b), c in zip(xy, si)])
        for mode, tr in (("links", 1), ("wide", 0)):
            r.set_option("traversal", tr)
            r.set_frame(**fr)
            gs = r.trace_samples(xy, si)
            rel = np.abs(gs - ref_s).max(1) / np.maximum(np.abs(ref_s).max(1), 1e-4)
            print("  samples %-6s: frac rel<1e-4 %.3f  <1e-2 %.3f  worst %.3g" % (
                mode, (rel < 1e-4).mean(), (rel < 1e-2).mean(), rel.max()), flush=True)
        # full frame timing per mode
        for mode, opts in (("links", {"traversal": 1}), ("wide/tiles", {"traversal": 0, "kernel": 1}),
                           ("wide/mega", {"traversal": 0, "kernel": 0})):
            for k, val in opts.items():
                r.set_option(k, val)
            r.set_frame(**fr)
            r.render_async(); r.sync()
            r.render_async(); r.sync()
            ms, n = r.last_render_ms()
            print("  full frame %-10s %.2f ms  %.1f Mpaths/s" % (mode, ms, 640 * 360 * 256 / ms / 1e3), flush=True)
    r.set_option("traversal", 1); r.set_option("counters", 1)
    for f in frames:
        v = o.setup_frame(f)
        r.set_frame(**pkg.scene_io.frame_from_view(v))
        r.read_counters()
        r.render_async(); r.sync()
        c = r.read_counters()
        p = c["paths"]
        print("counters frame", f, {k: round(val / p, 3) for k, val in c.items()}, flush=True)


if __name__ == "__main__":
    main()
