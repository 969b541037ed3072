#!/usr/bin/env python3
"""Full-frame timings of the other BASELINE configs from snapshots (no oracle): motion-blur frames at
1024 spp (configs[4]) and one production-size frame (1920x1080, 1024 spp, 5 bounces)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package(); sio = pkg.scene_io
static = sio.load_static(sio.static_path())
for tag, cfg, frames in (("motionblur", None, [372, 376, 1100]), ("production", pkg.Config.production(), [520])):
    if cfg is None:
        cfg = pkg.Config.testing(); cfg.spp = 1024
    r = pkg.Renderer(cfg, 0)
    r.upload_static(**static)
    for f in frames:
        fr = sio.load_frame(sio.frame_path(f, tag))
        r.set_frame(fr["subframes"], fr["dyn_instances"], fr["tlas_nodes"], fr["tlas_links"])
        r.render_async(); r.sync()
        r.render_async(); r.sync()
        ms, n = r.last_render_ms()
        paths = cfg.width * cfg.height * cfg.spp
        print("%s frame %d: %dx%d x %d spp, %d bounces: %.1f ms, %.1f Mpaths/s, %d launches, lanes %d, rounds %d, pool %.1f GB" % (
            tag, f, cfg.width, cfg.height, cfg.spp, cfg.max_bounces, ms, paths / ms / 1e3, n,
            r.get_stat("wave_lanes"), r.get_stat("wave_rounds"), r.get_stat("pool_bytes") / 1e9), flush=True)
    r.close()
