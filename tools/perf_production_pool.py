#!/usr/bin/env python3
"""Production-size frame (1920x1080, 1024 spp, 5 bounces) against the size of the path-state pool."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package(); sio = pkg.scene_io
static = sio.load_static(sio.static_path())
cfg = pkg.Config.production()
fr = sio.load_frame(sio.frame_path(520, "production"))
for budget_mb, lanes in ((16384, 256), (60000, 256), (120000, 256)):
    r = pkg.Renderer(cfg, 0)
    r.upload_static(**static)
    r.set_option("pool_budget_mb", budget_mb)
    r.set_option("lanes", lanes)
    r.set_frame(fr["subframes"], fr["dyn_instances"], fr["tlas_nodes"], fr["tlas_links"])
    r.render_async(); r.sync()
    ms, n = r.last_render_ms()
    paths = cfg.width * cfg.height * cfg.spp
    print("budget %6d MB: lanes %3d, rounds %3d, pool %.1f GB: %.1f ms, %.1f Mpaths/s, %d launches" % (
        budget_mb, r.get_stat("wave_lanes"), r.get_stat("wave_rounds"), r.get_stat("pool_bytes") / 1e9, ms, paths / ms / 1e3, n), flush=True)
    r.close()
