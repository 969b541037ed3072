#!/usr/bin/env python3
"""Developer tool: frame times against the queue-fetch threshold (option min_active), + validation."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package(); sio = pkg.scene_io
r = pkg.Renderer(pkg.Config.testing(), 0)
r.upload_static(**sio.load_static(sio.static_path()))
frames = {f: sio.load_frame(sio.frame_path(f)) for f in (0, 520, 1400)}
for ma in [int(a) for a in sys.argv[1:]] or [8]:
    r.set_option("min_active", ma)
    out = []
    for f, fr in frames.items():
        r.set_frame(fr["subframes"], fr["dyn_instances"], fr["tlas_nodes"], fr["tlas_links"])
        r.render_async(); r.sync()
        best = 1e9
        for _ in range(2):
            r.render_async(); r.sync()
            best = min(best, r.last_render_ms()[0])
        out.append("%d: %.2f ms" % (f, best))
    print("min_active %2d | %s" % (ma, " | ".join(out)), flush=True)
r.set_option("min_active", 8)
r.set_option("validate", 1)
for f in (520, 1000):
    fr = sio.load_frame(sio.frame_path(f))
    r.set_frame(fr["subframes"], fr["dyn_instances"], fr["tlas_nodes"], fr["tlas_links"])
    r.render_rect(96, 240, 96, 64, 0, 256, 1, tonemap=False)
    print("frame %d validate mismatches %d" % (f, r.get_stat("validate_mismatches")), flush=True)
