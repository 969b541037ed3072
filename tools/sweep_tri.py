#!/usr/bin/env python3
"""Developer tool: frame times against tri_threshold (for builds with a deeper pending-triangle list)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package(); sio = pkg.scene_io
r = pkg.Renderer(pkg.Config.testing(), 0)
r.upload_static(**sio.load_static(sio.static_path()))
frames = {f: sio.load_frame(sio.frame_path(f)) for f in (0, 520, 1400)}
for tt in [int(a) for a in sys.argv[1:]] or [8]:
    r.set_option("tri_threshold", tt)
    out = []
    for f, fr in frames.items():
        r.set_frame(fr["subframes"], fr["dyn_instances"], fr["tlas_nodes"], fr["tlas_links"])
        best = 1e9
        for _ in range(3):
            r.render_async(); r.sync()
            best = min(best, r.last_render_ms()[0])
        out.append("%d: %.2f" % (f, best))
    print("tri_threshold %2d | %s" % (tt, " | ".join(out)), flush=True)
