#!/usr/bin/env python3
"""Developer tool: run frame 520 at 64 spp with several scheduling parameters (use with a WF_STATS build)."""
import os, sys, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package(); sio = pkg.scene_io
cfg = pkg.Config.testing(); cfg.spp = int(os.environ.get("SPP", "64"))
r = pkg.Renderer(cfg, 0)
r.upload_static(**sio.load_static(sio.static_path()))
fr = sio.load_frame(sio.frame_path(520))
r.set_frame(fr["subframes"], fr["dyn_instances"], fr["tlas_nodes"], fr["tlas_links"])
r.render_async(); r.sync()
for ma, nt, nb, tt, xt in [(8,12,2,8,4),(4,12,2,8,4),(2,12,2,8,4),(1,12,2,8,4),(4,12,2,12,4),(4,12,2,16,6),(4,8,2,8,4),(4,16,4,8,4)]:
    for k, v in (("min_active", ma), ("node_threshold", nt), ("node_burst", nb), ("tri_threshold", tt), ("xform_threshold", xt)):
        r.set_option(k, v)
    sys.stderr.write("== r%d n%d b%d t%d x%d\n" % (ma, nt, nb, tt, xt)); sys.stderr.flush()
    r.render_async(); r.sync()
    print("r%d n%d b%d t%d x%d: %.2f ms" % (ma, nt, nb, tt, xt, r.last_render_ms()[0]), flush=True)
