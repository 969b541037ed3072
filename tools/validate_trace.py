#!/usr/bin/env python3
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package(); sio = pkg.scene_io
r = pkg.Renderer(pkg.Config.testing(), 0)
r.upload_static(**sio.load_static(sio.static_path()))
r.set_option("validate", 1)
for f in [int(a) for a in sys.argv[1:]] or [1000]:
    fr = sio.load_frame(sio.frame_path(f))
    r.set_frame(fr["subframes"], fr["dyn_instances"], fr["tlas_nodes"], fr["tlas_links"])
    r.render_rect(96, 240, 96, 64, 0, 256, 1, tonemap=False)
    r.render_rect(96, 240, 96, 64, 0, 256, 1, tonemap=False)
