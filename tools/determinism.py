#!/usr/bin/env python3
"""Developer tool: render one snapshot frame several times and count differing pixels."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package(); sio = pkg.scene_io
r = pkg.Renderer(pkg.Config.testing(), 0)
r.upload_static(**sio.load_static(sio.static_path()))
for f in [int(a) for a in sys.argv[1:]] or [520]:
    fr = sio.load_frame(sio.frame_path(f))
    imgs = []
    for i in range(4):
        r.set_frame(fr["subframes"], fr["dyn_instances"], fr["tlas_nodes"], fr["tlas_links"])
        imgs.append(r.render().copy())
    for i in range(1, 4):
        d = (imgs[i] != imgs[0]).any(axis=-1)
        print("frame %d run %d vs 0: %d pixels differ, max diff %d" % (f, i, d.sum(), np.abs(imgs[i].astype(int) - imgs[0].astype(int)).max()), flush=True)
