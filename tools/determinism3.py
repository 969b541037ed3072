#!/usr/bin/env python3
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package(); sio = pkg.scene_io
r = pkg.Renderer(pkg.Config.testing(), 0)
r.upload_static(**sio.load_static(sio.static_path()))
f = 1000
fr = sio.load_frame(sio.frame_path(f))
configs = [("default", {}), ("tri1", {"tri_threshold": 1}), ("xform1", {"xform_threshold": 1}), ("node1", {"node_threshold": 1}),
           ("tri1 xform1", {"tri_threshold": 1, "xform_threshold": 1}), ("tri1 node1", {"tri_threshold": 1, "node_threshold": 1}), ("burst1", {"node_burst": 1})]
for name, opts in configs:
    for k, v in {"kernel": 2, "bvh": 1, "tri_threshold": 8, "xform_threshold": 4, "node_threshold": 12, "node_burst": 2, "lanes": 256, **opts}.items():
        r.set_option(k, v)
    r.set_frame(fr["subframes"], fr["dyn_instances"], fr["tlas_nodes"], fr["tlas_links"])
    x0, y0, w, h = 96, 240, 96, 64
    a, _ = r.render_rect(x0, y0, w, h, 0, 256, 1, tonemap=False)
    nd = []
    for run in range(5):
        b, _ = r.render_rect(x0, y0, w, h, 0, 256, 1, tonemap=False)
        nd.append(int((a != b).any(axis=-1).sum()))
    print("%-16s differing pixels over 5 reruns of a %dx%d window: %s" % (name, w, h, nd), flush=True)
