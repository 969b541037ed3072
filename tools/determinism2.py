#!/usr/bin/env python3
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package(); sio = pkg.scene_io
r = pkg.Renderer(pkg.Config.testing(), 0)
r.upload_static(**sio.load_static(sio.static_path()))
f = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
fr = sio.load_frame(sio.frame_path(f))
r.set_frame(fr["subframes"], fr["dyn_instances"], fr["tlas_nodes"], fr["tlas_links"])
a, _ = r.render_rect(0, 0, 640, 360, 0, 256, 1, tonemap=False)
for run in range(3):
    b, _ = r.render_rect(0, 0, 640, 360, 0, 256, 1, tonemap=False)
    d = np.argwhere((a != b).any(axis=-1))
    print("run", run, "differing pixels:", len(d))
    for (y, x) in d[:6]:
        print("  px", x, y, a[y, x], b[y, x], "rel", np.abs(a[y, x] - b[y, x]).max() / max(np.abs(a[y, x]).max(), 1e-9))
        # which sample differs?
        for s in range(256):
            s1 = r.trace_samples([[x, y]], [s])[0]
        # per-sample via 1-sample rects, twice
        v1 = np.stack([r.render_rect(int(x), int(y), 1, 1, s, 1, 1, tonemap=False)[0][0, 0] for s in range(256)])
        v2 = np.stack([r.render_rect(int(x), int(y), 1, 1, s, 1, 1, tonemap=False)[0][0, 0] for s in range(256)])
        bad = np.argwhere((v1 != v2).any(axis=-1)).ravel()
        print("   per-sample rerun differences:", bad[:10], "sum1", v1.sum(0), "sum2", v2.sum(0))
