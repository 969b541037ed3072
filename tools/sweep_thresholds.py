#!/usr/bin/env python3
"""Developer tool: frame times against the scheduling thresholds of wf_trace_cw (options of ptgpu_set_option)."""
import os, sys, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg = ge.load_package(); sio = pkg.scene_io
r = pkg.Renderer(pkg.Config.testing(), 0)
r.upload_static(**sio.load_static(sio.static_path()))
frames = {f: sio.load_frame(sio.frame_path(f)) for f in (0, 520, 1400)}
base = dict(node_threshold=16, node_burst=2, tri_threshold=8, xform_threshold=4, min_active=8)
variants = [dict()] + [dict([kv]) for kv in (("node_threshold", 12), ("node_threshold", 20), ("node_threshold", 24), ("node_burst", 1), ("node_burst", 3), ("node_burst", 4),
            ("tri_threshold", 6), ("tri_threshold", 10), ("tri_threshold", 12), ("xform_threshold", 3), ("xform_threshold", 6), ("xform_threshold", 8),
            ("min_active", 6), ("min_active", 12))]
for v in variants:
    opts = dict(base); opts.update(v)
    for k, val in opts.items():
        r.set_option(k, val)
    out = []
    for f, fr in frames.items():
        r.set_frame(fr["subframes"], fr["dyn_instances"], fr["tlas_nodes"], fr["tlas_links"])
        best = 1e9
        for _ in range(3):
            r.render_async(); r.sync()
            best = min(best, r.last_render_ms()[0])
        out.append("%d: %.2f" % (f, best))
    print("%-24s | %s" % (v or "base", " | ".join(out)), flush=True)
