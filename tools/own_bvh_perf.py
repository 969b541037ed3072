#!/usr/bin/env python3
"""Frame times with the BLASes recovered from the reference's arrays (ptgpu_upload_static) against the
BLASes built by this library from the triangles (ptgpu_upload_meshes); no oracle needed."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package(); sio = pkg.scene_io
cfg = pkg.Config.testing()
st = sio.load_static(sio.static_path())
an = pkg.Animation(cfg)
meshes = sio.mesh_table(st["instances"], an.mesh_rows)
out = {}
for name in ("reference BLAS", "own BLAS"):
    r = pkg.Renderer(cfg, 0)
    t0 = time.perf_counter()
    if name == "own BLAS":
        r.upload_meshes(st["indices"], st["pos"], st["normal"], st["albedo"], st["material"], meshes, st["instances"])
    else:
        r.upload_static(**st)
    up = time.perf_counter() - t0
    line = []
    for f in (0, 520, 1400):
        an.set_frame(r, f)
        r.render_async(); r.sync()
        best = 1e9
        for _ in range(2):
            r.render_async(); r.sync()
            best = min(best, r.last_render_ms()[0])
        line.append("%d: %.2f ms" % (f, best))
        out[(name, f)] = r.fetch_bgra().copy()
    print("%-14s upload+build %.2f s | %s" % (name, up, " | ".join(line)), flush=True)
    r.close()
for f in (0, 520, 1400):
    a, b = out[("reference BLAS", f)], out[("own BLAS", f)]
    print("frame %d: identical pixels %.6f" % (f, (a == b).all(-1).mean()))
