#!/usr/bin/env python3
"""Per-path event counts on the REFERENCE's own BVH (links traversal with counters) for every
snapshot frame, and the algorithmic flops per path F of SURVEY.md 8(d):
F = 25 N_node + 56 N_tri + 61 N_blas + 9 N_ray + 1400 N_sky + 210 N_att + 445 N_bounce + 115 N_hit + 15 N_miss + 90
Writes profiles/flops_per_path.json (run under gpurun; the counts are exact integers, the kernel
visits exactly the nodes ray_query.hh:184-223 visits)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def flops(c):
    p = c["paths"]
    n = {k: v / p for k, v in c.items()}
    return (25 * n["node_visits"] + 56 * n["tri_tests"] + 61 * n["blas_enters"] + 9 * n["rays"] + 1400 * n["sky_marches"]
            + 210 * n["sky_attenuations"] + 445 * n["bounces"] + 115 * n["hits"] + 15 * n["misses"] + 90), n


def main():
    pkg = ge.load_package()
    sio = pkg.scene_io
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "flops_per_path.json")
    r = pkg.Renderer(pkg.Config.testing(), 0)
    r.upload_static(**sio.load_static(sio.static_path()))
    r.set_option("traversal", 1)
    r.set_option("counters", 1)
    res = {"formula": "F = 25 N_node + 56 N_tri + 61 N_blas + 9 N_ray + 1400 N_sky + 210 N_att + 445 N_bounce + 115 N_hit + 15 N_miss + 90",
           "config": "640x360x256spp, 4 bounces, stand-in terrain/pine/bunny", "frames": {}, "events_per_path": {}}
    for f in sio.available_frames():
        fr = sio.load_frame(sio.frame_path(f))
        r.set_frame(fr["subframes"], fr["dyn_instances"], fr["tlas_nodes"], fr["tlas_links"])
        r.read_counters()
        r.render_async()
        r.sync()
        c = r.read_counters()
        F, n = flops(c)
        res["frames"][str(f)] = round(F, 1)
        res["events_per_path"][str(f)] = {k: round(v, 4) for k, v in n.items() if k != "paths"}
        print(f, round(F, 1), {k: round(v, 3) for k, v in n.items()}, flush=True)
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as fh:
        json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
