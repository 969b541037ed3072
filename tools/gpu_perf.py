#!/usr/bin/env python3
"""Developer tool: full-frame timings of kernel variants on snapshot frames + parity vs oracle window."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    pkg = ge.load_package()
    sio = pkg.scene_io
    frames = [int(a) for a in sys.argv[1:]] or [0, 330, 520, 1400]
    cfg = pkg.Config.testing()
    r = pkg.Renderer(cfg, 0)
    r.upload_static(**sio.load_static(sio.static_path()))
    variants = [("default", {"kernel": 2})]
    try:
        if os.environ.get("PTGPU_NO_ORACLE"): raise RuntimeError()
        from oracle import refbind
        o = refbind.get("fast"); o.load_scene()
    except Exception as e:  # noqa
        o = None
    for f in frames:
        fr = sio.load_frame(sio.frame_path(f))
        r.set_frame(fr["subframes"], fr["dyn_instances"], fr["tlas_nodes"], fr["tlas_links"])
        ref = None
        if o is not None:
            o.setup_frame(f)
            ref = o.render_rect(160, 90, 320, 180, 0, 8, 32)
        for name, opts in variants:
            for k, v in opts.items():
                r.set_option(k, v)
            r.render_async(); r.sync()
            best = 1e9
            for _ in range(2):
                r.render_async(); r.sync()
                best = min(best, r.last_render_ms()[0])
            msg = "frame %4d %-8s %8.2f ms %7.1f Mpaths/s" % (f, name, best, 640 * 360 * 256 / best / 1e3)
            if ref is not None:
                g = r.render_rect(160, 90, 320, 180, 0, 8, 32)
                mae = np.abs(g[1][..., :3].astype(float) - ref[1][..., :3].astype(float)).mean()
                rel = abs(g[0].mean() - ref[0].mean()) / ref[0].mean()
                msg += "  | window MAE %.4f mean-rel %.2e" % (mae, rel)
            print(msg, flush=True)


if __name__ == "__main__":
    main()
