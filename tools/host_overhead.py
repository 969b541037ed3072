#!/usr/bin/env python3
"""Developer tool: wall time of each host-side call of bench.py's device-resident loop."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
if "--torch" in sys.argv:
    import torch
    torch.cuda.set_device(0)
    torch.cuda.synchronize()
sampler = None
if "--sampler" in sys.argv:
    import bench
    sampler = bench.ClockSampler(0)
pkg = ge.load_package(); sio = pkg.scene_io
cfg = pkg.Config.testing()
r = pkg.Renderer(cfg, 0)
r.upload_static(**sio.load_static(sio.static_path()))
an = pkg.Animation(cfg)
acc = {}
if sampler: sampler.start()
t_all = time.perf_counter()
def timed(name, fn):
    t = time.perf_counter(); out = fn(); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t; return out
for i in range(12):
    f = [0, 520, 1400, 100][i % 4]
    timed("anim.frame (host replay)", lambda: an.frame(f))
    timed("anim.set_frame", lambda: an.set_frame(r, f))
    timed("render_async", lambda: r.render_async())
    ms, n = timed("last_render_ms (wait)", lambda: r.last_render_ms())
    acc["device ms"] = acc.get("device ms", 0.0) + ms / 1e3
    timed("get_stat x3", lambda: (r.get_stat("trace_us"), r.get_stat("trace_launches"), r.get_stat("shade_us")))
print("wall per step %.3f ms (%s)" % (1e3 * (time.perf_counter() - t_all) / 12, " ".join(sys.argv[1:]) or "plain"))
for k, v in acc.items():
    print("%-28s %.3f ms per step" % (k, 1e3 * v / 12))
