import os, sys
sys.path.insert(0, os.getcwd())
import __graft_entry__ as ge
pkg = ge.load_package(); sio = pkg.scene_io
cfg = pkg.Config.testing()
r = pkg.Renderer(cfg, 0)
r.upload_static(**sio.load_static(sio.static_path()))
for f in (520, 0, 1400):
    fr = sio.load_frame(sio.frame_path(f))
    r.set_frame(fr["subframes"], fr["dyn_instances"], fr["tlas_nodes"], fr["tlas_links"])
    sys.stderr.write("== frame %d\n" % f); sys.stderr.flush()
    r.render_async(); r.sync()
    print("frame %d %.2f ms" % (f, r.last_render_ms()[0]), flush=True)
