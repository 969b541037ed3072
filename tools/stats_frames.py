"""Developer tool: lane / step census of the traversal kernel (a -DWF_STATS build of the library:
PTGPU_OUT=libptgpu_stats.so PTGPU_BUILD_DIR=build_stats PTGPU_NVCC_FLAGS=-DWF_STATS sh build.sh, loaded with
PTGPU_LIB=). With --validate every ray is also re-traced by the plain single-ray traversal, whose node and
triangle tests per ray are printed beside the scheduled kernel's: the price of postponing triangle tests."""
import os, sys
sys.path.insert(0, os.getcwd())
import __graft_entry__ as ge
pkg = ge.load_package(); sio = pkg.scene_io
cfg = pkg.Config.testing()
r = pkg.Renderer(cfg, 0)
opts = [a for a in sys.argv[1:] if "=" in a]
for kv in opts:
    k, v = kv.split("="); r.set_option(k, int(v))
r.upload_static(**sio.load_static(sio.static_path()))
if "--validate" in sys.argv:
    r.set_option("validate", 1)
frames = [int(a) for a in sys.argv[1:] if a.isdigit()] or [520, 0, 1400]
for f in frames:
    fr = sio.load_frame(sio.frame_path(f))
    r.set_frame(fr["subframes"], fr["dyn_instances"], fr["tlas_nodes"], fr["tlas_links"])
    sys.stderr.write("== frame %d\n" % f); sys.stderr.flush()
    r.render_async(); r.sync()
    print("frame %d %.2f ms" % (f, r.last_render_ms()[0]), flush=True)
