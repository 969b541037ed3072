"""diagnostic: why do pt_gpu's BMPs differ from the Python host's for the same frame?"""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
from oracle import refbind
REF = os.path.join(ROOT, "oracle", "_ref")
outs = []
for tag, extra in (("a", []), ("b", []), ("serial", ["--serial"])):
    d = "/tmp/dropin_%s" % tag
    os.makedirs(d, exist_ok=True)
    r = subprocess.run([os.path.join(REF, "pt_gpu"), "--gpus", "1", "--frames", "518", "519", "--out", d] + extra, cwd=REF, capture_output=True, text=True)
    print(tag, r.returncode, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:])
    outs.append(np.fromfile(os.path.join(d, "frame_0518.bmp"), dtype=np.uint8))
print("pt_gpu run a vs b: %d bytes differ; a vs serial: %d" % ((outs[0] != outs[1]).sum(), (outs[0] != outs[2]).sum()))
pkg = ge.load_package()
o = refbind.get("fast"); o.load_scene()
view = o.setup_frame(518)
for flat in (1, 0):
    r = pkg.Renderer(pkg.Config.testing(), 0)
    r.set_option("flat", flat)
    r.upload_static(**pkg.scene_io.static_from_view(view))
    r.set_frame(**pkg.scene_io.frame_from_view(view))
    a = r.render_bmp(); b = r.render_bmp()
    d = (a.astype(int) - outs[0].astype(int))
    print("python fresh ctx flat=%d: two renders differ in %d bytes; vs pt_gpu: %d bytes differ, max |diff| %d, hist %s" % (
        flat, (a != b).sum(), (d != 0).sum(), np.abs(d).max(), np.bincount(np.abs(d).ravel())[:6]))
    for opt in ("sort", "dyn_first"):
        r.set_option(opt, 0)
        c = r.render_bmp()
        print("   with %s=0: vs default %d bytes differ" % (opt, (c != a).sum()))
        r.set_option(opt, 1)
    r.close()
