"""The drop-in itself: the reference's host program with baseline_render swapped for the C ABI
(host/ptgpu_main.cc compiled together with the reference's scene.cc / bvh.cc / mesh.cc into oracle/_ref/pt_gpu,
see INTEGRATION.md), and the file bytes it writes.

  * pt_gpu renders frames with the reference's own load_scene() / setup_animation_frame() and writes
    output/frame_NNNN.bmp: the files equal, byte for byte, what the Python host gets from ptgpu_render_bmp for
    the same frames (same library, same inputs, deterministic kernels);
  * ptgpu_render_bmp's bytes equal the reference's write_bmp (bmp.cc:7-63) of the same BGRA frame: header,
    bottom-up rows, B,G,R order and row padding — checked on a width whose rows need padding, too.
"""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
PT_GPU = os.path.join(REF_DIR, "pt_gpu")


def same_as_write_bmp(ref, bmp, w, h):
    """Every byte write_bmp DEFINES must be equal. It never writes the reserved header bytes 6..9 nor the row
    padding (its buffer is `new uint8_t[]`, bmp.cc:17: indeterminate); the library writes zeros there."""
    assert bmp[6:10].tolist() == [0, 0, 0, 0]
    defined = np.ones(54, bool)
    defined[6:10] = False
    assert np.array_equal(ref[:54][defined], bmp[:54][defined]), (ref[:54].tolist(), bmp[:54].tolist())
    pitch = (w * 3 + 3) // 4 * 4
    a, b = ref[54:].reshape(h, pitch)[:, :w * 3], bmp[54:].reshape(h, pitch)[:, :w * 3]
    assert np.array_equal(a, b)


def dumped(d, name, dtype, cols):
    return np.fromfile(os.path.join(d, name), dtype=dtype).reshape(-1, cols)


def test_pt_gpu_writes_the_frames_the_library_renders(pkg, frames, oracle, tmp_path):
    """pt_gpu = the reference's host program (load_scene, setup_animation_frame, the frame loop of main.cc:78-102)
    with baseline_render swapped for the C ABI. Its two builds of the reference's host code (this binary and the
    oracle's shared library, -fPIC) round a few transforms differently under -ffast-math, so the comparison
    feeds pt_gpu's OWN input arrays (--dump) through the Python host: same library, same inputs -> same bytes."""
    if not os.path.exists(PT_GPU):
        pytest.skip("oracle/_ref/pt_gpu not built (make -C oracle, needs /root/reference)")
    out, dump = tmp_path / "output", tmp_path / "dump"
    out.mkdir()
    dump.mkdir()
    run = subprocess.run([PT_GPU, "--gpus", "1", "--frames", "518", "520", "--out", str(out), "--dump", str(dump)],
                         cwd=REF_DIR, capture_output=True, text=True, timeout=900)
    assert run.returncode == 0, run.stderr[-2000:]
    assert "RENDERED 2 FRAMES" in run.stdout, run.stdout[-2000:]
    assert sorted(os.listdir(out)) == ["frame_0518.bmp", "frame_0519.bmp"]
    d = str(dump)
    r = pkg.Renderer(pkg.Config.testing(), device=0)
    try:
        r.upload_static(nodes=dumped(d, "nodes.bin", np.float32, 6), links=dumped(d, "links.bin", np.uint32, 2),
                        indices=np.fromfile(os.path.join(d, "indices.bin"), dtype=np.uint32),
                        pos=dumped(d, "pos.bin", np.float32, 4), normal=dumped(d, "normal.bin", np.float32, 4),
                        albedo=dumped(d, "albedo.bin", np.float32, 4), material=dumped(d, "material.bin", np.float32, 4),
                        instances=dumped(d, "instances.bin", np.uint8, 160))
        for f in (518, 519):
            pre = "frame_%04d_" % f
            r.set_frame(subframes=dumped(d, pre + "subframes.bin", np.uint8, 160), dyn_instances=dumped(d, pre + "dyn_instances.bin", np.uint8, 160),
                        tlas_nodes=dumped(d, pre + "tlas_nodes.bin", np.float32, 6), tlas_links=dumped(d, pre + "tlas_links.bin", np.uint32, 2))
            mine = r.render_bmp()
            theirs = np.fromfile(out / ("frame_%04d.bmp" % f), dtype=np.uint8)
            assert theirs.size == 54 + 640 * 3 * 360
            assert np.array_equal(mine, theirs), "frame %d: %d bytes differ" % (f, int((mine != theirs).sum()))
            # and the oracle's own build of the same host code gives the same picture up to those roundings
            other = frames.use(f).render_bmp()
            assert np.abs(other[54:].astype(np.int32) - theirs[54:].astype(np.int32)).mean() < 0.5
    finally:
        r.close()


def test_bad_command_lines_are_rejected():
    if not os.path.exists(PT_GPU):
        pytest.skip("oracle/_ref/pt_gpu not built")
    for args in (["--gpus", "0"], ["--gpus", "-2"], ["--frames", "5", "3"], ["--frames", "-1", "4"], ["--step", "0"], ["--step", "x"]):
        run = subprocess.run([PT_GPU] + args, cwd=REF_DIR, capture_output=True, text=True, timeout=60)
        assert run.returncode == 2 and "usage" in run.stderr, (args, run.returncode)


@pytest.mark.parametrize("frame", [0, 520])
def test_bmp_bytes_equal_write_bmp(frames, oracle, tmp_path, frame):
    """bmp.cc:7-63 on the BGRA frame the library returns vs the BMP the library packs on the device"""
    r = frames.use(frame)
    bmp = r.render_bmp()
    bgra = r.fetch_bgra()
    path = str(tmp_path / "ref.bmp")
    oracle.write_bmp(path, bgra)
    ref = np.fromfile(path, dtype=np.uint8)
    assert ref.size == bmp.size
    same_as_write_bmp(ref, bmp, 640, 360)


def test_bmp_row_padding_matches_write_bmp(pkg, oracle, tmp_path):
    """a 638-pixel-wide frame: rows of 1914 bytes padded to 1916 (bmp.cc:15)"""
    cfg = pkg.Config.testing()
    cfg.width, cfg.height, cfg.spp = 638, 8, 8
    view = oracle.setup_frame(520)
    r = pkg.Renderer(cfg, device=0)
    try:
        r.set_option("flat", 0)    # a second full flat build is not what this test is about
        r.upload_static(**pkg.scene_io.static_from_view(view))
        r.set_frame(**pkg.scene_io.frame_from_view(view))
        bmp = r.render_bmp()
        bgra = r.fetch_bgra()
    finally:
        r.close()
    assert bmp.size == 54 + 1916 * 8
    path = str(tmp_path / "ref.bmp")
    oracle.write_bmp(path, bgra)
    ref = np.fromfile(path, dtype=np.uint8)
    same_as_write_bmp(ref, bmp, 638, 8)
    pad = bmp[54:].reshape(8, 1916)[:, 1914:]
    assert (pad == 0).all()                                     # the library's padding bytes are defined (zero)
