"""GPU parity tests: the CUDA path, called through the C ABI (include/ptgpu.h), against the oracle
(the unmodified reference compiled into oracle/_ref) on the same inputs.

Bars (BASELINE.json north_star):
  * integer work (pcg4d) bit-exact;
  * tonemap_pixel within 1 LSB;
  * traversal: same triangle as the reference ray query;
  * images: tonemapped MAE <= 1/255 per channel, image-mean linear radiance within 1e-3 relative
    (at enough samples), zero BAD frames under the restated validator (PSNR >= 32 dB after 2x
    downscale). Per-pixel linear radiance is NOT bit-comparable even CPU-vs-CPU (the oracle's own
    -ffast-math and strict builds differ on 18 % of pixels), so the oracle's
    fast-vs-strict distance is measured and printed beside the GPU-vs-oracle distance.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUDENT_ID = 152121358
# fraction of single samples that must agree with path_trace_pixel to 1e-3 (measured 0.981-0.999: the bar allows
# twice the measured misses; one rounding difference can flip a lobe choice and change a whole path)
BAR_SAMPLES = 0.96


def mae255(a_bgra, b_bgra):
    return np.abs(a_bgra[..., :3].astype(np.float64) - b_bgra[..., :3].astype(np.float64)).mean()


def mean_rel(a_rgb, b_rgb):
    return abs(float(a_rgb.mean(dtype=np.float64)) - float(b_rgb.mean(dtype=np.float64))) / max(float(b_rgb.mean(dtype=np.float64)), 1e-12)


# ---- integer and per-pixel kernels -------------------------------------------------------------------

def test_pcg4d_bit_exact(renderer, oracle):
    assert renderer.pcg4d([[0, 0, 0, STUDENT_ID]]).tolist() == [[3346572545, 3185534624, 3185534624, 1847258501]]
    assert renderer.pcg4d([[320, 180, 255, STUDENT_ID]]).tolist() == [[2700474327, 2096636365, 3411749299, 4208948998]]
    rng = np.random.RandomState(3)
    states = rng.randint(0, 2 ** 32, size=(4096, 4), dtype=np.uint64).astype(np.uint32)
    for steps in (1, 2, 7):
        got = renderer.pcg4d(states, steps)
        for i in range(0, 4096, 257):   # the oracle is called per state; check a spread subset
            s = tuple(int(x) for x in states[i])
            for _ in range(steps):
                s = oracle.pcg4d(s)
            assert tuple(int(x) for x in got[i]) == s
    # the empty input is a no-op
    assert renderer.pcg4d(np.zeros((0, 4), np.uint32)).shape == (0, 4)


def test_tonemap_known_answers_and_random(renderer, oracle):
    kat = [((0.18, 0.09, 0.045), (54, 92, 141, 255)), ((1, 0.5, 0.25), (165, 206, 232, 255)),
           ((0.01, 0.005, 0.0025), (2, 5, 12, 255)), ((16, 8, 4), (252, 255, 255, 255)), ((0, 0, 0), (0, 0, 0, 255))]
    got = renderer.tonemap([k for k, _ in kat])
    for g, (_, want) in zip(got, kat):
        assert np.abs(g.astype(int) - np.array(want)).max() <= 1, (g, want)
    rng = np.random.RandomState(5)
    rgb = np.exp(rng.uniform(np.log(1e-5), np.log(50.0), size=(2000, 3))).astype(np.float32)
    got = renderer.tonemap(rgb)
    want = np.array([oracle.tonemap(c) for c in rgb])
    d = np.abs(got.astype(int) - want.astype(int))
    assert d.max() <= 1                      # at most one LSB (powf vs the reference's double pow)
    assert (d > 0).mean() < 0.01
    assert (got[:, 3] == 255).all()


# ---- traversal ---------------------------------------------------------------------------------------

def camera_like_rays(n, seed):
    rng = np.random.RandomState(seed)
    o = np.stack([rng.uniform(-90, 90, n), rng.uniform(15, 70, n), rng.uniform(-90, 90, n)], 1)
    d = rng.normal(size=(n, 3))
    d[:, 1] = -np.abs(d[:, 1]) * 0.7
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3], rays[:, 3], rays[:, 4:7], rays[:, 7] = o, 0.0, d, 1e9
    return rays


@pytest.mark.parametrize("frame", [520, 1400])
def test_closest_hit_matches_reference_ray_query(frames, oracle, frame):
    """ray_query_initialize/proceed/confirm (ray_query.hh:111-290) vs both device traversals."""
    r = frames.use(frame)
    rays = camera_like_rays(600, frame)
    ref = [oracle.trace_closest(x[0:3], x[4:7], float(x[3]), float(x[7]), 0) for x in rays]
    ref_t = np.array([h["thit"] for h in ref], np.float32)
    assert (ref_t > 0).mean() > 0.5                       # the rays do hit the scene
    for mode in (1, 0):                                   # 1: reference link tables, 0: wide BVH
        r.set_option("traversal", mode)
        frames.current = None
        r = frames.use(frame)
        f, u = r.trace_closest(rays, 0)
        hit_ref, hit_gpu = ref_t > 0, f[:, 0] > 0
        assert (hit_ref == hit_gpu).mean() > 0.995
        both = hit_ref & hit_gpu
        same_tri = np.array([u[i, 0] == ref[i]["instance"] and u[i, 1] == ref[i]["primitive"] for i in range(len(ref))])
        assert same_tri[both].mean() > 0.995, mode
        sel = both & same_tri
        np.testing.assert_allclose(f[sel, 0], ref_t[sel], rtol=2e-4)
        bary = np.array([h["bary"] for h in ref], np.float32)
        # the flat static scene stores world-space vertices (coordinates up to 100: ulp 8e-6) where the
        # reference tests in instance space (a tree is a few units across): on the smallest leaf-card
        # triangles that is a few 1e-3 of a barycentric coordinate; the hit point itself is origin + t * dir
        db = np.abs(f[sel, 1:4] - bary[sel]).max(axis=1)
        assert np.percentile(db, 99) < 5e-3 and db.max() < 3e-2, (np.percentile(db, 99), db.max())
        bf = np.array([h["back_face"] for h in ref])
        assert (u[sel, 2].astype(bool) == bf[sel]).all()
    r.set_option("traversal", 0)
    frames.current = None


# ---- single samples -------------------------------------------------------------------------------------

@pytest.mark.parametrize("frame", [0, 330, 1400])
def test_path_trace_pixel_samples(frames, oracle, frame):
    """path_trace_pixel(xy, sample_index, ...) (path_tracer.hh:637) sample by sample. One rounding
    difference can flip a lobe choice and change a whole path, so the bar is: the large majority of
    samples agree to 1e-3, and the mean over the set agrees closely."""
    r = frames.use(frame)
    rng = np.random.RandomState(frame + 1)
    n = 1500
    xy = np.stack([rng.randint(0, 640, n), rng.randint(0, 360, n)], 1).astype(np.uint32)
    si = rng.randint(0, 256, n).astype(np.int32)
    ref = np.stack([oracle.trace_sample(int(a), int(b), int(c)) for (a, b), c in zip(xy, si)])
    got = r.trace_samples(xy, si)
    rel = np.abs(got - ref).max(1) / np.maximum(np.abs(ref).max(1), 1e-4)
    print("frame %d: %.4f of %d samples within 1e-3 of path_trace_pixel" % (frame, (rel < 1e-3).mean(), n))
    assert (rel < 1e-3).mean() > BAR_SAMPLES, (rel < 1e-3).mean()
    assert np.isfinite(got).all()
    # the three kernels run the same device functions; nvcc contracts FMAs per kernel, so they agree
    # to rounding (and exactly on most pixels), not bit for bit
    rgb_a, _ = r.render_rect(300, 170, 16, 8, 0, 8, 32)
    r.set_option("kernel", 0)
    rgb_b, _ = r.render_rect(300, 170, 16, 8, 0, 8, 32)
    r.set_option("kernel", 1)
    rgb_c, _ = r.render_rect(300, 170, 16, 8, 0, 8, 32)
    r.set_option("kernel", 2)
    # the tile kernel and the wavefront kernel find the same hits (the scheduled traversal is validated ray by
    # ray against the plain one the tile kernel uses): they differ by FMA contraction in the shading code only
    close = np.isclose(rgb_a, rgb_c, rtol=1e-3, atol=1e-7).all(axis=-1)
    assert close.mean() >= 0.97, close.mean()
    # the megakernel walks the instanced 4-wide BVH, not the flat world-space scene: the same hits, with t and
    # barycentrics equal to rounding only, so more of its paths take another turn somewhere
    close = np.isclose(rgb_a, rgb_b, rtol=1e-3, atol=1e-7).all(axis=-1)
    assert close.mean() >= 0.75, close.mean()


def test_sample_index_contract(frames, oracle):
    """README.md:52-57: RNG key (x, y, (uint)sample_index, STUDENT_ID), subframe sample_index/8,
    negative indices use subframe 0."""
    r = frames.use(375)
    xy = np.array([[320, 180]] * 4, np.uint32)
    si = np.array([17, 17, -3, 255], np.int32)
    got = r.trace_samples(xy, si)
    assert np.array_equal(got[0], got[1])
    for i in (0, 2, 3):
        ref = oracle.trace_sample(320, 180, int(si[i]))
        assert np.abs(got[i] - ref).max() <= 1e-3 * max(np.abs(ref).max(), 1e-3)


# ---- images -----------------------------------------------------------------------------------------------

@pytest.mark.parametrize("frame", [0, 330, 520, 1000, 1400, 1750])
def test_frame_window_parity(frames, oracle, oracle_strict, frame):
    """A 320x180 window, 16 samples spread over all motion-blur subframes (sample stride 16)."""
    r = frames.use(frame)
    x0, y0, w, h, n, stride = 160, 90, 320, 180, 16, 16
    g_rgb, g_bgra = r.render_rect(x0, y0, w, h, 0, n, stride)
    o_rgb, o_bgra = oracle.render_rect(x0, y0, w, h, 0, n, stride)
    oracle_strict.setup_frame(frame)
    s_rgb, s_bgra = oracle_strict.render_rect(x0, y0, w, h, 0, n, stride)
    gpu_mae, self_mae = mae255(g_bgra, o_bgra), mae255(s_bgra, o_bgra)
    gpu_rel, self_rel = mean_rel(g_rgb, o_rgb), mean_rel(s_rgb, o_rgb)
    print("frame %d: GPU-vs-oracle MAE %.4f/255, mean-rel %.2e | oracle fast-vs-strict MAE %.4f/255, mean-rel %.2e"
          % (frame, gpu_mae, gpu_rel, self_mae, self_rel))
    assert gpu_mae <= 1.0, gpu_mae                         # <= 1/255 per channel
    assert gpu_mae <= max(2.0 * self_mae, 0.3)             # and no worse than the oracle's own noise floor
    assert gpu_rel <= max(1e-3, 3.0 * self_rel), gpu_rel   # image-mean linear radiance
    from oracle import validator_np as V
    ref_png = V.make_reference_png_array(o_bgra[..., 2::-1])
    psnr, good = V.validate_frame(ref_png, g_bgra[..., 2::-1])
    assert good, psnr


def test_golden_frame_0(frames, oracle):
    """The reference's shipped output/frame_0000.bmp (frame 0 at the full 256 spp): the whole frame
    through ptgpu_render + ptgpu_render_bmp."""
    from helpers import GOLDEN, read_bmp_rgb
    gold = read_bmp_rgb(GOLDEN)
    r = frames.use(0)
    bgra = r.render()
    rgb = bgra[..., 2::-1]
    mae = np.abs(rgb.astype(np.float64) - gold.astype(np.float64)).mean()
    mse = ((rgb.astype(np.float64) - gold.astype(np.float64)) ** 2).mean()
    psnr = 10 * np.log10(255.0 ** 2 / mse)
    print("golden frame 0: MAE %.4f/255, PSNR %.1f dB" % (mae, psnr))
    assert mae <= 0.06 and psnr >= 55.0          # measured 0.026 and 62.6 dB
    assert (bgra[..., 3] == 255).all()
    # fused BMP packing = write_bmp of the BGRA frame (bmp.cc:15-52), byte for byte
    bmp = r.render_bmp()
    assert bmp.size == 54 + 640 * 3 * 360
    assert np.array_equal(read_bmp_rgb_bytes(bmp), rgb)
    assert bytes(bmp[:2]) == b"BM" and int.from_bytes(bytes(bmp[2:6]), "little") == bmp.size


def read_bmp_rgb_bytes(buf):
    w = int.from_bytes(bytes(buf[18:22]), "little")
    h = int.from_bytes(bytes(buf[22:26]), "little")
    pitch = (w * 3 + 3) // 4 * 4
    px = np.asarray(buf[54:54 + pitch * h]).reshape(h, pitch)[:, :w * 3].reshape(h, w, 3)
    return px[::-1, :, ::-1]


def test_render_frame_dropin_matches_split_calls(frames, oracle, pkg):
    """ptgpu_render_frame (the one-call drop-in for main.cc:88) == set_frame + render, and the
    explicit-range entry point gives the same image as the one that parses the reference TLAS."""
    r = frames.use(520)
    a = r.render().copy()
    v = oracle.setup_frame(520)
    fr = pkg.scene_io.frame_from_view(v)
    b = r.render_frame(fr["subframes"], fr["dyn_instances"], fr["tlas_nodes"], fr["tlas_links"])
    assert np.array_equal(a, b)
    # scene.cc:634-674: 1 frame-static extra (buddha; the logo is gone after frame 119) then, per
    # subframe, teapot + armadillo
    n_sub, n_dyn = fr["subframes"].shape[0], fr["dyn_instances"].shape[0]
    per = (n_dyn - 1) // n_sub
    begin = np.array([1 + i * per for i in range(n_sub)], np.uint32)
    r.set_frame_ranges(fr["subframes"], fr["dyn_instances"], begin, begin + per)
    c = r.render()
    frames.current = None
    assert np.array_equal(a, c)


def test_edge_cases(frames, pkg):
    r = frames.use(0)
    # ragged rectangle (not a multiple of any tile size), single sample, negative-free offsets
    rgb, bgra = r.render_rect(637, 357, 3, 3, 5, 1, 1)
    assert rgb.shape == (3, 3, 3) and np.isfinite(rgb).all()
    one = r.trace_samples([[637, 357]], [5])
    assert np.allclose(rgb[0, 0], one[0], rtol=0, atol=0)
    # fewer samples than sample lanes, and a sample stride that lands in the last subframe
    rgb2, _ = r.render_rect(0, 0, 9, 5, 255, 1, 1)
    assert np.isfinite(rgb2).all()
    # a sample beyond the frame's subframes is an error, not a silent clamp
    with pytest.raises(pkg.PtgpuError):
        r.render_rect(0, 0, 4, 4, 256, 1, 1)
    with pytest.raises(pkg.PtgpuError):
        r.render_rect(0, 0, 0, 4, 0, 1, 1)
    # rendering before a frame is set fails loudly
    r2 = pkg.Renderer(pkg.Config.testing(), 0)
    with pytest.raises(pkg.PtgpuError):
        r2.render()
    r2.close()


def test_scheduled_traversal_equals_plain_traversal(frames):
    """Every ray of every round re-traced with the plain single-ray traversal (option "validate"):
    the burst-scheduled kernel (postponed triangle tests, parked instance groups) must store the same
    hit. Frames 520 and 1000 contain the dragon/buddha meshes with degenerate triangles, which once
    produced spurious, scheduling-dependent hits."""
    for frame in (520, 1000):
        r = frames.use(frame)
        r.set_option("validate", 1)
        try:
            a, _ = r.render_rect(96, 240, 96, 64, 0, 64, 4, tonemap=False)
            assert r.get_stat("validate_mismatches") == 0
            assert r.get_stat("wave_lanes") == 64 and r.get_stat("wave_rounds") <= 8
            # and the image does not depend on how the warps were scheduled
            r.set_option("validate", 0)
            # (plain_trace: the queues traced by the plain one-thread-per-ray loop instead of the scheduled kernel)
            for opts in ({"node_threshold": 1, "tri_threshold": 1, "xform_threshold": 1}, {"node_burst": 1, "min_active": 1},
                         {"plain_trace": 1}, {"plain_trace": 2}, {"lanes": 8}):
                for k, v in opts.items():
                    r.set_option(k, v)
                b, _ = r.render_rect(96, 240, 96, 64, 0, 64, 4, tonemap=False)
                if "lanes" in opts:   # another summation order over the samples: rounding-level differences
                    np.testing.assert_allclose(a, b, rtol=2e-5, atol=1e-7)
                else:
                    assert np.array_equal(a, b), opts
                for k, v in {"node_threshold": 16, "tri_threshold": 8, "xform_threshold": -1, "node_burst": -1, "min_active": -1,
                             "plain_trace": 0, "lanes": 256}.items():
                    r.set_option(k, v)
        finally:
            r.set_option("validate", 0)


def _variant_case(pkg, variant, cfg, frame, window, samples):
    from oracle import refbind
    if not refbind.available(variant):
        pytest.skip("oracle variant %s not built" % variant)
    o = refbind.get(variant)
    o.load_scene()
    view = o.setup_frame(frame)
    assert (o.config.width, o.config.height, o.config.spp, o.config.max_bounces) == (cfg.width, cfg.height, cfg.spp, cfg.max_bounces)
    r = pkg.Renderer(cfg, device=0)
    try:
        r.upload_static(**pkg.scene_io.static_from_view(view))
        r.set_frame(**pkg.scene_io.frame_from_view(view))
        x0, y0, w, h = window
        s_begin, s_count, s_stride = samples
        g_rgb, g_bgra = r.render_rect(x0, y0, w, h, s_begin, s_count, s_stride)
        o_rgb, o_bgra = o.render_rect(x0, y0, w, h, s_begin, s_count, s_stride)
        ms, _ = r.last_render_ms()
    finally:
        r.close()
    return mae255(g_bgra, o_bgra), mean_rel(g_rgb, o_rgb), ms, view["subframes"].shape[0]


@pytest.mark.parametrize("frame", [376, 1100])
def test_motion_blur_config_1024spp(pkg, frame):
    """BASELINE.json configs[4]: motion-blur-heavy frames at 4x spp (1024 spp = 128 subframes per
    frame, scene.cc:648-650): frame 376 is inside the 97-degree camera whip-pan (scene.cc:382-383),
    frame 1100 in the bunny dash (scene.cc:541-563). 16 samples spread over all 128 subframes."""
    cfg = pkg.Config.testing()
    cfg.spp = 1024
    mae, rel, ms, n_sub = _variant_case(pkg, "mb", cfg, frame, (160, 90, 320, 180), (0, 16, 64))
    print("motion blur frame %d (128 subframes): MAE %.4f/255, image-mean rel %.2e, %.2f ms" % (frame, mae, rel, ms))
    assert n_sub == 128
    assert mae <= 1.0 and rel <= 2e-3


def test_high_poly_stress_scene(pkg, oracle):
    """BASELINE.json configs[3]: buddha + dragon + armadillo (132,547 triangles) side by side above the
    terrain, built through the reference's own add_instance/build_tlas (oracle/ref_harness.cc), one
    static instance and three dynamic ones through the C ABI; parity on a window + full-frame timing."""
    view = oracle.setup_stress_scene()
    try:
        assert view["n_static_instances"] == 1 and view["instances"].shape[0] == 4
        r = pkg.Renderer(pkg.Config.testing(), device=0)
        try:
            r.upload_static(**pkg.scene_io.static_from_view(view))
            r.set_frame(**pkg.scene_io.frame_from_view(view))
            x0, y0, w, h = 80, 100, 480, 240
            g_rgb, g_bgra = r.render_rect(x0, y0, w, h, 0, 16, 16)
            o_rgb, o_bgra = oracle.render_rect(x0, y0, w, h, 0, 16, 16)
            mae, rel = mae255(g_bgra, o_bgra), mean_rel(g_rgb, o_rgb)
            r.render_async(); r.sync()
            r.render_async(); r.sync()
            ms, _ = r.last_render_ms()
            print("stress scene: MAE %.4f/255, image-mean rel %.2e; full frame 640x360x256: %.1f ms = %.1f Mpaths/s"
                  % (mae, rel, ms, 640 * 360 * 256 / ms / 1e3))
            assert mae <= 1.0 and rel <= 1e-3
        finally:
            r.close()
    finally:
        oracle.restore_scene()


def test_production_config_window(pkg):
    """The production settings of config.hh:21-25 (1920x1080, 1024 spp, 5 bounces): a 384x216 window
    of frame 520, 8 samples spread over the 128 subframes, against the oracle built at that config."""
    cfg = pkg.Config.production()
    mae, rel, ms, n_sub = _variant_case(pkg, "prod", cfg, 520, (768, 432, 384, 216), (0, 8, 128))
    print("production frame 520 window: MAE %.4f/255, image-mean rel %.2e, %.2f ms" % (mae, rel, ms))
    assert n_sub == 128
    assert mae <= 1.0 and rel <= 2e-3


def test_animation_module_renders_the_same_frames(frames, oracle, pkg):
    """N1: ptgpu_set_animation_frame (own keyframe replay, no reference TLAS) vs ptgpu_set_frame on what
    the reference's setup_animation_frame produced: same picture (transforms agree to float rounding)."""
    import os
    if not os.path.exists(pkg.animation.default_path()):
        pytest.skip("scenes/_cache/animation.json not built")
    an = pkg.Animation(pkg.Config.testing())
    for frame in (100, 1100):
        r = frames.use(frame)
        a_rgb, a_bgra = r.render_rect(160, 90, 320, 180, 0, 32, 8)
        an.set_frame(r, frame)
        frames.current = None
        b_rgb, b_bgra = r.render_rect(160, 90, 320, 180, 0, 32, 8)
        o_rgb, o_bgra = oracle.render_rect(160, 90, 320, 180, 0, 32, 8)
        print("animation module frame %d: vs reference-arrays path MAE %.4f/255; vs oracle MAE %.4f/255, mean-rel %.2e"
              % (frame, mae255(a_bgra, b_bgra), mae255(b_bgra, o_bgra), mean_rel(b_rgb, o_rgb)))
        # the module's matrices agree with the reference's to float rounding (tests/test_animation_cpu.py);
        # a last-bit difference can still flip individual paths, so the two renders are compared statistically
        assert mae255(a_bgra, b_bgra) <= 1.0 and mean_rel(a_rgb, b_rgb) <= 2e-3
        assert mae255(b_bgra, o_bgra) <= 1.0 and mean_rel(b_rgb, o_rgb) <= 2e-3
    an.close()


def test_validator_on_full_frames(frames, oracle):
    """validator.py's rule (restated in oracle/validator_np.py) on whole frames at the full 256 spp: the
    oracle-rendered frame is turned into the half-size reference PNG array the validator expects, the
    GPU frame goes through ptgpu_render_bmp + the BMP reader. No BAD frame, MAE <= 1/255. Frame 0 is
    cheap for the oracle (black sky); frame 1750 is the end card at dusk."""
    from helpers import read_bmp_rgb
    from oracle import validator_np as V
    import tempfile
    for frame in (0, 1750):
        r = frames.use(frame)
        bmp = r.render_bmp()
        with tempfile.NamedTemporaryFile(suffix=".bmp") as f:
            f.write(bmp.tobytes()); f.flush()
            own = read_bmp_rgb(f.name)
        _, o_bgra = oracle.render_frame()
        ref_rgb = o_bgra[..., 2::-1]
        psnr, good = V.validate_frame(V.make_reference_png_array(ref_rgb), own)
        mae = np.abs(own.astype(np.float64) - ref_rgb.astype(np.float64)).mean()
        print("validator frame %04d: PSNR %.2f dB %s, full-frame MAE %.4f/255" % (frame, psnr, "GOOD" if good else "BAD", mae))
        assert good and mae <= 1.0


def test_full_size_properties(frames, oracle):
    """At BASELINE.json's full size (640x360x256 spp) the oracle is too slow to compare everything,
    so use size-independent properties: determinism (two runs bit-identical), linearity of the
    accumulation (mean over samples 0..255 == mean of the two half sets), and the oracle on a
    strided pixel subset."""
    r = frames.use(1400)
    a = r.render().copy()
    b = r.render()
    assert np.array_equal(a, b)
    full, _ = r.render_rect(200, 100, 64, 32, 0, 256, 1, tonemap=False)
    even, _ = r.render_rect(200, 100, 64, 32, 0, 128, 2, tonemap=False)
    odd, _ = r.render_rect(200, 100, 64, 32, 1, 128, 2, tonemap=False)
    np.testing.assert_allclose(full, 0.5 * (even + odd), rtol=2e-5, atol=1e-6)
    # oracle at the full 256 spp on 3 rows
    for y in (40, 180, 300):
        o_rgb, o_bgra = oracle.render_rect(0, y, 640, 1, 0, 256, 1)
        assert mae255(a[y:y + 1], o_bgra) <= 1.0


@pytest.mark.parametrize("frame", [520, 1400])
def test_every_10th_row_at_full_spp(frames, oracle, frame):
    """Content frames at the full 256 spp on every 10th row (36 rows, 5.9 M paths): the image-mean bar of
    BASELINE.json (1e-3 relative) and the tonemapped MAE, asserted at the sample count the metric is quoted on."""
    r = frames.use(frame)
    rows = list(range(0, 360, 10))
    g_rgb = np.concatenate([r.render_rect(0, y, 640, 1, 0, 256, 1)[0] for y in rows])
    g_bgra = np.concatenate([r.render_rect(0, y, 640, 1, 0, 256, 1)[1] for y in rows])
    o = [oracle.render_rect(0, y, 640, 1, 0, 256, 1) for y in rows]
    o_rgb, o_bgra = np.concatenate([x[0] for x in o]), np.concatenate([x[1] for x in o])
    rel, mae = mean_rel(g_rgb, o_rgb), mae255(g_bgra, o_bgra)
    print("frame %d, every 10th row at 256 spp: image-mean rel %.2e, MAE %.4f/255" % (frame, rel, mae))
    assert rel <= 1e-3, rel          # measured 1.5e-5 and 8e-5
    assert mae <= 0.25, mae          # measured 0.09 and 0.11


def test_device_validator_matches_the_restated_validator(frames):
    """ptgpu_validate_frame (validator.py:41-52 on the device-resident frame) against
    oracle/validator_np.py on the fetched frame: same PSNR to rounding, same verdict. References: the
    golden frame 0 downscaled (a GOOD case), the same with noise, and a wrong image (BAD)."""
    from helpers import GOLDEN, read_bmp_rgb
    from oracle import validator_np as V
    r = frames.use(0)
    own = r.render()[..., 2::-1]
    gold_half = V.make_reference_png_array(read_bmp_rgb(GOLDEN))
    rng = np.random.RandomState(11)
    noisy = np.clip(gold_half.astype(np.int32) + rng.randint(-40, 41, gold_half.shape), 0, 255).astype(np.uint8)
    wrong = rng.randint(0, 256, gold_half.shape).astype(np.uint8)
    verdicts = []
    for ref in (gold_half, noisy, wrong):
        want_psnr, want_good = V.validate_frame(ref, own)
        psnr, good = r.validate_frame(ref)
        print("device validator: %.6f dB vs %.6f dB, good %s" % (psnr, want_psnr, good))
        assert abs(psnr - want_psnr) <= 1e-9 * max(1.0, abs(want_psnr)) and good == want_good
        verdicts.append(good)
    assert verdicts[0] and not verdicts[2]
    with pytest.raises(ValueError):
        r.validate_frame(gold_half[:-1])


def test_scene_from_meshes_gives_the_same_hits(pkg, oracle, renderer):
    """ptgpu_upload_meshes (BLASes built here from the triangles, no reference BVH) against
    ptgpu_upload_static (BLASes recovered from bvh.cc's arrays) and against the oracle's ray query: the
    closest hit of a ray does not depend on the tree it was found with."""
    from test_abi_cpu import _mesh_table
    frame = 520
    view = oracle.setup_frame(frame)
    st = pkg.scene_io.static_from_view(view)
    an = pkg.Animation(pkg.Config.testing())
    sub, dyn, b, e = an.frame(frame)
    own = pkg.Renderer(pkg.Config.testing(), device=0)
    own.upload_meshes(st["indices"], st["pos"], st["normal"], st["albedo"], st["material"], _mesh_table(pkg, oracle, st), st["instances"])
    own.set_frame_ranges(sub, dyn, b, e)
    renderer.set_frame_ranges(sub, dyn, b, e)
    rays = camera_like_rays(4000, frame)
    f_ref, u_ref = renderer.trace_closest(rays, 0)
    f_own, u_own = own.trace_closest(rays, 0)
    assert (f_ref[:, 0] > 0).mean() > 0.5
    assert np.array_equal(f_ref[:, 0] > 0, f_own[:, 0] > 0)
    same = (u_ref[:, :2] == u_own[:, :2]).all(1)
    print("own-built BLAS: %d rays, %d hits, same triangle %.4f, identical t %.4f" % (
        len(rays), int((f_ref[:, 0] > 0).sum()), same.mean(), (f_ref[:, 0] == f_own[:, 0]).mean()))
    assert same.all() and np.array_equal(f_ref, f_own)
    o_ref = [oracle.trace_closest(x[0:3], x[4:7], float(x[3]), float(x[7]), 0) for x in rays[:400]]
    o_t = np.array([h["thit"] for h in o_ref], np.float32)
    hit = (o_t > 0) & (f_own[:400, 0] > 0)
    assert ((o_t > 0) == (f_own[:400, 0] > 0)).mean() > 0.995
    np.testing.assert_allclose(f_own[:400, 0][hit], o_t[hit], rtol=2e-4)
    # whole-path check: the same window rendered with either tree
    a_rgb, a_bgra = renderer.render_rect(200, 120, 96, 64, 0, 64, 4)
    b_rgb, b_bgra = own.render_rect(200, 120, 96, 64, 0, 64, 4)
    print("own-built BLAS window: identical pixels %.4f, MAE %.5f/255" % ((a_bgra == b_bgra).all(-1).mean(), mae255(a_bgra, b_bgra)))
    assert mae255(a_bgra, b_bgra) <= 0.02
    # what needs the reference's link tables says so
    with pytest.raises(pkg.PtgpuError):
        own.set_option("traversal", 1)
    with pytest.raises(pkg.PtgpuError):
        own.set_frame(**pkg.scene_io.frame_from_view(view))
    own.close()
    an.close()


def test_scene_from_obj_files_without_reference_mesh_or_bvh_code(pkg, oracle, renderer):
    """SURVEY.md N4 + N2 + N1 together: OBJ/MTL files -> ptgpu_meshes_* -> ptgpu_upload_meshes (own BLASes)
    -> ptgpu_anim frame state -> render, against the same frame rendered from the reference's arrays. Only
    the placement of the static instances and the height repaint of the terrain (load_scene,
    scene.cc:141-269) still come from the reference here."""
    import os
    from oracle import refbind
    from test_meshes_cpu import SCENE_MESHES
    frame = 1000
    view = oracle.setup_frame(frame)
    st = pkg.scene_io.static_from_view(view)
    ms = pkg.MeshSet()
    for name in SCENE_MESHES:
        ms.load_obj(name, os.path.join(refbind.REF_DIR, "data", name + ".obj"))
    a = ms.arrays()
    nt = ms.meshes["terrain"][0]
    a["albedo"][:nt] = st["albedo"][:nt]        # load_scene's repaint of the terrain (scene.cc:155-163)
    a["material"][:nt] = st["material"][:nt]
    an = pkg.Animation(pkg.Config.testing())
    sub, dyn, b, e = an.frame(frame)
    own = pkg.Renderer(pkg.Config.testing(), device=0)
    own.upload_meshes(a["indices"], a["pos"], a["normal"], a["albedo"], a["material"], ms.table(), st["instances"])
    own.set_frame_ranges(sub, dyn, b, e)
    renderer.set_frame_ranges(sub, dyn, b, e)
    a_rgb, a_bgra = renderer.render_rect(160, 100, 128, 64, 0, 64, 4)
    b_rgb, b_bgra = own.render_rect(160, 100, 128, 64, 0, 64, 4)
    same = (a_bgra == b_bgra).all(-1).mean()
    print("scene from OBJ files: identical pixels %.4f, MAE %.5f/255, mean-rel %.2e" % (same, mae255(a_bgra, b_bgra), mean_rel(b_rgb, a_rgb)))
    # normals agree to 1 ulp with the oracle build's, so a few paths may differ; the image does not
    assert same > 0.95 and mae255(a_bgra, b_bgra) <= 0.05 and mean_rel(b_rgb, a_rgb) <= 1e-4   # measured: 0.9705, 0.016, 2e-6
    own.close(); an.close(); ms.close()
