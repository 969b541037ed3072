"""CPU suite for the N1 frame-setup module (csrc/frame_setup.cu through the C ABI): the per-frame
scene state it produces must equal what the reference's setup_animation_frame leaves in the scene
(scene.cc:271-718), minus the TLASes this library does not need."""
import time

import numpy as np
import pytest

FRAMES = [0, 59, 119, 120, 140, 300, 375, 520, 660, 980, 1100, 1114, 1115, 1400, 1740, 1799]


@pytest.fixture(scope="module")
def anim(pkg):
    import os
    if not os.path.exists(pkg.animation.default_path()):
        pytest.skip("scenes/_cache/animation.json not built")
    return pkg.Animation(pkg.Config.testing())


def test_frame_state_matches_setup_animation_frame(pkg, oracle, anim):
    assert anim.n_subframes == 32 and anim.frame_count == 1800
    for f in FRAMES:
        v = oracle.setup_frame(f)
        ns = v["n_static_instances"]
        ref_sub, ref_dyn = v["subframes"].copy(), v["instances"][ns:].copy()
        sub, dyn, b, e = anim.frame(f)
        # same instances in the same order: frame-static extras, then per subframe (scene.cc:634-674)
        assert dyn.shape == ref_dyn.shape, f
        assert np.array_equal(dyn[:, :24], ref_dyn[:, :24]), f          # (bvh, mesh) handles: exact
        t, rt = dyn[:, 32:].copy().view(np.float32), ref_dyn[:, 32:].copy().view(np.float32)
        np.testing.assert_allclose(t, rt, rtol=2e-6, atol=3e-5)          # transform + inverse (coordinates ~100)
        # camera floats without the float3 padding lanes (the reference leaves those uninitialised)
        keep = [0, 1, 2, 4, 5, 6, 8, 9, 10, 12, 13, 14, 16, 17, 18, 19, 21]
        cam, rcam = sub[:, 16:104].copy().view(np.float32)[:, keep], ref_sub[:, 16:104].copy().view(np.float32)[:, keep]
        np.testing.assert_allclose(cam, rcam, rtol=2e-6, atol=3e-5)
        assert np.array_equal(sub[:, 96:100].copy().view(np.int32), ref_sub[:, 96:100].copy().view(np.int32))  # aperture_polygon
        lk = [0, 1, 2, 4, 5, 6]
        light, rlight = sub[:, 112:140].copy().view(np.float32)[:, lk], ref_sub[:, 112:140].copy().view(np.float32)[:, lk]
        np.testing.assert_allclose(light, rlight, rtol=0, atol=1e-6)
        # ranges: contiguous, in order, covering everything after the extras
        n_extra = int(b[0])
        assert n_extra in (1, 2) and (b[1:] == e[:-1]).all() and int(e[-1]) == dyn.shape[0]
        # the reference's TLAS leaves of subframe i hold exactly extras + range i (what ptgpu_set_frame parses)
        tl = ref_sub[:, :8].copy().view(np.uint32)
        links = v["links"]
        for i in (0, 17, 31):
            cnt, off = int(tl[i, 0]), int(tl[i, 1])
            leaves = links[8 * off: 8 * off + cnt, 0]
            ids = sorted(int(x & 0x7FFFFFFF) - ns for x in leaves[(leaves & 0x80000000) != 0] if (x & 0x7FFFFFFF) >= ns)
            assert ids == list(range(n_extra)) + list(range(int(b[i]), int(e[i]))), (f, i)


def test_frame_setup_is_pure_and_fast(anim):
    a = anim.frame(1000)
    anim.frame(3)
    b = anim.frame(1000)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    t0 = time.perf_counter()
    for f in range(0, 1800, 18):
        anim.frame(f)
    per_frame = (time.perf_counter() - t0) / 100
    # the reference's setup_animation_frame takes ~75 ms per frame on one core
    assert per_frame < 5e-3, per_frame


def test_production_subframe_count(pkg):
    import os
    if not os.path.exists(pkg.animation.default_path()):
        pytest.skip("scenes/_cache/animation.json not built")
    a = pkg.Animation(pkg.Config.production())
    assert a.n_subframes == 128 and a.max_instances == 2 + 5 * 128
    sub, dyn, b, e = a.frame(520)
    assert sub.shape == (128, 160) and dyn.shape[0] == 1 + 2 * 128
    np.testing.assert_allclose(sub[0, 16 + 64:16 + 68].copy().view(np.float32)[0], 1920 / 1080, rtol=1e-6)  # aspect_ratio
