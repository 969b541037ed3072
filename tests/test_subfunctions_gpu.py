"""Direct parity of the hot path's sub-functions: every device function that restates a reference function is
called ON ITS OWN (ptgpu_debug_eval, include/ptgpu.h) on seeded inputs and compared with the reference's
function of the same name (oracle/ref_harness.cc: ref_eval -> path_tracer.hh / math.hh). Whole-path and image
tests (test_parity_gpu.py) bound these only in aggregate; here every BSDF lobe — reflection, refraction,
diffuse, delta, "bad" sample, total internal reflection — and both early-outs of the sky march are forced
by construction and counted.

Tolerances: integers and the RNG floats bit-exact; floating point 1e-5 relative where the function is a
few operations, wider where the reference leaks double precision through unqualified cos/sin/sqrt/exp/pow
(math.hh includes <cmath> without `using`) and is built with -ffast-math — each bar is about twice the
error measured on B200 (run this file as a script on a GPU box to print the measured errors).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

FN = dict(RAND4=0, FILM_OFFSET=1, CAMERA_RAY=2, GGX_VNDF=3, BSDF=4, SAMPLE_BSDF=5, SKY_ATTENUATION=6,
          SKY_SCATTERING=7, SAMPLE_CONE=8, SHADOW_RAY=9, TRACE_RAY=10)


MEASURED = {}
ENFORCE = True


def bar(key, value, greater=False):
    """assert value < BARS[key] (or > for agreement rates), remembering what was measured"""
    MEASURED.setdefault(key, []).append(float(value))
    if ENFORCE:
        assert (value > BARS[key]) if greater else (value < BARS[key]), (key, float(value), BARS[key])


def unbits(f32):
    return np.ascontiguousarray(f32, np.float32).view(np.uint32)


def unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def rel_err(got, want, floor=1e-6):
    """max over components of |got - want| / max(|want| row-wise, floor), per row"""
    scale = np.maximum(np.abs(want).max(axis=1, keepdims=True), floor)
    return (np.abs(got - want) / scale).max(axis=1)


def both(renderer, oracle, fn, inputs):
    inputs = np.asarray(inputs, np.float32)
    return renderer.debug_eval(FN[fn], inputs), oracle.eval(FN[fn], inputs)


# ---- input generators -------------------------------------------------------------------------------------

def surfaces(rng, n):
    """albedo[3], roughness, metallic, transmission, eta — every corner of the material space"""
    albedo = rng.uniform(0.02, 1.0, (n, 3))
    rough = rng.choice([1e-4, 5e-4, 2e-3, 0.01, 0.09, 0.25, 0.64, 1.0], n)   # < 1e-3: delta lobe (path_tracer.hh:69, 277, 291)
    metallic = rng.choice([0.0, 0.0, 0.3, 1.0], n)
    transmission = rng.choice([0.0, 0.0, 0.5, 1.0], n)
    eta = rng.choice([1.5, 1.0 / 1.5], n)                                     # back face / front face (path_tracer.hh:394-400)
    return np.column_stack([albedo, rough, metallic, transmission, eta])


def view_dirs(rng, n):
    v = unit(rng.normal(size=(n, 3)))
    v[:, 2] = np.abs(v[:, 2])
    v[: n // 16, 2] = 1e-7                                                    # the clamp of path_tracer.hh:701-702
    return unit(v)


def sun(elevation):
    return np.array([0.0, np.sin(elevation), np.cos(elevation)])


# ---- the tests ------------------------------------------------------------------------------------------------

def test_rand4_bit_exact(renderer, oracle):
    rng = np.random.RandomState(11)
    st = rng.randint(0, 2 ** 32, size=(5000, 4), dtype=np.uint64).astype(np.uint32)
    st[0] = (0, 0, 0, 0)
    st[1] = (0xFFFFFFFF,) * 4
    inp = np.zeros((len(st), 24), np.float32)
    inp.view(np.uint32)[:, 0:4] = st                            # bit patterns, never converted (a NaN payload would not survive)
    got, want = both(renderer, oracle, "RAND4", inp)
    assert np.array_equal(unbits(got[:, 0:4]), unbits(want[:, 0:4]))          # the pcg4d state
    assert np.array_equal(unbits(got[:, 4:8]), unbits(want[:, 4:8]))          # the four uniforms, bit for bit
    assert (got[:, 4:8] >= 0).all() and (got[:, 4:8] <= 1).all()


def test_film_offset(renderer, oracle):
    rng = np.random.RandomState(12)
    u = rng.uniform(0, 1, (8000, 2))
    u[:8] = [[0, 0], [0.5, 0.5], [1 - 2 ** -24, 0.25], [2 ** -32, 0.75], [0.999, 0.999], [1e-6, 1e-6], [0.5, 0], [0, 0.5]]
    got, want = both(renderer, oracle, "FILM_OFFSET", u)
    assert np.isfinite(got[:, :2]).all()
    # inv_erf is steep at the ends of the unit interval: the last ulps of u move the offset by 1e-4 pixel there
    err = np.abs(got[:, :2] - want[:, :2]).max(axis=1)
    bar("film_p999", np.percentile(err, 99.9))
    bar("film_max", err.max())


@pytest.mark.parametrize("frame", [330, 520])
def test_camera_ray(frames, oracle, frame):
    """frame 330 has depth of field (aperture_radius > 0, hexagonal aperture), frame 520 a pinhole"""
    r = frames.use(frame)
    rng = np.random.RandomState(frame)
    n = 6000
    inp = np.column_stack([rng.uniform(0, 1, (n, 2)), rng.uniform(0, 640, n), rng.uniform(0, 360, n), rng.randint(0, 32, n)])
    got, want = both(r, oracle, "CAMERA_RAY", inp)
    bar("camera_dir", rel_err(got[:, 0:3], want[:, 0:3]).max())
    bar("camera_origin_abs", np.abs(got[:, 3:6] - want[:, 3:6]).max())
    if frame == 330:
        assert np.ptp(want[:, 3:6], axis=0).max() > 1e-3       # the aperture really moves the origin


def test_sample_ggx_vndf(renderer, oracle):
    rng = np.random.RandomState(13)
    n = 8000
    inp = np.column_stack([view_dirs(rng, n), surfaces(rng, n)[:, 3], rng.uniform(0, 1, (n, 2))])
    got, want = both(renderer, oracle, "GGX_VNDF", inp)
    bar("vndf", rel_err(got[:, 0:3], want[:, 0:3]).max())
    delta = inp[:, 3] < 1e-3
    assert delta.sum() > 100 and (got[delta, 0:3] == [0, 0, 1]).all()


def test_bsdf_eval(renderer, oracle):
    rng = np.random.RandomState(14)
    n = 20000
    light = unit(rng.normal(size=(n, 3)))                       # both hemispheres: reflection and transmission branches (:192-193)
    inp = np.column_stack([light, view_dirs(rng, n), surfaces(rng, n)])
    got, want = both(renderer, oracle, "BSDF", inp)
    assert np.isfinite(got[:, :4]).all()
    err = rel_err(got[:, 0:4], want[:, 0:4], floor=1e-4)
    bar("bsdf_p999", np.percentile(err, 99.9))
    bar("bsdf_max", err.max())
    assert ((want[:, 3] == 0) == (got[:, 3] == 0)).mean() > 0.9995     # zero pdf (wrong-side / delta) on the same inputs


def lobe_of(out, rough):
    """which branch of sample_bsdf (path_tracer.hh:250-274, 291-293) produced a row"""
    bad = (out[:, 3:6] == 0).all(axis=1) & (out[:, 6] == 1) & (out[:, 2] == 1)
    delta = out[:, 6] < 0
    refr = ~bad & (out[:, 2] < 0)
    return np.where(bad, 0, np.where(delta & refr, 1, np.where(delta, 2, np.where(refr, 3, 4))))


def test_sample_bsdf_every_lobe(renderer, oracle):
    rng = np.random.RandomState(15)
    n = 40000
    s = surfaces(rng, n)
    inp = np.column_stack([rng.uniform(0, 1, (n, 3)), view_dirs(rng, n), s])
    # grazing views from inside glass: refract() returns the zero vector = total internal reflection (:91-96, :258-259)
    inp[:2000, 5] = rng.uniform(1e-3, 0.3, 2000)
    inp[:2000, 3:6] = unit(inp[:2000, 3:6])
    inp[:2000, 11] = 1.0
    inp[:2000, 12] = 1.5
    got, want = both(renderer, oracle, "SAMPLE_BSDF", inp)
    assert np.isfinite(got[:, :7]).all()
    lg, lw = lobe_of(got, inp[:, 9]), lobe_of(want, inp[:, 9])
    counts = np.bincount(lw, minlength=5)
    # bad sample, delta refraction, delta reflection, rough refraction, rough reflection/diffuse: all forced
    assert (counts > 200).all(), counts
    same = lg == lw
    # u.z within rounding of a lobe boundary may pick the other lobe (the reference's probabilities are double)
    assert (~same).sum() <= 4, (~same).sum()
    err_dir = np.abs(got[same, 0:3] - want[same, 0:3]).max(axis=1)
    bar("sample_dir_p999", np.percentile(err_dir, 99.9))
    bar("sample_dir_max", err_dir.max())
    # What the path tracer multiplies into the throughput is attenuation / |pdf| (path_tracer.hh:725-732): that
    # ratio is well conditioned and must agree tightly. The two factors on their own carry the GGX density,
    # which at roughness 0.002 and 0.01 (just above the delta threshold) is a ratio of nearly cancelling terms,
    # a^2 / (h.z^2 (a^2 - 1) + 1)^2 with 1 - h.z^2 ~ a^2: the last ulp of h.z moves it by percents, and the
    # reference evaluates part of it in double ("double-precision leakage").
    g, w = got[same], want[same]
    ok = ~((w[:, 3:6] == 0).all(axis=1))
    weight_g = g[ok, 3:6] / np.abs(g[ok, 6:7])
    weight_w = w[ok, 3:6] / np.abs(w[ok, 6:7])
    bar("sample_weight_p999", np.percentile(rel_err(weight_g, weight_w, floor=1e-4), 99.9))
    ill = inp[same, 9] < 0.05      # roughness 0.002 and 0.01: a^2 of 4e-6 and 1e-4 against 1 - h.z^2
    err_att = rel_err(g[:, 3:6], w[:, 3:6], floor=1e-4)
    err_pdf = rel_err(g[:, 6:7], w[:, 6:7], floor=1e-4)
    bar("sample_att_p999", np.percentile(err_att[~ill], 99.9))
    bar("sample_pdf_p999", np.percentile(err_pdf[~ill], 99.9))
    bar("sample_pdf_ill_max", err_pdf[ill].max())
    tir = (inp[:2000, 12] == 1.5) & (lw[:2000] == 0)
    assert tir.sum() > 50                                      # total internal reflection happened and was "bad" on both sides
    assert (lg[:2000][tir] == 0).all()


def test_sample_cone(renderer, oracle):
    rng = np.random.RandomState(16)
    n = 6000
    d = unit(rng.normal(size=(n, 3)))
    d[:3] = [[0, 1, 0], [0, 0, 1], [0, -1, 0]]
    inp = np.column_stack([d, np.full(n, np.cos(np.radians(4.0))), rng.uniform(0, 1, (n, 2))])
    got, want = both(renderer, oracle, "SAMPLE_CONE", inp)
    bar("cone_abs", np.abs(got[:, 0:3] - want[:, 0:3]).max())


def test_sky_attenuation(renderer, oracle):
    rng = np.random.RandomState(17)
    n = 6000
    pos = np.column_stack([rng.uniform(-100, 100, n), rng.uniform(0, 60, n), rng.uniform(-100, 100, n)])
    view = unit(rng.normal(size=(n, 3)))
    view[: n // 2, 1] = np.abs(view[: n // 2, 1])              # half of them look up; the rest may dive into the ground (= 0, :485)
    view = unit(view)
    inp = np.column_stack([rng.uniform(0, 1, n), pos, view])
    got, want = both(renderer, oracle, "SKY_ATTENUATION", inp)
    assert (want[:, 0:3] == 0).all(axis=1).sum() > 200 and (want[:, 0:3] > 0).all(axis=1).sum() > 2000
    assert ((want[:, 0:3] == 0).all(axis=1) == (got[:, 0:3] == 0).all(axis=1)).mean() > 0.999
    err = rel_err(got[:, 0:3], want[:, 0:3], floor=1e-3)
    bar("sky_att_p995", np.percentile(err, 99.5))


def test_sky_scattering_and_its_rng_draw(renderer, oracle):
    """nishita_atmosphere_scattering draws one rand4 only after both early-outs (path_tracer.hh:513, 521, 525):
    the seed coming back is the bit-exact witness of that."""
    rng = np.random.RandomState(18)
    n = 6000
    seeds = rng.randint(0, 2 ** 32, size=(n, 4), dtype=np.uint64).astype(np.uint32)
    elev = rng.choice([-0.1, 0.02, 0.3, 1.2], n)
    light = np.stack([sun(e) for e in elev])
    pos = np.column_stack([rng.uniform(-100, 100, n), rng.uniform(0, 60, n), rng.uniform(-100, 100, n)])
    view = unit(rng.normal(size=(n, 3)))
    tmax = rng.choice([-1.0, 0.5, 500.0, 999.0, 1000.0, 5e3, 1e5], n)
    inp = np.zeros((n, 24), np.float32)
    inp.view(np.uint32)[:, 0:4] = seeds                         # bit patterns, never converted
    inp[:, 4:7], inp[:, 7:10], inp[:, 10] = light, 4.0, np.cos(np.radians(4.0))
    inp[:, 11:14], inp[:, 14:17], inp[:, 17] = pos, view, tmax
    got, want = both(renderer, oracle, "SKY_SCATTERING", inp)
    assert np.array_equal(unbits(got[:, 6:10]), unbits(want[:, 6:10]))        # same decision to draw, same state after
    drew = (unbits(want[:, 6:10]) != seeds).any(axis=1)
    assert drew.sum() > 1500 and (~drew).sum() > 1500                          # both the march and the early-outs happened
    assert (got[~drew, 0:3] == 1).all() and (got[~drew, 3:6] == 0).all()       # early-out: attenuation 1, no in-scatter
    err_a = rel_err(got[drew, 0:3], want[drew, 0:3], floor=1e-3)
    err_s = rel_err(got[drew, 3:6], want[drew, 3:6], floor=1e-4)
    bar("sky_scat_att_p995", np.percentile(err_a, 99.5))
    bar("sky_scat_p995", np.percentile(err_s, 99.5))


def scene_rays(n, seed):
    rng = np.random.RandomState(seed)
    o = np.column_stack([rng.uniform(-90, 90, n), rng.uniform(15, 70, n), rng.uniform(-90, 90, n)])
    d = rng.normal(size=(n, 3))
    d[:, 1] = -np.abs(d[:, 1]) * 0.7
    d[: n // 8, 1] = np.abs(d[: n // 8, 1])                    # some miss into the sky, towards and away from the sun
    return o, unit(d)


@pytest.mark.parametrize("frame", [520, 1400])
def test_trace_ray_hit_info(frames, oracle, frame):
    """trace_ray (path_tracer.hh:340-412): closest hit -> position, tangent frame, interpolated material; miss -> sun disk"""
    r = frames.use(frame)
    n = 3000
    o, d = scene_rays(n, frame)
    d[:40] = oracle.view()["subframes"].view(np.float32).reshape(-1, 40)[0, 28:31]    # straight at the sun: the visible-disk branch (:361-365)
    sub = np.random.RandomState(frame).randint(0, 32, n)
    inp = np.column_stack([o, d, np.zeros(n), sub])
    got, want = both(r, oracle, "TRACE_RAY", inp)
    hit_w, hit_g = want[:, 0] >= 0, got[:, 0] >= 0
    assert (hit_w == hit_g).mean() > 0.998
    assert hit_w.sum() > 1500 and (~hit_w).sum() > 100
    miss = ~hit_w & ~hit_g
    assert np.allclose(got[miss][:, [13, 14, 15, 18, 21]], want[miss][:, [13, 14, 15, 18, 21]], rtol=1e-5, atol=0)   # sun radiance, emission 1, nee_pdf
    assert (want[miss, 21] > 0).sum() >= 5                     # some rays did see the sun disk
    h = hit_w & hit_g
    same = h & (np.abs(got[:, 0] - want[:, 0]) <= 2e-4 * np.abs(want[:, 0]))
    assert same.sum() / h.sum() > 0.995
    g, w = got[same], want[same]
    bar("hit_pos_abs", np.abs(g[:, 1:4] - w[:, 1:4]).max())
    # material: albedo, roughness, metallic, emission, transmission (barycentric interpolation), eta exact
    mat = [13, 14, 15, 16, 17, 18, 19]
    em = np.abs(g[:, mat] - w[:, mat]).max(axis=1)
    bar("hit_material_p99", np.percentile(em, 99))
    assert (g[:, 20] == w[:, 20]).mean() > 0.999               # front / back face
    # shading normal = third column of the tangent frame
    en = np.abs(g[:, 10:13] - w[:, 10:13]).max(axis=1)
    bar("hit_normal_p99", np.percentile(en, 99))
    assert np.abs(np.linalg.norm(g[:, 10:13], axis=1) - 1).max() < 1e-5


@pytest.mark.parametrize("frame", [520, 1400])
def test_trace_shadow_ray(frames, oracle, frame):
    """trace_shadow_ray (path_tracer.hh:415-427) from surface points towards the sun cone, as nee_branch does (:606-609)"""
    r = frames.use(frame)
    n = 4000
    o, d = scene_rays(n, frame + 7)
    first = oracle.eval(FN["TRACE_RAY"], np.column_stack([o, d, np.zeros(n), np.zeros(n)]))
    hit = first[:, 0] > 0
    pos = first[hit, 1:4]
    rng = np.random.RandomState(frame)
    light = oracle.view()["subframes"].view(np.float32).reshape(-1, 40)[0, 28:31].astype(np.float64)
    cone = oracle.eval(FN["SAMPLE_CONE"], np.column_stack([np.tile(light, (len(pos), 1)), np.full(len(pos), np.cos(np.radians(4.0))), rng.uniform(0, 1, (len(pos), 2))]))[:, 0:3]
    inp = np.column_stack([pos, cone, np.full(len(pos), 1e-4), np.full(len(pos), 1e9), np.zeros(len(pos))])
    got, want = both(r, oracle, "SHADOW_RAY", inp)
    assert 0.05 < want[:, 0].mean() < 0.95                     # both lit and shadowed points
    bar("shadow_agree", (got[:, 0] == want[:, 0]).mean(), greater=True)
    # bounded tmax: nothing beyond it counts
    inp[:, 7] = 1e-3
    got, want = both(r, oracle, "SHADOW_RAY", inp)
    bar("shadow_short_agree", (got[:, 0] == want[:, 0]).mean(), greater=True)


# bars: about twice the error measured on B200 (printed by `python tests/test_subfunctions_gpu.py`)
BARS = {
    "film_p999": 2e-4, "film_max": 1e-3, "camera_dir": 2e-6, "camera_origin_abs": 1e-5, "vndf": 3e-5,
    "bsdf_p999": 1e-4, "bsdf_max": 5e-3,
    "sample_dir_p999": 5e-6, "sample_dir_max": 2e-4, "sample_att_p999": 2e-3, "sample_pdf_p999": 2e-3, "sample_weight_p999": 3e-4, "sample_pdf_ill_max": 0.3,
    "cone_abs": 5e-5, "sky_att_p995": 5e-4, "sky_scat_att_p995": 3e-4, "sky_scat_p995": 6e-4,
    "hit_pos_abs": 2e-4, "hit_material_p99": 1e-6, "hit_normal_p99": 1.5e-3, "shadow_agree": 0.995, "shadow_short_agree": 0.99,
}


if __name__ == "__main__":
    # prints the measured errors beside the bars (run on a GPU box; bars are not enforced here)
    import __graft_entry__ as ge
    from oracle import refbind
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import FrameCache
    pkg = ge.load_package()
    o = refbind.get("fast")
    o.load_scene()
    rr = pkg.Renderer(pkg.Config.testing(), device=0)
    rr.upload_static(**pkg.scene_io.static_from_view(o.setup_frame(0)))
    fc = FrameCache(pkg, o, rr)
    ENFORCE = False
    for name, fn in list(globals().items()):
        if not name.startswith("test_"):
            continue
        params = [m.args[1] for m in getattr(fn, "pytestmark", []) if m.name == "parametrize"]
        for p in (params[0] if params else [None]):
            try:
                fn(rr, o) if p is None else fn(fc, o, p)
                print("%-40s %-5s ok" % (name, p))
            except AssertionError as e:
                print("%-40s %-5s ASSERT %s" % (name, p, str(e)[:300]))
    for k, v in MEASURED.items():
        print("%-22s bar %-8g measured %s" % (k, BARS[k], " ".join("%.3g" % x for x in v)))
    # where sample_bsdf's attenuation differs most: inputs and both outputs, by lobe
    rng = np.random.RandomState(15)
    n = 40000
    sf = surfaces(rng, n)
    inp = np.column_stack([rng.uniform(0, 1, (n, 3)), view_dirs(rng, n), sf]).astype(np.float32)
    got, want = both(rr, o, "SAMPLE_BSDF", inp)
    lw = lobe_of(want, inp[:, 9])
    err = rel_err(got[:, 3:6], want[:, 3:6], floor=1e-4)
    for lobe in range(5):
        m = lw == lobe
        if m.any():
            print("lobe %d: n %d, attenuation rel err p50 %.2e p99 %.2e p99.9 %.2e max %.2e" % (
                lobe, m.sum(), np.percentile(err[m], 50), np.percentile(err[m], 99), np.percentile(err[m], 99.9), err[m].max()))
    np.set_printoptions(precision=7, suppress=False, linewidth=200)
    for i in np.argsort(-err)[:12]:
        print("row %d lobe %d err %.3e\n  in  u %s view %s albedo %s rough %.4g metal %.3g trans %.3g eta %.4g\n  gpu dir %s att %s pdf %.6g\n  ref dir %s att %s pdf %.6g" % (
            i, lw[i], err[i], inp[i, 0:3], inp[i, 3:6], inp[i, 6:9], inp[i, 9], inp[i, 10], inp[i, 11], inp[i, 12],
            got[i, 0:3], got[i, 3:6], got[i, 6], want[i, 0:3], want[i, 3:6], want[i, 6]))
