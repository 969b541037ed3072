"""CPU suite: the C-ABI library loads, exports every symbol include/ptgpu.h declares, fails loudly
without a GPU, and its host-side BVH flattening is structurally correct. No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ptgpu.h")).read()
    return sorted(set(re.findall(r"\b(ptgpu_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_list_agree(pkg):
    assert declared_symbols() == sorted(pkg.capi.SYMBOLS)


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.load_library()
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_default_config_is_shipped_config_hh(pkg):
    lib = pkg.load_library()
    cfg = pkg.Config()
    lib.ptgpu_default_config(C.byref(cfg))
    assert (cfg.width, cfg.height, cfg.spp, cfg.max_bounces) == (640, 360, 256, 4)
    assert cfg.student_id == 152121358 and cfg.samples_per_subframe == 8
    assert pkg.Config.production().subframes == 128 and pkg.Config.testing().subframes == 32


def test_create_fails_loudly_without_gpu(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.PtgpuError) as e:
        pkg.Renderer(pkg.Config.testing(), 0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg_dir = os.path.join(ROOT, "path-tracing...but-on-the-lumi-cluster_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hh", ".cc", ".sh")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for line in text.splitlines():
                    code = line.split("#")[0].split("//")[0]
                    assert "refbind" not in code and "libptref" not in code, (f, line)
                    assert not re.search(r"\bimport\s+oracle|from\s+oracle", code), (f, line)


def test_host_flatten_of_the_reference_bvhs(pkg, oracle):
    """bvh.cc's link-table BVHs -> 4-wide layout: every triangle of all 18 meshes reachable exactly
    once, boxes nested, all 885 static instances in the static TLAS, stack bound within capacity."""
    lib = pkg.load_library()
    v = oracle.setup_frame(0)
    st = pkg.scene_io.static_from_view(v)
    arrs = {k: np.ascontiguousarray(a) for k, a in st.items()}
    out = (C.c_uint64 * 8)()
    err = C.create_string_buffer(512)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    rc = lib.ptgpu_host_flatten_check(
        p(arrs["nodes"]), arrs["nodes"].shape[0], p(arrs["links"]), arrs["links"].shape[0],
        p(arrs["indices"]), arrs["indices"].shape[0], p(arrs["pos"]), arrs["pos"].shape[0],
        p(arrs["instances"]), arrs["instances"].shape[0], out, err, 512)
    assert rc == 0, err.value
    n_blas, n_nodes, n_tris, n_tlas, stack, bad, cap, _ = list(out)
    assert n_blas == 18                                  # scene.cc:139-182 loads 18 meshes
    assert n_tris == arrs["indices"].shape[0] // 3       # every triangle, once
    assert bad == 0 and 0 < stack <= cap
    assert n_nodes < arrs["nodes"].shape[0] // 4         # 4-wide collapse shrinks the node count
    assert n_tlas < 885


def test_host_flatten_rejects_corrupt_links(pkg, oracle):
    lib = pkg.load_library()
    v = oracle.setup_frame(0)
    st = pkg.scene_io.static_from_view(v)
    arrs = {k: np.ascontiguousarray(a).copy() for k, a in st.items()}
    arrs["links"][0, 0] = 0  # root of the first BLAS accepts itself
    out = (C.c_uint64 * 8)()
    err = C.create_string_buffer(512)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    rc = lib.ptgpu_host_flatten_check(
        p(arrs["nodes"]), arrs["nodes"].shape[0], p(arrs["links"]), arrs["links"].shape[0],
        p(arrs["indices"]), arrs["indices"].shape[0], p(arrs["pos"]), arrs["pos"].shape[0],
        p(arrs["instances"]), arrs["instances"].shape[0], out, err, 512)
    assert rc != 0 and len(err.value) > 0


def test_scene_snapshot_roundtrip(pkg, oracle, tmp_path):
    sio = pkg.scene_io
    v = oracle.setup_frame(520)
    sp, fp = str(tmp_path / "static.npz"), str(tmp_path / "frame.npz")
    sio.save_static(sp, v)
    sio.save_frame(fp, v, 520)
    st, fr = sio.load_static(sp), sio.load_frame(fp)
    live_s, live_f = sio.static_from_view(v), sio.frame_from_view(v)
    for k in sio.STATIC_KEYS:
        assert np.array_equal(st[k], live_s[k]), k
    for k in sio.FRAME_KEYS:
        assert np.array_equal(fr[k], live_f[k]), k
    assert fr["frame"] == 520
    # reference layout facts the C ABI relies on (include/ptgpu.h)
    assert st["links"].shape[0] == 8 * st["nodes"].shape[0]
    tl = fr["subframes"][:, :8].copy().view(np.uint32)      # subframe.tlas {node_count, node_offset}
    assert (tl[:, 1] >= st["nodes"].shape[0]).all()


def _mesh_table(pkg, oracle, static):
    names = ["logo", "buddha", "teapot", "armadillo", "dragon", "bunny", "end"]   # the per-frame meshes (scene.cc:634-674)
    extra = [tuple(oracle.find_mesh(n)[:4]) for n in names]
    return pkg.scene_io.mesh_table(static["instances"], extra)


def test_own_blas_builder(pkg, oracle):
    """ptgpu_upload_meshes' host side. Every BLAS built from the triangles alone (binned /
    full-sweep SAH, optimal 8-wide collapse): each triangle of each mesh reachable exactly once, child
    boxes enclose their subtrees, all static instances in the TLAS, stack bound within capacity."""
    lib = pkg.load_library()
    st = pkg.scene_io.static_from_view(oracle.setup_frame(0))
    arrs = {k: np.ascontiguousarray(a) for k, a in st.items()}
    meshes = _mesh_table(pkg, oracle, st)
    assert meshes.shape[0] >= 14
    out = (C.c_uint64 * 8)()
    err = C.create_string_buffer(512)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    rc = lib.ptgpu_host_build_check(
        p(arrs["indices"]), arrs["indices"].shape[0], p(arrs["pos"]), arrs["pos"].shape[0],
        p(meshes), meshes.shape[0], p(arrs["instances"]), arrs["instances"].shape[0], out, err, 512)
    assert rc == 0, err.value
    n_blas, _, n_tris, n_tlas, stack, bad, cap, n_cw = list(out)
    assert n_blas == meshes.shape[0] and n_tris == int(meshes[:, 1].sum())
    assert bad == 0 and 0 < stack <= cap and n_cw < n_tris
    # a mesh that runs past the index buffer is refused, with a message
    broken = meshes.copy()
    broken[-1, 1] += 10 ** 6
    rc = lib.ptgpu_host_build_check(
        p(arrs["indices"]), arrs["indices"].shape[0], p(arrs["pos"]), arrs["pos"].shape[0],
        p(broken), broken.shape[0], p(arrs["instances"]), arrs["instances"].shape[0], out, err, 512)
    assert rc != 0 and b"index buffer" in err.value
    # an instance whose mesh is not in the table is refused
    rc = lib.ptgpu_host_build_check(
        p(arrs["indices"]), arrs["indices"].shape[0], p(arrs["pos"]), arrs["pos"].shape[0],
        p(meshes[1:]), meshes.shape[0] - 1, p(arrs["instances"]), arrs["instances"].shape[0], out, err, 512)
    assert rc != 0 and b"does not match" in err.value


def test_oversize_dynamic_sets_are_rejected(pkg):
    """the kernels hold a subframe's per-frame instances as one 24-bit group / 16 stack entries;
    ptgpu_set_frame_ranges must refuse more instead of corrupting the traversal stack."""
    lib = pkg.load_library()
    err = C.create_string_buffer(256)

    def check(begin, end, n_dyn):
        b, e = np.asarray(begin, np.uint32), np.asarray(end, np.uint32)
        return lib.ptgpu_host_check_dynamic_ranges(b.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p), len(b), n_dyn, err, 256)

    assert check([2, 5], [5, 7], 7) == 0                   # the animation: prefix 2 (logo, buddha) + 2-3 per subframe
    assert check([0], [16], 16) == 0                       # exactly the limit
    assert check([0], [17], 17) != 0 and b"17 dynamic instances" in err.value
    assert check([10, 12], [12, 20], 20) != 0              # prefix 10 + 8 of its own = 18
    assert check([3], [2], 4) != 0 and b"bad range" in err.value
    assert check([0], [5], 4) != 0                         # range beyond the instance array


def test_flat_scene_build_of_a_small_scene(pkg, oracle):
    """The flat static scene (all static instances as world-space triangles under one 8-wide BVH) of the stress
    scene's static part (the terrain, 73 730 triangles): every triangle a leaf exactly once, vertices inside their
    quantised leaf boxes and every ancestor's slot box."""
    lib = pkg.load_library()
    v = oracle.setup_stress_scene()
    try:
        st = pkg.scene_io.static_from_view(v)
        arrs = {k: np.ascontiguousarray(a) for k, a in st.items()}
        out = (C.c_uint64 * 8)()
        err = C.create_string_buffer(512)
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        rc = lib.ptgpu_host_flat_check(
            p(arrs["nodes"]), arrs["nodes"].shape[0], p(arrs["links"]), arrs["links"].shape[0],
            p(arrs["indices"]), arrs["indices"].shape[0], p(arrs["pos"]), arrs["pos"].shape[0],
            p(arrs["instances"]), arrs["instances"].shape[0], out, err, 512)
    finally:
        oracle.restore_scene()
    assert rc == 0, err.value
    n_tris, n_nodes, n_top, depth, bad = list(out)[:5]
    assert n_tris == 73730 and bad == 0
    assert 73730 / 8 < n_nodes < 73730 / 3 and 0 < depth <= 20
