"""OBJ/MTL loader (csrc/mesh_loader.cc) against the reference's load_mesh (mesh.cc:104-265):
the 18 meshes of load_scene (scene.cc:139-182) loaded in the same order from the same files must give the
oracle's mesh buffers — indices and handles exactly, attributes to float rounding (the oracle build may
contract or reassociate the normalisation and the emission scaling)."""
import os

import numpy as np
import pytest

SCENE_MESHES = ["terrain", "leaf_tree", "maple_tree", "pine_tree", "tropical_tree", "willow_tree",
                "rock0", "rock1", "rock2", "rock3", "rock4", "armadillo", "buddha", "bunny", "dragon", "teapot",
                "end", "logo"]   # scene.cc:139-182


@pytest.fixture(scope="module")
def loaded(pkg, oracle):
    from oracle import refbind
    ms = pkg.MeshSet()
    for name in SCENE_MESHES:
        ms.load_obj(name, os.path.join(refbind.REF_DIR, "data", name + ".obj"))
    yield ms
    ms.close()


def test_mesh_handles_and_indices_match_load_mesh(loaded, oracle):
    v = oracle.setup_frame(0)
    for name in SCENE_MESHES:
        assert list(loaded.meshes[name]) == oracle.find_mesh(name)[:4], name      # mesh handle (mesh.hh:18-28)
    a = loaded.arrays()
    assert np.array_equal(a["indices"], v["indices"])                             # same de-duplication, same numbering


def test_vertex_attributes_match_load_mesh(loaded, oracle):
    v = oracle.setup_frame(0)
    a = loaded.arrays()
    assert a["pos"].shape == v["pos"].shape
    assert np.array_equal(a["pos"][:, :3], v["pos"][:, :3])                       # parsed numbers: exact
    np.testing.assert_allclose(a["normal"][:, :3], v["normal"][:, :3], rtol=0, atol=2.5e-7)
    # load_scene repaints the terrain's non-water vertices by height after loading it (scene.cc:141-163):
    # compare the water vertices of the terrain and every vertex of the other 17 meshes
    nt = loaded.meshes["terrain"][0]
    keep = np.ones(a["pos"].shape[0], bool)
    keep[:nt] = a["material"][:nt, 2] != 0
    assert 0 < keep[:nt].sum() < nt
    assert np.array_equal(a["albedo"][keep], v["albedo"][keep])
    np.testing.assert_allclose(a["material"][keep], v["material"][keep], rtol=3e-7, atol=0)
    # the emissive logo and the glass/water materials are in there
    assert a["material"][:, 3].max() > 0 and a["material"][:, 2].max() > 0


def test_loader_errors_and_small_cases(pkg, tmp_path):
    ms = pkg.MeshSet()
    with pytest.raises(pkg.PtgpuError, match="Unable to open"):
        ms.load_obj("nope", tmp_path / "missing.obj")
    # a face with position//normal, a negative-free index layout, and a material library that is missing
    obj = tmp_path / "t.obj"
    obj.write_text("mtllib t.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 2\nusemtl red\nf 1//1 2//1 3//1\nf 3//1 2//1 1//1\n")
    with pytest.raises(pkg.PtgpuError, match="t.mtl"):
        ms.load_obj("t", obj)
    (tmp_path / "t.mtl").write_text("newmtl red\nKd 0.5 0.25 0.125\nd 0.75\nPr 0.3\nPm 1\nKe 0.25 0 0\nTf 0.1 0.9 0.2\n")
    m = ms.load_obj("t", obj)
    a = ms.arrays()
    assert m[:2] == (3, 2)                                        # 3 distinct index groups, 2 triangles
    assert a["indices"][-6:].tolist() == [0, 1, 2, 2, 1, 0]
    assert np.allclose(a["normal"][-1, :3], [0, 0, 1])            # vn is normalised (mesh.cc:176)
    assert np.allclose(a["albedo"][-1], [0.5, 0.25, 0.125, 0.75])
    # material = roughness, metallicness, max transmission, max(emission / max(albedo, emission)) (mesh.cc:236-250)
    assert np.allclose(a["material"][-1], [0.3, 1.0, 0.9, 0.5])
    assert ms.table().shape == (1, 4)
    ms.close()
