"""Scene snapshots: the arrays that cross the render seam, stored as .npz files.

A snapshot is exactly what the reference's caller holds in host memory when it reaches
main.cc:88 — the static arrays after load_scene() and the per-frame arrays after
setup_animation_frame(f) — so that the benchmark and tools can feed the C ABI on a machine where
the reference's scene code and assets are not present. Snapshots are written by
oracle/make_snapshots.py (test infrastructure, run where /root/reference is mounted) into
scenes/_cache/ (git-ignored, travels to the GPU box). This module only reads/writes arrays.
"""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CACHE = os.path.join(ROOT, "scenes", "_cache")

STATIC_KEYS = ("nodes", "links", "indices", "pos", "normal", "albedo", "material", "instances")
FRAME_KEYS = ("subframes", "dyn_instances", "tlas_nodes", "tlas_links")


def static_path(tag="testing"):
    """The static arrays do not depend on config.hh (resolution/spp only enter the per-frame camera and
    the subframe count), so every config shares the one snapshot."""
    return os.path.join(CACHE, "static_testing.npz")


def frame_path(frame, tag="testing"):
    return os.path.join(CACHE, "frame_%s_%04d.npz" % (tag, frame))


def save_static(path, view):
    """view: dict of the seam arrays plus n_static_nodes / n_static_instances (any frame; with
    n_static_nodes marking the end of the BLAS region)."""
    n = view["n_static_nodes"]
    ns = view["n_static_instances"]
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez_compressed(
        path,
        nodes=view["nodes"][:n], links=view["links"][:8 * n], indices=view["indices"],
        pos=view["pos"], normal=view["normal"], albedo=view["albedo"], material=view["material"],
        instances=view["instances"][:ns])


def save_frame(path, view, frame):
    n = view["n_static_nodes"]
    ns = view["n_static_instances"]
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez_compressed(
        path, frame=np.int32(frame),
        subframes=view["subframes"], dyn_instances=view["instances"][ns:],
        tlas_nodes=view["nodes"][n:], tlas_links=view["links"][8 * n:])


def load_static(path):
    with np.load(path) as z:
        return {k: np.ascontiguousarray(z[k]) for k in STATIC_KEYS}


def load_frame(path):
    with np.load(path) as z:
        d = {k: np.ascontiguousarray(z[k]) for k in FRAME_KEYS}
        d["frame"] = int(z["frame"])
        return d


def frame_from_view(view):
    """The per-frame arrays out of a live oracle view (same split as save_frame)."""
    n = view["n_static_nodes"]
    ns = view["n_static_instances"]
    return {"subframes": view["subframes"], "dyn_instances": view["instances"][ns:],
            "tlas_nodes": view["nodes"][n:], "tlas_links": view["links"][8 * n:]}


def static_from_view(view):
    n = view["n_static_nodes"]
    ns = view["n_static_instances"]
    return {"nodes": view["nodes"][:n], "links": view["links"][:8 * n], "indices": view["indices"],
            "pos": view["pos"], "normal": view["normal"], "albedo": view["albedo"],
            "material": view["material"], "instances": view["instances"][:ns]}


def available_frames(tag="testing"):
    if not os.path.isdir(CACHE):
        return []
    out = []
    for name in sorted(os.listdir(CACHE)):
        if name.startswith("frame_%s_" % tag) and name.endswith(".npz"):
            out.append(int(name[len("frame_%s_" % tag):-4]))
    return out


def mesh_table(instances, extra=()):
    """(n, 4) uint32 table of the distinct meshes (vertex_count, triangle_count, index_offset,
    base_vertex_offset) named by `instances` ((m, 160) uint8 tlas_instance records: bvh at byte 0, mesh at
    byte 8, bvh.hh:69-81) plus `extra` rows (meshes only per-frame instances use), ordered by index_offset."""
    inst = np.ascontiguousarray(instances, dtype=np.uint8).reshape(-1, 160)
    rows = {tuple(int(x) for x in r) for r in inst[:, 8:24].copy().view(np.uint32).reshape(-1, 4)}
    rows |= {tuple(int(x) for x in r) for r in extra}
    return np.array(sorted(rows, key=lambda r: r[2]), dtype=np.uint32).reshape(-1, 4)
