"""Animation: host-side mirror of ptgpu_anim_* (csrc/frame_setup.cu) — the per-frame scene state of
the reference's setup_animation_frame (scene.cc:271-718) without its TLAS builds.

The keyframe rows and mesh handles are data: scenes/_cache/animation.json, written at build time by
oracle/extract_animation.py from the mounted reference. This module only loads that file and calls
the C ABI; nothing here touches oracle/.
"""
import ctypes as C
import json
import os

import numpy as np

from .capi import Config, PtgpuError, load_library
from .scene_io import CACHE

MESH_ORDER = ["logo", "buddha", "teapot", "armadillo", "dragon", "bunny", "end"]   # enum ptgpu_anim_mesh


def default_path():
    return os.path.join(CACHE, "animation.json")


class Animation:
    def __init__(self, config=None, path=None):
        self.lib = load_library()
        self.config = config or Config.testing()
        path = path or default_path()
        if not os.path.exists(path):
            raise FileNotFoundError("%s missing: run __graft_entry__.build() where the reference is mounted" % path)
        with open(path) as f:
            data = json.load(f)
        keys = np.zeros(len(data["keys"]), dtype=[("start", "<f4"), ("duration", "<f4"), ("from", "<f4"), ("to", "<f4"), ("var", "<i4")])
        for i, k in enumerate(data["keys"]):
            keys[i] = (k[0], k[1], k[2], k[3], int(k[4]))
        meshes = np.zeros((len(MESH_ORDER), 6), dtype=np.uint32)
        for i, name in enumerate(MESH_ORDER):
            meshes[i, :4] = data["meshes"][name]["mesh"]
            meshes[i, 4:] = data["meshes"][name]["blas"]
        self.mesh_rows = [tuple(int(x) for x in meshes[i, :4]) for i in range(len(MESH_ORDER))]   # the per-frame meshes
        self.handle = C.c_void_p()
        rc = self.lib.ptgpu_anim_create(C.byref(self.handle), keys.ctypes.data_as(C.c_void_p), keys.shape[0],
                                        meshes.ctypes.data_as(C.c_void_p), C.byref(self.config))
        if rc != 0:
            raise PtgpuError("ptgpu_anim_create failed")
        self.n_subframes = int(self.lib.ptgpu_anim_subframe_count(self.handle))
        self.max_instances = int(self.lib.ptgpu_anim_max_instances(self.handle))
        self.frame_count = int(self.lib.ptgpu_anim_frame_count(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.ptgpu_anim_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def frame(self, frame):
        """(subframes (n,160) u8, dyn_instances (m,160) u8, dyn_begin (n,) u32, dyn_end (n,) u32)."""
        sub = np.zeros((self.n_subframes, 160), np.uint8)
        dyn = np.zeros((self.max_instances, 160), np.uint8)
        b = np.zeros(self.n_subframes, np.uint32)
        e = np.zeros(self.n_subframes, np.uint32)
        n = C.c_size_t()
        rc = self.lib.ptgpu_anim_frame(self.handle, int(frame), sub.ctypes.data_as(C.c_void_p), dyn.ctypes.data_as(C.c_void_p),
                                       C.byref(n), b.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise PtgpuError("ptgpu_anim_frame failed")
        return sub, dyn[:n.value], b, e

    def set_frame(self, renderer, frame):
        """ptgpu_set_animation_frame: frame state straight into a Renderer's context."""
        rc = self.lib.ptgpu_set_animation_frame(renderer.ctx, self.handle, int(frame))
        if rc != 0:
            raise PtgpuError("ptgpu_set_animation_frame: %s" % self.lib.ptgpu_last_error(renderer.ctx).decode())
