"""MeshSet: host-side mirror of ptgpu_meshes_* (csrc/mesh_loader.cc) — OBJ/MTL files to the mesh buffers of
the render path, the reference's load_mesh (mesh.cc:104-265) without its mesh.cc. Needs no GPU."""
import ctypes as C

import numpy as np

from .capi import PtgpuError, load_library


class MeshSet:
    def __init__(self):
        self.lib = load_library()
        self.handle = C.c_void_p()
        if self.lib.ptgpu_meshes_create(C.byref(self.handle)) != 0:
            raise PtgpuError("ptgpu_meshes_create failed")
        self.meshes = {}     # name -> (vertex_count, triangle_count, index_offset, base_vertex_offset)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.ptgpu_meshes_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_obj(self, name, path):
        """load_mesh(mb, path): appends the mesh, returns its handle as a 4-tuple (mesh.hh:18-28)."""
        m = (C.c_uint32 * 4)()
        if self.lib.ptgpu_meshes_load_obj(self.handle, str(path).encode(), m) != 0:
            raise PtgpuError("ptgpu_meshes_load_obj: %s" % self.lib.ptgpu_meshes_last_error(self.handle).decode())
        self.meshes[name] = tuple(int(x) for x in m)
        return self.meshes[name]

    def _view(self, ptr, count, dtype, cols):
        if not ptr or count == 0:
            return np.zeros((0, cols) if cols > 1 else (0,), dtype)
        buf = (C.c_uint8 * (count * cols * 4)).from_address(ptr)
        a = np.frombuffer(buf, dtype=dtype).copy()
        return a.reshape(-1, cols) if cols > 1 else a

    def arrays(self):
        """Copies of the buffers: indices (n,), pos/normal/albedo/material (v, 4) float32 (16-byte vectors)."""
        L, h = self.lib, self.handle
        ni, nv = L.ptgpu_meshes_index_count(h), L.ptgpu_meshes_vertex_count(h)
        return {"indices": self._view(L.ptgpu_meshes_indices(h), ni, np.uint32, 1),
                "pos": self._view(L.ptgpu_meshes_pos(h), nv, np.float32, 4),
                "normal": self._view(L.ptgpu_meshes_normal(h), nv, np.float32, 4),
                "albedo": self._view(L.ptgpu_meshes_albedo(h), nv, np.float32, 4),
                "material": self._view(L.ptgpu_meshes_material(h), nv, np.float32, 4)}

    def table(self):
        """(n, 4) uint32 mesh table in load order, as ptgpu_upload_meshes takes it."""
        return np.array(list(self.meshes.values()), dtype=np.uint32).reshape(-1, 4)
