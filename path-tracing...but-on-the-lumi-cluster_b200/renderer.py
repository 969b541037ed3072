"""Renderer: host-side mirror of the reference's render seam (main.cc:82-101) over the C ABI.

    r = Renderer(Config.testing(), device=0)
    r.upload_static(nodes, links, indices, pos, normal, albedo, material, instances)   # load_scene()
    r.set_frame(subframes, dyn_instances, tlas_nodes, tlas_links, tlas_node_base)      # setup_animation_frame()
    bgra = r.render()                                                                  # baseline_render()

All arrays are numpy arrays in the reference's memory layout (include/ptgpu.h); nothing is
computed on the host.
"""
import ctypes as C

import numpy as np

from .capi import CNT_COUNT, CNT_NAMES, Config, PtgpuError, load_library


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _c(a, dtype=None):
    a = np.ascontiguousarray(a) if dtype is None else np.ascontiguousarray(a, dtype=dtype)
    return a


class Renderer:
    def __init__(self, config=None, device=0):
        self.lib = load_library()
        self.config = config or Config.testing()
        self.ctx = C.c_void_p()
        rc = self.lib.ptgpu_create(C.byref(self.ctx), device, C.byref(self.config))
        if rc != 0:
            msg = self.lib.ptgpu_last_error(None).decode()
            self.ctx = C.c_void_p()
            raise PtgpuError("ptgpu_create: " + msg)
        self.device = device

    # -- plumbing ------------------------------------------------------------------------
    def _check(self, rc, what):
        if rc != 0:
            raise PtgpuError("%s: %s" % (what, self.lib.ptgpu_last_error(self.ctx).decode()))

    def close(self):
        if getattr(self, "ctx", None) is not None and self.ctx:
            self.lib.ptgpu_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key, value):
        self._check(self.lib.ptgpu_set_option(self.ctx, key.encode(), int(value)), "ptgpu_set_option")

    # -- scene ---------------------------------------------------------------------------
    def upload_static(self, nodes, links, indices, pos, normal, albedo, material, instances):
        """Everything load_scene() produces (scene.cc:135-269); `instances` = the static ones."""
        nodes = _c(nodes, np.float32).reshape(-1, 6)
        links = _c(links, np.uint32).reshape(-1, 2)
        indices = _c(indices, np.uint32)
        pos, normal = _c(pos, np.float32).reshape(-1, 4), _c(normal, np.float32).reshape(-1, 4)
        albedo, material = _c(albedo, np.float32).reshape(-1, 4), _c(material, np.float32).reshape(-1, 4)
        instances = _c(instances, np.uint8).reshape(-1, 160)
        self._check(self.lib.ptgpu_upload_static(
            self.ctx, _ptr(nodes), nodes.shape[0], _ptr(links), links.shape[0],
            _ptr(indices), indices.shape[0], _ptr(pos), _ptr(normal), _ptr(albedo), _ptr(material),
            pos.shape[0], _ptr(instances), instances.shape[0]), "ptgpu_upload_static")
        self.n_static_nodes = nodes.shape[0]
        self.n_static = instances.shape[0]

    def upload_meshes(self, indices, pos, normal, albedo, material, meshes, instances):
        """ptgpu_upload_meshes: the static scene without the reference's BVH arrays; `meshes` is an
        (n, 4) uint32 table of (vertex_count, triangle_count, index_offset, base_vertex_offset)."""
        indices = _c(indices, np.uint32)
        pos, normal = _c(pos, np.float32).reshape(-1, 4), _c(normal, np.float32).reshape(-1, 4)
        albedo, material = _c(albedo, np.float32).reshape(-1, 4), _c(material, np.float32).reshape(-1, 4)
        meshes = _c(meshes, np.uint32).reshape(-1, 4)
        instances = _c(instances, np.uint8).reshape(-1, 160)
        self._check(self.lib.ptgpu_upload_meshes(
            self.ctx, _ptr(indices), indices.shape[0], _ptr(pos), _ptr(normal), _ptr(albedo), _ptr(material),
            pos.shape[0], _ptr(meshes), meshes.shape[0], _ptr(instances), instances.shape[0]), "ptgpu_upload_meshes")
        self.n_static_nodes = 0
        self.n_static = instances.shape[0]

    def set_frame(self, subframes, dyn_instances, tlas_nodes, tlas_links, tlas_node_base=None):
        """Everything setup_animation_frame() produces (scene.cc:271-718)."""
        subframes = _c(subframes, np.uint8).reshape(-1, 160)
        dyn = _c(dyn_instances, np.uint8).reshape(-1, 160)
        tn = _c(tlas_nodes, np.float32).reshape(-1, 6)
        tl = _c(tlas_links, np.uint32).reshape(-1, 2)
        base = self.n_static_nodes if tlas_node_base is None else tlas_node_base
        self._check(self.lib.ptgpu_set_frame(
            self.ctx, _ptr(subframes), subframes.shape[0], _ptr(dyn), dyn.shape[0],
            _ptr(tn), _ptr(tl), tn.shape[0], base), "ptgpu_set_frame")

    def set_frame_ranges(self, subframes, dyn_instances, dyn_begin, dyn_end):
        subframes = _c(subframes, np.uint8).reshape(-1, 160)
        dyn = _c(dyn_instances, np.uint8).reshape(-1, 160)
        b, e = _c(dyn_begin, np.uint32), _c(dyn_end, np.uint32)
        self._check(self.lib.ptgpu_set_frame_ranges(
            self.ctx, _ptr(subframes), subframes.shape[0], _ptr(dyn), dyn.shape[0], _ptr(b), _ptr(e)),
            "ptgpu_set_frame_ranges")

    # -- render --------------------------------------------------------------------------
    def render(self, out=None):
        """baseline_render (main.cc:12-46): (H, W, 4) uint8 BGRA, row 0 = top."""
        c = self.config
        out = np.empty((c.height, c.width, 4), np.uint8) if out is None else out
        self._check(self.lib.ptgpu_render(self.ctx, _ptr(out)), "ptgpu_render")
        return out

    def render_bmp(self, out=None):
        """baseline_render + write_bmp's packing (bmp.cc:15-52): the complete BMP file bytes."""
        n = self.lib.ptgpu_bmp_size(self.ctx)
        out = np.empty(n, np.uint8) if out is None else out
        self._check(self.lib.ptgpu_render_bmp(self.ctx, _ptr(out)), "ptgpu_render_bmp")
        return out

    def render_frame(self, subframes, dyn_instances, tlas_nodes, tlas_links, out=None):
        c = self.config
        out = np.empty((c.height, c.width, 4), np.uint8) if out is None else out
        subframes = _c(subframes, np.uint8).reshape(-1, 160)
        dyn = _c(dyn_instances, np.uint8).reshape(-1, 160)
        tn = _c(tlas_nodes, np.float32).reshape(-1, 6)
        tl = _c(tlas_links, np.uint32).reshape(-1, 2)
        self._check(self.lib.ptgpu_render_frame(
            self.ctx, _ptr(subframes), subframes.shape[0], _ptr(dyn), dyn.shape[0],
            _ptr(tn), _ptr(tl), tn.shape[0], self.n_static_nodes, _ptr(out)), "ptgpu_render_frame")
        return out

    def render_rect(self, x0, y0, w, h, s_begin, s_count, s_stride=1, tonemap=True):
        rgb = np.empty((h, w, 3), np.float32)
        bgra = np.empty((h, w, 4), np.uint8) if tonemap else None
        self._check(self.lib.ptgpu_render_rect(self.ctx, x0, y0, w, h, s_begin, s_count, s_stride,
                                               _ptr(rgb), _ptr(bgra)), "ptgpu_render_rect")
        return rgb, bgra

    def trace_samples(self, xy, sample_index):
        xy = _c(xy, np.uint32).reshape(-1, 2)
        si = _c(sample_index, np.int32).reshape(-1)
        out = np.empty((xy.shape[0], 3), np.float32)
        self._check(self.lib.ptgpu_trace_samples(self.ctx, _ptr(xy), _ptr(si), xy.shape[0], _ptr(out)),
                    "ptgpu_trace_samples")
        return out

    def tonemap(self, rgb):
        rgb = _c(rgb, np.float32).reshape(-1, 3)
        out = np.empty((rgb.shape[0], 4), np.uint8)
        self._check(self.lib.ptgpu_tonemap(self.ctx, _ptr(rgb), rgb.shape[0], _ptr(out)), "ptgpu_tonemap")
        return out

    def trace_closest(self, rays, subframe=0):
        rays = _c(rays, np.float32).reshape(-1, 8)
        of = np.empty((rays.shape[0], 4), np.float32)
        ou = np.empty((rays.shape[0], 3), np.uint32)
        self._check(self.lib.ptgpu_trace_closest(self.ctx, _ptr(rays), rays.shape[0], subframe, _ptr(of), _ptr(ou)),
                    "ptgpu_trace_closest")
        return of, ou

    def pcg4d(self, states, steps=1):
        s = np.array(states, dtype=np.uint32).reshape(-1, 4).copy()
        self._check(self.lib.ptgpu_pcg4d(self.ctx, _ptr(s), s.shape[0], steps), "ptgpu_pcg4d")
        return s

    def debug_eval(self, fn, inputs):
        """ptgpu_debug_eval: one device function of the path per row of `inputs` (n x <= 24 floats, uint32
        values as float bits; PTGPU_FN_* of include/ptgpu.h). Returns n x 32 floats."""
        src = np.asarray(inputs, np.float32)
        a = np.zeros((src.shape[0], 24), np.float32)
        a[:, :src.shape[1]] = src
        out = np.zeros((src.shape[0], 32), np.float32)
        self._check(self.lib.ptgpu_debug_eval(self.ctx, int(fn), _ptr(a), a.shape[0], _ptr(out)), "ptgpu_debug_eval")
        return out

    # -- device-resident rendering ---------------------------------------------------------
    def render_async(self):
        self._check(self.lib.ptgpu_render_async(self.ctx), "ptgpu_render_async")

    def sync(self):
        self._check(self.lib.ptgpu_sync(self.ctx), "ptgpu_sync")

    def fetch_bgra(self, out=None):
        c = self.config
        out = np.empty((c.height, c.width, 4), np.uint8) if out is None else out
        self._check(self.lib.ptgpu_fetch_bgra(self.ctx, _ptr(out)), "ptgpu_fetch_bgra")
        return out

    def fetch_bmp(self, out=None):
        n = self.lib.ptgpu_bmp_size(self.ctx)
        out = np.empty(n, np.uint8) if out is None else out
        self._check(self.lib.ptgpu_fetch_bmp(self.ctx, _ptr(out)), "ptgpu_fetch_bmp")
        return out

    def validate_frame(self, ref_rgb_half):
        """validator.py on the device-resident frame: (psnr, good) against the half-size RGB reference."""
        ref = np.ascontiguousarray(ref_rgb_half, dtype=np.uint8)
        hh, hw = (self.config.height + 1) // 2, (self.config.width + 1) // 2
        if ref.shape != (hh, hw, 3):
            raise ValueError("reference image must be %dx%dx3 uint8, got %r" % (hh, hw, ref.shape))
        psnr = C.c_double()
        good = C.c_int32()
        self._check(self.lib.ptgpu_validate_frame(self.ctx, _ptr(ref), C.byref(psnr), C.byref(good)), "ptgpu_validate_frame")
        return psnr.value, bool(good.value)

    def last_render_ms(self):
        ms = C.c_float()
        n = C.c_int32()
        self._check(self.lib.ptgpu_last_render_ms(self.ctx, C.byref(ms), C.byref(n)), "ptgpu_last_render_ms")
        return ms.value, n.value

    def read_counters(self):
        out = (C.c_uint64 * CNT_COUNT)()
        self._check(self.lib.ptgpu_read_counters(self.ctx, out), "ptgpu_read_counters")
        return {n: int(out[i]) for i, n in enumerate(CNT_NAMES)}

    def get_stat(self, key):
        out = C.c_uint64()
        self._check(self.lib.ptgpu_get_stat(self.ctx, key.encode(), C.byref(out)), "ptgpu_get_stat")
        return int(out.value)

    def scene_stats(self):
        out = (C.c_uint64 * 8)()
        self._check(self.lib.ptgpu_scene_stats(self.ctx, out), "ptgpu_scene_stats")
        keys = ["device_scene_bytes", "frame_bytes", "wide_nodes", "triangles", "static_instances",
                "tlas_nodes", "reference_layout_bytes", "blas_count"]
        return {k: int(out[i]) for i, k in enumerate(keys)}
