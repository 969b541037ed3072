"""TEST INFRASTRUCTURE — ctypes binding of the oracle (oracle/_ref/libptref*.so).

The oracle is the UNMODIFIED reference renderer compiled by oracle/Makefile from
/root/reference/{bvh,mesh,scene,bmp}.cc + oracle/ref_harness.cc. It is the checker for the
CUDA path and the CPU baseline of bench.py. Product code must never import this module:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

VARIANTS = {
    "fast": "libptref.so",            # reference Makefile flags, shipped TESTING config
    "strict": "libptref_strict.so",   # -O2, no fast-math (noise floor / non-AVX2 fallback)
    "prod": "libptref_prod.so",       # production config.hh:21-25
    "mb": "libptref_mb.so",           # configs[4]: 640x360 at 1024 spp
}


class RefConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("width", "height", "spp", "max_bounces")] + \
               [("student_id", C.c_uint32)] + \
               [(n, C.c_int32) for n in (
                   "samples_per_subframe", "subframe_count", "framerate",
                   "sz_bvh", "sz_bvh_node", "sz_bvh_link", "sz_tlas_instance", "sz_mesh",
                   "sz_subframe", "sz_camera", "sz_light", "sz_float3", "sz_float4")]


class RefSceneView(C.Structure):
    _fields_ = [
        ("nodes", C.c_void_p), ("n_nodes", C.c_uint64),
        ("links", C.c_void_p), ("n_links", C.c_uint64),
        ("indices", C.c_void_p), ("n_indices", C.c_uint64),
        ("pos", C.c_void_p), ("normal", C.c_void_p), ("albedo", C.c_void_p),
        ("material", C.c_void_p), ("n_verts", C.c_uint64),
        ("instances", C.c_void_p), ("n_instances", C.c_uint64),
        ("n_static_instances", C.c_uint64),
        ("subframes", C.c_void_p), ("n_subframes", C.c_uint64),
        ("n_static_nodes", C.c_uint64),
    ]


def available(variant="fast"):
    return os.path.exists(os.path.join(REF_DIR, VARIANTS[variant])) and \
        os.path.exists(os.path.join(REF_DIR, "data", "terrain.obj"))


def _cpu_has_v3():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = set(line.split(":", 1)[1].split())
                    return {"avx2", "fma", "bmi2", "movbe", "f16c"} <= flags
    except OSError:
        pass
    return False


def _as_np(ptr, nbytes, dtype):
    if not ptr or nbytes == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_uint8 * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype)


class Oracle:
    """One loaded reference renderer (one `scene` object inside the .so)."""

    def __init__(self, variant="fast"):
        if variant != "strict" and not _cpu_has_v3():
            # the fast builds target x86-64-v3; a host without AVX2/FMA gets the strict build
            if variant != "fast":
                raise RuntimeError("oracle variant %r needs an x86-64-v3 CPU" % variant)
            variant = "strict"
        path = os.path.join(REF_DIR, VARIANTS[variant])
        if not os.path.exists(path):
            raise FileNotFoundError(
                "%s missing: run `make -C oracle` where /root/reference is mounted" % path)
        self.variant = variant
        # RTLD_LOCAL + a private copy of the globals per .so: variants can coexist in one process
        self.lib = C.CDLL(path, mode=os.RTLD_LOCAL) if hasattr(os, "RTLD_LOCAL") else C.CDLL(path)
        L = self.lib
        L.ref_get_config.argtypes = [C.POINTER(RefConfig)]
        L.ref_load_scene.argtypes = [C.c_char_p]
        L.ref_load_scene.restype = C.c_int
        L.ref_frame_count.restype = C.c_uint32
        L.ref_setup_frame.argtypes = [C.c_uint32]
        L.ref_setup_frame.restype = C.c_int
        L.ref_get_view.argtypes = [C.POINTER(RefSceneView)]
        L.ref_find_mesh.argtypes = [C.c_char_p, C.POINTER(C.c_uint32)]
        L.ref_find_mesh.restype = C.c_int
        L.ref_trace_sample.argtypes = [C.c_uint32, C.c_uint32, C.c_int32, C.POINTER(C.c_float)]
        L.ref_render_rect.argtypes = [C.c_int32] * 7 + [C.c_void_p, C.c_void_p, C.c_int32]
        L.ref_tonemap.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_uint8)]
        L.ref_pcg4d.argtypes = [C.POINTER(C.c_uint32)]
        L.ref_rand4.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_float)]
        L.ref_trace_closest.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float,
                                        C.c_float, C.c_uint32, C.POINTER(C.c_float),
                                        C.POINTER(C.c_uint32)]
        L.ref_setup_stress_scene.restype = C.c_int
        L.ref_restore_scene.restype = C.c_int
        L.ref_write_bmp.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.ref_eval.argtypes = [C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]
        L.ref_eval.restype = C.c_int
        self.config = RefConfig()
        L.ref_get_config(C.byref(self.config))
        self.loaded = False
        self.frame = None

    # -- scene ---------------------------------------------------------------------------
    def load_scene(self):
        if not self.loaded:
            rc = self.lib.ref_load_scene(REF_DIR.encode())
            if rc != 0:
                raise RuntimeError("ref_load_scene failed: %d" % rc)
            self.loaded = True
        return self

    def setup_frame(self, frame):
        self.load_scene()
        self.lib.ref_setup_frame(int(frame))
        self.frame = int(frame)
        return self.view()

    def setup_stress_scene(self):
        """BASELINE.json configs[3]: terrain + buddha + dragon + armadillo, fixed camera (ref_harness.cc).
        The view then has 1 static instance and 3 dynamic ones. Call restore_scene() afterwards."""
        self.load_scene()
        if self.lib.ref_setup_stress_scene() != 0:
            raise RuntimeError("ref_setup_stress_scene failed")
        self.frame = "stress"
        return self.view()

    def restore_scene(self):
        self.lib.ref_restore_scene()
        self.frame = None

    def frame_count(self):
        return int(self.lib.ref_frame_count())

    def view(self):
        """Raw pointers + numpy views (no copies) of the arrays crossing the seam (main.cc:29-37).
        Views are valid until the next setup_frame (scene.cc:274-277 reallocates)."""
        v = RefSceneView()
        self.lib.ref_get_view(C.byref(v))
        c = self.config
        return {
            "raw": v,
            "nodes": _as_np(v.nodes, v.n_nodes * c.sz_bvh_node, np.float32).reshape(-1, 6),
            "links": _as_np(v.links, v.n_links * c.sz_bvh_link, np.uint32).reshape(-1, 2),
            "indices": _as_np(v.indices, v.n_indices * 4, np.uint32),
            "pos": _as_np(v.pos, v.n_verts * 16, np.float32).reshape(-1, 4),
            "normal": _as_np(v.normal, v.n_verts * 16, np.float32).reshape(-1, 4),
            "albedo": _as_np(v.albedo, v.n_verts * 16, np.float32).reshape(-1, 4),
            "material": _as_np(v.material, v.n_verts * 16, np.float32).reshape(-1, 4),
            "instances": _as_np(v.instances, v.n_instances * c.sz_tlas_instance, np.uint8)
            .reshape(-1, c.sz_tlas_instance),
            "subframes": _as_np(v.subframes, v.n_subframes * c.sz_subframe, np.uint8)
            .reshape(-1, c.sz_subframe),
            "n_static_instances": int(v.n_static_instances),
            "n_static_nodes": int(v.n_static_nodes),
        }

    def find_mesh(self, name):
        out = (C.c_uint32 * 6)()
        if self.lib.ref_find_mesh(name.encode(), out) != 0:
            raise KeyError(name)
        return list(out)

    # -- compute -------------------------------------------------------------------------
    def trace_sample(self, x, y, sample_index):
        out = (C.c_float * 3)()
        self.lib.ref_trace_sample(x, y, sample_index, out)
        return np.array(out[:], dtype=np.float32)

    def render_rect(self, x0, y0, w, h, s_begin, s_count, s_stride=1, nthreads=0, tonemap=True):
        rgb = np.empty((h, w, 3), dtype=np.float32)
        bgra = np.empty((h, w, 4), dtype=np.uint8) if tonemap else None
        self.lib.ref_render_rect(x0, y0, w, h, s_begin, s_count, s_stride,
                                 rgb.ctypes.data, bgra.ctypes.data if tonemap else None, nthreads)
        return rgb, bgra

    def render_frame(self, spp=None, nthreads=0):
        c = self.config
        return self.render_rect(0, 0, c.width, c.height, 0, spp or c.spp, 1, nthreads)

    def tonemap(self, rgb):
        a = (C.c_float * 3)(*[float(x) for x in rgb])
        o = (C.c_uint8 * 4)()
        self.lib.ref_tonemap(a, o)
        return tuple(o[:])

    def pcg4d(self, state):
        s = (C.c_uint32 * 4)(*[int(x) for x in state])
        self.lib.ref_pcg4d(s)
        return tuple(s[:])

    def rand4(self, state):
        s = (C.c_uint32 * 4)(*[int(x) for x in state])
        f = (C.c_float * 4)()
        self.lib.ref_rand4(s, f)
        return tuple(s[:]), np.array(f[:], dtype=np.float32)

    def trace_closest(self, o, d, tmin=0.0, tmax=1e9, subframe=0):
        of = (C.c_float * 3)(*[float(x) for x in o])
        df = (C.c_float * 3)(*[float(x) for x in d])
        rf = (C.c_float * 4)()
        ru = (C.c_uint32 * 3)()
        self.lib.ref_trace_closest(of, df, tmin, tmax, subframe, rf, ru)
        return {"thit": rf[0], "bary": (rf[1], rf[2], rf[3]), "instance": ru[0],
                "primitive": ru[1], "back_face": bool(ru[2])}

    def eval(self, fn, inputs):
        """One reference sub-function per row of `inputs` (n x <=24 floats; uint32 values as float bits):
        the PTGPU_FN_* codes and layouts of include/ptgpu.h. Returns n x 32 floats."""
        a = np.zeros((len(inputs), 24), np.float32)
        src = np.asarray(inputs, np.float32)
        a[:, :src.shape[1]] = src
        out = np.zeros((len(inputs), 32), np.float32)
        rc = self.lib.ref_eval(int(fn), a.ctypes.data, len(a), out.ctypes.data)
        if rc != 0:
            raise RuntimeError("ref_eval(%d) failed: %d" % (fn, rc))
        return out

    def write_bmp(self, path, bgra):
        h, w = bgra.shape[:2]
        a = np.ascontiguousarray(bgra)
        self.lib.ref_write_bmp(path.encode(), w, h, a.ctypes.data)


_cache = {}


def get(variant="fast"):
    """Process-wide oracle per variant (load_scene costs seconds)."""
    if variant not in _cache:
        _cache[variant] = Oracle(variant)
    return _cache[variant]
