"""ctypes declarations of include/ptgpu.h. Fails loudly when libptgpu.so is missing."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
CNT_NAMES = ["paths", "rays", "node_visits", "tri_tests", "blas_enters", "bounces",
             "shadow_rays", "sky_marches", "sky_attenuations", "hits", "misses"]
CNT_COUNT = 16

# every symbol include/ptgpu.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "ptgpu_default_config", "ptgpu_create", "ptgpu_destroy", "ptgpu_last_error",
    "ptgpu_upload_static", "ptgpu_set_frame", "ptgpu_set_frame_ranges",
    "ptgpu_render", "ptgpu_render_bmp", "ptgpu_bmp_size", "ptgpu_render_frame",
    "ptgpu_render_rect", "ptgpu_trace_samples", "ptgpu_tonemap", "ptgpu_trace_closest",
    "ptgpu_pcg4d", "ptgpu_render_async", "ptgpu_fetch_bgra", "ptgpu_fetch_bmp", "ptgpu_sync",
    "ptgpu_last_render_ms", "ptgpu_set_option", "ptgpu_read_counters", "ptgpu_scene_stats",
    "ptgpu_host_flatten_check", "ptgpu_get_stat", "ptgpu_validate_frame",
    "ptgpu_upload_meshes", "ptgpu_host_build_check", "ptgpu_host_flat_check", "ptgpu_debug_eval", "ptgpu_host_check_dynamic_ranges", "ptgpu_host_prepare_static", "ptgpu_warm_up",
    "ptgpu_meshes_create", "ptgpu_meshes_destroy", "ptgpu_meshes_last_error", "ptgpu_meshes_load_obj",
    "ptgpu_meshes_index_count", "ptgpu_meshes_vertex_count", "ptgpu_meshes_indices", "ptgpu_meshes_pos",
    "ptgpu_meshes_normal", "ptgpu_meshes_albedo", "ptgpu_meshes_material",
    "ptgpu_anim_create", "ptgpu_anim_destroy", "ptgpu_anim_subframe_count", "ptgpu_anim_max_instances",
    "ptgpu_anim_frame_count", "ptgpu_anim_frame", "ptgpu_set_animation_frame",
]


class LibraryMissing(RuntimeError):
    pass


class PtgpuError(RuntimeError):
    pass


class Config(C.Structure):
    """ptgpu_config: the compile-time constants of the reference's config.hh, at run time."""
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("spp", C.c_int32),
                ("max_bounces", C.c_int32), ("student_id", C.c_uint32),
                ("samples_per_subframe", C.c_int32)]

    @classmethod
    def testing(cls):      # config.hh:14-18 (as shipped)
        return cls(640, 360, 256, 4, 152121358, 8)

    @classmethod
    def production(cls):   # config.hh:21-25
        return cls(1920, 1080, 1024, 5, 152121358, 8)

    @property
    def subframes(self):
        return (self.spp + self.samples_per_subframe - 1) // self.samples_per_subframe


def lib_path():
    # PTGPU_LIB: developer override to load an experimental build of the same library
    return os.environ.get("PTGPU_LIB") or os.path.join(HERE, "libptgpu.so")


_lib = None


def load_library():
    """dlopen libptgpu.so (built in-tree by build.sh / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise LibraryMissing(
            "%s not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the render path)" % path)
    L = C.CDLL(path)
    vp, sz, u32p, f32p, u8p = C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_float), C.POINTER(C.c_uint8)
    L.ptgpu_default_config.argtypes = [C.POINTER(Config)]
    L.ptgpu_default_config.restype = None
    L.ptgpu_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(Config)]
    L.ptgpu_destroy.argtypes = [vp]
    L.ptgpu_destroy.restype = None
    L.ptgpu_last_error.argtypes = [vp]
    L.ptgpu_last_error.restype = C.c_char_p
    L.ptgpu_upload_static.argtypes = [vp, vp, sz, vp, sz, vp, sz, vp, vp, vp, vp, sz, vp, sz]
    L.ptgpu_set_frame.argtypes = [vp, vp, sz, vp, sz, vp, vp, sz, sz]
    L.ptgpu_set_frame_ranges.argtypes = [vp, vp, sz, vp, sz, vp, vp]
    L.ptgpu_render.argtypes = [vp, vp]
    L.ptgpu_render_bmp.argtypes = [vp, vp]
    L.ptgpu_bmp_size.argtypes = [vp]
    L.ptgpu_bmp_size.restype = sz
    L.ptgpu_render_frame.argtypes = [vp, vp, sz, vp, sz, vp, vp, sz, sz, vp]
    L.ptgpu_render_rect.argtypes = [vp] + [C.c_int32] * 7 + [vp, vp]
    L.ptgpu_trace_samples.argtypes = [vp, vp, vp, sz, vp]
    L.ptgpu_tonemap.argtypes = [vp, vp, sz, vp]
    L.ptgpu_trace_closest.argtypes = [vp, vp, sz, C.c_uint32, vp, vp]
    L.ptgpu_pcg4d.argtypes = [vp, vp, sz, C.c_int32]
    L.ptgpu_render_async.argtypes = [vp]
    L.ptgpu_fetch_bgra.argtypes = [vp, vp]
    L.ptgpu_fetch_bmp.argtypes = [vp, vp]
    L.ptgpu_sync.argtypes = [vp]
    L.ptgpu_last_render_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_int32)]
    L.ptgpu_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.ptgpu_read_counters.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.ptgpu_scene_stats.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.ptgpu_anim_create.argtypes = [C.POINTER(vp), vp, sz, vp, C.POINTER(Config)]
    L.ptgpu_anim_destroy.argtypes = [vp]
    L.ptgpu_anim_destroy.restype = None
    L.ptgpu_anim_subframe_count.argtypes = [vp]
    L.ptgpu_anim_subframe_count.restype = sz
    L.ptgpu_anim_max_instances.argtypes = [vp]
    L.ptgpu_anim_max_instances.restype = sz
    L.ptgpu_anim_frame_count.argtypes = [vp]
    L.ptgpu_anim_frame_count.restype = C.c_uint32
    L.ptgpu_anim_frame.argtypes = [vp, C.c_uint32, vp, vp, C.POINTER(sz), vp, vp]
    L.ptgpu_set_animation_frame.argtypes = [vp, vp, C.c_uint32]
    L.ptgpu_get_stat.argtypes = [vp, C.c_char_p, C.POINTER(C.c_uint64)]
    L.ptgpu_host_flatten_check.argtypes = [vp, sz, vp, sz, vp, sz, vp, sz, vp, sz, C.POINTER(C.c_uint64), C.c_char_p, sz]
    L.ptgpu_warm_up.argtypes = [C.c_int]
    L.ptgpu_host_prepare_static.argtypes = [vp, sz, vp, sz, vp, sz, vp, sz, vp, sz, C.c_int32, C.c_char_p, sz]
    L.ptgpu_host_check_dynamic_ranges.argtypes = [vp, vp, sz, sz, C.c_char_p, sz]
    L.ptgpu_debug_eval.argtypes = [vp, C.c_int32, vp, sz, vp]
    L.ptgpu_host_flat_check.argtypes = [vp, sz, vp, sz, vp, sz, vp, sz, vp, sz, C.POINTER(C.c_uint64), C.c_char_p, sz]
    L.ptgpu_upload_meshes.argtypes = [vp, vp, sz, vp, vp, vp, vp, sz, vp, sz, vp, sz]
    L.ptgpu_host_build_check.argtypes = [vp, sz, vp, sz, vp, sz, vp, sz, C.POINTER(C.c_uint64), C.c_char_p, sz]
    L.ptgpu_validate_frame.argtypes = [vp, vp, C.POINTER(C.c_double), C.POINTER(C.c_int32)]
    L.ptgpu_meshes_create.argtypes = [C.POINTER(vp)]
    L.ptgpu_meshes_destroy.argtypes = [vp]
    L.ptgpu_meshes_destroy.restype = None
    L.ptgpu_meshes_last_error.argtypes = [vp]
    L.ptgpu_meshes_last_error.restype = C.c_char_p
    L.ptgpu_meshes_load_obj.argtypes = [vp, C.c_char_p, vp]
    for fn in ("ptgpu_meshes_index_count", "ptgpu_meshes_vertex_count"):
        getattr(L, fn).argtypes = [vp]
        getattr(L, fn).restype = sz
    for fn in ("ptgpu_meshes_indices", "ptgpu_meshes_pos", "ptgpu_meshes_normal", "ptgpu_meshes_albedo", "ptgpu_meshes_material"):
        getattr(L, fn).argtypes = [vp]
        getattr(L, fn).restype = vp
    untyped = ("ptgpu_default_config", "ptgpu_destroy", "ptgpu_last_error", "ptgpu_bmp_size", "ptgpu_anim_destroy",
               "ptgpu_anim_subframe_count", "ptgpu_anim_max_instances", "ptgpu_anim_frame_count", "ptgpu_meshes_destroy",
               "ptgpu_meshes_last_error", "ptgpu_meshes_index_count", "ptgpu_meshes_vertex_count", "ptgpu_meshes_indices",
               "ptgpu_meshes_pos", "ptgpu_meshes_normal", "ptgpu_meshes_albedo", "ptgpu_meshes_material")
    for name in SYMBOLS:
        if name not in untyped:
            getattr(L, name).restype = C.c_int
    _lib = L
    return L
