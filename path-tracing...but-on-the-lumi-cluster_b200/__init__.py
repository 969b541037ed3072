"""B200-native rendering hot path of the path tracer (host-side mirror of the C ABI).

The directory name is not a Python identifier; load it with `__graft_entry__.load_package()`
(alias `ptb200`). Everything here is plumbing around `libptgpu.so` (csrc/, include/ptgpu.h):
there is no CPU fallback, and nothing in this package touches `oracle/`.
"""
from .capi import (Config, LibraryMissing, PtgpuError, lib_path, load_library)  # noqa: F401
from .renderer import Renderer  # noqa: F401
from . import scene_io  # noqa: F401
from . import sharding  # noqa: F401
from . import capi  # noqa: F401
from .animation import Animation  # noqa: F401
from .meshes import MeshSet  # noqa: F401
from . import validate  # noqa: F401
