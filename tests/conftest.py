import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (runs under gpurun / at round end)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as ge
    return ge.load_package()


@pytest.fixture(scope="session")
def oracle():
    """The unmodified reference compiled into oracle/_ref (fast build, shipped TESTING config)."""
    from oracle import refbind
    if not refbind.available("fast"):
        pytest.skip("oracle/_ref not built (make -C oracle, needs /root/reference)")
    o = refbind.get("fast")
    o.load_scene()
    return o


@pytest.fixture(scope="session")
def oracle_strict():
    from oracle import refbind
    if not refbind.available("strict"):
        pytest.skip("oracle/_ref strict build missing")
    o = refbind.get("strict")
    o.load_scene()
    return o


@pytest.fixture(scope="session")
def renderer(pkg, oracle):
    """One context on cuda:0 with the static scene of the oracle uploaded through the C ABI."""
    view = oracle.setup_frame(0)
    r = pkg.Renderer(pkg.Config.testing(), device=0)
    r.upload_static(**pkg.scene_io.static_from_view(view))
    yield r
    r.close()


class FrameCache:
    """setup_animation_frame through the oracle + ptgpu_set_frame, once per requested frame."""

    def __init__(self, pkg, oracle, renderer):
        self.pkg, self.oracle, self.renderer = pkg, oracle, renderer
        self.current = None

    def use(self, frame):
        if self.current != frame or self.oracle.frame != frame:
            view = self.oracle.setup_frame(frame)
            self.renderer.set_frame(**self.pkg.scene_io.frame_from_view(view))
            self.current = frame
        return self.renderer


@pytest.fixture(scope="session")
def frames(pkg, oracle, renderer):
    return FrameCache(pkg, oracle, renderer)
