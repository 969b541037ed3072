"""Shared test helpers."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the reference's shipped output/frame_0000.bmp (frame 0, 640x360, 256 spp), copied as a fixture
GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_frame_0000.bmp")


def read_bmp_rgb(path):
    """24-bit bottom-up BMP (what bmp.cc writes) -> (H, W, 3) uint8 RGB, row 0 = top."""
    raw = np.fromfile(path, dtype=np.uint8)
    assert raw[0] == ord("B") and raw[1] == ord("M")
    off = int(np.frombuffer(raw[10:14].tobytes(), "<u4")[0])
    w = int(np.frombuffer(raw[18:22].tobytes(), "<u4")[0])
    h = int(np.frombuffer(raw[22:26].tobytes(), "<u4")[0])
    pitch = (w * 3 + 3) // 4 * 4
    px = raw[off:off + pitch * h].reshape(h, pitch)[:, :w * 3].reshape(h, w, 3)
    return px[::-1, :, ::-1].copy()
