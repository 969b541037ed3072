"""The package's validator (validate.py: validator.py without scikit-image) against the numpy restatement
in oracle/validator_np.py, its PNG / BMP readers against files written here, and its report format."""
import io
import os
import struct
import zlib

import numpy as np
import pytest

from helpers import GOLDEN, read_bmp_rgb


def write_png(path, img, filters):
    """Minimal PNG writer for the tests: 8-bit, non-interlaced, one filter type per row from `filters`."""
    h, w, c = img.shape
    ctype = {1: 0, 2: 4, 3: 2, 4: 6}[c]
    raw = bytearray()
    prev = np.zeros(w * c, np.int32)
    for y in range(h):
        line = img[y].reshape(-1).astype(np.int32)
        f = filters[y % len(filters)]
        left = np.concatenate([np.zeros(c, np.int32), line[:-c]])
        upleft = np.concatenate([np.zeros(c, np.int32), prev[:-c]])
        if f == 0:
            enc = line
        elif f == 1:
            enc = line - left
        elif f == 2:
            enc = line - prev
        elif f == 3:
            enc = line - ((left + prev) >> 1)
        else:
            p = left + prev - upleft
            pa, pb, pc = np.abs(p - left), np.abs(p - prev), np.abs(p - upleft)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, upleft))
            enc = line - pred
        raw.append(f)
        raw += bytes((enc & 255).astype(np.uint8))
        prev = line

    def chunk(kind, body):
        return struct.pack(">I", len(body)) + kind + body + struct.pack(">I", zlib.crc32(kind + body) & 0xFFFFFFFF)
    data = zlib.compress(bytes(raw))
    with open(path, "wb") as fh:
        fh.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ctype, 0, 0, 0)) +
                 chunk(b"IDAT", data[:len(data) // 2]) + chunk(b"IDAT", data[len(data) // 2:]) + chunk(b"IEND", b""))


def test_png_reader_all_filters(pkg, tmp_path):
    rng = np.random.RandomState(5)
    for c in (1, 3, 4):
        img = rng.randint(0, 256, (23, 17, c)).astype(np.uint8)
        for filters in ([0], [1], [2], [3], [4], [0, 1, 2, 3, 4]):
            p = tmp_path / ("t%d_%s.png" % (c, "".join(map(str, filters))))
            write_png(p, img, filters)
            assert np.array_equal(pkg.validate.read_png(str(p)), img), (c, filters)
    with pytest.raises(ValueError):
        (tmp_path / "bad.png").write_bytes(b"not a png")
        pkg.validate.read_png(str(tmp_path / "bad.png"))


def test_bmp_reader_matches_the_golden_frame(pkg):
    assert np.array_equal(pkg.validate.read_bmp(GOLDEN), read_bmp_rgb(GOLDEN))


def test_validator_numbers_and_report_format(pkg, tmp_path, monkeypatch):
    from oracle import validator_np as V
    gold = read_bmp_rgb(GOLDEN)
    ref_dir, own_dir = tmp_path / "ref", tmp_path / "own"
    ref_dir.mkdir(); own_dir.mkdir()
    rng = np.random.RandomState(9)
    half = V.make_reference_png_array(gold)
    noisy = np.clip(half.astype(np.int32) + rng.randint(-60, 61, half.shape), 0, 255).astype(np.uint8)
    write_png(ref_dir / "0000.png", half, [4])
    write_png(ref_dir / "0001.png", noisy, [1, 2])
    write_png(ref_dir / "0002.png", half, [0])
    for i in (0, 1):                                            # frame 2 is missing on the own side
        with open(GOLDEN, "rb") as src, open(own_dir / ("frame_%04d.bmp" % i), "wb") as dst:
            dst.write(src.read())
    monkeypatch.chdir(tmp_path)
    out = io.StringIO()
    ok = pkg.validate.main(["validator.py", str(ref_dir), str(own_dir)], frame_count=3, out=out)
    lines = out.getvalue().splitlines()
    want0, good0 = V.validate_frame(half, gold)
    want1, good1 = V.validate_frame(noisy, gold)
    assert good0 and not good1 and ok is False
    assert lines[0] == "0000: " + str(want0) + " GOOD"           # validator.py:48-54
    assert lines[1] == "0001: " + str(want1) + " BAD, BROKEN IMAGE?"
    assert lines[2] == "0002: (missing image)"
    assert "Validation result: failure." in lines and ("Min PSNR: " + str(min(want0, want1))) in lines
    assert (tmp_path / "validation_result.txt").read_text().startswith("0000: ")
    # an incomplete reference directory stops the run as in validator.py:34-37
    out = io.StringIO()
    assert pkg.validate.main(["validator.py", str(ref_dir), str(own_dir)], frame_count=5, out=out) is None
    assert "Reference files are incomplete, quitting!!!" in out.getvalue()
