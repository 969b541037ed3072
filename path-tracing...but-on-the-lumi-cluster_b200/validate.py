"""validate: the reference's validator.py (validator.py:1-79) without scikit-image.

    python tools/validator.py reference_directory own_directory      (same arguments, validator.py:15-17)

For every frame i in 0..1799 it reads `reference_directory/NNNN.png` (the course's half-size reference
image) and `own_directory/frame_NNNN.bmp` (what main.cc:93-101 writes), box-downscales the own frame by 2
(skimage.transform.downscale_local_mean, zero padded), truncates it to 8 bits, and reports the PSNR (data
range 255) with GOOD at >= 32 dB, in the reference's own line format; the summary also goes to
validation_result.txt. The PNG and BMP readers are the small subsets those files need (8-bit gray / RGB /
RGBA / palette, non-interlaced PNG; 24-bit bottom-up BMP as bmp.cc:15-52 writes it). For frames that are
still on the GPU use ptgpu_validate_frame (include/ptgpu.h) instead: same arithmetic, no read-back.
"""
import os
import struct
import sys
import zlib

import numpy as np

FRAME_COUNT = 1800       # validator.py:10
ACCEPT_MIN_PSNR = 32     # validator.py:11
RESIZE_FACTOR = 2        # validator.py:12


def read_png(path):
    """(H, W, C) uint8 array of an 8-bit non-interlaced PNG (C = 1, 2, 3 or 4; palette images -> RGB)."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:8] != b"\x89PNG\r\n\x1a\n":
        raise ValueError("%s: not a PNG file" % path)
    pos, idat, palette, header = 8, [], None, None
    while pos + 8 <= len(data):
        (length,), kind = struct.unpack(">I", data[pos:pos + 4]), data[pos + 4:pos + 8]
        body = data[pos + 8:pos + 8 + length]
        pos += 12 + length
        if kind == b"IHDR":
            header = struct.unpack(">IIBBBBB", body)
        elif kind == b"PLTE":
            palette = np.frombuffer(body, np.uint8).reshape(-1, 3)
        elif kind == b"IDAT":
            idat.append(body)
        elif kind == b"IEND":
            break
    if header is None:
        raise ValueError("%s: no IHDR chunk" % path)
    w, h, depth, ctype, _, _, interlace = header
    if depth != 8 or interlace != 0 or ctype not in (0, 2, 3, 4, 6):
        raise ValueError("%s: only 8-bit non-interlaced PNG is supported (depth %d, colour type %d, interlace %d)" % (path, depth, ctype, interlace))
    channels = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    stride = w * channels
    raw = zlib.decompress(b"".join(idat))
    if len(raw) != h * (stride + 1):
        raise ValueError("%s: image data has %d bytes, expected %d" % (path, len(raw), h * (stride + 1)))
    rows = np.frombuffer(raw, np.uint8).reshape(h, stride + 1)
    out = np.zeros((h, stride), np.uint8)
    prev = np.zeros(stride, np.int32)
    for y in range(h):
        ftype, line = int(rows[y, 0]), rows[y, 1:].astype(np.int32)
        if ftype == 0:
            cur = line
        elif ftype == 2:                                   # up
            cur = (line + prev) & 255
        elif ftype == 1:                                   # sub: a running sum per channel
            cur = (np.cumsum(line.reshape(-1, channels), axis=0) & 255).reshape(-1)
        else:                                              # average / Paeth depend on the reconstructed left pixel
            cur = np.zeros(stride, np.int32)
            for x in range(stride):
                a = cur[x - channels] if x >= channels else 0
                b = prev[x]
                if ftype == 3:
                    pred = (a + b) >> 1
                elif ftype == 4:
                    c = prev[x - channels] if x >= channels else 0
                    p = a + b - c
                    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                    pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                else:
                    raise ValueError("%s: unknown PNG filter %d" % (path, ftype))
                cur[x] = (line[x] + pred) & 255
        out[y] = cur
        prev = cur
    img = out.reshape(h, w, channels)
    if ctype == 3:
        if palette is None:
            raise ValueError("%s: palette image without PLTE" % path)
        img = palette[img[..., 0]]
    return img


def read_bmp(path):
    """(H, W, 3) uint8 RGB, row 0 = top, of an uncompressed 24-bit BMP (bmp.cc:15-52)."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:2] != b"BM":
        raise ValueError("%s: not a BMP file" % path)
    offset = struct.unpack("<I", data[10:14])[0]
    w, h = struct.unpack("<ii", data[18:26])
    bpp, compression = struct.unpack("<H", data[28:30])[0], struct.unpack("<I", data[30:34])[0]
    if bpp != 24 or compression != 0:
        raise ValueError("%s: only uncompressed 24-bit BMP is supported" % path)
    pitch = (w * 3 + 3) // 4 * 4
    rows = np.frombuffer(data, np.uint8, count=pitch * abs(h), offset=offset).reshape(abs(h), pitch)[:, :w * 3]
    img = rows.reshape(abs(h), w, 3)[..., ::-1]          # BGR -> RGB
    return img[::-1] if h > 0 else img                    # positive height = bottom-up


def downscale_local_mean(img, f=RESIZE_FACTOR):
    """skimage.transform.downscale_local_mean(img, (f, f, 1)): block mean, zero padded to a multiple of f."""
    h, w, c = img.shape
    ph, pw = (-h) % f, (-w) % f
    if ph or pw:
        img = np.pad(img, ((0, ph), (0, pw), (0, 0)))
        h, w = h + ph, w + pw
    return img.astype(np.float64).reshape(h // f, f, w // f, f, c).mean(axis=(1, 3))


def psnr(ref, own):
    """skimage.metrics.peak_signal_noise_ratio for uint8 images (data range 255)."""
    if ref.shape != own.shape:
        raise ValueError("Input images must have the same dimensions.")
    err = np.mean((ref.astype(np.float64) - own.astype(np.float64)) ** 2)
    return float("inf") if err == 0 else 10.0 * np.log10(255.0 * 255.0 / err)


def validate_frame(ref_img, own_img):
    """validator.py:41-52 for one frame: (psnr, good)."""
    own = downscale_local_mean(own_img).astype(np.uint8)
    if ref_img.ndim == 3 and ref_img.shape[2] == 4 and own.shape[2] == 3:
        raise ValueError("Input images must have the same dimensions.")   # what skimage says for RGBA vs RGB
    p = psnr(ref_img, own)
    return p, p >= ACCEPT_MIN_PSNR


class FrameVerdict:
    """One line of the report: a frame that is missing, or its PSNR and whether it passes."""

    def __init__(self, index, psnr_db=None):
        self.index = index
        self.psnr = psnr_db
        self.missing = psnr_db is None
        self.good = (not self.missing) and psnr_db >= ACCEPT_MIN_PSNR

    def line(self):
        # the strings are the reference tool's output format (validator.py:31-54): other scripts parse them
        name = str(self.index).zfill(4)
        if self.missing:
            return name + ": (missing image)"
        return name + ": " + str(self.psnr) + (" GOOD" if self.good else " BAD, BROKEN IMAGE?")


def validate_directory(ref_dir, own_dir, frame_count=FRAME_COUNT):
    """Verdicts for frames 0 .. frame_count-1 of `own_dir` (frame_NNNN.bmp) against `ref_dir` (NNNN.png).
    Raises FileNotFoundError naming the first reference image that does not exist."""
    for i in range(frame_count):
        ref = os.path.join(ref_dir, "%04d.png" % i)
        if not os.path.exists(ref):
            raise FileNotFoundError(ref_dir + "/" + "%04d.png" % i)
    verdicts = []
    for i in range(frame_count):
        own = os.path.join(own_dir, "frame_%04d.bmp" % i)
        if not os.path.exists(own):
            verdicts.append(FrameVerdict(i))
        else:
            verdicts.append(FrameVerdict(i, validate_frame(read_png(os.path.join(ref_dir, "%04d.png" % i)), read_bmp(own))[0]))
    return verdicts


def summary(verdicts):
    """The closing block of the report (validator.py:58-66)."""
    scored = [v.psnr for v in verdicts if not v.missing]
    ok = all(v.good for v in verdicts)
    return ("Validation result: " + ("successful" if ok else "failure") + ".\n" +
            "Sum PSNR: " + str(sum(scored) if scored else 0) + "\n" +
            "Min PSNR: " + str(min(scored) if scored else 1000) + "\n" +
            "Max PSNR: " + str(max(scored) if scored else 0) + "\n")


def main(argv=None, frame_count=FRAME_COUNT, out=sys.stdout):
    """Command line of the reference's validator.py: prints one line per frame and the summary, writes
    validation_result.txt, returns True / False (None when the run could not start)."""
    argv = sys.argv if argv is None else argv
    if len(argv) != 3:
        print("Usage: " + argv[0] + " reference_directory own_directory", file=out)
        return None
    try:
        verdicts = validate_directory(argv[1], argv[2], frame_count)
    except FileNotFoundError as e:
        print("Reference files are incomplete, quitting!!!", file=out)
        print(str(e) + " is missing.", file=out)
        return None
    body = "".join(v.line() + "\n" for v in verdicts)
    tail = summary(verdicts)
    out.write(body)
    print(tail, file=out)
    with open("validation_result.txt", "w") as f:
        f.write(body + tail)
    return all(v.good for v in verdicts)


if __name__ == "__main__":
    main()
