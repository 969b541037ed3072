"""TEST INFRASTRUCTURE — numpy restatement of the reference's validator.py (skimage is not installed).

Follows validator.py:41-52 literally: the own frame is box-downscaled by RESIZE_FACTOR (2) in x
and y (skimage.transform.downscale_local_mean = mean over non-overlapping blocks), cast to uint8
by truncation (astype), compared with the reference image by PSNR with data range 255
(skimage.metrics.peak_signal_noise_ratio for uint8 inputs), and flagged BAD below 32 dB.
The course's reference PNGs are not in the repo; the tests feed it oracle-rendered frames instead.
"""
import numpy as np

ACCEPT_MIN_PSNR = 32   # validator.py:11
RESIZE_FACTOR = 2      # validator.py:12


def downscale_local_mean(img, f=RESIZE_FACTOR):
    h, w, c = img.shape
    ph, pw = (-h) % f, (-w) % f
    if ph or pw:  # skimage pads with zeros (cval=0) up to a multiple of the factor
        img = np.pad(img, ((0, ph), (0, pw), (0, 0)))
        h, w = h + ph, w + pw
    return img.astype(np.float64).reshape(h // f, f, w // f, f, c).mean(axis=(1, 3))


def psnr_uint8(ref, own):
    err = np.mean((ref.astype(np.float64) - own.astype(np.float64)) ** 2)
    if err == 0:
        return float("inf")
    return 10.0 * np.log10(255.0 * 255.0 / err)


def validate_frame(ref_half_rgb, own_rgb):
    """ref_half_rgb: (H/2, W/2, 3) uint8 reference image; own_rgb: (H, W, 3) uint8 own frame.
    Returns (psnr, good) as validator.py:43-52 computes them."""
    own = downscale_local_mean(own_rgb).astype(np.uint8)
    p = psnr_uint8(ref_half_rgb, own)
    return p, p >= ACCEPT_MIN_PSNR


def make_reference_png_array(full_rgb):
    """What a course reference PNG would hold for a frame rendered at full size: the half-size
    block mean rounded to 8 bits (the PNGs are half the size of the rendered frames, validator.py:44)."""
    return np.clip(np.rint(downscale_local_mean(full_rgb)), 0, 255).astype(np.uint8)
