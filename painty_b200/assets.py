"""Host-side asset preparation (the role painty's io::imRead + ScaledMat + PaddedMat play in front
of the hot path). Nothing here runs on the device; it produces the f64 blobs handed to the C ABI.

Reference behaviour mirrored (citations relative to /root/reference):
  * io::imRead(gray, convertFrom_sRGB=true)        painty/io/src/ImageIO.cxx:78-113
  * ColorConverter::srgb2rgb                         painty/core/Color.hxx:189-195
  * FootprintBrush::setRadius (width/sizeMap/pad)    painty/renderer/FootprintBrush.hxx:46-63
  * ScaledMat = cv::resize(INTER_LANCZOS4)           painty/image/Mat.hxx:141-147
  * PaddedMat                                        painty/image/Mat.hxx:116-129
"""
import functools
import math
import os

import numpy as np

_ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "assets",
                       "painty_assets.npz")


@functools.lru_cache(maxsize=1)
def _npz():
    return dict(np.load(_ASSETS))


def srgb_to_linear(s):
    """Color.hxx:189-195 — s <= 0.0404482362771082 ? s/12.92 : ((s+0.055)/1.055)^2.4 (f64)."""
    s = np.asarray(s, dtype=np.float64)
    return np.where(s <= 0.0404482362771082, s / 12.92, np.power((s + 0.055) / 1.055, 2.4))


@functools.lru_cache(maxsize=1)
def footprint_full():
    """1024x1024 f64 linearised footprint (what imRead hands to ScaledMat)."""
    return srgb_to_linear(_npz()["footprint_u8"].astype(np.float64) * (1.0 / 0xFF))


@functools.lru_cache(maxsize=1)
def thickness_map():
    """171x800 f64 linearised stroke thickness sample (BrushStrokeSample.cxx:164)."""
    return np.ascontiguousarray(srgb_to_linear(_npz()["thickness_u16"].astype(np.float64) * (1.0 / 0xFFFF)))


def palette(name="lindemeier_measured"):
    """(K[n,3], S[n,3]) f64 of one of the shipped palettes."""
    z = _npz()
    return z[name + "_K"].copy(), z[name + "_S"].copy()


def footprint_geometry(radius):
    """(width, size_map, pad, footprint_side) exactly as FootprintBrush::setRadius computes them."""
    width = int(2.0 * math.ceil(radius) + 1.0)
    size_map = int(math.ceil(math.sqrt(2.0) * width))
    pad = (size_map - width) // 2
    return width, size_map, pad, width + 2 * pad


def is_safe_radius(radius):
    """True when the padded footprint is as wide as the pickup map (no OOB read)."""
    _, size_map, _, side = footprint_geometry(radius)
    return side == size_map


def snap_to_safe_radius(radius):
    """Nearest radius (searching outward in steps of 1) whose footprint is OOB-free."""
    r = float(radius)
    for k in range(0, 64):
        for cand in (r + k, r - k):
            if cand >= 1.0 and is_safe_radius(cand):
                return cand
    raise ValueError("no safe radius near %r" % radius)


@functools.lru_cache(maxsize=64)
def scaled_footprint(width):
    """cv2.resize(footprint_full, (width,width), INTER_LANCZOS4) on the f64 array."""
    import cv2

    return np.ascontiguousarray(cv2.resize(footprint_full(), (width, width), interpolation=cv2.INTER_LANCZOS4))


def baked_footprint(radius):
    """The padded (side x side) f64 footprint FootprintBrush::setRadius builds for `radius`."""
    width, _, pad, side = footprint_geometry(radius)
    out = np.zeros((side, side), dtype=np.float64)
    out[pad:pad + width, pad:pad + width] = scaled_footprint(width)
    return out
