"""Host-side asset preparation (the role painty's io::imRead + ScaledMat + PaddedMat play in front
of the hot path). Nothing here runs on the device; it produces the f64 blobs handed to the C ABI.

Reference behaviour mirrored (citations relative to /root/reference):
  * io::imRead(gray, convertFrom_sRGB=true)        painty/io/src/ImageIO.cxx:78-113
  * ColorConverter::srgb2rgb                         painty/core/Color.hxx:189-195
  * FootprintBrush::setRadius (width/sizeMap/pad)    painty/renderer/FootprintBrush.hxx:46-63
  * ScaledMat = cv::resize(INTER_LANCZOS4)           painty/image/Mat.hxx:141-147
  * PaddedMat                                        painty/image/Mat.hxx:116-129
  * TextureBrushDictionary::loadHeightMap            painty/renderer/src/TextureBrushDictionary.cxx:71-79
  * CanvasGpu::clear (canvas pattern substrate)      painty/renderer/src/CanvasGpu.cxx:27-40
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


_TEXTURES = os.path.join(os.path.dirname(_ASSETS), "painty_textures.npz")


@functools.lru_cache(maxsize=1)
def brush_textures():
    """The brush-texture dictionary's inputs: list of (name, size key, length key, height map f64) for data/textures in
    file-name order. loadHeightMap = imRead(gray, no sRGB) = u16 * (1/0xffff) as f64, then cv::normalize(NORM_MINMAX, 0, 1)
    (TextureBrushDictionary.cxx:71-79); evaluated with the container's cv2 like the LANCZOS4 footprints."""
    import cv2

    z = np.load(_TEXTURES)
    out = []
    for key in sorted(k for k in z.files if k.startswith("tex_")):
        u16 = z[key].astype(np.uint16) * np.uint16(257)
        gray = u16.astype(np.float64) * (1.0 / 0xFFFF)
        gray = cv2.normalize(gray, None, 0.0, 1.0, cv2.NORM_MINMAX)
        tok = key[4:].split("_")
        out.append((key[4:], int(tok[0]), int(tok[1]), np.ascontiguousarray(gray)))
    return out


def canvas_pattern(rows, cols):
    """Substrate reflectance R0 of the sbr renderer's canvas (CanvasGpu.cxx:27-40): canvas_patterns/0.png read as linear RGB
    (u8 / 0xff, srgb2rgb), converted to float32, LANCZOS4-scaled to the canvas size (ScaledMat); returned as f64 [rows,cols,3]."""
    import cv2

    pat = np.load(_TEXTURES)["canvas_pattern_u8"]
    lin = srgb_to_linear(pat.astype(np.float64) * (1.0 / 0xFF)).astype(np.float32)
    if lin.shape[:2] != (rows, cols):
        lin = cv2.resize(lin, (cols, rows), interpolation=cv2.INTER_LANCZOS4)
    return np.ascontiguousarray(lin.astype(np.float64))


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
