"""TEST INFRASTRUCTURE ONLY. ctypes front end for the two CPU checkers.

  Cpu("port")  -> oracle/libpainty_oracle.so   (our restatement; always available after `make`)
  Cpu("ref")   -> oracle/_ref/libpainty_ref.so (unmodified reference headers; built only where
                  /root/reference exists, travels prebuilt to the GPU box)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. Both back ends expose the same methods; images are AoS float64 numpy arrays
(K,S,R: [rows, cols, 3]; V,h: [rows, cols]) — the reference's own boundary layout.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PD = C.POINTER(C.c_double)
FOOTPRINT_PATH = "./data/footprint/footprint.png"  # FootprintBrush.hxx:49
SAMPLE_DIR = "data/sample_0"  # BrushStrokeSample.cxx:17


def _p(a):
    return None if a is None else a.ctypes.data_as(_PD)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def build(force=False):
    """Compile the checkers (g++ only). Building the checker is not using it."""
    import subprocess

    if force or not os.path.exists(os.path.join(_HERE, "libpainty_oracle.so")) or (
            os.path.isdir("/root/reference/painty") and not os.path.exists(os.path.join(_HERE, "_ref", "libpainty_ref.so"))):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))


def have_ref():
    return os.path.exists(os.path.join(_HERE, "_ref", "libpainty_ref.so"))


class Cpu:
    def __init__(self, kind="port"):
        assert kind in ("port", "ref")
        self.kind = kind
        self.px = "ora_" if kind == "port" else "ref_"
        path = os.path.join(_HERE, "libpainty_oracle.so") if kind == "port" else os.path.join(_HERE, "_ref", "libpainty_ref.so")
        if not os.path.exists(path):
            build()
        self.lib = C.CDLL(path)
        self._registered_widths = set()
        self._registered_tmap = False

    def fn(self, name, restype=None, argtypes=None):
        f = getattr(self.lib, self.px + name)
        f.restype = restype
        if argtypes is not None:
            f.argtypes = argtypes
        return f

    # ---- scalars --------------------------------------------------------------------------------
    def compute_reflectance(self, K, S, R0, d):
        out = np.zeros(3)
        self.fn("compute_reflectance", None, [_PD, _PD, _PD, C.c_double, _PD])(_p(_f64(K)), _p(_f64(S)), _p(_f64(R0)), d, _p(out))
        return out

    def coth(self, x):
        return self.fn("coth", C.c_double, [C.c_double])(x)

    def acoth(self, x):
        return self.fn("acoth", C.c_double, [C.c_double])(x)

    def compute_scattering_absorption(self, Rb, Rw):
        K, S = np.zeros(3), np.zeros(3)
        rc = self.fn("compute_scattering_absorption", C.c_int, [_PD] * 4)(_p(_f64(Rb)), _p(_f64(Rw)), _p(K), _p(S))
        if rc:
            raise ValueError("invalid_argument")
        return K, S

    def catmull_rom(self, p_1, p0, p1, p2, t, derivative=False):
        out = np.zeros(2)
        self.fn("catmull_rom_d1" if derivative else "catmull_rom", None, [_PD] * 4 + [C.c_double, _PD])(
            _p(_f64(p_1)), _p(_f64(p0)), _p(_f64(p1)), _p(_f64(p2)), t, _p(out))
        return out

    def catmull_rom_scalar(self, a, b, c, d, t):
        return self.fn("catmull_rom_scalar", C.c_double, [C.c_double] * 5)(a, b, c, d, t)

    def cubic_scalar(self, a, b, c, d, t):
        return self.fn("cubic_scalar", C.c_double, [C.c_double] * 5)(a, b, c, d, t)

    def spline_eval(self, pts, u, kind=0):
        pts = _f64(pts)
        out = np.zeros(2)
        self.fn("spline_eval", None, [C.c_int, _PD, C.c_double, C.c_int, _PD])(len(pts), _p(pts), u, kind, _p(out))
        return out

    def mvc_interpolate(self, polygon, values, pos):
        polygon, values = _f64(polygon).reshape(-1, 2), _f64(values).reshape(-1, 2)
        out = np.zeros(2)
        rc = self.fn("mvc_interpolate", C.c_int, [C.c_int, _PD, C.c_int, _PD, C.c_double, C.c_double, _PD])(
            len(polygon), _p(polygon), len(values), _p(values), pos[0], pos[1], _p(out))
        if rc:
            raise ValueError("invalid_argument")
        return out

    def interpolate_bilinear(self, m, x, y):
        m = _f64(m)
        return self.fn("interpolate_bilinear", C.c_double, [_PD, C.c_int, C.c_int, C.c_double, C.c_double])(
            _p(m), m.shape[0], m.shape[1], x, y)

    # ---- whole image ----------------------------------------------------------------------------
    def compose(self, K, S, V, R0):
        K, S, V, R0 = _f64(K), _f64(S), _f64(V), _f64(R0)
        out = np.empty_like(R0)
        self.fn("compose", None, [C.c_int, C.c_int] + [_PD] * 5)(V.shape[0], V.shape[1], _p(K), _p(S), _p(V), _p(R0), _p(out))
        return out

    def compose_onto(self, K, S, V, R0):
        K, S, V = _f64(K), _f64(S), _f64(V)
        R = _f64(R0).copy()
        self.fn("compose_onto", None, [C.c_int, C.c_int] + [_PD] * 4)(V.shape[0], V.shape[1], _p(K), _p(S), _p(V), _p(R))
        return R

    def compose_timed(self, K, S, V, R0, threads=1):
        """seconds for one ComputeReflectance pass over the given pixels (row-split over threads)."""
        K, S, V, R0 = _f64(K), _f64(S), _f64(V), _f64(R0)
        out = np.empty_like(R0)
        t = self.fn("compose_timed", C.c_double, [C.c_int64] + [_PD] * 5 + [C.c_int])(V.size, _p(K), _p(S), _p(V), _p(R0), _p(out), threads)
        return t, out

    def qrgb32(self, rgb):
        """linear RGB [.., 3] f64 -> uint32 0xffRRGGBB like DigitalCanvas::updateCanvas."""
        rgb = _f64(rgb)
        out = np.empty(rgb.shape[:-1], dtype=np.uint32)
        self.fn("qrgb32", None, [C.c_int64, _PD, C.c_void_p])(out.size, _p(rgb), out.ctypes.data_as(C.c_void_p))
        return out

    def rgb2lab(self, rgb):
        """linear RGB [.., 3] f64 -> CIELab (D65) like convertColor(rgb_2_CIELab) (Color.hxx:248-252)."""
        rgb = _f64(rgb)
        out = np.empty_like(rgb)
        self.fn("rgb2lab", None, [C.c_int64, _PD, _PD])(rgb.size // 3, _p(rgb), _p(out))
        return out

    def lab_scaled(self, rgb, rows, cols):
        """The planner's read-back prep (PictureTargetSbrPainter.cxx:334-341): ScaledMat(convertColor(rgb, rgb_2_CIELab),
        rows, cols) — the colour conversion by this checker, cv::resize(INTER_LANCZOS4) by the container's cv2 (OpenCV is
        the reference's own resize; the three channels are resized independently)."""
        import cv2

        lab = self.rgb2lab(rgb)
        if lab.shape[:2] == (rows, cols):
            return lab
        return np.ascontiguousarray(cv2.resize(lab, (cols, rows), interpolation=cv2.INTER_LANCZOS4))

    def bgr(self, rgb, bits=8, srgb=True):
        """io::imSave's pixel path (port only: cv::Mat::convertTo is restated, OpenCV is not available)."""
        rgb = _f64(rgb)
        out = np.empty(rgb.shape, dtype=np.uint8 if bits == 8 else np.uint16)
        f = self.lib.ora_bgr
        f.restype, f.argtypes = None, [C.c_int64, _PD, C.c_int, C.c_int, C.c_void_p]
        f(rgb.size // 3, _p(rgb), bits, int(srgb), out.ctypes.data_as(C.c_void_p))
        return out

    # ---- asset registration (ref only; no-ops for the port) -------------------------------------
    def _ensure_footprint(self, radius):
        if self.kind != "ref":
            return
        from painty_b200 import assets

        width = assets.footprint_geometry(radius)[0]
        if not self._registered_widths:
            dummy = np.zeros((1, 1))
            self.lib.ref_register_image(FOOTPRINT_PATH.encode(), 1, 1, _p(dummy))
        if width not in self._registered_widths:
            fp = assets.scaled_footprint(width)
            self.lib.ref_register_resize(width, width, _p(fp))
            self._registered_widths.add(width)

    def _ensure_tmap(self, tmap):
        if self.kind == "ref":
            tmap = _f64(tmap)
            self.lib.ref_register_image((SAMPLE_DIR + "/thickness_map.png").encode(), tmap.shape[0], tmap.shape[1], _p(tmap))

    def canvas(self, rows, cols):
        return CpuCanvas(self, rows, cols)

    def footprint_brush(self, radius):
        return CpuFootprintBrush(self, radius)

    def texture_brush(self, tmap=None):
        return CpuTextureBrush(self, tmap)

    def stroke_sample_at(self, tmap, x, y):
        tmap = _f64(tmap)
        if self.kind == "ref":
            self._ensure_tmap(tmap)
            f = self.lib.ref_stroke_sample_at
            f.restype = C.c_double
            f.argtypes = [C.c_char_p, C.c_double, C.c_double]
            return f(SAMPLE_DIR.encode(), x, y)
        return self.fn("stroke_sample_at", C.c_double, [_PD, C.c_int, C.c_int, C.c_double, C.c_double])(
            _p(tmap), tmap.shape[0], tmap.shape[1], x, y)


class CpuCanvas:
    def __init__(self, cpu, rows, cols):
        self.cpu, self.rows, self.cols = cpu, rows, cols
        self.h = C.c_void_p(cpu.fn("canvas_create", C.c_void_p, [C.c_int, C.c_int])(rows, cols))

    def __del__(self):
        if getattr(self, "h", None):
            self.cpu.fn("canvas_destroy", None, [C.c_void_p])(self.h)
            self.h = None

    def clear(self):
        self.cpu.fn("canvas_clear", None, [C.c_void_p])(self.h)

    def set_background(self, R0):
        self.cpu.fn("canvas_set_background", None, [C.c_void_p, _PD])(self.h, _p(_f64(R0)))

    def dry(self):
        self.cpu.fn("canvas_dry", None, [C.c_void_p])(self.h)

    def set_layer(self, K, S, V):
        self.cpu.fn("canvas_set_layer", None, [C.c_void_p, _PD, _PD, _PD])(self.h, _p(_f64(K)), _p(_f64(S)), _p(_f64(V)))

    def get(self):
        """dict(K,S,V,R0,h) AoS f64."""
        r, c = self.rows, self.cols
        o = dict(K=np.empty((r, c, 3)), S=np.empty((r, c, 3)), V=np.empty((r, c)), R0=np.empty((r, c, 3)), h=np.empty((r, c)))
        self.cpu.fn("canvas_get", None, [C.c_void_p] + [_PD] * 5)(self.h, _p(o["K"]), _p(o["S"]), _p(o["V"]), _p(o["R0"]), _p(o["h"]))
        return o

    def compose(self):
        out = np.empty((self.rows, self.cols, 3))
        self.cpu.fn("canvas_compose", None, [C.c_void_p, _PD])(self.h, _p(out))
        return out

    def render(self):
        out = np.empty((self.rows, self.cols, 3))
        self.cpu.fn("canvas_render", None, [C.c_void_p, _PD])(self.h, _p(out))
        return out


class CpuFootprintBrush:
    def __init__(self, cpu, radius):
        self.cpu = cpu
        if cpu.kind == "ref":
            cpu._ensure_footprint(radius)
            self.h = C.c_void_p(cpu.fn("fbrush_create", C.c_void_p, [C.c_double])(radius))
            assert self.h, "reference FootprintBrush ctor threw"
        else:
            self.h = C.c_void_p(cpu.fn("fbrush_create", C.c_void_p, [])())
            self.set_radius(radius)

    def __del__(self):
        if getattr(self, "h", None):
            self.cpu.fn("fbrush_destroy", None, [C.c_void_p])(self.h)
            self.h = None

    def set_radius(self, radius):
        from painty_b200 import assets

        if self.cpu.kind == "ref":
            self.cpu._ensure_footprint(radius)
            rc = self.cpu.fn("fbrush_set_radius", C.c_int, [C.c_void_p, C.c_double])(self.h, radius)
            assert rc == 0
        else:
            fp = assets.baked_footprint(radius)
            self.cpu.fn("fbrush_set_radius", C.c_int, [C.c_void_p, C.c_double, C.c_int, _PD])(self.h, radius, fp.shape[0], _p(fp))

    def dip(self, K, S):
        self.cpu.fn("fbrush_dip", None, [C.c_void_p, _PD, _PD])(self.h, _p(_f64(K)), _p(_f64(S)))

    def set_rates(self, pickup, deposition):
        self.cpu.fn("fbrush_set_rates", None, [C.c_void_p, C.c_double, C.c_double])(self.h, pickup, deposition)

    def set_use_snapshot(self, use):
        self.cpu.fn("fbrush_set_use_snapshot", None, [C.c_void_p, C.c_int])(self.h, int(use))

    def size_map(self):
        return self.cpu.fn("fbrush_size_map", C.c_int, [C.c_void_p])(self.h)

    def footprint_size(self):
        return self.cpu.fn("fbrush_footprint_size", C.c_int, [C.c_void_p])(self.h)

    def pickup_map(self):
        n = self.size_map()
        K, S, V = np.empty((n, n, 3)), np.empty((n, n, 3)), np.empty((n, n))
        self.cpu.fn("fbrush_get_pickup_map", None, [C.c_void_p, _PD, _PD, _PD])(self.h, _p(K), _p(S), _p(V))
        return K, S, V

    def counters(self):
        """(visited, active) stroke-pixels so far — port only."""
        v, a = C.c_uint64(0), C.c_uint64(0)
        self.cpu.fn("fbrush_counters", None, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)])(self.h, C.byref(v), C.byref(a))
        return v.value, a.value

    def imprint_batch(self, canvas, cx, cy, theta):
        cx, cy, theta = _f64(cx), _f64(cy), _f64(theta)
        return self.cpu.fn("fbrush_imprint_batch", C.c_double, [C.c_void_p, C.c_void_p, C.c_int, _PD, _PD, _PD])(
            self.h, canvas.h, len(cx), _p(cx), _p(cy), _p(theta))


class CpuTextureBrush:
    def __init__(self, cpu, tmap=None):
        from painty_b200 import assets

        self.cpu = cpu
        tmap = _f64(assets.thickness_map() if tmap is None else tmap)
        if cpu.kind == "ref":
            cpu._ensure_tmap(tmap)
            self.h = C.c_void_p(cpu.fn("tbrush_create", C.c_void_p, [C.c_char_p, C.c_int])(SAMPLE_DIR.encode(), 0))
            assert self.h
        else:
            self.h = C.c_void_p(cpu.fn("tbrush_create", C.c_void_p, [C.c_int, C.c_int, _PD])(tmap.shape[0], tmap.shape[1], _p(tmap)))

    def __del__(self):
        if getattr(self, "h", None):
            self.cpu.fn("tbrush_destroy", None, [C.c_void_p])(self.h)
            self.h = None

    def set_radius(self, r):
        self.cpu.fn("tbrush_set_radius", None, [C.c_void_p, C.c_double])(self.h, r)

    def dip(self, K, S):
        self.cpu.fn("tbrush_dip", None, [C.c_void_p, _PD, _PD])(self.h, _p(_f64(K)), _p(_f64(S)))

    def set_thickness_scale(self, s):
        self.cpu.fn("tbrush_set_thickness_scale", None, [C.c_void_p, C.c_double])(self.h, s)

    def enable_smudge(self, enable):
        self.cpu.fn("tbrush_enable_smudge", None, [C.c_void_p, C.c_int])(self.h, int(enable))

    def paint_stroke(self, canvas, path):
        path = _f64(path).reshape(-1, 2)
        return self.cpu.fn("tbrush_paint_stroke", C.c_double, [C.c_void_p, C.c_void_p, C.c_int, _PD])(self.h, canvas.h, len(path), _p(path))
