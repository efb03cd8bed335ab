"""ctypes binding of libpainty_b200.so — a thin Python mirror of the reference's renderer classes
(same names and argument meaning as painty::Canvas / PaintLayer / FootprintBrush / TextureBrush /
Renderer) used by the tests and bench.py. The product is the CUDA library behind the C ABI in
include/painty_b200.h; the drop-in C++ façade for painty itself lives in include/painty/.

There is no CPU fallback: importing works anywhere (so that symbol checks can run without a GPU), but
creating a Context without a B200 raises PaintyError.
"""
import ctypes as C
import os as _os

_os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")  # see pb_context_create: no load-time syncs behind persistent kernels
import os

import numpy as np

from . import assets

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpainty_b200.so")
_PD = C.POINTER(C.c_double)
_VP = C.c_void_p
F32, F64 = 0, 1


class PaintyError(RuntimeError):
    pass


class pb_stroke(C.Structure):
    _fields_ = [("radius", C.c_double), ("K", C.c_double * 3), ("S", C.c_double * 3), ("first_imprint", C.c_int64),
                ("n_imprints", C.c_int64)]


class pb_tstroke(C.Structure):
    _fields_ = [("radius", C.c_double), ("K", C.c_double * 3), ("S", C.c_double * 3), ("thickness_scale", C.c_double),
                ("first_vertex", C.c_int64), ("n_vertices", C.c_int32), ("texture_id", C.c_int32)]


STROKE_DTYPE = np.dtype([("radius", "<f8"), ("K", "<f8", 3), ("S", "<f8", 3), ("first_imprint", "<i8"), ("n_imprints", "<i8")])
TSTROKE_DTYPE = np.dtype([("radius", "<f8"), ("K", "<f8", 3), ("S", "<f8", 3), ("thickness_scale", "<f8"),
                          ("first_vertex", "<i8"), ("n_vertices", "<i4"), ("texture_id", "<i4")])
assert STROKE_DTYPE.itemsize == C.sizeof(pb_stroke) and TSTROKE_DTYPE.itemsize == C.sizeof(pb_tstroke)

_lib = None


def lib():
    """Load the in-tree CUDA library; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PaintyError("%s is missing: run `python -m painty_b200.build` (there is no CPU fallback)" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.pb_last_error.restype = C.c_char_p
        _lib.pb_context_stream.restype = _VP
        _lib.pb_context_launch_count.restype = C.c_int64
        _lib.pb_fbrush_get_pickup_rate.restype = C.c_double
        _lib.pb_fbrush_get_deposition_rate.restype = C.c_double
    return _lib


def _chk(rc):
    if rc != 0:
        raise PaintyError(lib().pb_last_error().decode())


def _p(a):
    return None if a is None else a.ctypes.data_as(_PD)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _d3(a):
    return (C.c_double * 3)(*[float(x) for x in a])


# ---- host-side scalar calls ---------------------------------------------------------------------
def ComputeReflectance(K, S, R0, d):
    out = (C.c_double * 3)()
    _chk(lib().pb_compute_reflectance(_d3(K), _d3(S), _d3(R0), C.c_double(d), out))
    return np.array(out[:])


def ComputeScatteringAndAbsorption(Rb, Rw):
    K, S = (C.c_double * 3)(), (C.c_double * 3)()
    rc = lib().pb_compute_scattering_absorption(_d3(Rb), _d3(Rw), K, S)
    if rc:
        raise ValueError(lib().pb_last_error().decode())  # std::invalid_argument in the reference
    return np.array(K[:]), np.array(S[:])


def mixed(K1, S1, v1, K2, S2, v2):
    K, S = (C.c_double * 3)(), (C.c_double * 3)()
    _chk(lib().pb_paint_mixed(_d3(K1), _d3(S1), C.c_double(v1), _d3(K2), _d3(S2), C.c_double(v2), K, S))
    return np.array(K[:]), np.array(S[:])


def mixSinglePaint(baseK, baseS, weights):
    baseK, baseS, w = _f64(baseK), _f64(baseS), _f64(weights)
    K, S = (C.c_double * 3)(), (C.c_double * 3)()
    rc = lib().pb_paint_mix_single(len(baseK), _p(baseK), _p(baseS), len(w), _p(w), K, S)
    if rc:
        raise ValueError(lib().pb_last_error().decode())
    return np.array(K[:]), np.array(S[:])


def expand_stroke(path, mode=0):
    """Stroke -> (cx, cy, theta) imprints. mode 0 = FootprintBrush::paintStroke, 1 = GUI mouse-move loop."""
    path = _f64(path).reshape(-1, 2)
    n = C.c_int64(0)
    _chk(lib().pb_expand_stroke(mode, len(path), _p(path), C.c_int64(0), None, None, None, C.byref(n)))
    cx, cy, th = np.empty(n.value), np.empty(n.value), np.empty(n.value)
    _chk(lib().pb_expand_stroke(mode, len(path), _p(path), n, _p(cx), _p(cy), _p(th), C.byref(n)))
    return cx, cy, th


def expand_strokes(first_vertex, n_vertices, path_xy, mode=0):
    """Batch form of expand_stroke: -> (cx, cy, theta, first_imprint[n], n_imprints[n])."""
    fv = np.ascontiguousarray(first_vertex, dtype=np.int64)
    nv = np.ascontiguousarray(n_vertices, dtype=np.int32)
    path_xy = _f64(path_xy).reshape(-1, 2)
    n = len(fv)
    fi, ni, total = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64), C.c_int64(0)
    args = (mode, C.c_int64(n), fv.ctypes.data_as(_VP), nv.ctypes.data_as(_VP), _p(path_xy))
    _chk(lib().pb_expand_stroke_batch(*args, C.c_int64(0), None, None, None, None, None, C.byref(total)))
    cx, cy, th = np.empty(total.value), np.empty(total.value), np.empty(total.value)
    _chk(lib().pb_expand_stroke_batch(*args, C.c_int64(total.value), _p(cx), _p(cy), _p(th), fi.ctypes.data_as(_VP),
                                      ni.ctypes.data_as(_VP), C.byref(total)))
    return cx, cy, th, fi, ni


def plan_dependencies(rows, cols, box, allowed):
    """Host-side dataflow plan: (offsets[n+1], preds) CSR of the strokes each stroke has to wait for."""
    box = np.ascontiguousarray(box, dtype=np.int32).reshape(-1, 4)
    allowed = np.ascontiguousarray(allowed, dtype=np.int32).reshape(-1, 4)
    n = len(box)
    offsets = np.zeros(n + 1, dtype=np.int64)
    total = C.c_int64(0)
    _chk(lib().pb_plan_dependencies(rows, cols, C.c_int64(n), box.ctypes.data_as(_VP), allowed.ctypes.data_as(_VP),
                                    offsets.ctypes.data_as(_VP), C.c_int64(0), None, C.byref(total)))
    preds = np.zeros(max(total.value, 1), dtype=np.int32)
    _chk(lib().pb_plan_dependencies(rows, cols, C.c_int64(n), box.ctypes.data_as(_VP), allowed.ctypes.data_as(_VP),
                                    offsets.ctypes.data_as(_VP), C.c_int64(total.value), preds.ctypes.data_as(_VP), C.byref(total)))
    return offsets, preds[:total.value]


def plan_segments(rows, cols, first, count, side, radius, cx, cy, segment_length=64, use_snapshot=True, single=None):
    """Segment-level dataflow plan of a footprint stroke list (see pb_plan_segments).

    Returns (seg_first[n+1], seg_len[n], seg_off[segments+1], pred_stroke, pred_need)."""
    first = np.ascontiguousarray(first, dtype=np.int64)
    count = np.ascontiguousarray(count, dtype=np.int64)
    side = np.ascontiguousarray(side, dtype=np.int32)
    radius = np.ascontiguousarray(radius, dtype=np.float64)
    cx = np.ascontiguousarray(cx, dtype=np.float64)
    cy = np.ascontiguousarray(cy, dtype=np.float64)
    single = None if single is None else np.ascontiguousarray(single, dtype=np.uint8)
    n = len(first)
    seg_first = np.zeros(n + 1, dtype=np.int32)
    seg_len = np.zeros(max(n, 1), dtype=np.int32)
    total = C.c_int64(0)

    def call(seg_cap, seg_off, pred_cap, ps, pn):
        _chk(lib().pb_plan_segments(rows, cols, C.c_int64(n), first.ctypes.data_as(_VP), count.ctypes.data_as(_VP),
                                    side.ctypes.data_as(_VP), radius.ctypes.data_as(_VP),
                                    None if single is None else single.ctypes.data_as(_VP), cx.ctypes.data_as(_VP),
                                    cy.ctypes.data_as(_VP), int(segment_length), int(bool(use_snapshot)),
                                    seg_first.ctypes.data_as(_VP), seg_len.ctypes.data_as(_VP), C.c_int64(seg_cap),
                                    seg_off.ctypes.data_as(_VP), C.c_int64(pred_cap), ps.ctypes.data_as(_VP),
                                    pn.ctypes.data_as(_VP), C.byref(total)))

    dummy = np.zeros(1, dtype=np.int32)
    call(0, dummy, 0, dummy, dummy)
    n_seg, n_pred = int(seg_first[n]), int(total.value)
    seg_off = np.zeros(n_seg + 1, dtype=np.int32)
    ps = np.zeros(max(n_pred, 1), dtype=np.int32)
    pn = np.zeros(max(n_pred, 1), dtype=np.int32)
    call(n_seg, seg_off, n_pred, ps, pn)
    return seg_first, seg_len[:n], seg_off, ps[:n_pred], pn[:n_pred]


def plan_claim_order(rows, cols, first, count, side, radius, cx, cy, pool, run, cost, slots, segment_length=64,
                     use_snapshot=True, single=None, return_makespan=False):
    """Host list-scheduling of a footprint stroke list (see pb_plan_claim_order). `slots` = per pool, the list of
    concurrent strokes of each of its runs. Returns the global claim sequence (stroke indices)."""
    n = len(first)
    arr = lambda a, t: np.ascontiguousarray(a, dtype=t)
    first, count, side, radius = arr(first, np.int64), arr(count, np.int64), arr(side, np.int32), arr(radius, np.float64)
    cx, cy, pool, run, cost = arr(cx, np.float64), arr(cy, np.float64), arr(pool, np.int32), arr(run, np.int32), arr(cost, np.float64)
    rpp = arr([len(x) for x in slots], np.int32)
    flat = arr([v for x in slots for v in x], np.int32)
    order = np.zeros(max(n, 1), dtype=np.int32)
    single = None if single is None else arr(single, np.uint8)
    makespan = C.c_double(0.0)
    p = lambda a: None if a is None else a.ctypes.data_as(_VP)
    _chk(lib().pb_plan_claim_order(rows, cols, C.c_int64(n), p(first), p(count), p(side), p(radius), p(single), p(cx), p(cy),
                                   int(segment_length), int(bool(use_snapshot)), p(pool), p(run), p(cost), len(slots),
                                   p(rpp), p(flat), p(order), C.byref(makespan)))
    return (order[:n], makespan.value) if return_makespan else order[:n]


# ---- device objects -----------------------------------------------------------------------------
class Context:
    def __init__(self, device=0, precision=F32):
        self.h = _VP()
        _chk(lib().pb_context_create(device, precision, C.byref(self.h)))
        self.precision = precision
        self.device = device
        self.dtype = np.float64 if precision == F64 else np.float32

    def close(self):
        if self.h:
            lib().pb_context_destroy(self.h)
            self.h = _VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        _chk(lib().pb_context_synchronize(self.h))

    @property
    def stream(self):
        return lib().pb_context_stream(self.h)

    @property
    def launches(self):
        return lib().pb_context_launch_count(self.h)

    def km_compose_planes(self, n, K, S, V, R0, R):
        """Raw streaming compose on device pointers (ints): K,S,R0,R are 3 pointers each, V one."""
        a3 = lambda x: (_VP * 3)(*[_VP(int(p)) for p in x])
        _chk(lib().pb_km_compose_planes(self.h, C.c_int64(n), a3(K), a3(S), _VP(int(V)), a3(R0), a3(R)))

    def km_compose_stacked_planes(self, n, K, S, V, R0, R):
        """K,S: lists (per layer) of 3 pointers; V: list of pointers; R0,R: 3 pointers."""
        L = len(V)
        Ka = (_VP * (3 * L))(*[_VP(int(p)) for l in K for p in l])
        Sa = (_VP * (3 * L))(*[_VP(int(p)) for l in S for p in l])
        Va = (_VP * L)(*[_VP(int(p)) for p in V])
        a3 = lambda x: (_VP * 3)(*[_VP(int(p)) for p in x])
        _chk(lib().pb_km_compose_stacked_planes(self.h, C.c_int64(n), L, Ka, Sa, Va, a3(R0), a3(R)))


class PaintLayer:
    """painty::PaintLayer<vec3> (renderer/PaintLayer.hxx); ctor is (rows, cols)."""

    def __init__(self, ctx, rows, cols):
        self.ctx, self.rows, self.cols = ctx, rows, cols
        self.h = _VP()
        _chk(lib().pb_layer_create(ctx.h, rows, cols, C.byref(self.h)))

    def __del__(self):
        if getattr(self, "h", None) and self.ctx.h and lib is not None:
            lib().pb_layer_destroy(self.h)
            self.h = None

    def getRows(self):
        return self.rows

    def getCols(self):
        return self.cols

    def clear(self):
        _chk(lib().pb_layer_clear(self.h))

    def upload(self, K, S, V):
        _chk(lib().pb_layer_upload(self.h, _p(_f64(K)), _p(_f64(S)), _p(_f64(V))))

    def download(self):
        K, S, V = np.empty((self.rows, self.cols, 3)), np.empty((self.rows, self.cols, 3)), np.empty((self.rows, self.cols))
        _chk(lib().pb_layer_download(self.h, _p(K), _p(S), _p(V)))
        return K, S, V

    def composeOnto(self, R0):
        """In place on a host AoS f64 array, reallocated to ones when the size differs (PaintLayer.hxx:81-96)."""
        if R0 is None or R0.shape[:2] != (self.rows, self.cols):
            R0 = np.ones((self.rows, self.cols, 3))
        R0 = _f64(R0)
        _chk(lib().pb_layer_compose_onto(self.h, _p(R0)))
        return R0

    def copyTo(self, other):
        _chk(lib().pb_layer_copy(self.h, other.h))
        other.rows, other.cols = self.rows, self.cols


class Canvas:
    """painty::Canvas<vec3> (renderer/Canvas.hxx); ctor is (rows, cols). band=(row_begin,row_end,halo)
    creates one row band of a larger canvas for multi-GPU sharding."""

    def __init__(self, ctx, rows, cols, band=None):
        self.ctx, self.rows, self.cols = ctx, rows, cols
        self.h = _VP()
        if band is None:
            _chk(lib().pb_canvas_create(ctx.h, rows, cols, C.byref(self.h)))
        else:
            _chk(lib().pb_canvas_create_band(ctx.h, rows, cols, band[0], band[1], band[2], C.byref(self.h)))
        first, n = C.c_int(0), C.c_int(0)
        lib().pb_canvas_stored_rows(self.h, C.byref(first), C.byref(n))
        self.store_first, self.store_rows = first.value, n.value
        self.band = band if band is not None else (0, rows, 0)

    def __del__(self):
        if getattr(self, "h", None) and self.ctx.h and lib is not None:
            lib().pb_canvas_destroy(self.h)
            self.h = None

    def clear(self):
        _chk(lib().pb_canvas_clear(self.h))

    def setBackground(self, R0):
        R0 = _f64(R0)
        assert R0.shape == (self.store_rows, self.cols, 3)
        _chk(lib().pb_canvas_set_background(self.h, _p(R0)))

    def compose_lab_scaled(self, rows, cols):
        """The sbr planner's read-back prep on the device: compose -> CIELab -> LANCZOS4 resize to rows x cols (host AoS f64)."""
        out = np.empty((rows, cols, 3))
        _chk(lib().pb_canvas_compose_lab_scaled(self.h, int(rows), int(cols), _p(out)))
        return out

    def dryCanvas(self):
        _chk(lib().pb_canvas_dry(self.h))

    def upload_layer(self, K, S, V):
        _chk(lib().pb_canvas_upload_layer(self.h, _p(_f64(K)), _p(_f64(S)), _p(_f64(V))))

    def download(self, which="KSVRh"):
        r, c = self.store_rows, self.cols
        o = {}
        if "K" in which:
            o["K"] = np.empty((r, c, 3))
        if "S" in which:
            o["S"] = np.empty((r, c, 3))
        if "V" in which:
            o["V"] = np.empty((r, c))
        if "R" in which:
            o["R0"] = np.empty((r, c, 3))
        if "h" in which:
            o["h"] = np.empty((r, c))
        _chk(lib().pb_canvas_download(self.h, _p(o.get("K")), _p(o.get("S")), _p(o.get("V")), _p(o.get("R0")), _p(o.get("h"))))
        return o

    def compose(self, out=None):
        """Renderer::compose(canvas) -> host AoS f64 [rows, cols, 3]."""
        if out is None:
            out = np.empty((self.store_rows, self.cols, 3))
        _chk(lib().pb_canvas_compose(self.h, _p(out)))
        return out

    def render(self):
        """Renderer::render(canvas): compose + directional-light relighting -> host AoS f64 [rows, cols, 3]."""
        out = np.empty((self.store_rows, self.cols, 3))
        _chk(lib().pb_canvas_render(self.h, _p(out)))
        return out

    def compose_qrgb32(self):
        """compose + rgb2srgb + 8-bit like DigitalCanvas::updateCanvas: uint32 0xffRRGGBB [rows, cols]."""
        out = np.empty((self.store_rows, self.cols), dtype=np.uint32)
        _chk(lib().pb_canvas_compose_qrgb32(self.h, out.ctypes.data_as(_VP)))
        return out

    def compose_bgr(self, bits=8, srgb=True):
        """compose + (rgb2srgb) + convertTo(8|16 bit) + RGB2BGR like io::imSave before the encoder."""
        out = np.empty((self.store_rows, self.cols, 3), dtype=np.uint8 if bits == 8 else np.uint16)
        _chk(lib().pb_canvas_compose_bgr(self.h, bits, int(srgb), out.ctypes.data_as(_VP)))
        return out

    def compose_device(self, d_out_ptr, plane_stride):
        _chk(lib().pb_canvas_compose_device(self.h, _VP(int(d_out_ptr)), C.c_int64(plane_stride)))

    def compose_band_device(self, d_out_ptr, plane_stride):
        _chk(lib().pb_canvas_compose_band_device(self.h, _VP(int(d_out_ptr)), C.c_int64(plane_stride)))

    def device_planes(self):
        planes = (_VP * 11)()
        n = C.c_int64(0)
        _chk(lib().pb_canvas_device_planes(self.h, planes, C.byref(n)))
        return [int(p) for p in planes], n.value


class Renderer:
    """painty::Renderer<vec3>: compose (renderer/Renderer.hxx:26-53) and render (:60-156)."""

    def render(self, canvas):
        return canvas.render()

    def compose(self, a, R0=None):
        if isinstance(a, Canvas):
            return a.compose()
        R0 = _f64(R0)
        out = np.empty_like(R0)
        _chk(lib().pb_layer_compose(a.h, _p(R0), _p(out)))
        return out


class FootprintBrush:
    """painty::FootprintBrush<vec3> (renderer/FootprintBrush.hxx). The footprint image is prepared on the
    host by painty_b200.assets (= painty's own imRead + ScaledMat + PaddedMat)."""

    def __init__(self, ctx, radius):
        self.ctx = ctx
        self.h = _VP()
        _chk(lib().pb_fbrush_create(ctx.h, C.byref(self.h)))
        self.setRadius(radius)

    def __del__(self):
        if getattr(self, "h", None) and self.ctx.h and lib is not None:
            lib().pb_fbrush_destroy(self.h)
            self.h = None

    def setRadius(self, radius):
        acted = C.c_int(0)
        _chk(lib().pb_fbrush_set_radius(self.h, C.c_double(radius), 0, None, C.byref(acted)))
        if acted.value:
            fp = assets.baked_footprint(radius)
            _chk(lib().pb_fbrush_set_radius(self.h, C.c_double(radius), fp.shape[0], _p(fp), C.byref(acted)))

    def register_radius(self, radius):
        fp = assets.baked_footprint(radius)
        _chk(lib().pb_fbrush_register_footprint(self.h, C.c_double(radius), fp.shape[0], _p(fp)))

    def dip(self, paint):
        _chk(lib().pb_fbrush_dip(self.h, _d3(paint[0]), _d3(paint[1])))

    def clean(self):
        _chk(lib().pb_fbrush_clean(self.h))

    def setPickupRate(self, r):
        lib().pb_fbrush_set_pickup_rate(self.h, C.c_double(r))

    def setDepositionRate(self, r):
        lib().pb_fbrush_set_deposition_rate(self.h, C.c_double(r))

    def getPickupRate(self):
        return lib().pb_fbrush_get_pickup_rate(self.h)

    def getDepositionRate(self):
        return lib().pb_fbrush_get_deposition_rate(self.h)

    def setUseSnapshotBuffer(self, use):
        lib().pb_fbrush_set_use_snapshot(self.h, int(bool(use)))

    def getUseSnapshotBuffer(self):
        return bool(lib().pb_fbrush_get_use_snapshot(self.h))

    def updateSnapshot(self, canvas):
        _chk(lib().pb_fbrush_update_snapshot(self.h, canvas.h))

    def getPickupMap(self):
        n = lib().pb_fbrush_size_map(self.h)
        K, S, V = np.empty((n, n, 3)), np.empty((n, n, 3)), np.empty((n, n))
        _chk(lib().pb_fbrush_pickup_map(self.h, _p(K), _p(S), _p(V)))
        return K, S, V

    def getSnapshot(self, canvas):
        r, c = canvas.store_rows, canvas.cols
        K, S, V = np.empty((r, c, 3)), np.empty((r, c, 3)), np.empty((r, c))
        _chk(lib().pb_fbrush_snapshot_download(self.h, _p(K), _p(S), _p(V)))
        return K, S, V

    def imprint(self, center, theta, canvas):
        self.imprint_batch(canvas, [center[0]], [center[1]], [theta])

    def imprint_batch(self, canvas, cx, cy, theta):
        cx, cy, theta = _f64(cx), _f64(cy), _f64(theta)
        _chk(lib().pb_fbrush_imprint_batch(self.h, canvas.h, C.c_int64(len(cx)), _p(cx), _p(cy), _p(theta)))

    def paintStroke(self, path, canvas, mode=0):
        cx, cy, th = expand_stroke(path, mode)
        self.imprint_batch(canvas, cx, cy, th)

    def stroke_batch(self, canvas, strokes, cx, cy, theta):
        """strokes: numpy structured array of STROKE_DTYPE (submission order)."""
        strokes = np.ascontiguousarray(strokes, dtype=STROKE_DTYPE)
        cx, cy, theta = _f64(cx), _f64(cy), _f64(theta)
        _chk(lib().pb_fbrush_stroke_batch(self.h, canvas.h, C.c_int64(len(strokes)), strokes.ctypes.data_as(_VP),
                                          C.c_int64(len(cx)), _p(cx), _p(cy), _p(theta)))

    def plan_stroke_batch(self, canvas, strokes, cx, cy, theta, dist_desc=None):
        """Host half of stroke_batch (pb_fbrush_plan_stroke_batch): no stream is touched, so it may run on another host
        thread while the device executes an earlier batch. Returns a BatchPlan for run_batch_plan."""
        strokes = np.ascontiguousarray(strokes, dtype=STROKE_DTYPE)
        cx, cy, theta = _f64(cx), _f64(cy), _f64(theta)
        h = _VP()
        _chk(lib().pb_fbrush_plan_stroke_batch(self.h, canvas.h, C.byref(dist_desc) if dist_desc is not None else None,
                                               C.c_int64(len(strokes)), strokes.ctypes.data_as(_VP), C.c_int64(len(cx)), _p(cx), _p(cy),
                                               _p(theta), C.byref(h)))
        return BatchPlan(h)

    def run_batch_plan(self, canvas, plan, dist_desc=None):
        _chk(lib().pb_fbrush_run_batch_plan(self.h, canvas.h, C.byref(dist_desc) if dist_desc is not None else None, plan.h))

    def enable_visited_count(self, enable=True):
        lib().pb_fbrush_enable_visited_count(self.h, int(enable))

    def counters(self):
        v, a = C.c_uint64(0), C.c_uint64(0)
        _chk(lib().pb_fbrush_counters(self.h, C.byref(v), C.byref(a)))
        return v.value, a.value

    def batch_stats(self):
        """Host-side figures of the last stroke / imprint batch (pb_fbrush_batch_stats)."""
        out = (C.c_double * 12)()
        _chk(lib().pb_fbrush_batch_stats(self.h, out))
        keys = _STATS_KEYS
        return {k: float(v) for k, v in zip(keys, out)}


class BandImage:
    """Assembled reflectance image of a band-sharded canvas (pb_band_image): 3 planes of rows x cols in one allocation."""

    def __init__(self, ctx, rows, cols):
        self.ctx, self.rows, self.cols = ctx, rows, cols
        self.h = _VP()
        _chk(lib().pb_band_image_create(ctx.h, rows, cols, C.byref(self.h)))
        base, stride = _VP(), C.c_int64(0)
        _chk(lib().pb_band_image_device(self.h, C.byref(base), C.byref(stride)))
        self.base, self.plane_stride_bytes = base.value, stride.value

    def __del__(self):
        if getattr(self, "h", None) and self.ctx.h and lib is not None:
            lib().pb_band_image_destroy(self.h)
            self.h = None

    def download(self):
        out = np.empty((self.rows, self.cols, 3))
        _chk(lib().pb_band_image_download(self.h, _p(out)))
        return out


def compose_gather(canvas, images):
    """Compose `canvas`' own band and store it into every image of `images` (BandImage objects or raw base pointers that
    share one plane stride) at the band's rows."""
    bases = [im.base if isinstance(im, BandImage) else int(im[0]) for im in images]
    stride = images[0].plane_stride_bytes if isinstance(images[0], BandImage) else int(images[0][1])
    arr = (_VP * len(bases))(*bases)
    _chk(lib().pb_canvas_compose_gather(canvas.h, len(bases), arr, C.c_int64(stride)))


def lanczos4_taps(src, dst):
    """(offsets[dst], weights[dst, 8]) of one axis of cv::resize(INTER_LANCZOS4) as the library evaluates it (host only)."""
    ofs = np.zeros(dst, dtype=np.int32)
    w = np.zeros((dst, 8), dtype=np.float32)
    _chk(lib().pb_lanczos4_taps(int(src), int(dst), ofs.ctypes.data_as(_VP), w.ctypes.data_as(_VP)))
    return ofs, w


_STATS_KEYS = ("plan_ms", "imprint_constants_ms", "strokes_planned", "segments", "wait_entries", "model_ms", "strokes_this_rank",
               "launches_this_rank", "resident_clusters_small", "resident_clusters_mid", "resident_clusters_large",
               "threads_per_cluster_max")


class BatchPlan:
    """Host-side plan of a footprint stroke batch (pb_batch_plan)."""

    def __init__(self, h):
        self.h = h

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:
            lib().pb_batch_plan_destroy(self.h)
            self.h = None

    def stats(self):
        out = (C.c_double * 12)()
        _chk(lib().pb_batch_plan_stats(self.h, out))
        keys = _STATS_KEYS
        return {k: float(v) for k, v in zip(keys, out)}


class TextureBrushDictionary:
    """Host-side painty::TextureBrushDictionary (renderer/src/TextureBrushDictionary.cxx): groups of brush textures by
    (size key, length key) and the lookup rule up to the random draw, which the caller makes and records."""

    def __init__(self, size_keys, length_keys, rows, cols):
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        sk, lk, r, c = i32(size_keys), i32(length_keys), i32(rows), i32(cols)
        self.h = _VP()
        _chk(lib().pb_texdict_create(len(sk), sk.ctypes.data_as(_VP), lk.ctypes.data_as(_VP), r.ctypes.data_as(_VP),
                                     c.ctypes.data_as(_VP), C.byref(self.h)))

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:
            lib().pb_texdict_destroy(self.h)
            self.h = None

    def lookup(self, path, brush_size):
        """-> (size group, length group, candidate entry indices)."""
        path = _f64(path).reshape(-1, 2)
        i0, i1, n = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        cand = np.zeros(256, dtype=np.int32)
        _chk(lib().pb_texdict_lookup(self.h, len(path), _p(path), C.c_double(brush_size), C.byref(i0), C.byref(i1), len(cand),
                                     cand.ctypes.data_as(_VP), C.byref(n)))
        return i0.value, i1.value, cand[:n.value].copy()


class TextureBrush:
    """painty::TextureBrush<vec3> with smudge disabled (renderer/TextureBrush.hxx)."""

    def __init__(self, ctx, thickness_map=None):
        self.ctx = ctx
        tm = _f64(assets.thickness_map() if thickness_map is None else thickness_map)
        self.h = _VP()
        _chk(lib().pb_tbrush_create(ctx.h, tm.shape[0], tm.shape[1], _p(tm), C.byref(self.h)))

    def __del__(self):
        if getattr(self, "h", None) and self.ctx.h and lib is not None:
            lib().pb_tbrush_destroy(self.h)
            self.h = None

    def setRadius(self, r):
        _chk(lib().pb_tbrush_set_radius(self.h, C.c_double(r)))

    def dip(self, paint):
        lib().pb_tbrush_dip(self.h, _d3(paint[0]), _d3(paint[1]))

    def setThicknessScale(self, s):
        lib().pb_tbrush_set_thickness_scale(self.h, C.c_double(s))

    def enableSmudge(self, enable):
        lib().pb_tbrush_enable_smudge(self.h, int(bool(enable)))

    def addTexture(self, thickness_map):
        """Upload one more thickness texture (TextureBrushDictionary entry); returns its id for pb_tstroke.texture_id."""
        tm = _f64(thickness_map)
        tid = C.c_int(0)
        _chk(lib().pb_tbrush_add_texture(self.h, tm.shape[0], tm.shape[1], _p(tm), C.byref(tid)))
        return tid.value

    def selectTexture(self, texture_id):
        """The texture paintStroke samples (BrushStrokeSample::setThicknessMap's role); 0 = the constructor's map."""
        _chk(lib().pb_tbrush_select_texture(self.h, int(texture_id)))

    def paintStroke(self, path, canvas):
        path = _f64(path).reshape(-1, 2)
        _chk(lib().pb_tbrush_paint_stroke(self.h, canvas.h, len(path), _p(path)))

    def stroke_batch(self, canvas, strokes, path_xy):
        strokes = np.ascontiguousarray(strokes, dtype=TSTROKE_DTYPE)
        path_xy = _f64(path_xy).reshape(-1, 2)
        _chk(lib().pb_tbrush_stroke_batch(self.h, canvas.h, C.c_int64(len(strokes)), strokes.ctypes.data_as(_VP),
                                          C.c_int64(len(path_xy)), _p(path_xy)))

    def counters(self):
        p = C.c_uint64(0)
        _chk(lib().pb_tbrush_counters(self.h, C.byref(p)))
        return p.value
