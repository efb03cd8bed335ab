"""torch.distributed glue for one canvas sharded into row bands over the GPUs of a node (one process per GPU).

Plumbing only: it creates this rank's band canvas, exchanges CUDA IPC handles of the brush's canvas-record / snapshot /
dirty-map / flag allocations with `all_gather_object`, maps the peers' memory (NVLink) and fills the `pb_dist_desc` that
`pb_fbrush_stroke_batch_dist` consumes. The rendering itself is the C ABI / CUDA library.
"""
import ctypes as C

import numpy as np

from . import api

_VP = C.c_void_p
MAX_BANDS = 8


class pb_dist_desc(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("rows_per_band", C.c_int32), ("reserved", C.c_int32),
                ("canvas_base", _VP * MAX_BANDS), ("canvas_stride", C.c_int64 * MAX_BANDS),
                ("snapshot_base", _VP * MAX_BANDS), ("snapshot_stride", C.c_int64 * MAX_BANDS),
                ("dirty_base", _VP * MAX_BANDS), ("flags_base", _VP * MAX_BANDS)]


def rows_per_band(rows, world):
    return (rows + world - 1) // world


class DistCanvas:
    """This rank's band of a rows x cols canvas. Rank r owns rows [r*rpb, min((r+1)*rpb, rows))."""

    def __init__(self, ctx, rows, cols, dist):
        self.ctx, self.rows, self.cols, self.dist = ctx, rows, cols, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        assert self.world <= MAX_BANDS
        self.rpb = rows_per_band(rows, self.world)
        self.row_begin = min(self.rank * self.rpb, rows)
        self.row_end = min(self.row_begin + self.rpb, rows)
        self.canvas = api.Canvas(ctx, rows, cols, band=(self.row_begin, self.row_end, 0))
        self._desc = {}
        self._imported = []

    def _export(self, ptr):
        h = (C.c_ubyte * 64)()
        api._chk(api.lib().pb_ipc_export(self.ctx.h, _VP(ptr), h))
        return bytes(h)

    def _import(self, handle):
        p = _VP()
        api._chk(api.lib().pb_ipc_import(self.ctx.h, (C.c_ubyte * 64).from_buffer_copy(handle), C.byref(p)))
        self._imported.append(p.value)
        return p.value

    def attach(self, brush):
        """Exchange the peer mappings for this (canvas, brush) pair. Collective: every rank must call it."""
        lib = api.lib()
        wbase, sbase, dbase, fbase = _VP(), _VP(), _VP(), _VP()
        api._chk(lib.pb_fbrush_dist_storage(brush.h, self.canvas.h, C.byref(wbase), C.byref(sbase), C.byref(dbase), C.byref(fbase)))
        mine = dict(canvas=self._export(wbase.value), snapshot=self._export(sbase.value), dirty=self._export(dbase.value),
                    flags=self._export(fbase.value))
        everyone = [None] * self.world
        self.dist.all_gather_object(everyone, mine)
        d = pb_dist_desc()
        d.world, d.rank, d.rows_per_band = self.world, self.rank, self.rpb
        own = dict(canvas=wbase.value, snapshot=sbase.value, dirty=dbase.value, flags=fbase.value)
        for r, e in enumerate(everyone):
            for key, arr in (("canvas", d.canvas_base), ("snapshot", d.snapshot_base), ("dirty", d.dirty_base), ("flags", d.flags_base)):
                arr[r] = own[key] if r == self.rank else self._import(e[key])
        self._desc[id(brush)] = d
        self.dist.barrier()
        return d

    def plan(self, brush, strokes, cx, cy, theta):
        """Host half of stroke_batch: every rank plans the same global stroke list (submission order). Touches no stream —
        call it on a helper thread to overlap the planning of the next batch with the execution of the current one."""
        d = self._desc.get(id(brush)) or self.attach(brush)
        return brush.plan_stroke_batch(self.canvas, strokes, cx, cy, theta, dist_desc=d)

    def stroke_batch(self, brush, strokes=None, cx=None, cy=None, theta=None, plan=None, after_launch=None):
        """Every rank passes the same global stroke list (submission order), or a plan made from it by plan().
        after_launch() runs once this rank's kernels are queued (e.g. to start planning the next batch)."""
        d = self._desc.get(id(brush)) or self.attach(brush)
        lib = api.lib()
        api._chk(lib.pb_fbrush_dist_begin(brush.h, self.canvas.h))  # planes -> records of this band (synchronises the stream)
        if plan is None:
            plan = self.plan(brush, strokes, cx, cy, theta)
        self.dist.barrier()  # every band is ready (cleared / snapshot taken / converted) before any kernel touches peer rows
        brush.run_batch_plan(self.canvas, plan, dist_desc=d)
        if after_launch is not None:
            after_launch()
        self.ctx.synchronize()
        self.dist.barrier()  # all GPUs have finished writing into each other's bands
        api._chk(lib.pb_fbrush_dist_end(brush.h, self.canvas.h))  # records -> planes

    # ---- final assembly of the reflectance image: compose with a peer-store epilogue (no separate collective) ----
    def attach_image(self, root=None):
        """Create the assembled-image buffers and exchange their peer mappings. root=None: every rank receives the image
        (all-gather); root=r: only rank r does. Collective."""
        lib = api.lib()
        if getattr(self, "_image", None):  # re-attach: nobody may still be writing into the old image
            self.ctx.synchronize()
            self.dist.barrier()
            lib.pb_band_image_destroy(self._image)
        self._image_root = root
        self._image = _VP()
        mine = None
        if root is None or root == self.rank:
            api._chk(lib.pb_band_image_create(self.ctx.h, self.rows, self.cols, C.byref(self._image)))
            base, stride = _VP(), C.c_int64(0)
            api._chk(lib.pb_band_image_device(self._image, C.byref(base), C.byref(stride)))
            self._image_base, self._image_stride = base.value, stride.value
            mine = (self._export(base.value), stride.value)
        everyone = [None] * self.world
        self.dist.all_gather_object(everyone, mine)
        self._image_dst = []
        for r, e in enumerate(everyone):
            if e is None:
                continue
            self._image_stride = e[1]
            self._image_dst.append(self._image_base if r == self.rank else self._import(e[0]))
        self.dist.barrier()

    def compose_gather(self):
        """Renderer::compose of this rank's band, stored straight into the destination images over NVLink. The images are
        complete after every rank's context has synchronised and a process-group barrier (finish_gather)."""
        if not hasattr(self, "_image_dst"):
            self.attach_image()
        arr = (_VP * len(self._image_dst))(*self._image_dst)
        api._chk(api.lib().pb_canvas_compose_gather(self.canvas.h, len(self._image_dst), arr, C.c_int64(self._image_stride)))

    def finish_gather(self):
        self.ctx.synchronize()
        self.dist.barrier()

    def image_device(self):
        """(device base pointer, plane stride in bytes) of this rank's assembled image, or None on a non-root rank."""
        return (self._image_base, self._image_stride) if self._image else None

    def download_image(self, out=None):
        """Assembled reflectance as host AoS f64 [rows, cols, 3] (ranks that hold an image)."""
        out = np.empty((self.rows, self.cols, 3)) if out is None else out
        api._chk(api.lib().pb_band_image_download(self._image, api._p(out)))
        return out

    def close(self):
        self.ctx.synchronize()
        self.dist.barrier()
        if getattr(self, "_image", None):
            api.lib().pb_band_image_destroy(self._image)
            self._image = None
        for p in self._imported:
            api.lib().pb_ipc_close(self.ctx.h, _VP(p))
        self._imported = []
