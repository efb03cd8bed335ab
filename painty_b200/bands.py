"""Row-band sharding of a canvas over the GPUs of one node (one process per GPU): host-side routing helpers.

What shards, and how exactly:

* KM compose / dry: independent pixels -> every rank composes its own rows, no exchange. The final image is assembled by
  the compose kernel itself: `pb_canvas_compose_gather` stores the band's reflectance rows straight into the destination
  ranks' images over NVLink peer mappings (`painty_b200/dist.py: DistCanvas.compose_gather`); `gather_bands` below is the
  torch.distributed collective used by the CPU (gloo) tests of the host logic.
* Texture-brush strokes (no smudge): a pixel's result depends only on the earlier strokes that cover that pixel, so a
  rank applies — in submission order — every stroke whose bounding box meets its band; the kernel clips to the
  stored rows. No halo, no exchange, bit-identical to the single-canvas result (`route_texture_strokes`).
* Footprint-brush strokes carry state along the whole stroke (pickup map, snapshot buffer), so a stroke is executed
  entirely by the rank that owns its first imprint. The routing and the cross-GPU dataflow live in the library
  (`pb_fbrush_stroke_batch_dist`, csrc/capi.cu + csrc/imprint.cu): strokes that leave their band ("straddlers") stage the
  neighbour rows of each 64-imprint dataflow segment in local windows (pulled and pushed over NVLink per segment) or access
  them directly through peer mappings; stroke order across GPUs is kept by progress words polled through peer memory.
  `route_footprint_strokes` / `stroke_levels` restate that routing rule on the host for the CPU tests
  (tests/test_bands_cpu.py); the product path does not call them.

Nothing here touches a GPU: it is host-side planning, testable with the gloo backend.
"""
import math

import numpy as np

from . import assets


def band_ranges(rows, world):
    """Equal row bands [begin, end) per rank; the first rows % world bands get one extra row."""
    base, extra = divmod(rows, world)
    out, r0 = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((r0, r0 + n))
        r0 += n
    return out


def footprint_stroke_rows(radius, cy):
    """Inclusive row interval a footprint stroke reads or writes: the union of updateSnapshot's allowed boxes
    (FootprintBrush.hxx:298-305), i.e. centres +- ((side-1)/2 + radius) with a 2-row safety margin."""
    side = assets.footprint_geometry(radius)[3]
    m = (side - 1) // 2 + radius + 2.0
    return int(math.floor(float(np.min(cy)) - m)), int(math.ceil(float(np.max(cy)) + m))


def texture_stroke_rows(radius, path):
    """Inclusive row interval of TextureBrush::paintStroke's bounding box (TextureBrush.hxx:58-84): vertices,
    the two end extensions by `radius`, then +- radius."""
    y = np.asarray(path, dtype=np.float64).reshape(-1, 2)[:, 1]
    return int(math.floor(float(y.min()) - 2.0 * radius - 1.0)), int(math.ceil(float(y.max()) + 2.0 * radius + 1.0))


def route_texture_strokes(strokes, rows, world):
    """strokes: list of dict(radius, path). Returns per rank the indices (in submission order) of the strokes whose
    bounding box meets the rank's band."""
    bands = band_ranges(rows, world)
    out = [[] for _ in range(world)]
    radius = 0.0
    for i, s in enumerate(strokes):
        if not abs(radius - s["radius"]) < 0.5:  # TextureBrush::setRadius fuzzy rule (TextureBrush.hxx:33-41)
            radius = s["radius"]
        if len(np.asarray(s["path"]).reshape(-1, 2)) < 2:
            continue
        lo, hi = texture_stroke_rows(radius, s["path"])
        for r, (b, e) in enumerate(bands):
            if hi >= b and lo < e:
                out[r].append(i)
    return out


def route_footprint_strokes(radii, cy_per_stroke, rows, world):
    """Owner rank per stroke (band of its first imprint, clamped into the canvas) and a straddler mask (the stroke's
    row interval leaves the owner's band)."""
    bands = band_ranges(rows, world)
    starts = np.array([b for b, _ in bands])
    owner = np.zeros(len(radii), dtype=np.int64)
    straddles = np.zeros(len(radii), dtype=bool)
    for i, (r, cy) in enumerate(zip(radii, cy_per_stroke)):
        if len(cy) == 0:
            continue
        y0 = min(max(int(cy[0]), 0), rows - 1)
        o = int(np.searchsorted(starts, y0, side="right") - 1)
        owner[i] = o
        lo, hi = footprint_stroke_rows(r, cy)
        straddles[i] = max(lo, 0) < bands[o][0] or min(hi, rows - 1) >= bands[o][1]
    return owner, straddles


def stroke_levels(regions, rows, cols, tile=64):
    """Dataflow level (wave index) of each stroke: 1 + the largest level of any earlier stroke whose region
    (x0,y0,x1,y1 inclusive, clipped) shares a tile with it. Strokes of one level are pairwise disjoint."""
    tx, ty = (cols + tile - 1) // tile, (rows + tile - 1) // tile
    lvl = np.zeros((ty, tx), dtype=np.int64)
    out = np.zeros(len(regions), dtype=np.int64)
    for i, (x0, y0, x1, y1) in enumerate(regions):
        if x1 < x0 or y1 < y0:
            continue
        sl = (slice(max(y0, 0) // tile, min(y1, rows - 1) // tile + 1), slice(max(x0, 0) // tile, min(x1, cols - 1) // tile + 1))
        out[i] = int(lvl[sl].max()) + 1
        lvl[sl] = out[i]
    return out


def gather_bands(band_planes, rows, cols, world, dist=None):
    """Assemble the full image from per-rank band results. band_planes: tensor [3, band_rows*cols] of this rank.
    Bands may differ by one row, so every rank pads to the largest band before the all_gather. Returns
    [3, rows, cols] on every rank (world == 1: no communication)."""
    import torch

    bands = band_ranges(rows, world)
    if world == 1 or dist is None:
        return band_planes.reshape(3, rows, cols)
    max_rows = max(e - b for b, e in bands)
    rank = dist.get_rank()
    mine = band_planes.reshape(3, -1, cols)
    if mine.shape[1] < max_rows:
        mine = torch.cat([mine, mine.new_zeros((3, max_rows - mine.shape[1], cols))], dim=1)
    mine = mine.contiguous()
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    assert bands[rank][1] - bands[rank][0] == band_planes.numel() // (3 * cols)
    return torch.cat([p[:, :e - b] for p, (b, e) in zip(parts, bands)], dim=1)
