// Index geometry of the footprint-brush imprint, shared by the imprint kernel and by host test hooks (pb_imprint_hits,
// pb_ring_rects) so that the arithmetic the device runs is the arithmetic the CPU tests check against a brute-force
// restatement of FootprintBrush.hxx:88-114 (cell of a pixel) and :278-319 (snapshot ring).
//
//  * hits_exact / hits_fast: the <= 2 canvas pixels whose rotated + rounded position is one pickup-map cell.
//    hits_exact decides with the reference's f64 expressions (same operation order, no FMA). hits_fast decides in
//    single precision and reports "ambiguous" whenever a candidate lies within eps of a rounding boundary (eps is a
//    bound of the float error, see hits_fast); the caller then falls back to hits_exact. Both paths therefore give
//    the reference's answer bit for bit; the fast one runs ~6x fewer instructions and no FP64.
//  * ring_rects: the snapshot ring "allowed box minus open interior" of an imprint as a list of <= 8 rectangles —
//    either the whole ring (4 rectangles) or only the part that was not in the previous imprint's ring (incremental
//    update, O(perimeter)). The dirty map is flat (byte index == pixel index) and is scanned in aligned 32-bit words.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define PB_HD __host__ __device__ __forceinline__
#else
#define PB_HD inline
#endif

namespace pb {

// Per-imprint constants, computed on the host in f64 with the reference's libm (FootprintBrush.hxx:95-96).
struct alignas(16) DevImprint {
  double cx, cy, c, s;  // centre, cos(-theta), sin(-theta)
  float fc, fs;         // (float) c, s
  int32_t ix, iy;       // floor(cx), floor(cy)
  int32_t flags;        // kIm*
  int32_t pad[3];
};
constexpr int kImFastXY = 1;  // (int)(col + cx) == col + ix + carry for every integer col (same for y), see make_imprint
constexpr int kImFracX  = 2;  // cx has a fractional part
constexpr int kImFracY  = 4;
constexpr int kImBorder = 8;  // the footprint box overhangs the left or top border: pixels of column/row 0 can be hit twice

// trunc(col + cx) for integer col: with ix = floor(cx), f = cx - ix in [0,1) the sum is (col + ix) + f, so truncation
// toward zero gives col + ix when that is >= 0 or f == 0, and col + ix + 1 otherwise. The double addition rounds, which
// can only change the truncated value when f is within an ulp of 0 or 1 — those imprints are flagged "not fast".
inline DevImprint make_imprint(double cx, double cy, double theta, int wr) {
  DevImprint im{};
  im.cx = cx, im.cy = cy;
  im.c = std::cos(-theta), im.s = std::sin(-theta);
  im.fc = static_cast<float>(im.c), im.fs = static_cast<float>(im.s);
  const double flx = std::floor(cx), fly = std::floor(cy);
  const double fx = cx - flx, fy = cy - fly;
  auto safe = [](double v, double f) { return std::fabs(v) < 1048576.0 && (f == 0.0 || (f > 1e-6 && f < 1.0 - 1e-6)); };
  int flags = 0;
  if (safe(cx, fx) && safe(cy, fy)) {
    flags |= kImFastXY;
    im.ix = static_cast<int32_t>(flx), im.iy = static_cast<int32_t>(fly);
  }
  if (fx > 0.0) flags |= kImFracX;
  if (fy > 0.0) flags |= kImFracY;
  if ((cx - wr < 0.0) || (cy - wr < 0.0)) flags |= kImBorder;
  im.flags = flags;
  return im;
}

struct PixelHits {
  int px[2], py[2];  // canvas pixel, row-major order of the reference's (row, col) loop
  int n;             // 0..2, or -1: hits_fast could not decide
  // append without a dynamically indexed store (keeps the struct in registers on the device)
  PB_HD void push(int x, int y) {
    if (n == 0) px[0] = x, py[0] = y;
    if (n == 1) px[1] = x, py[1] = y;
    ++n;
  }
};

// Exact: the 2x2 lattice neighbourhood of the cell's pre-image, float pre-filter with a wide margin, then the reference's
// f64 forward expression (:95-100) for the survivors. `ph` selects the border phase ((fy >= 0) * 2 + (fx >= 0)), -1 = all.
PB_HD void hits_exact(const DevImprint& im, int wr, int mx, int my, int rows, int cols, int ph, PixelHits& h) {
  h.n = 0;
  h.px[0] = h.px[1] = h.py[0] = h.py[1] = 0;
  const float fc = im.fc, fs = im.fs;
  const float u = static_cast<float>(mx - wr), v = static_cast<float>(my - wr);
  const float colf = fmaf(u, fc, v * fs), rowf = fmaf(v, fc, -(u * fs));
  const int c0 = static_cast<int>(floorf(colf)), r0 = static_cast<int>(floorf(rowf));
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int dr = 0; dr < 2; ++dr) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int dc = 0; dc < 2; ++dc) {
      const int row = r0 + dr, col = c0 + dc;
      const float ec = static_cast<float>(col) - colf, er = static_cast<float>(row) - rowf;
      const float du = fmaf(ec, fc, -(er * fs)), dv = fmaf(ec, fs, er * fc);
      if (fabsf(du) > 0.51f || fabsf(dv) > 0.51f) continue;
      if (col < -wr || col > wr || row < -wr || row > wr) continue;
      const double rc = col * im.c - row * im.s;
      const double rr = col * im.s + row * im.c;
      if (static_cast<int>(round(rc + wr)) != mx || static_cast<int>(round(rr + wr)) != my) continue;
      const double fx = col + im.cx, fy = row + im.cy;
      const int px = static_cast<int>(fx), py = static_cast<int>(fy);  // trunc toward zero (:92-93)
      if (py < 0 || px < 0 || px >= cols || py >= rows) continue;
      if (ph >= 0 && ((fy >= 0.0 ? 2 : 0) + (fx >= 0.0 ? 1 : 0)) != ph) continue;
      h.push(px, py);  // a unit cell cannot hold more than 2 lattice points (min distance 1 < diagonal sqrt 2)
    }
  }
  if (h.n > 2) h.n = 2;
}

// Single precision. (u, v) = (mx - wr, my - wr) as floats. A candidate pixel (col, row) maps to the cell iff its rotated
// offset from the cell centre (du, dv) lies in [-0.5, 0.5) x [-0.5, 0.5) up to the tie rule of round(); du and dv are
// evaluated here with float c and s, whose error is bounded by
//   |u| + |v| rounding steps of 2^-24 each on the inverse rotation  +  (|col| + |row| + |u|) * 2^-24 for float(c), float(s)
//   <= 5 * wr * 6e-8 = wr * 3e-7,
// so with eps >= wr * 1e-6 (the caller passes lo = 0.5 - eps, hi = 0.5 + eps) every candidate outside the band
// [lo, hi] is decided exactly like the f64 expression; candidates inside the band make the call return n = -1.
// Loop-range and map-range checks are the caller's business: it only passes cells whose pre-image lies inside the loop
// range (max active-cell radius <= wr - 2, checked when the footprint is registered).
PB_HD void hits_fast(float fc, float fs, int ix, int iy, int flags, float u, float v, float lo, float hi, int rows, int cols,
                     int ph, PixelHits& h) {
  h.n = 0;
  h.px[0] = h.px[1] = h.py[0] = h.py[1] = 0;
  if (!(flags & kImFastXY)) {
    h.n = -1;
    return;
  }
  const float colf = fmaf(u, fc, v * fs), rowf = fmaf(v, fc, -(u * fs));
  const float c0f = floorf(colf), r0f = floorf(rowf);
  const float ec = c0f - colf, er = r0f - rowf;  // in (-1, 0]
  // rotated offset of candidate (c0 + dc, r0 + dr) from the cell centre: (du0, dv0) + dc * (fc, fs) + dr * (-fs, fc)
  const float du0 = fmaf(ec, fc, -(er * fs)), dv0 = fmaf(ec, fs, er * fc);
  unsigned inside = 0u, band = 0u;  // bit 2 * dr + dc: row-major order of the candidates
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 4; ++i) {
    const int dc = i & 1, dr = i >> 1;
    const float a = fabsf(du0 + (dc ? fc : 0.0f) - (dr ? fs : 0.0f));
    const float b = fabsf(dv0 + (dc ? fs : 0.0f) + (dr ? fc : 0.0f));
    const float m = fmaxf(a, b);
    if (m < lo) inside |= 1u << i;
    if (!(m < lo) && !(m > hi)) band |= 1u << i;
  }
  if (band) {
    h.n = -1;
    return;
  }
  const int c0 = static_cast<int>(c0f) + ix, r0 = static_cast<int>(r0f) + iy;
  const bool fracx = (flags & kImFracX) != 0, fracy = (flags & kImFracY) != 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 2; ++k) {  // a unit cell holds at most 2 lattice points
    if (inside == 0u) break;
#if defined(__CUDA_ARCH__)
    const int i = __ffs(static_cast<int>(inside)) - 1;
#else
    const int i = __builtin_ctz(inside);
#endif
    inside &= inside - 1u;
    const int tx = c0 + (i & 1), ty = r0 + (i >> 1);
    const int px = tx + ((tx < 0 && fracx) ? 1 : 0), py = ty + ((ty < 0 && fracy) ? 1 : 0);
    if (py < 0 || px < 0 || px >= cols || py >= rows) continue;
    if (ph >= 0 && ((ty >= 0 ? 2 : 0) + (tx >= 0 ? 1 : 0)) != ph) continue;
    h.push(px, py);
  }
}

// ---- snapshot ring ----------------------------------------------------------------------------------------------
struct RingGeom {
  int tlx, tly, brx, bry;  // footprint box corners (exclusive interior bounds)
  int ax0, ay0, ax1, ay1;  // allowed box, clipped to canvas and stored rows
};

// FootprintBrush.hxx:290-305 for centre (cx, cy), half side wr (== hr), brush radius; rows [row_lo, row_hi] are stored.
PB_HD RingGeom ring_geom(double cx, double cy, int wr, double radius, int cols, int row_lo, int row_hi) {
  RingGeom g;
  g.tlx = static_cast<int>(cx - wr), g.tly = static_cast<int>(cy - wr);
  g.brx = static_cast<int>(cx + wr), g.bry = static_cast<int>(cy + wr);
  const int ax0 = static_cast<int>(cx - wr - radius), ay0 = static_cast<int>(cy - wr - radius);
  const int ax1 = static_cast<int>(cx + wr + radius), ay1 = static_cast<int>(cy + wr + radius);
  g.ax0 = ax0 > 0 ? ax0 : 0;
  g.ay0 = ay0 > row_lo ? ay0 : row_lo;
  g.ax1 = ax1 < cols - 1 ? ax1 : cols - 1;
  g.ay1 = ay1 < row_hi ? ay1 : row_hi;
  return g;
}

PB_HD bool in_ring(const RingGeom& g, int row, int col) {
  if (col < g.ax0 || col > g.ax1 || row < g.ay0 || row > g.ay1) return false;
  return !(row > g.tly && row < g.bry && col > g.tlx && col < g.brx);
}

struct Rect {
  int x0, y0, x1, y1;  // inclusive; empty if x1 < x0 or y1 < y0
};

// Calls f(Rect) for the non-empty parts of (A \ B) clipped to the allowed box of `clip`: top, bottom, left, right.
template <typename F>
PB_HD void rect_difference(const Rect& A, const Rect& B, const RingGeom& clip, F&& f) {
  auto emit = [&](int x0, int y0, int x1, int y1) {
    x0 = x0 > clip.ax0 ? x0 : clip.ax0;
    y0 = y0 > clip.ay0 ? y0 : clip.ay0;
    x1 = x1 < clip.ax1 ? x1 : clip.ax1;
    y1 = y1 < clip.ay1 ? y1 : clip.ay1;
    if (x1 >= x0 && y1 >= y0) f(Rect{x0, y0, x1, y1});
  };
  if (A.x1 < A.x0 || A.y1 < A.y0) return;
  if (B.x1 < B.x0 || B.y1 < B.y0 || B.x0 > A.x1 || B.x1 < A.x0 || B.y0 > A.y1 || B.y1 < A.y0) {
    emit(A.x0, A.y0, A.x1, A.y1);
    return;
  }
  emit(A.x0, A.y0, A.x1, B.y0 - 1);
  emit(A.x0, B.y1 + 1, A.x1, A.y1);
  const int my0 = A.y0 > B.y0 ? A.y0 : B.y0, my1 = A.y1 < B.y1 ? A.y1 : B.y1;
  emit(A.x0, my0, B.x0 - 1, my1);
  emit(B.x1 + 1, my0, A.x1, my1);
}

// The rectangles a ring pass has to scan, in a fixed order (host and device enumerate the same list):
//  * prev == nullptr: the whole ring of g — allowed box minus open interior (<= 4 rectangles);
//  * otherwise ring(g) minus ring(prev), or a superset of it (<= 8 rectangles): the pixels that entered the allowed box
//    plus the pixels that left the interior. Pixels of ring(g) that already were in ring(prev) were copied by the
//    previous pass and cannot have been touched since: the previous imprint only touches pixels of its own open interior
//    (compact footprints), and no other stroke writes them within a dataflow segment.
template <typename F>
PB_HD void ring_rects(const RingGeom& g, const RingGeom* prev, F&& f) {
  const Rect allowed{g.ax0, g.ay0, g.ax1, g.ay1};
  const Rect interior{g.tlx + 1, g.tly + 1, g.brx - 1, g.bry - 1};
  if (prev == nullptr) {
    rect_difference(allowed, interior, g, f);
    return;
  }
  const Rect allowed_prev{prev->ax0, prev->ay0, prev->ax1, prev->ay1};
  const Rect interior_prev{prev->tlx + 1, prev->tly + 1, prev->brx - 1, prev->bry - 1};
  rect_difference(allowed, allowed_prev, g, f);
  rect_difference(interior_prev, interior, g, f);
}

// A rectangle is scanned as (rows) x (nw words): word j of row y covers the flat dirty-map bytes
// [4 * (((y_local * pitch + x0) >> 2) + j), + 4). nw words cover [x0, x1] of a row whatever its alignment; bytes outside
// the ring are masked with in_ring by the scanner.
PB_HD int rect_words(const Rect& r) { return (r.x1 - r.x0 + 3) / 4 + 1; }
// item q in [0, rows * nw) -> (row, j); inv_nw ~ 1 / nw (any float within a few ulp works, the quotient is corrected)
PB_HD void rect_item(const Rect& r, int nw, float inv_nw, int q, int& row, int& j) {
  int rr = static_cast<int>((static_cast<float>(q) + 0.5f) * inv_nw);
  if (rr * nw > q) --rr;
  if ((rr + 1) * nw <= q) ++rr;
  row = r.y0 + rr;
  j   = q - rr * nw;
}


// The items (dirty-map words) of one ring pass, numbered 0 .. total-1 across the rectangles of ring_rects in their fixed
// order. The device walks them with one thread per item and stride; the host test hook walks them all.
struct RingList {
  Rect r[8];
  int n, total;
};
PB_HD void ring_list(const RingGeom& g, const RingGeom* prev, RingList& out) {
  out.n = 0, out.total = 0;
  ring_rects(g, prev, [&](const Rect& r) {
    if (out.n < 8) {
      out.r[out.n++] = r;
      out.total += (r.y1 - r.y0 + 1) * rect_words(r);
    }
  });
}
// item t (0 <= t < total) -> canvas row, first column x0 of its rectangle and word index j within the row of the rectangle:
// the word covers the flat dirty-map bytes [4 * (((row_local * pitch + x0) >> 2) + j), + 4)
PB_HD void ring_list_item(const RingList& rl, int t, int& row, int& x0, int& j) {
  int ri = 0, nw = rect_words(rl.r[0]), cnt = (rl.r[0].y1 - rl.r[0].y0 + 1) * nw;
  while (t >= cnt && ri + 1 < rl.n) {
    t -= cnt;
    ++ri;
    nw  = rect_words(rl.r[ri]);
    cnt = (rl.r[ri].y1 - rl.r[ri].y0 + 1) * nw;
  }
#if defined(__CUDA_ARCH__)
  const float inv = __frcp_rn(static_cast<float>(nw));
#else
  const float inv = 1.0f / static_cast<float>(nw);
#endif
  rect_item(rl.r[ri], nw, inv, t, row, j);
  x0 = rl.r[ri].x0;
}

}  // namespace pb
