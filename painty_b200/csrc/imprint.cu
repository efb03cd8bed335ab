// Footprint-brush imprint engine for sm_100a: batched, order-preserving pickup/deposit.
//
// Restates painty/renderer/FootprintBrush.hxx:73-143 (imprint), :278-319 (updateSnapshot),
// :331-340 (blend), :349-384 (pickupPaint), :393-431 (depositPaint) on SoA planes in HBM.
//
// Parallel decomposition (nothing like the reference's serial double loop):
//   * a STROKE (dip -> setRadius -> chain of imprints) is owned by one persistent thread-block CLUSTER
//     (1..16 CTAs, ~2 active footprint cells per thread); strokes are popped from a queue in submission
//     order. Dependencies are tracked per SEGMENT of a stroke (64 imprints): a segment waits until the earlier
//     strokes whose footprint+snapshot region overlaps its own have published enough progress (host-built
//     lists of (stroke, segments needed), schedule.hpp) — a dataflow schedule that keeps the reference's
//     stroke order wherever it is observable while letting overlapping strokes follow each other closely.
//     Consecutive imprints of a stroke are a true dependency chain; the cluster-wide hardware barrier between
//     them is the latency floor;
//   * inside an imprint a thread owns ACTIVE pickup-map cells (footprint height > 0, ~14.5 % of the
//     padded square, compacted once per radius). Its pickup-map state (7 values per cell) lives in the
//     CTA's shared memory for the whole stroke. For each imprint the thread inverts the rotation to find
//     the <= 2 canvas pixels whose rotated+rounded position is its cell, checks each candidate with the
//     reference's exact f64 forward expression, and applies pickup+deposit to them in row-major order —
//     exactly the order in which the reference's (row, col) loop hits a shared pickup cell. No two
//     threads ever touch the same cell;
//   * canvas pixels are hit at most once per imprint except at the left/top border, where C++
//     truncation folds column/row (-1,0) onto 0. Those imprints run in <= 4 barrier-separated
//     phases ordered by (row-negative?, col-negative?) which reproduces the row-major order;
//   * updateSnapshot copies a canvas ring of ~2x the footprint area on EVERY imprint in the reference. Here a
//     byte-per-pixel dirty map records where snapshot and canvas can differ (only pixels touched by an
//     imprint since their last copy); the ring pass scans the map with 32-bit loads and copies just the dirty
//     pixels — bit-identical result, ~50x less traffic;
//   * per-imprint constants (centre, cos/sin(-theta)) are computed on the host in f64 with the same libm
//     as the reference; all index maths on the device is IEEE f64 without FMA contraction, so every
//     round()/trunc() decision is bit-identical to the CPU's;
//   * canvas, snapshot and dirty planes are accessed with L2-only loads/stores (ld/st.global.cg): they are
//     shared between SMs, and the 126 MB L2 keeps the working set of the running strokes resident.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>
#include <type_traits>

#include "imprint.cuh"
#include "ring_words.hpp"

namespace cg = cooperative_groups;

namespace pb {
namespace {

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// progress words: (batch epoch << 32) | segments completed
__device__ __forceinline__ long long ld_acquire64(const long long* p) {
  long long v;
  asm volatile("ld.acquire.gpu.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release64(long long* p, long long v) {
  asm volatile("st.release.gpu.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long ld_acquire64_sys(const long long* p) {
  long long v;
  asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release64_sys(long long* p, long long v) {
  asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// FootprintBrush.hxx:331-340: blend(va, a, vb, b) = vt > MinVolume ? (va*a + vb*b)/vt : a, applied to the six
// K/S components that share one (va, vb). FP64 keeps the reference's exact operation order. FP32 mode evaluates
// the six quotients with one reciprocal (error ~1e-7 relative, far inside the 1e-4 reflectance budget).
struct Blend6d {
  double va, vb, vt;
  bool on;
  __device__ __forceinline__ Blend6d(double a, double b) : va(a), vb(b), vt(a + b), on(a + b > kMinVolume) {}
  __device__ __forceinline__ double operator()(double a, double b) const { return on ? (va * a + vb * b) / vt : a; }
};
struct Blend6f {
  float wa, wb;
  bool on;
  __device__ __forceinline__ Blend6f(float a, float b) {
    const float vt = a + b;
    on             = vt > static_cast<float>(kMinVolume);
    const float r  = __fdividef(1.0f, vt);  // rcp.approx, 1 ulp
    wa             = a * r;
    wb             = b * r;
  }
  __device__ __forceinline__ float operator()(float a, float b) const { return on ? fmaf(wa, a, wb * b) : a; }
};
template <typename T>
struct BlendSel;
template <>
struct BlendSel<float> {
  using type = Blend6f;
};
template <>
struct BlendSel<double> {
  using type = Blend6d;
};

template <typename T>
struct OpCtx {
  T pickup_rate, deposition_rate, cap;
  T paintK[3], paintS[3];
};

// View of one row band: plane pointers, dirty map and row pitches, indexed with (band-local row) * pitch + column.
// Single GPU: built from the launch parameters (band is a compile-time 0, everything folds into constant-bank
// loads). Multi GPU: one view per band in shared memory, rebuilt per stroke — the executor's own band, peer
// mappings, or local staging windows (virtual base pointers so that the same index arithmetic works).
template <typename T>
struct Band {
  T* can[kLayerPlanes];
  T* src[kLayerPlanes];
  unsigned char* dirty;
  unsigned char* touched;  // staging windows only: pixels to push back to their owner
  int pitch, dpitch;
  __device__ __forceinline__ Band() {}
  __device__ __forceinline__ Band(const ImprintLaunch& L, int band) {
#pragma unroll
    for (int k = 0; k < kLayerPlanes; ++k) {
      can[k] = static_cast<T*>(L.canvas[band][k]);
      src[k] = static_cast<T*>(L.snapshot[band][k]);
    }
    dirty   = L.dirty[band];
    touched = nullptr;
    pitch   = L.cols;
    dpitch  = L.dirty_pitch;
  }
  // the executor's own band, from dedicated launch fields (compile-time offsets into the constant bank)
  __device__ __forceinline__ explicit Band(const ImprintLaunch& L) {
#pragma unroll
    for (int k = 0; k < kLayerPlanes; ++k) {
      can[k] = static_cast<T*>(L.own_canvas[k]);
      src[k] = static_cast<T*>(L.own_snapshot[k]);
    }
    dirty   = L.own_dirty;
    touched = nullptr;
    pitch   = L.cols;
    dpitch  = L.dirty_pitch;
  }
};
// VIEWS = the stroke may touch rows of other bands (multi GPU, straddling strokes): pixels are addressed through the
// per-band views. Otherwise everything lies in the executor's own band and folds into constant-bank operands.
template <typename T, bool VIEWS>
__device__ __forceinline__ Band<T> band_view(const ImprintLaunch& L, const Band<T>* views, int band) {
  if (VIEWS) return views[band];
  return Band<T>(L);
}
// Owner band of a canvas row. Almost every row a straddling stroke touches still lies in the executor's own band.
__device__ __forceinline__ int band_of(const ImprintLaunch& L, int row) {
  const int b0 = L.my_band * L.rows_per_band;
  return (row >= b0 && row < b0 + L.rows_per_band) ? L.my_band : row / L.rows_per_band;
}

// One (canvas pixel, pickup cell) interaction = pickupPaint (:349-384) then depositPaint (:393-431).
// Split in two so that a thread can put the loads of all its interactions in flight before it computes any.
template <typename T>
struct OpData {
  T cK[3], cS[3], sK[3], sS[3], vSrc, vCan;
};

template <typename T>
__device__ __forceinline__ void op_load(const Band<T>& C, int ci, OpData<T>& d) {
  const bool own_src = C.src[PV] == C.can[PV];  // snapshot buffer disabled: pickup source is the canvas itself
  d.vSrc = __ldcg(C.src[PV] + ci);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    d.cK[k] = __ldcg(C.can[PK + k] + ci);
    d.cS[k] = __ldcg(C.can[PS + k] + ci);
  }
  d.vCan = own_src ? d.vSrc : __ldcg(C.can[PV] + ci);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    d.sK[k] = own_src ? d.cK[k] : __ldcg(C.src[PK + k] + ci);
    d.sS[k] = own_src ? d.cS[k] : __ldcg(C.src[PS + k] + ci);
  }
}

template <typename T>
__device__ __forceinline__ void op_finish(const OpCtx<T>& P, const Band<T>& C, int ci, T fh, const OpData<T>& d, T* pick, int ps,
                                          int slot) {
  using Blend = typename BlendSel<T>::type;
  const bool own_src = C.src[PV] == C.can[PV];
  T vCan = d.vCan;
  T vP   = pick[PV * ps + slot];
  T pK[3], pS[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    pK[k] = pick[(PK + k) * ps + slot];
    pS[k] = pick[(PS + k) * ps + slot];
  }
  // pickup
  const T leave = P.pickup_rate * d.vSrc * fh;
  if (leave > static_cast<T>(kMinVolume)) {
    const T remain = d.vSrc - leave;
    __stcg(C.src[PV] + ci, remain);
    if (own_src) vCan = remain;
    const Blend bl(vP, leave);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      pK[k] = bl(pK[k], d.sK[k]);
      pS[k] = bl(pS[k], d.sS[k]);
    }
    vP = vP + leave;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      pick[(PK + k) * ps + slot] = pK[k];
      pick[(PS + k) * ps + slot] = pS[k];
    }
  }
  // deposit
  const T vFree = fmax(static_cast<T>(0), P.cap - vP);
  const Blend b_src(vP, vFree);
  const T vLeave       = P.deposition_rate * vP * fh;
  pick[PV * ps + slot] = vP - vLeave;
  const T vB           = P.cap * fh;
  const Blend b_can(vB, vCan);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    __stcg(C.can[PK + k] + ci, b_can(b_src(pK[k], P.paintK[k]), d.cK[k]));
    __stcg(C.can[PS + k] + ci, b_can(b_src(pS[k], P.paintS[k]), d.cS[k]));
  }
  __stcg(C.can[PV] + ci, vB + vCan);
}

// The <= 2 canvas pixels whose rotated + rounded position is pickup cell (mx,my), in row-major (row, col) order.
// Inverse rotation gives the centre (colf,rowf) of the cell's pre-image, a unit square rotated by theta: a lattice
// point can only map to this cell if its rotated offset from the cell centre is within 0.5 in both axes. A float
// pre-filter with a 0.01 margin leaves 1-2 of the 2x2 neighbourhood; those are decided by the reference's exact
// f64 forward expression (:95-100), same operation order, no FMA. Returns the number of hits of phase `ph`.
struct Hits {
  int ci[2];   // pixel index in the stored planes (< 2^31, like the reference's int32 K(i))
  int dof[2];  // byte offset in the dirty map
  int band[2];
  int n;
};
template <typename T, bool MULTI>
__device__ __forceinline__ Hits find_hits(const ImprintLaunch& L, const Band<T>* views, const DevImprint& im, float fc, float fs,
                                          int wr, int mx, int my, bool border, int ph, int row_lo, int row_hi) {
  Hits h;
  h.n = 0;
  h.ci[0] = h.ci[1] = h.dof[0] = h.dof[1] = h.band[0] = h.band[1] = 0;
  const float u = static_cast<float>(mx - wr), v = static_cast<float>(my - wr);
  const float colf = fmaf(u, fc, v * fs), rowf = fmaf(v, fc, -u * fs);
  const int c0 = static_cast<int>(floorf(colf)), r0 = static_cast<int>(floorf(rowf));
#pragma unroll
  for (int dr = 0; dr < 2; ++dr) {
#pragma unroll
    for (int dc = 0; dc < 2; ++dc) {
      const int row = r0 + dr, col = c0 + dc;
      const float ec = static_cast<float>(col) - colf, er = static_cast<float>(row) - rowf;
      const float du = fmaf(ec, fc, -er * fs), dv = fmaf(ec, fs, er * fc);
      if (fabsf(du) > 0.51f || fabsf(dv) > 0.51f) continue;
      if (col < -wr || col > wr || row < -wr || row > wr) continue;
      const double rc = col * im.c - row * im.s;
      const double rr = col * im.s + row * im.c;
      if (static_cast<int>(round(rc + wr)) != mx || static_cast<int>(round(rr + wr)) != my) continue;
      const double fx = col + im.cx, fy = row + im.cy;
      const int px = static_cast<int>(fx), py = static_cast<int>(fy);  // trunc toward zero (:92-93)
      if (py < 0 || px < 0 || px >= L.cols || py >= L.rows) continue;
      if (border && ((fy >= 0.0 ? 2 : 0) + (fx >= 0.0 ? 1 : 0)) != ph) continue;
      int band = 0, lrow = py - row_lo, pitch = L.cols, dpitch = L.dirty_pitch;
      if (MULTI) {
        band   = band_of(L, py);  // the row's owner GPU
        lrow   = py - band * L.rows_per_band;
        pitch  = views[band].pitch;
        dpitch = views[band].dpitch;
      } else if (py < row_lo || py > row_hi) {
        continue;  // band canvas without peers: rows outside the stored window are not ours
      }
      const int ci = lrow * pitch + px, dof = lrow * dpitch + px;
      if (h.n == 0) h.ci[0] = ci, h.dof[0] = dof, h.band[0] = band;
      if (h.n == 1) h.ci[1] = ci, h.dof[1] = dof, h.band[1] = band;
      ++h.n;  // a unit cell cannot hold more than 2 lattice points (min distance 1 < diagonal sqrt 2)
    }
  }
  h.n = min(h.n, 2);
  return h;
}

// updateSnapshot(canvas, centre) (:278-319): copy canvas -> snapshot on the ring "allowed box minus open
// interior". Only pixels flagged in the dirty map can differ, so the pass scans the map (one 32-bit word = 4
// pixels) and copies just those. The imprint chain is latency bound, hence every thread first issues ALL of its
// word loads (kScanBatch independent L2 requests in flight), then all pixel loads of a dirty word, then stores.
constexpr int kScanBatch = 8;

template <typename T, bool MULTI>
__device__ __forceinline__ void ring_word(const ImprintLaunch& L, const Band<T>* views, const RingGeom& g, int row, int wi,
                                          unsigned word) {
  const bool mid = row > g.tly && row < g.bry;
  const int band = MULTI ? band_of(L, row) : 0;
  const int lrow = MULTI ? row - band * L.rows_per_band : row - L.store_first;
  const Band<T> C = band_view<T, MULTI>(L, views, band);
  unsigned char* drow = C.dirty + static_cast<int64_t>(lrow) * C.dpitch;
  const int64_t rbase = static_cast<int64_t>(lrow) * C.pitch;
  bool need[4];
  T v[4][kLayerPlanes];
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const int col = wi * 4 + b;
    need[b] = ((word >> (8 * b)) & 0xffu) != 0u && col >= g.ax0 && col <= g.ax1 && !(mid && col > g.tlx && col < g.brx);
    if (need[b]) {
#pragma unroll
      for (int k = 0; k < kLayerPlanes; ++k) v[b][k] = __ldcg(C.can[k] + rbase + col);
    }
  }
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    if (need[b]) {
      const int col = wi * 4 + b;
#pragma unroll
      for (int k = 0; k < kLayerPlanes; ++k) __stcg(C.src[k] + rbase + col, v[b][k]);
      __stcg(drow + col, static_cast<unsigned char>(0));
      if (MULTI && C.touched) __stcg(C.touched + static_cast<int64_t>(lrow) * C.dpitch + col, static_cast<unsigned char>(1));
    }
  }
}

template <typename T, bool MULTI>
__device__ __forceinline__ void ring_scan(const ImprintLaunch& L, const Band<T>* views, const RingGeom& g, int gt, int gstride) {
  const RingWords rw(g);  // exactly the words outside the open interior (ring_words.hpp)
  const int total = rw.total;
  for (int base = gt; base < total; base += gstride * kScanBatch) {
    unsigned word[kScanBatch];
    int rows[kScanBatch], wis[kScanBatch];
#pragma unroll
    for (int u = 0; u < kScanBatch; ++u) {
      const int i = base + u * gstride;
      word[u]     = 0u;
      if (i < total) {
        int row, wi;
        rw.at(i, row, wi);
        rows[u] = row;
        wis[u]  = wi;
        const int band = MULTI ? band_of(L, row) : 0;
        const int lrow = MULTI ? row - band * L.rows_per_band : row - L.store_first;
        const unsigned char* dbase = MULTI ? views[band].dirty : L.own_dirty;
        const int dpitch           = MULTI ? views[band].dpitch : L.dirty_pitch;
        word[u] = __ldcg(reinterpret_cast<const unsigned*>(dbase + static_cast<int64_t>(lrow) * dpitch) + wi);
      }
    }
#pragma unroll
    for (int u = 0; u < kScanBatch; ++u)
      if (word[u] != 0u) ring_word<T, MULTI>(L, views, g, rows[u], wis[u], word[u]);
  }
}

// ---- multi-GPU staging windows ------------------------------------------------------------------------------------
// Layout of one window in the stroke slot's scratch: 7 canvas planes | 7 snapshot planes (npx elements each) | dirty
// bytes | touched bytes, npx = rows * cols (cols a multiple of 4).
template <typename T>
struct Window {
  T* can[kLayerPlanes];
  T* src[kLayerPlanes];
  unsigned char *dirty, *touched;
  int band, row0, rows, ox, cols;
  __device__ __forceinline__ Window(const ImprintLaunch& L, const DevStroke& st, int group, int w) {
    band = st.win_band[w], row0 = st.win_row0[w], rows = st.win_rows[w], ox = st.win_ox, cols = st.win_cols;
    unsigned char* base = L.win_scratch + static_cast<int64_t>(group) * L.win_stride + static_cast<int64_t>(w) * (L.win_stride / 2);
    const int64_t npx   = static_cast<int64_t>(rows) * cols;
    T* planes           = reinterpret_cast<T*>(base);
#pragma unroll
    for (int k = 0; k < kLayerPlanes; ++k) {
      can[k] = planes + k * npx;
      src[k] = planes + (kLayerPlanes + k) * npx;
    }
    dirty   = base + 2 * kLayerPlanes * npx * static_cast<int64_t>(sizeof(T));
    touched = dirty + npx;
  }
};

// Build the per-band views of a stroke and pull its windows from the neighbours' HBM (coalesced rows over NVLink).
template <typename T>
__device__ __forceinline__ void stage_in(const ImprintLaunch& L, const DevStroke& st, Band<T>* views, int group, int sgt, int gstride,
                                         int tid) {
  if (tid < L.n_bands) {
    Band<T> v(L, tid);
    if (st.flags & 8) {
      for (int w = 0; w < 2; ++w) {
        if (st.win_band[w] != tid) continue;
        const Window<T> W(L, st, group, w);
        const int64_t off = static_cast<int64_t>(W.row0) * W.cols + W.ox;  // virtual base: index = lrow * cols + px
#pragma unroll
        for (int k = 0; k < kLayerPlanes; ++k) {
          v.can[k] = W.can[k] - off;
          v.src[k] = W.src[k] - off;
        }
        v.dirty   = W.dirty - off;
        v.touched = W.touched - off;
        v.pitch   = W.cols;
        v.dpitch  = W.cols;
      }
    }
    views[tid] = v;
  }
  if (!(st.flags & 8)) return;
  constexpr int CH = 16 / static_cast<int>(sizeof(T));  // pixels per 16-byte chunk
  const bool vec   = (L.cols & 3) == 0;                 // row starts of the peer planes are 16 B aligned
  for (int w = 0; w < 2; ++w) {
    if (st.win_band[w] < 0) continue;
    const Window<T> W(L, st, group, w);
    const int npx = W.rows * W.cols;
    if (vec) {
      const int cpr = W.cols / 4;  // 4-pixel groups per window row
      for (int i = sgt; i < W.rows * cpr; i += gstride) {
        const int lr = i / cpr, c = (i - lr * cpr) * 4, wi = lr * W.cols + c;
        const int64_t gi = static_cast<int64_t>(W.row0 + lr) * L.cols + W.ox + c;
        uint4 a[kLayerPlanes][4 / CH], b[kLayerPlanes][4 / CH];
#pragma unroll
        for (int k = 0; k < kLayerPlanes; ++k) {
#pragma unroll
          for (int j = 0; j < 4 / CH; ++j) {
            a[k][j] = __ldcg(reinterpret_cast<const uint4*>(static_cast<const T*>(L.canvas[W.band][k]) + gi) + j);
            b[k][j] = __ldcg(reinterpret_cast<const uint4*>(static_cast<const T*>(L.snapshot[W.band][k]) + gi) + j);
          }
        }
        const unsigned dflags =
          __ldcg(reinterpret_cast<const unsigned*>(L.dirty[W.band] + static_cast<int64_t>(W.row0 + lr) * L.dirty_pitch + W.ox + c));
#pragma unroll
        for (int k = 0; k < kLayerPlanes; ++k) {
#pragma unroll
          for (int j = 0; j < 4 / CH; ++j) {
            __stcg(reinterpret_cast<uint4*>(W.can[k] + wi) + j, a[k][j]);
            __stcg(reinterpret_cast<uint4*>(W.src[k] + wi) + j, b[k][j]);
          }
        }
        __stcg(reinterpret_cast<unsigned*>(W.dirty + wi), dflags);
        __stcg(reinterpret_cast<unsigned*>(W.touched + wi), 0u);
      }
      continue;
    }
    for (int i = sgt; i < npx; i += gstride) {
      const int lr = i / W.cols, c = i - lr * W.cols, px = W.ox + c;
      __stcg(W.touched + i, static_cast<unsigned char>(0));
      if (px >= L.cols) continue;
      const int64_t gi = static_cast<int64_t>(W.row0 + lr) * L.cols + px;
      T a[kLayerPlanes], b[kLayerPlanes];
#pragma unroll
      for (int k = 0; k < kLayerPlanes; ++k) {
        a[k] = __ldcg(static_cast<const T*>(L.canvas[W.band][k]) + gi);
        b[k] = __ldcg(static_cast<const T*>(L.snapshot[W.band][k]) + gi);
      }
      const unsigned char dflag = __ldcg(L.dirty[W.band] + static_cast<int64_t>(W.row0 + lr) * L.dirty_pitch + px);
#pragma unroll
      for (int k = 0; k < kLayerPlanes; ++k) {
        __stcg(W.can[k] + i, a[k]);
        __stcg(W.src[k] + i, b[k]);
      }
      __stcg(W.dirty + i, dflag);
    }
  }
}

// Push back only what this stroke changed: untouched pixels may meanwhile have been refreshed by a commuting stroke's
// snapshot ring on their owner (an idempotent copy this window must not undo).
template <typename T>
__device__ __forceinline__ void stage_out(const ImprintLaunch& L, const DevStroke& st, int group, int sgt, int gstride) {
  for (int w = 0; w < 2; ++w) {
    if (st.win_band[w] < 0) continue;
    const Window<T> W(L, st, group, w);
    const int nwords = W.rows * W.cols / 4;  // the touched map is scanned 4 pixels at a time
    const int cpr    = W.cols / 4;
    for (int i = sgt; i < nwords; i += gstride) {
      const unsigned t = __ldcg(reinterpret_cast<const unsigned*>(W.touched) + i);
      if (t == 0u) continue;
      const int lr = i / cpr, c = (i - lr * cpr) * 4;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        if (((t >> (8 * b)) & 0xffu) == 0u) continue;
        const int wi = lr * W.cols + c + b, px = W.ox + c + b;
        const int64_t gi = static_cast<int64_t>(W.row0 + lr) * L.cols + px;
        T va[kLayerPlanes], vb[kLayerPlanes];
#pragma unroll
        for (int k = 0; k < kLayerPlanes; ++k) {
          va[k] = __ldcg(W.can[k] + wi);
          vb[k] = __ldcg(W.src[k] + wi);
        }
#pragma unroll
        for (int k = 0; k < kLayerPlanes; ++k) {
          __stcg(static_cast<T*>(L.canvas[W.band][k]) + gi, va[k]);
          __stcg(static_cast<T*>(L.snapshot[W.band][k]) + gi, vb[k]);
        }
        __stcg(L.dirty[W.band] + static_cast<int64_t>(W.row0 + lr) * L.dirty_pitch + px, __ldcg(W.dirty + wi));
      }
    }
  }
}

constexpr int kRegCells = 2;  // cells per thread whose geometry is kept in registers across the stroke

template <typename T, bool CL, int MAXB, bool MULTI>
__global__ void __launch_bounds__(MAXB, 1) imprint_kernel(const ImprintLaunch L) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ long long s_stroke;
  __shared__ unsigned long long s_active;
  __shared__ Band<T> s_view[MULTI ? kMaxBands : 1];
  const Band<T>* views = MULTI ? s_view : nullptr;

  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, bd = blockDim.x;
  // A stroke is owned by `G` clusters of `csize0` CTAs. Within a cluster the hardware barrier synchronises; across
  // the clusters of a group a monotonic counter in L2 does (arrive = fence + atomicAdd by one thread per cluster,
  // wait = acquire-poll), bracketed by two cluster barriers — the classic grid-sync construction at group scope.
  const int csize0 = CL ? static_cast<int>(cluster.num_blocks()) : 1;
  const int G      = CL ? L.group : 1;
  const int cid    = static_cast<int>(blockIdx.x) / csize0;  // cluster index in the grid
  const int group  = cid / G, grank = cid % G;
  const int crank  = (CL ? static_cast<int>(cluster.block_rank()) : 0) + grank * csize0;  // CTA rank within the group
  const int csize  = csize0 * G;
  const int gstride = csize * bd;
  const int sgt     = crank * bd + tid;  // contiguous numbering for the coalesced dirty-map scan
  bool remote   = false;  // current stroke touches rows of another GPU: barriers need system-scope fences
  auto sync_all = [&]() {
    if (MULTI && remote) __threadfence_system();
    if (CL) {
      cluster.sync();
      if (G > 1) {
        if (cluster.block_rank() == 0 && tid == 0) {
          __threadfence();
          const unsigned ticket = atomicAdd(L.group_bar + group, 1u);
          const unsigned target = (ticket / static_cast<unsigned>(G) + 1u) * static_cast<unsigned>(G);
          while (static_cast<unsigned>(ld_acquire(reinterpret_cast<const int*>(L.group_bar + group))) < target) {
          }
        }
        cluster.sync();
      }
    } else {
      __syncthreads();
    }
  };

  OpCtx<T> C;
  C.pickup_rate     = static_cast<T>(L.pickup_rate);
  C.deposition_rate = static_cast<T>(L.deposition_rate);
  C.cap             = static_cast<T>(L.capacity);

  if (tid == 0) s_active = 0ull;
  unsigned long long my_active = 0;
  const int row_lo = L.store_first, row_hi = L.store_first + L.store_rows - 1;

  for (;;) {
    sync_all();
    if (crank == 0 && tid == 0) {
      const long long ticket = atomicAdd(L.queue, 1);
      s_stroke = (ticket < L.n_strokes && L.order != nullptr) ? static_cast<long long>(L.order[ticket]) : ticket;
      if (G > 1) __stcg(L.group_stroke + group, s_stroke);
    }
    sync_all();
    const int64_t si = G > 1 ? __ldcg(L.group_stroke + group) : (CL ? *cluster.map_shared_rank(&s_stroke, 0) : s_stroke);
    if (si >= L.n_strokes) break;
    const DevStroke st = L.strokes[si];

    // Dataflow wait. A stroke is cut into SEGMENTS of seg_len imprints; segment k may start once every earlier
    // stroke has finished the segments whose region meets segment k's (host-built lists of (stroke, segments
    // needed)); strokes publish their progress at segment boundaries. The wait of segment 0 comes first because
    // the staging windows below snapshot the neighbours' rows (windowed strokes have a single segment).
    auto seg_wait = [&](int k) {
      const int pb0 = L.seg_off[st.seg_begin + k], pb1 = L.seg_off[st.seg_begin + k + 1];
      if (pb1 == pb0) return;
      if (crank == 0) {
        for (int p = pb0 + tid; p < pb1; p += bd) {
          const int2 pr        = L.preds[p];
          const long long want = (static_cast<long long>(L.epoch) << 32) | static_cast<unsigned>(pr.y);
          if (MULTI) {  // the predecessor may run on another GPU: poll its progress word through NVLink
            const long long* flag = L.done[pr.x >> 27] + (pr.x & 0x7ffffff);
            while (ld_acquire64_sys(flag) < want) __nanosleep(256);
          } else {
            const long long* flag = L.done[0] + pr.x;
            while (ld_acquire64(flag) < want) __nanosleep(64);
          }
        }
      }
      sync_all();
    };
    auto seg_publish = [&](int done_segments) {  // call after a sync_all: every CTA's stores precede it
      if (crank == 0 && tid == 0) {
        const long long v = (static_cast<long long>(L.epoch) << 32) | static_cast<unsigned>(done_segments);
        if (MULTI) {
          __threadfence_system();
          st_release64_sys(L.done[L.my_band] + L.flag_offset + si, v);
        } else {
          __threadfence();
          st_release64(L.done[0] + L.flag_offset + si, v);
        }
      }
    };
    seg_wait(0);

#pragma unroll
    for (int k = 0; k < 3; ++k) {
      C.paintK[k] = static_cast<T>(st.paintK[k]);
      C.paintS[k] = static_cast<T>(st.paintS[k]);
    }
    remote       = MULTI && (st.flags & 4) != 0;
    if (MULTI && (st.flags & 12)) stage_in<T>(L, st, s_view, group, sgt, gstride, tid), sync_all();
    const int nA = st.n_active;
    const int wr = (st.side - 1) / 2;  // == hr (square footprint), FootprintBrush.hxx:75-78
    const T* fhs = static_cast<const T*>(st.fh);
    // The compacted cell list is cut into csize contiguous chunks (neighbouring cells -> neighbouring lanes ->
    // neighbouring canvas pixels); cell = cell0 + k*bd of this thread lives in shared-memory slot tid + k*bd.
    const int per_cta   = (nA + csize - 1) / csize;
    const int cell_end  = min(nA, (crank + 1) * per_cta);
    const int cell0     = crank * per_cta + tid;
    const int my_cells  = cell_end > cell0 ? (cell_end - cell0 + bd - 1) / bd : 0;
    const int cta_cells = ((per_cta + bd - 1) / bd) * bd;
    const bool in_smem  = cta_cells <= L.smem_cells;
    T* pick      = in_smem ? reinterpret_cast<T*>(smem_raw) : static_cast<T*>(L.scratch) + blockIdx.x * L.scratch_stride;
    const int ps = in_smem ? L.smem_cells : static_cast<int>(L.scratch_stride / kLayerPlanes);

    // per-thread cell geometry in registers; dip() = clean pickup map (:150-166) or continue with the brush's map
    int cmx[kRegCells], cmy[kRegCells];
    T cfh[kRegCells];
#pragma unroll
    for (int k = 0; k < kRegCells; ++k) {
      cmx[k] = cmy[k] = 0;
      cfh[k]          = static_cast<T>(0);
    }
    for (int k = 0; k < my_cells; ++k) {
      const int cell = cell0 + k * bd, slot = tid + k * bd;
      const uint32_t xy = st.xy[cell];
      if (k < kRegCells) {
#pragma unroll
        for (int q = 0; q < kRegCells; ++q)
          if (q == k) {
            cmx[q] = static_cast<int>(xy & 0xffffu);
            cmy[q] = static_cast<int>(xy >> 16);
            cfh[q] = fhs[cell];
          }
      }
      if (st.flags & 1) {
        const int64_t mi = static_cast<int64_t>(xy >> 16) * st.size_map + (xy & 0xffffu);
#pragma unroll
        for (int q = 0; q < kLayerPlanes; ++q) pick[q * ps + slot] = static_cast<const T*>(L.pick_dense[q])[mi];
      } else {
#pragma unroll
        for (int q = 0; q < kLayerPlanes; ++q) pick[q * ps + slot] = static_cast<T>(0);
      }
    }

    // The imprint chain, instantiated twice in the multi-GPU kernel: strokes that stay inside the executor's band
    // (the large majority) address it directly like the single-GPU kernel; only straddling strokes pay for the
    // per-band views.
    auto imprint_chain = [&](auto views_tag) {
      constexpr bool VIEWS = decltype(views_tag)::value;
      DevImprint nxt = st.n_imprints > 0 ? L.imprints[st.first_imprint] : DevImprint{0, 0, 1, 0};
      int seg_k = 0, seg_next = st.seg_len;  // next segment boundary (imprint index)
      for (int ii = 0; ii < st.n_imprints; ++ii) {
        if (ii == seg_next) {
          ++seg_k;
          seg_next += st.seg_len;
          seg_publish(seg_k);
          seg_wait(seg_k);
        }
        const DevImprint im = nxt;
        if (ii + 1 < st.n_imprints) nxt = L.imprints[st.first_imprint + ii + 1];  // prefetch (hidden behind this imprint)

        if (L.use_snapshot) {
          // updateSnapshot(canvas, centre) (:278-319): refresh the ring allowed-box \ open interior
          RingGeom g;
          g.tlx = static_cast<int>(im.cx - wr), g.tly = static_cast<int>(im.cy - wr);
          g.brx = static_cast<int>(im.cx + wr), g.bry = static_cast<int>(im.cy + wr);
          g.ax0 = max(static_cast<int>(im.cx - wr - st.radius), 0);
          g.ay0 = max(max(static_cast<int>(im.cy - wr - st.radius), 0), VIEWS ? 0 : row_lo);
          g.ax1 = min(static_cast<int>(im.cx + wr + st.radius), L.cols - 1);
          g.ay1 = min(min(static_cast<int>(im.cy + wr + st.radius), L.rows - 1), VIEWS ? L.rows - 1 : row_hi);
          ring_scan<T, VIEWS>(L, views, g, sgt, gstride);
        }
        // left/top overhang: canvas pixels of column/row 0 can be hit twice (B#11) -> ordered phases
        const bool border = (im.cx - wr < 0.0) || (im.cy - wr < 0.0);
        const float fc = static_cast<float>(im.c), fs = static_cast<float>(im.s);
        sync_all();

        const int n_phase = border ? 4 : 1;
        for (int ph = 0; ph < n_phase; ++ph) {
          // CPP cells per pass: all loads of up to 2*CPP interactions are in flight before any is computed
          constexpr int CPP = MAXB <= 256 ? 2 : 1;
          for (int k = 0; k < my_cells; k += CPP) {
            int mx[CPP], my[CPP], slot[CPP];
            T fh[CPP];
            bool have[CPP];
  #pragma unroll
            for (int q = 0; q < CPP; ++q) {
              have[q] = k + q < my_cells;
              slot[q] = tid + (k + q) * bd;
              if (k + q < kRegCells) {
                mx[q] = cmx[(k + q) & 1], my[q] = cmy[(k + q) & 1], fh[q] = cfh[(k + q) & 1];
              } else if (have[q]) {
                const int cell    = cell0 + (k + q) * bd;
                const uint32_t xy = st.xy[cell];
                mx[q] = static_cast<int>(xy & 0xffffu), my[q] = static_cast<int>(xy >> 16), fh[q] = fhs[cell];
              } else {
                mx[q] = my[q] = 0, fh[q] = static_cast<T>(0);
              }
            }
            Hits h[CPP];
            OpData<T> d[CPP][2];
  #pragma unroll
            for (int q = 0; q < CPP; ++q) {
              h[q].n = 0;
              if (have[q]) h[q] = find_hits<T, VIEWS>(L, views, im, fc, fs, wr, mx[q], my[q], border, ph, row_lo, row_hi);
  #pragma unroll
              for (int j = 0; j < 2; ++j)
                if (j < h[q].n) op_load(band_view<T, VIEWS>(L, views, h[q].band[j]), h[q].ci[j], d[q][j]);
            }
  #pragma unroll
            for (int q = 0; q < CPP; ++q) {
  #pragma unroll
              for (int j = 0; j < 2; ++j) {
                if (j < h[q].n) {
                  const Band<T> B = band_view<T, VIEWS>(L, views, h[q].band[j]);
                  op_finish(C, B, h[q].ci[j], fh[q], d[q][j], pick, ps, slot[q]);
                  if (B.dirty) __stcg(B.dirty + h[q].dof[j], static_cast<unsigned char>(1));
                  if (VIEWS && B.touched) __stcg(B.touched + h[q].dof[j], static_cast<unsigned char>(1));
                  ++my_active;
                }
              }
            }
          }
          sync_all();
        }
      }

    };
    if constexpr (MULTI) {
      if (st.flags & 12) {
        imprint_chain(std::true_type{});
      } else {
        imprint_chain(std::false_type{});
      }
    } else {
      imprint_chain(std::false_type{});
    }

    if (st.flags & 2) {
      for (int k = 0; k < my_cells; ++k) {
        const int cell = cell0 + k * bd, slot = tid + k * bd;
        const uint32_t xy = st.xy[cell];
        const int64_t mi  = static_cast<int64_t>(xy >> 16) * st.size_map + (xy & 0xffffu);
#pragma unroll
        for (int q = 0; q < kLayerPlanes; ++q) static_cast<T*>(L.pick_dense[q])[mi] = pick[q * ps + slot];
      }
    }
    sync_all();
    if (MULTI && (st.flags & 8)) {  // push the touched window pixels back into their owners' HBM
      stage_out<T>(L, st, group, sgt, gstride);
      __threadfence_system();
      sync_all();
    }
    seg_publish(kStrokeDone);
  }

  if (my_active) atomicAdd(&s_active, my_active);
  __syncthreads();
  if (tid == 0 && s_active) atomicAdd(L.counters, s_active);
  if (CL) cluster.sync();  // rank 0's shared memory must outlive the last remote read of s_stroke
}

__global__ void __launch_bounds__(256) count_visited_kernel(const DevStroke* strokes, int64_t n_strokes,
                                                            const DevImprint* imprints, int rows, int cols,
                                                            unsigned long long* counter) {
  __shared__ unsigned long long s_sum;
  if (threadIdx.x == 0) s_sum = 0;
  __syncthreads();
  unsigned long long mine = 0;
  for (int64_t si = blockIdx.y; si < n_strokes; si += gridDim.y) {
    const DevStroke st = strokes[si];
    const int wr = (st.side - 1) / 2, w = 2 * wr + 1;
    for (int ii = blockIdx.x; ii < st.n_imprints; ii += gridDim.x) {
      const DevImprint im = imprints[st.first_imprint + ii];
      for (int i = threadIdx.x; i < w * w; i += blockDim.x) {
        const int row = i / w - wr, col = i % w - wr;
        const int px = static_cast<int>(col + im.cx), py = static_cast<int>(row + im.cy);
        const double rc = col * im.c - row * im.s;
        const double rr = col * im.s + row * im.c;
        const int mx = static_cast<int>(round(rc + wr)), my = static_cast<int>(round(rr + wr));
        if (py < 0 || px < 0 || px >= cols || py >= rows) continue;
        if (my < 0 || mx < 0 || mx >= st.size_map || my >= st.size_map) continue;
        ++mine;
      }
    }
  }
  if (mine) atomicAdd(&s_sum, mine);
  __syncthreads();
  if (threadIdx.x == 0 && s_sum) atomicAdd(counter, s_sum);
}

// variants by maximum block size: smaller CTAs get a larger register budget (no spills on the critical path).
// A 1024-thread variant (64 registers, ~1.5 KB of spill traffic per thread) was measured and is slower: 24.3 vs
// 21.1 us per imprint at r = 129, 8.71 vs 8.14 s on the 10k-stroke workload. So is a 384-thread variant with two
// cells in flight per thread (168 registers): 23.1 us at r = 129, 18.5 vs 16.6 us at r = 112 — for the large
// footprints more resident warps beat more independent work per thread.
template <typename T, bool CL, bool MULTI>
const void* kernel_ptr_b(int block) {
  if (block <= 256) return reinterpret_cast<const void*>(imprint_kernel<T, CL, 256, MULTI>);
  return reinterpret_cast<const void*>(imprint_kernel<T, CL, 512, MULTI>);
}
template <typename T>
const void* kernel_ptr_t(bool cl, int block, bool multi) {
  if (multi) return cl ? kernel_ptr_b<T, true, true>(block) : kernel_ptr_b<T, false, true>(block);
  return cl ? kernel_ptr_b<T, true, false>(block) : kernel_ptr_b<T, false, false>(block);
}
const void* kernel_ptr(int precision, bool cl, int block, bool multi) {
  return precision == PB_F64 ? kernel_ptr_t<double>(cl, block, multi) : kernel_ptr_t<float>(cl, block, multi);
}

}  // namespace

// A stroke is latency bound (a chain of dependent imprints), so it is spread thin: CTAs of 128..1024 threads on
// up to 16 SMs (non-portable cluster size), about one active cell per thread.
// Launch classes (a run of consecutive strokes of one class shares a launch):
//   1  : tiny footprints (<= 256 active cells), one CTA per stroke
//   16 : cluster of 16 CTAs x 128 threads (<= 4096 cells, <= 2 per thread)
//   17 : cluster of 16 CTAs x 256 or 512 threads for the large footprints
int imprint_cluster_class(int n_active) { return n_active <= 256 ? 1 : (n_active <= 4096 ? 16 : 17); }

void imprint_plan(pb_context* ctx, int max_active, ImprintLaunch& L, size_t& smem_bytes) {
  const int cls     = imprint_cluster_class(max_active);
  const int cluster = cls == 1 ? 1 : 16;
  // Large footprints are bound by the per-SM L2 sector rate of their scattered SoA accesses. PB_IMPRINT_GROUP=G lets
  // G clusters (G x 16 SMs) cooperate on one stroke (1.8x faster per stroke at G = 4), but on the sbr workload the
  // cluster slots are worth more as concurrent strokes (measured: 2.31 s vs 2.50 s per 2000 strokes), so G = 1.
  const char* genv = std::getenv("PB_IMPRINT_GROUP");
  int group        = (cls == 17 && genv) ? std::min(8, std::max(1, std::atoi(genv))) : 1;
  // <= 256 threads per CTA: that kernel variant keeps two cells' interactions (56 loads) in flight without
  // spills; without groups the largest footprints (> 2 cells per thread at 256) trade that for twice the warps.
  int block = 128;
  if (cluster == 1) {
    block = max_active <= 128 ? 128 : 256;
  } else if (group > 1) {
    block = max_active <= group * 16 * 128 * 2 ? 128 : 256;
  } else {
    block = max_active <= 4096 ? 128 : (max_active <= 8192 ? 256 : 512);
  }
  const size_t es     = ctx->esize();
  const int per_cta   = (std::max(max_active, 1) + cluster * group - 1) / (cluster * group);
  const int cta_cells = (per_cta + block - 1) / block * block;
  const size_t budget = 200 * 1024;  // dynamic shared memory for this CTA's slice of the pickup map
  size_t need         = static_cast<size_t>(cta_cells) * kLayerPlanes * es;
  if (need <= budget) {
    L.smem_cells     = cta_cells;
    L.scratch_stride = 0;
  } else {  // only reachable for footprints beyond ~59k active cells (radius > 200): global scratch
    L.smem_cells     = 0;
    need             = 0;
    L.scratch_stride = static_cast<int64_t>(cta_cells) * kLayerPlanes;
  }
  smem_bytes = need;
  L.block    = block;
  L.cluster  = cluster;
  L.group    = group;
  // occupancy queries and attribute changes cost milliseconds: do them once per launch shape
  static std::mutex cache_mutex;  // contexts of different devices may plan from different host threads
  std::lock_guard<std::mutex> lock(cache_mutex);
  static std::map<std::tuple<int, int, int, int, size_t>, int> cache;
  const bool multi = L.n_bands > 1;
  const auto key   = std::make_tuple(ctx->device, ctx->precision * 2 + (multi ? 1 : 0), cluster * 100 + group, block, smem_bytes);
  auto it        = cache.find(key);
  if (it != cache.end()) {
    L.grid = it->second;
    return;
  }
  const void* fn = kernel_ptr(ctx->precision, cluster > 1, block, multi);
  PB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  if (cluster > 8) PB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  if (cluster == 1) {
    int per_sm = 0;
    PB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, block, smem_bytes));
    PB_REQUIRE(per_sm >= 1, "imprint kernel does not fit on an SM");
    L.grid = ctx->sm_count * per_sm;
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim            = dim3(static_cast<unsigned>(cluster * ctx->sm_count));
    cfg.blockDim           = dim3(static_cast<unsigned>(block));
    cfg.dynamicSmemBytes   = smem_bytes;
    cudaLaunchAttribute attr[1];
    attr[0].id               = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = static_cast<unsigned>(cluster);
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs                = attr;
    cfg.numAttrs             = 1;
    int n_clusters           = 0;
    PB_CUDA(cudaOccupancyMaxActiveClusters(&n_clusters, fn, &cfg));
    PB_REQUIRE(n_clusters >= group, "imprint kernel: the cooperating clusters of one stroke do not fit on the device");
    L.grid = (n_clusters / group) * group * cluster;
  }
  cache[key] = L.grid;
}

void imprint_launch(pb_context* ctx, const ImprintLaunch& L, size_t smem_bytes) {
  if (L.n_strokes <= 0) return;
  const void* fn = kernel_ptr(ctx->precision, L.cluster > 1, L.block, L.n_bands > 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim            = dim3(static_cast<unsigned>(L.grid));
  cfg.blockDim           = dim3(static_cast<unsigned>(L.block));
  cfg.dynamicSmemBytes   = smem_bytes;
  cfg.stream             = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id               = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(L.cluster);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs                = attr;
  cfg.numAttrs             = L.cluster > 1 ? 1 : 0;
  void* args[]             = {const_cast<ImprintLaunch*>(&L)};
  PB_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
  ctx->launches++;
}

void imprint_count_visited(pb_context* ctx, const DevStroke* strokes, int64_t n_strokes, const DevImprint* imprints,
                           int rows, int cols, unsigned long long* counter) {
  if (n_strokes <= 0) return;
  dim3 grid(64, static_cast<unsigned>(std::min<int64_t>(n_strokes, 2048)));
  count_visited_kernel<<<grid, 256, 0, ctx->stream>>>(strokes, n_strokes, imprints, rows, cols, counter);
  PB_CUDA(cudaGetLastError());
  ctx->launches++;
}

}  // namespace pb
