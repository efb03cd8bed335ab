// Footprint-brush imprint engine for sm_100a: batched, order-preserving pickup/deposit.
//
// Restates painty/renderer/FootprintBrush.hxx:73-143 (imprint), :278-319 (updateSnapshot),
// :331-340 (blend), :349-384 (pickupPaint), :393-431 (depositPaint) on SoA planes in HBM.
//
// Parallel decomposition (nothing like the reference's serial double loop):
//   * a STROKE (dip -> setRadius -> chain of imprints) is owned by one persistent thread-block CLUSTER
//     (1..16 CTAs); strokes are popped from a queue in a host-planned order. Dependencies are tracked per SEGMENT
//     of a stroke (64 imprints): a segment waits until the earlier strokes whose footprint+snapshot region overlaps
//     its own have published enough progress (host-built lists of (stroke, segments needed), schedule.hpp).
//     Consecutive imprints of a stroke are a true dependency chain: ONE cluster-wide barrier per imprint;
//   * inside an imprint a thread owns ACTIVE pickup-map cells (footprint height > 0, ~14.5 % of the padded square,
//     compacted once per radius). Cell state (7 pickup values, height, coordinates) lives in the CTA's shared memory
//     for the whole stroke. Per imprint a thread first turns its cells into a list of INTERACTIONS (cell, canvas
//     pixel): the inverse rotation gives the 2x2 pixel neighbourhood of the cell's pre-image; a single-precision test
//     decides which of them round onto the cell and falls back to the reference's f64 expression whenever a candidate
//     is within the float error bound of a rounding boundary (imprint_geom.hpp) — bit-identical decisions, no FP64 on
//     the common path. A cell has 0, 1 or 2 interactions (1 on average), listed in the row-major order of the
//     reference's loop; the list of the NEXT imprint is built before the barrier (it depends on no pixel data);
//   * the interactions are then processed two at a time: all 28 loads in flight, pickup + deposit, stores. Canvas
//     pixels are unique per imprint except at the left/top border, where C++ truncation folds column/row (-1,0) onto
//     0: those imprints run in <= 4 barrier-separated phases which reproduce the row-major order;
//   * updateSnapshot copies a canvas ring of ~2x the footprint area on EVERY imprint in the reference. Here a
//     byte-per-pixel dirty map records where snapshot and canvas can differ; the first imprint after a wait scans
//     the whole ring, every later one only the pixels that ENTERED the ring since the previous imprint (a few
//     one-pixel strips, handled by the last two warps of each CTA) — bit-identical result. For compact footprints
//     (every touched pixel lies in the open interior of the footprint box) ring pixels and touched pixels of one imprint
//     are disjoint, so the ring pass needs no barrier of its own;
//   * per-imprint constants (centre, cos/sin(-theta), float copies, floor of the centre) come from the host in f64
//     with the same libm as the reference;
//   * canvas, snapshot and dirty planes are accessed with L2-only loads/stores (ld/st.global.cg): they are
//     shared between SMs, and the 126 MB L2 keeps the working set of the running strokes resident.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>
#include <type_traits>

#include "imprint.cuh"

namespace cg = cooperative_groups;

namespace pb {
namespace {

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

// A pixel record: 8 elements (Kr Kg Kb Sr Sg Sb V 0). FP32: 32 bytes = one 256-bit L2-only access (sm_100a
// LDG/STG.E.256) and exactly one sector; FP64: two of them.
template <typename T>
struct Rec {
  T v[kRecord];
};
// L1C = false: L2-only accesses (ld/st.global.cg) — the data is shared between the SMs of a cluster. L1C = true: default
// caching (L1): used by strokes that run on ONE CTA (no cluster) and stay in the executor's own band. Everything such a
// stroke reads was written by its own SM, or by another stroke whose completion it acquired (ld.acquire on the progress
// word + bar.sync: the acquire drops stale L1 lines), so L1 hits are coherent — and consecutive imprints overlap by ~98 %,
// which turns almost every record access into an L1 hit.
template <bool L1C>
__device__ __forceinline__ Rec<float> ld_rec(const float* p) {
  Rec<float> r;
  if (L1C) {
    asm volatile("ld.global.ca.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(p)
                 : "memory");
  } else {
    asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(p)
                 : "memory");
  }
  return r;
}
template <bool L1C>
__device__ __forceinline__ void st_rec(float* p, const Rec<float>& r) {
  if (L1C) {
    asm volatile("st.global.wb.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]),
                 "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7])
                 : "memory");
  } else {
    asm volatile("st.global.cg.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]),
                 "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7])
                 : "memory");
  }
}
template <bool L1C>
__device__ __forceinline__ Rec<double> ld_rec(const double* p) {
  Rec<double> r;
  if (L1C) {
    asm volatile("ld.global.ca.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3]) : "l"(p) : "memory");
    asm volatile("ld.global.ca.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.v[4]), "=d"(r.v[5]), "=d"(r.v[6]), "=d"(r.v[7]) : "l"(p + 4) : "memory");
  } else {
    asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3]) : "l"(p) : "memory");
    asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.v[4]), "=d"(r.v[5]), "=d"(r.v[6]), "=d"(r.v[7]) : "l"(p + 4) : "memory");
  }
  return r;
}
template <bool L1C>
__device__ __forceinline__ void st_rec(double* p, const Rec<double>& r) {
  if (L1C) {
    asm volatile("st.global.wb.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(r.v[0]), "d"(r.v[1]), "d"(r.v[2]), "d"(r.v[3]) : "memory");
    asm volatile("st.global.wb.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p + 4), "d"(r.v[4]), "d"(r.v[5]), "d"(r.v[6]), "d"(r.v[7]) : "memory");
  } else {
    asm volatile("st.global.cg.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(r.v[0]), "d"(r.v[1]), "d"(r.v[2]), "d"(r.v[3]) : "memory");
    asm volatile("st.global.cg.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p + 4), "d"(r.v[4]), "d"(r.v[5]), "d"(r.v[6]), "d"(r.v[7]) : "memory");
  }
}
// the non-templated forms = L2-only (staging windows, peer memory)
template <typename T>
__device__ __forceinline__ Rec<T> ld_rec(const T* p) {
  return ld_rec<false>(p);
}
template <typename T>
__device__ __forceinline__ void st_rec(T* p, const Rec<T>& r) {
  st_rec<false>(p, r);
}
// scalar accesses with the same choice of cache operator
template <bool L1C>
__device__ __forceinline__ unsigned ld_word(const unsigned* p) {
  unsigned v;
  if (L1C)
    asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  else
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
template <bool L1C>
__device__ __forceinline__ void st_byte(unsigned char* p, unsigned char v) {
  const unsigned w = v;
  if (L1C)
    asm volatile("st.global.wb.u8 [%0], %1;" ::"l"(p), "r"(w) : "memory");
  else
    asm volatile("st.global.cg.u8 [%0], %1;" ::"l"(p), "r"(w) : "memory");
}
template <bool L1C, typename T>
__device__ __forceinline__ void st_scalar(T* p, T v) {
  if (L1C) {
    *reinterpret_cast<volatile T*>(p) = v;
  } else {
    __stcg(p, v);
  }
}

// View of one row band: canvas records, snapshot records and dirty map, all indexed with (band-local row) * pitch +
// column. Single GPU: built from the launch parameters (everything folds into constant-bank loads). Multi GPU: one view
// per band in shared memory, rebuilt per dataflow segment — the executor's own band, peer mappings, or local staging
// windows (virtual base pointers so that the same index arithmetic works).
template <typename T>
struct Band {
  T* can;
  T* src;
  unsigned char* dirty;
  unsigned char* touched;  // staging windows only: pixels to push back to their owner
  int pitch;
  __device__ __forceinline__ Band() {}
  __device__ __forceinline__ Band(const ImprintLaunch& L, int band) {
    can     = static_cast<T*>(L.canvas[band]);
    src     = static_cast<T*>(L.snapshot[band]);
    dirty   = L.dirty[band];
    touched = nullptr;
    pitch   = L.cols;
  }
  // the executor's own band, from dedicated launch fields (compile-time offsets into the constant bank)
  __device__ __forceinline__ explicit Band(const ImprintLaunch& L) {
    can     = static_cast<T*>(L.own_canvas);
    src     = static_cast<T*>(L.own_snapshot);
    dirty   = L.own_dirty;
    touched = nullptr;
    pitch   = L.cols;
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
  Rec<T> can, src;
};

template <typename T, bool L1C>
__device__ __forceinline__ void op_load(const Band<T>& C, int ci, OpData<T>& d) {
  const bool own_src = C.src == C.can;  // snapshot buffer disabled: pickup source is the canvas itself
  d.can = ld_rec<L1C>(C.can + static_cast<int64_t>(ci) * kRecord);
  d.src = own_src ? d.can : ld_rec<L1C>(C.src + static_cast<int64_t>(ci) * kRecord);
}

template <typename T, bool L1C>
__device__ __forceinline__ void op_finish(const OpCtx<T>& P, const Band<T>& C, int ci, T fh, const OpData<T>& d, T* pick, int ps,
                                          int slot) {
  using Blend = typename BlendSel<T>::type;
  const bool own_src = C.src == C.can;
  T vCan = d.can.v[PV];
  T vP   = pick[PV * ps + slot];
  T pK[3], pS[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    pK[k] = pick[(PK + k) * ps + slot];
    pS[k] = pick[(PS + k) * ps + slot];
  }
  // pickup
  const T vSrc  = d.src.v[PV];
  const T leave = P.pickup_rate * vSrc * fh;
  if (leave > static_cast<T>(kMinVolume)) {
    const T remain = vSrc - leave;
    if (own_src) {
      vCan = remain;
    } else {
      st_scalar<L1C>(C.src + static_cast<int64_t>(ci) * kRecord + PV, remain);
    }
    const Blend bl(vP, leave);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      pK[k] = bl(pK[k], d.src.v[PK + k]);
      pS[k] = bl(pS[k], d.src.v[PS + k]);
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
  Rec<T> out;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    out.v[PK + k] = b_can(b_src(pK[k], P.paintK[k]), d.can.v[PK + k]);
    out.v[PS + k] = b_can(b_src(pS[k], P.paintS[k]), d.can.v[PS + k]);
  }
  out.v[PV] = vB + vCan;
  out.v[7]  = static_cast<T>(0);
  st_rec<L1C>(C.can + static_cast<int64_t>(ci) * kRecord, out);
}

// The rare path of the hit test (a candidate within the float error of a rounding boundary): the reference's f64
// expression. Kept out of line: it is 4x larger than the common path and runs for well under 1 % of the cells.
__device__ __noinline__ int3 hits_exact_cold(const DevImprint* imprints, int64_t index, int wr, int mx, int my, int rows, int cols,
                                             int ph) {
  const DevImprint full = imprints[index];
  PixelHits h;
  hits_exact(full, wr, mx, my, rows, cols, ph, h);
  return make_int3(h.n, (h.py[0] << 16) | h.px[0], (h.py[1] << 16) | h.px[1]);  // canvas sides are < 65536
}

// updateSnapshot(canvas, centre) (:278-319): copy canvas -> snapshot on the ring "allowed box minus open
// interior". Only pixels flagged in the dirty map can differ, so the pass scans the map (one 32-bit word = 4
// pixels) and copies just those. The rectangles to scan come from ring_rects (imprint_geom.hpp): the whole ring, or
// only what entered it since the previous imprint. Items (words) are numbered across the rectangles; a thread takes
// items t0, t0 + stride, ...
// Items (words) are numbered across the rectangles; a thread takes items t0, t0 + stride, ... in BATCHES of kRingBatch: the
// dirty words of a batch are loaded together, then the dirty ring pixels among them are copied four records at a time
// (kRingCopy loads in flight, then the stores) — the pass costs a thread a few L2 round trips per batch instead of one
// per word plus one per copied pixel.
constexpr int kRingBatch = 4;
constexpr int kRingCopy  = 2;  // records in flight while copying (register budget of the 512-thread CTAs)
struct RingCtx {
  int rows_per_band, my_band, store_first, pitch;
};
__device__ __forceinline__ int ring_band_of(const RingCtx& X, int row) {
  const int b0 = X.my_band * X.rows_per_band;
  return (row >= b0 && row < b0 + X.rows_per_band) ? X.my_band : row / X.rows_per_band;
}
template <typename T, bool VIEWS, bool L1C>
__device__ __forceinline__ void ring_scan(T* own_can, T* own_src, unsigned char* own_dirty, const Band<T>* views, const RingCtx X,
                                       const RingGeom g, const RingGeom prev, const bool has_prev, int t0, int stride) {
  // the rectangle list is small and indexed dynamically: it lives in local memory (L1)
  RingList rl;
  ring_list(g, has_prev ? &prev : nullptr, rl);
  const int total = rl.total;
#pragma unroll 1
  for (int base = t0; base < total; base += kRingBatch * stride) {
    int row[kRingBatch], w[kRingBatch];
    unsigned word[kRingBatch];
#pragma unroll
    for (int q = 0; q < kRingBatch; ++q) {
      const int t = base + q * stride;
      word[q] = 0u, row[q] = 0, w[q] = 0;
      if (t < total) {
        int x0, j;
        ring_list_item(rl, t, row[q], x0, j);
        int lrow = row[q] - X.store_first, pitch = X.pitch;
        const unsigned char* dbase = own_dirty;
        if (VIEWS) {  // the segment's window: indexed with global rows
          lrow  = row[q];
          pitch = views[0].pitch;
          dbase = views[0].dirty;
        }
        w[q]    = ((lrow * pitch + x0) >> 2) + j;
        word[q] = ld_word<L1C>(reinterpret_cast<const unsigned*>(dbase) + w[q]);
      }
    }
    // dirty ring pixels of the batch: bit 4 * q + b = byte b of word q
    unsigned todo = 0u;
#pragma unroll
    for (int q = 0; q < kRingBatch; ++q) {
      if (word[q] == 0u) continue;
      int lrow = row[q] - X.store_first, pitch = X.pitch;
      if (VIEWS) {
        lrow  = row[q];
        pitch = views[0].pitch;
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        if (((word[q] >> (8 * b)) & 0xffu) != 0u && in_ring(g, row[q], 4 * w[q] + b - lrow * pitch)) todo |= 1u << (4 * q + b);
      }
    }
#pragma unroll 1
    while (todo) {
      Rec<T> rec[kRingCopy];
      int f[kRingCopy], bnd[kRingCopy];
#pragma unroll
      for (int k = 0; k < kRingCopy; ++k) {
        f[k] = -1, bnd[k] = 0;
        if (todo) {
          const int bit = __ffs(static_cast<int>(todo)) - 1;
          todo &= todo - 1u;
          const int q = bit >> 2;
          int wq = w[0], rq = row[0];  // select w[q], row[q] without dynamic register indexing
#pragma unroll
          for (int z = 1; z < kRingBatch; ++z) {
            if (q == z) wq = w[z], rq = row[z];
          }
          f[k]         = 4 * wq + (bit & 3);
          const T* can = own_can;
          if (VIEWS) can = views[0].can;
          rec[k] = ld_rec<L1C>(can + static_cast<int64_t>(f[k]) * kRecord);
        }
      }
#pragma unroll
      for (int k = 0; k < kRingCopy; ++k) {
        if (f[k] >= 0) {
          T* src                 = own_src;
          unsigned char* dirty   = own_dirty;
          unsigned char* touched = nullptr;
          if (VIEWS) {
            src     = views[0].src;
            dirty   = views[0].dirty;
            touched = views[0].touched;
          }
          st_rec<L1C>(src + static_cast<int64_t>(f[k]) * kRecord, rec[k]);
          st_byte<L1C>(dirty + f[k], static_cast<unsigned char>(0));
          if (VIEWS && touched) __stcg(touched + f[k], static_cast<unsigned char>(1));
        }
      }
    }
  }
}

// ---- multi-GPU staging windows ------------------------------------------------------------------------------------
// A straddling stroke works on a LOCAL copy of its current dataflow segment's whole region (the union of the allowed
// boxes of the segment's imprints, rows of the executor's own band and of its neighbours alike): one window per stroke
// slot, laid out as canvas records | snapshot records (npx records each) | dirty bytes | touched bytes, npx = rows * cols
// (cols a multiple of 4). Inside the segment every pixel access goes to the window through ONE view with a uniform pitch
// — the same index arithmetic and the same code as a stroke inside the band, no per-pixel band selection. The window is
// filled when the segment starts (own rows from local HBM, neighbour rows over NVLink) and the pixels the segment
// touched are written back when it ends.
template <typename T>
struct Window {
  T* can;
  T* src;
  unsigned char *dirty, *touched;
  int x0, y0, rows, cols;  // the stroke's frame
  __device__ __forceinline__ Window(const ImprintLaunch& L, const DevStroke& st, int slot) {
    x0 = st.win_x0, y0 = st.win_y0, rows = st.win_rows, cols = st.win_cols;
    unsigned char* base = L.win_scratch + static_cast<int64_t>(slot) * L.win_stride;
    const int64_t npx   = static_cast<int64_t>(rows) * cols;
    can     = reinterpret_cast<T*>(base);
    src     = can + npx * kRecord;
    dirty   = base + 2 * kRecord * npx * static_cast<int64_t>(sizeof(T));
    touched = dirty + npx;
  }
};

// The view of a stroke's window: virtual base pointers, index = global row * window cols + column.
template <typename T>
__device__ __forceinline__ void stage_view(const ImprintLaunch& L, const DevStroke& st, Band<T>* view, int slot) {
  const Window<T> W(L, st, slot);
  const int64_t off = static_cast<int64_t>(W.y0) * W.cols + W.x0;
  Band<T> v;
  v.can     = W.can - off * kRecord;
  v.src     = W.src - off * kRecord;
  v.dirty   = W.dirty - off;
  v.touched = W.touched - off;
  v.pitch   = W.cols;
  *view     = v;
}

// Pull one rectangle (canvas coordinates, inside the frame) from the bands' HBM into the window; touched flags off.
template <typename T>
__device__ __forceinline__ void stage_pull(const ImprintLaunch& L, const Window<T>& W, const Rect& r, int sgt, int gstride) {
  const int rc = r.x1 - r.x0 + 1, rr = r.y1 - r.y0 + 1;
  if (rc <= 0 || rr <= 0) return;
  const int npx = rr * rc;
  // two pixels per thread and iteration: the six loads of a pair are in flight together
#pragma unroll 1
  for (int i0 = sgt; i0 < npx; i0 += 2 * gstride) {
    Rec<T> a[2], b[2];
    unsigned char dflag[2];
    int wi[2];
    bool on[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = i0 + u * gstride;
      wi[u]       = -1;
      dflag[u]    = 0;
      on[u]       = false;
      if (i >= npx) continue;
      const int lr = i / rc, c = i - lr * rc, px = r.x0 + c, gy = r.y0 + lr;
      wi[u] = (gy - W.y0) * W.cols + (px - W.x0);
      if (px >= L.cols || gy >= L.rows) continue;  // column padding outside the canvas: only the flags are initialised
      on[u]            = true;
      const int band   = gy / L.rows_per_band;
      const int64_t gi = static_cast<int64_t>(gy - band * L.rows_per_band) * L.cols + px;
      a[u]     = ld_rec(static_cast<const T*>(L.canvas[band]) + gi * kRecord);
      b[u]     = ld_rec(static_cast<const T*>(L.snapshot[band]) + gi * kRecord);
      dflag[u] = __ldcg(L.dirty[band] + gi);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (wi[u] < 0) continue;
      __stcg(W.touched + wi[u], static_cast<unsigned char>(0));
      __stcg(W.dirty + wi[u], dflag[u]);
      if (on[u]) {
        st_rec(W.can + static_cast<int64_t>(wi[u]) * kRecord, a[u]);
        st_rec(W.src + static_cast<int64_t>(wi[u]) * kRecord, b[u]);
      }
    }
  }
}

// Bring the window up to date for a segment: its whole region (`full`: first segment, or other strokes may have written
// into the region since the previous segment), or only what was not in the previous segment's region.
template <typename T>
__device__ __forceinline__ void stage_in(const ImprintLaunch& L, const DevStroke& st, const DevWindow& dw, const DevWindow* prev,
                                         bool full, int slot, int sgt, int gstride) {
  if (dw.cols <= 0 || dw.rows <= 0) return;
  const Window<T> W(L, st, slot);
  const Rect cur{dw.x0, dw.y0, dw.x0 + dw.cols - 1, dw.y0 + dw.rows - 1};
  if (full || prev == nullptr || prev->cols <= 0 || prev->rows <= 0) {
    stage_pull<T>(L, W, cur, sgt, gstride);
    return;
  }
  const Rect old{prev->x0, prev->y0, prev->x0 + prev->cols - 1, prev->y0 + prev->rows - 1};
  RingGeom clip;  // rect_difference clips to an "allowed" box: the current region itself
  clip.ax0 = cur.x0, clip.ay0 = cur.y0, clip.ax1 = cur.x1, clip.ay1 = cur.y1;
  clip.tlx = clip.tly = clip.brx = clip.bry = 0;
  rect_difference(cur, old, clip, [&](const Rect& r) { stage_pull<T>(L, W, r, sgt, gstride); });
}

// Write back what this segment changed, and only that: untouched pixels may meanwhile have been refreshed by a commuting
// stroke's snapshot ring on their owner (an idempotent copy this window must not undo). The touched flags are cleared, so
// a pixel is written back once per segment that touches it.
template <typename T>
__device__ __forceinline__ void stage_out(const ImprintLaunch& L, const DevStroke& st, const DevWindow& dw, int slot, int sgt,
                                          int gstride) {
  if (dw.cols <= 0 || dw.rows <= 0) return;
  const Window<T> W(L, st, slot);
  const int cpr    = dw.cols / 4;  // the touched map is scanned 4 pixels at a time (x0, cols and the frame are 4-aligned)
  const int nwords = dw.rows * cpr;
#pragma unroll 1
  for (int i = sgt; i < nwords; i += gstride) {
    const int lr = i / cpr, c = (i - lr * cpr) * 4, gy = dw.y0 + lr;
    const int wbase = (gy - W.y0) * W.cols + (dw.x0 + c - W.x0);
    const unsigned t = __ldcg(reinterpret_cast<const unsigned*>(W.touched + wbase));
    if (t == 0u) continue;
    __stcg(reinterpret_cast<unsigned*>(W.touched + wbase), 0u);
    const int band = gy / L.rows_per_band;
#pragma unroll 1
    for (int b = 0; b < 4; ++b) {
      if (((t >> (8 * b)) & 0xffu) == 0u) continue;
      const int wi = wbase + b, px = dw.x0 + c + b;
      const int64_t gi = static_cast<int64_t>(gy - band * L.rows_per_band) * L.cols + px;
      const Rec<T> va = ld_rec(W.can + static_cast<int64_t>(wi) * kRecord);
      const Rec<T> vb = ld_rec(W.src + static_cast<int64_t>(wi) * kRecord);
      st_rec(static_cast<T*>(L.canvas[band]) + gi * kRecord, va);
      st_rec(static_cast<T*>(L.snapshot[band]) + gi * kRecord, vb);
      __stcg(L.dirty[band] + gi, __ldcg(W.dirty + wi));
    }
  }
}

// 16-byte asynchronous copy global -> shared (no registers held while the data is in flight)
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// SCR: the per-cell state lives in a per-CTA global scratch area instead of shared memory (footprints beyond ~60 000
// active cells in FP64 mode); everything else is identical.
// MODE: 0 = single GPU; 1 = multi GPU, strokes inside the executor's band only; 2 = multi GPU incl. straddling strokes
template <typename T, bool CL, int MAXB, int MODE, bool SCR>
__global__ void __launch_bounds__(MAXB, 1) imprint_kernel(const ImprintLaunch L) {
  constexpr bool MULTI = MODE != 0;
  // interactions in flight per thread: 16 registers of records each in FP32 mode, 32 in FP64 mode; as many as fit
  // without spilling (512-thread CTAs leave 128 registers per thread)
  constexpr int kInFlight = sizeof(T) == 4 ? (MAXB > 256 ? 3 : 4) : (MAXB > 256 ? 1 : 2);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ long long s_stroke;
  __shared__ unsigned long long s_active;
  __shared__ Band<T> s_view[1];  // MODE 2: the view of the current segment's staging window
  // Register diet (one CTA of 512 threads leaves 128 registers per thread): everything that is constant over a stroke
  // or an imprint lives in shared memory and is re-read where it is used — the stroke record, the paint constants, and a
  // ring of four imprint records (previous, current, next, and the one being prefetched with cp.async).
  __shared__ DevStroke s_st;
  __shared__ OpCtx<T> s_ctx;
  __shared__ DevImprint s_im[4];
  const Band<T>* views = MODE == 2 ? s_view : nullptr;

  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, bd = blockDim.x;
  const int csize   = CL ? static_cast<int>(cluster.num_blocks()) : 1;
  const int crank   = CL ? static_cast<int>(cluster.block_rank()) : 0;
  auto sync_all = [&]() {
    if (CL) {
      cluster.sync();
    } else {
      __syncthreads();
    }
  };
  // the same barrier in two halves: arrive (release: this thread's stores) ... independent work ... wait (acquire)
  auto sync_arrive = [&]() {
    if (CL) asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  };
  auto sync_wait = [&]() {
    if (CL) {
      asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
    } else {
      __syncthreads();
    }
  };
  // numbering of the threads of a stroke for coalesced scans: all of them, or only the ring threads (the last
  // ring_threads threads of every CTA; they own the fewest cells)
  auto scan_id     = [&]() { return crank * bd + tid; };
  auto scan_stride = [&]() { return csize * bd; };
  auto slot_id     = [&]() { return static_cast<int>(blockIdx.x) / csize; };  // stroke slot (cluster index in the grid)

  // Per-cell state of this CTA: 7 pickup planes, heights, cell coordinates relative to the map centre; then the
  // interaction lists (always in shared memory).
  const int cc = L.cta_cells;
  unsigned char* cell_base;
  unsigned char* list_base;
  if constexpr (SCR) {
    cell_base = static_cast<unsigned char*>(L.scratch) + static_cast<int64_t>(blockIdx.x) * L.scratch_stride;
    list_base = smem_raw;
  } else {
    cell_base = smem_raw;
    list_base = smem_raw + static_cast<size_t>(cc) * (8 * sizeof(T) + sizeof(float2));
  }
  T* const pick     = reinterpret_cast<T*>(cell_base);
  T* const fhp      = pick + static_cast<int64_t>(kLayerPlanes) * cc;
  float2* const uvp = reinterpret_cast<float2*>(fhp + cc);
  const int ps      = cc;
  int* const lci          = reinterpret_cast<int*>(list_base);
  unsigned char* const lk = list_base + static_cast<size_t>(2 * L.chunk_cells) * bd * sizeof(int);

  if (tid == 0) s_active = 0ull;
  const int row_lo = L.store_first, row_hi = L.store_first + L.store_rows - 1;

  for (;;) {
    sync_all();
    if (crank == 0 && tid == 0) {
      const long long ticket = atomicAdd(L.queue, 1);
      s_stroke = (ticket < L.n_strokes && L.order != nullptr) ? static_cast<long long>(L.order[ticket]) : ticket;
    }
    sync_all();
    const int64_t si = CL ? *cluster.map_shared_rank(&s_stroke, 0) : s_stroke;
    if (si >= L.n_strokes) break;
    static_assert(sizeof(DevStroke) == 144, "DevStroke is copied in nine 16-byte pieces");
    if (tid < 9) cp_async16(reinterpret_cast<char*>(&s_st) + 16 * tid, reinterpret_cast<const char*>(L.strokes + si) + 16 * tid);
    if (tid < 9) cp_async_wait_all();
    __syncthreads();
    const DevStroke& st = s_st;

    // Dataflow wait. A stroke is cut into SEGMENTS of seg_len imprints; segment k may start once every earlier
    // stroke has finished the segments whose region meets segment k's (host-built lists of (stroke, segments
    // needed)); strokes publish their progress at segment boundaries. Returns whether there was anything to wait for.
    auto seg_wait = [&](int k) -> bool {
      const int pb0 = L.seg_off[st.seg_begin + k], pb1 = L.seg_off[st.seg_begin + k + 1];
      if (pb1 == pb0) return false;
      if (crank == 0) {
        for (int p = pb0 + tid; p < pb1; p += bd) {
          const int2 pr        = L.preds[p];
          const long long want = (static_cast<long long>(L.epoch) << 32) | static_cast<unsigned>(pr.y);
          // Watchdog: a wait that lasts longer than L.watchdog_ns (default 60 s, PB_IMPRINT_WATCHDOG_S; 0 = off) cannot be a
          // legitimate dependency: report it and stop the kernel instead of hanging the device.
          unsigned polls = 0;
          unsigned long long t0 = 0;
          auto stuck = [&](long long have) {
            if (L.watchdog_ns == 0 || (++polls & 1023u) != 0u) return false;
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            if (now - t0 < L.watchdog_ns) return false;
            printf("painty_b200 imprint watchdog: band %d stroke %lld (flag %d, segment %d, kernel mode %d) waits for band %d flag %d: "
                   "wants %llx, sees %llx\n",
                   L.my_band, static_cast<long long>(si), st.flag_index, k, MODE, MULTI ? (pr.x >> 27) : 0, MULTI ? (pr.x & 0x7ffffff) : pr.x,
                   static_cast<unsigned long long>(want), static_cast<unsigned long long>(have));
            __trap();
            return true;
          };
          if (MULTI) {  // the predecessor may run on another GPU: poll its progress word through NVLink
            const long long* flag = L.done[pr.x >> 27] + (pr.x & 0x7ffffff);
            long long have;
            while ((have = ld_acquire64_sys(flag)) < want) {
              if (stuck(have)) break;
              __nanosleep(256);
            }
          } else {
            const long long* flag = L.done[0] + pr.x;
            long long have;
            while ((have = ld_acquire64(flag)) < want) {
              if (stuck(have)) break;
              __nanosleep(64);
            }
          }
        }
      }
      sync_all();
      return true;
    };
    auto seg_publish = [&](int done_segments) {  // call after a sync_all: every CTA's stores precede it
      if (crank == 0 && tid == 0) {
        const long long v = (static_cast<long long>(L.epoch) << 32) | static_cast<unsigned>(done_segments);
        if (MULTI) {
          __threadfence_system();
          st_release64_sys(L.done[L.my_band] + st.flag_index, v);
        } else {
          __threadfence();
          st_release64(L.done[0] + st.flag_index, v);
        }
      }
    };
    seg_wait(0);

    // imprint records 0 and 1 -> ring slots 0 and 1; paint constants
    if (tid < 8 && (tid >> 2) < st.n_imprints)
      cp_async16(reinterpret_cast<char*>(&s_im[tid >> 2]) + 16 * (tid & 3),
                 reinterpret_cast<const char*>(L.imprints + st.first_imprint + (tid >> 2)) + 16 * (tid & 3));
    if (tid == 0) {
      s_ctx.pickup_rate     = static_cast<T>(L.pickup_rate);
      s_ctx.deposition_rate = static_cast<T>(L.deposition_rate);
      s_ctx.cap             = static_cast<T>(L.capacity);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        s_ctx.paintK[k] = static_cast<T>(st.paintK[k]);
        s_ctx.paintS[k] = static_cast<T>(st.paintS[k]);
      }
    }
    const bool two_phase = (st.flags & kStrokeTwoPhase) != 0;
    const int wr = (st.side - 1) / 2;  // == hr (square footprint), FootprintBrush.hxx:75-78
    // The compacted cell list is cut into csize contiguous chunks (neighbouring cells -> neighbouring lanes ->
    // neighbouring canvas pixels); cell = cell0 + k*bd of this thread lives in slot tid + k*bd.
    int my_cells;
    {
      const int nA       = st.n_active;
      const int per_cta  = (nA + csize - 1) / csize;
      const int cell_end = min(nA, (crank + 1) * per_cta);
      const int cell0    = crank * per_cta + tid;
      my_cells           = cell_end > cell0 ? (cell_end - cell0 + bd - 1) / bd : 0;
      const T* fhs       = static_cast<const T*>(st.fh);
      // dip() = clean pickup map (:150-166), or continue with the brush's map
      for (int k = 0; k < my_cells; ++k) {
        const int cell = cell0 + k * bd, slot = tid + k * bd;
        const uint32_t xy = st.xy[cell];
        const int mx = static_cast<int>(xy & 0xffffu), my = static_cast<int>(xy >> 16);
        uvp[slot] = make_float2(static_cast<float>(mx - wr), static_cast<float>(my - wr));
        fhp[slot] = fhs[cell];
        if (st.flags & kStrokeLoadPick) {
          const int64_t mi = static_cast<int64_t>(my) * st.size_map + mx;
#pragma unroll
          for (int q = 0; q < kLayerPlanes; ++q) pick[q * ps + slot] = static_cast<const T*>(L.pick_dense[q])[mi];
        } else {
#pragma unroll
          for (int q = 0; q < kLayerPlanes; ++q) pick[q * ps + slot] = static_cast<T>(0);
        }
      }
    }
    if (tid < 8) cp_async_wait_all();
    __syncthreads();
    int my_active = 0;

    // The imprint chain, instantiated twice in the multi-GPU kernel: strokes that stay inside the executor's band
    // (the large majority) address it directly like the single-GPU kernel; only straddling strokes pay for the
    // per-band views.
    auto imprint_chain = [&](auto views_tag) {
      constexpr bool VIEWS = decltype(views_tag)::value;
      constexpr bool L1C   = !CL && !VIEWS;  // one CTA, own band: L1-cached pixel accesses (see ld_rec)
      if (st.n_imprints <= 0) return;
      const DevWindow* wins = VIEWS ? L.windows + st.seg_begin : nullptr;  // one staging window per dataflow segment

      // Interactions (cell, pixel) of `chunk` of this thread's cells for imprint ii, border phase ph (-1 = all).
      auto build_list = [&](int ii, int chunk, int ph) -> int {
        const DevImprint& im = s_im[ii & 3];
        const float fc = im.fc, fs = im.fs, lo = 0.5f - st.eps, hi = 0.5f + st.eps;
        const int ix = im.ix, iy = im.iy, flags = im.flags;
        int n        = 0;
        const int k0 = chunk * L.chunk_cells, k1 = min(my_cells, k0 + L.chunk_cells);
        for (int k = k0; k < k1; ++k) {
          const float2 uv = uvp[tid + k * bd];
          PixelHits h;
          if (two_phase) {
            h.n = -1;
          } else {
            hits_fast(fc, fs, ix, iy, flags, uv.x, uv.y, lo, hi, L.rows, L.cols, ph, h);
          }
          if (h.n < 0) {  // a candidate within the float error of a rounding boundary: the reference's f64 expression
            const int3 e = hits_exact_cold(L.imprints, st.first_imprint + ii, wr, static_cast<int>(uv.x) + wr,
                                           static_cast<int>(uv.y) + wr, L.rows, L.cols, ph);
            h.n = e.x;
            h.px[0] = e.y & 0xffff, h.py[0] = e.y >> 16;
            h.px[1] = e.z & 0xffff, h.py[1] = e.z >> 16;
          }
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (j < h.n) {
              const int py = h.py[j], px = h.px[j];
              int band = 0, lrow = py - row_lo, pitch = L.cols;
              bool ok = true;
              if (VIEWS) {  // the segment's window is indexed with global rows
                lrow  = py;
                pitch = views[0].pitch;
              } else if (py < row_lo || py > row_hi) {
                ok = false;  // band canvas without peers: rows outside the stored window are not ours
              }
              if (ok) {
                lci[n * bd + tid] = lrow * pitch + px;
                lk[n * bd + tid]  = static_cast<unsigned char>((k - k0) | (band << 4));
                ++n;
              }
            }
          }
        }
        return n;
      };
      // pickup + deposit of the listed interactions, kInFlight at a time (all record loads of the group in flight
      // before any compute). Interactions of one cell are adjacent in the list and finish in list order.
      auto process = [&](int n, int chunk) {
        const int k0 = chunk * L.chunk_cells;
#pragma unroll 1
        for (int j = 0; j < n; j += kInFlight) {
          int ci[kInFlight];
          unsigned e[kInFlight];
          OpData<T> d[kInFlight];
#pragma unroll
          for (int q = 0; q < kInFlight; ++q) {
            ci[q] = 0, e[q] = 0u;
            if (j + q < n) {
              ci[q] = lci[(j + q) * bd + tid];
              e[q]  = lk[(j + q) * bd + tid];
              op_load<T, L1C>(band_view<T, VIEWS>(L, views, static_cast<int>(e[q] >> 4)), ci[q], d[q]);
            }
          }
#pragma unroll
          for (int q = 0; q < kInFlight; ++q) {
            if (j + q < n) {
              const Band<T> B = band_view<T, VIEWS>(L, views, static_cast<int>(e[q] >> 4));
              const int slot  = tid + (k0 + static_cast<int>(e[q] & 15u)) * bd;
              op_finish<T, L1C>(s_ctx, B, ci[q], fhp[slot], d[q], pick, ps, slot);
              if (B.dirty) st_byte<L1C>(B.dirty + ci[q], static_cast<unsigned char>(1));
              if (VIEWS && B.touched) __stcg(B.touched + ci[q], static_cast<unsigned char>(1));
              ++my_active;
            }
          }
        }
      };
      auto first_phase = [&](int ii) { return (s_im[ii & 3].flags & kImBorder) ? 0 : -1; };
      auto geom_of     = [&](int ii) {
        const DevImprint& im = s_im[ii & 3];
        return ring_geom(im.cx, im.cy, wr, st.radius, L.cols, VIEWS ? 0 : row_lo, VIEWS ? L.rows - 1 : row_hi);
      };

      if (VIEWS) {
        if (tid == 0) stage_view<T>(L, st, s_view, slot_id());
        stage_in<T>(L, st, wins[0], nullptr, true, slot_id(), scan_id(), scan_stride());
        sync_all();
      }
      // The chain is a flat loop over UNITS (imprint ii, border phase ph, cell chunk): process the unit's list, build
      // the list of the next unit (geometry only, so the one that starts the next imprint is built before the barrier),
      // and synchronise whenever the next unit starts a new phase or imprint. One build site, one process site.
      const int chunks = max((my_cells + L.chunk_cells - 1) / L.chunk_cells, 1);
      const bool tracing = L.trace != nullptr && si == 0 && crank == 0 && (tid == 0 || tid == bd - 1);
      auto stamp = [&](int i, int slot) {
        if (tracing && i < kTraceImprints) L.trace[(i * 2 + (tid != 0 ? 1 : 0)) * kTraceStamps + slot] = clock64();
      };
      int ii = 0, ph = first_phase(0), chunk = 0;
      int n_list = build_list(0, 0, ph);
      bool need_full = true, imprint_start = true;
      int seg_k = 0, seg_next = st.seg_len;  // next segment boundary (imprint index)
      for (;;) {
        if (imprint_start) {
          stamp(ii, 0);
          if (ii == seg_next) {
            ++seg_k;
            seg_next += st.seg_len;
            if (wins) {  // write the pixels the finished segment touched back into their owners' HBM
              stage_out<T>(L, st, wins[seg_k - 1], slot_id(), scan_id(), scan_stride());
              __threadfence_system();
              sync_all();
            }
            seg_publish(seg_k);
            const bool waited = seg_wait(seg_k);
            if (waited) need_full = true;
            if (wins) {  // after a wait other strokes may have written into the region: pull all of it again
              stage_in<T>(L, st, wins[seg_k], wins + (seg_k - 1), waited, slot_id(), scan_id(), scan_stride());
              sync_all();
              need_full = true;  // newly pulled pixels bring their own dirty flags
            }
          }
          // imprint record ii + 2 -> ring slot (ii + 2) & 3 (nobody reads that slot during this imprint); completed
          // before this imprint's barrier
          if (tid < 4 && ii + 2 < st.n_imprints)
            cp_async16(reinterpret_cast<char*>(&s_im[(ii + 2) & 3]) + 16 * tid,
                       reinterpret_cast<const char*>(L.imprints + st.first_imprint + ii + 2) + 16 * tid);
          if (L.use_snapshot) {
            const int rt_local = tid - (bd - L.ring_threads);
            const RingCtx X{L.rows_per_band, L.my_band, L.store_first, L.cols};
            T* const oc = static_cast<T*>(L.own_canvas);
            T* const os = static_cast<T*>(L.own_snapshot);
            if (two_phase || need_full) {
              ring_scan<T, VIEWS, L1C>(oc, os, L.own_dirty, views, X, geom_of(ii), RingGeom{}, false, scan_id(), scan_stride());
            } else if (rt_local >= 0) {
              ring_scan<T, VIEWS, L1C>(oc, os, L.own_dirty, views, X, geom_of(ii), geom_of(ii - 1), true,
                                  crank * L.ring_threads + rt_local, csize * L.ring_threads);
            }
            need_full = false;
            if (two_phase) sync_all();
          }
          imprint_start = false;
          stamp(ii, 1);
        }
        const int ii_now = ii;
        process(n_list, chunk);
        stamp(ii_now, 2);
        // next unit; left/top overhang: canvas pixels of column/row 0 can be hit twice (B#11) -> 4 ordered phases
        bool barrier = false, done = false;
        if (chunk + 1 < chunks) {
          ++chunk;
        } else {
          chunk   = 0;
          barrier = true;
          if (ph >= 0 && ph < 3) {
            ++ph;
          } else if (ii + 1 < st.n_imprints) {
            ++ii;
            ph            = first_phase(ii);
            imprint_start = true;
          } else {
            done = true;
          }
        }
        // the barrier's latency (store acknowledgements, arrival of the other CTAs) overlaps building the next list,
        // which depends on no pixel data. The arrive is a release (MEMBAR.ALL.GPU in the SASS: the warp waits for its
        // stores' acknowledgements); building the list BEFORE the arrive instead (hiding the acknowledgement rather than
        // the others' arrival) measured slower: r = 151 14.5 against 12.8 us per imprint, bench step 3.38 against 3.12 s.
        if (barrier) {
          if (tid < 4) cp_async_wait_all();
          sync_arrive();
        }
        if (!done) n_list = build_list(ii, chunk, ph);
        stamp(ii_now, 3);
        if (barrier) sync_wait();
        stamp(ii_now, 4);
        if (done) break;
      }
      if (wins) {
        stage_out<T>(L, st, wins[seg_k], slot_id(), scan_id(), scan_stride());
        __threadfence_system();
      }
    };
    if constexpr (MODE == 2) {
      imprint_chain(std::true_type{});  // this variant is only launched for straddling strokes (capi.cu: run_plan)
    } else {
      imprint_chain(std::false_type{});
    }

    if (st.flags & kStrokeStorePick) {
      for (int k = 0; k < my_cells; ++k) {
        const int slot   = tid + k * bd;
        const float2 uv  = uvp[slot];
        const int64_t mi = static_cast<int64_t>(static_cast<int>(uv.y) + wr) * st.size_map + (static_cast<int>(uv.x) + wr);
#pragma unroll
        for (int q = 0; q < kLayerPlanes; ++q) static_cast<T*>(L.pick_dense[q])[mi] = pick[q * ps + slot];
      }
    }
    {  // active stroke-pixels of this stroke: one shared-memory atomic per warp
      const int warp_sum = __reduce_add_sync(0xffffffffu, my_active);
      if ((tid & 31) == 0 && warp_sum) atomicAdd(&s_active, static_cast<unsigned long long>(warp_sum));
    }
    sync_all();
    seg_publish(kStrokeDone);
  }

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

// variants by maximum block size: smaller CTAs get a larger register budget. The scratch variant only exists for the
// shape very large footprints use (clusters of 512-thread CTAs).
template <typename T, bool CL, int MODE>
const void* kernel_ptr_b(int block, bool scr) {
  if constexpr (CL) {
    if (scr) return reinterpret_cast<const void*>(imprint_kernel<T, true, 512, MODE, true>);
  }
  if (block <= 256) return reinterpret_cast<const void*>(imprint_kernel<T, CL, 256, MODE, false>);
  return reinterpret_cast<const void*>(imprint_kernel<T, CL, 512, MODE, false>);
}
template <typename T>
const void* kernel_ptr_t(bool cl, int block, int mode, bool scr) {
  if (mode == 2) return cl ? kernel_ptr_b<T, true, 2>(block, scr) : kernel_ptr_b<T, false, 2>(block, scr);
  if (mode == 1) return cl ? kernel_ptr_b<T, true, 1>(block, scr) : kernel_ptr_b<T, false, 1>(block, scr);
  return cl ? kernel_ptr_b<T, true, 0>(block, scr) : kernel_ptr_b<T, false, 0>(block, scr);
}
// mode: 0 single GPU, 1 multi GPU without / 2 with the band-view chain
const void* kernel_ptr(int precision, bool cl, int block, int mode, bool scr) {
  return precision == PB_F64 ? kernel_ptr_t<double>(cl, block, mode, scr) : kernel_ptr_t<float>(cl, block, mode, scr);
}
int launch_mode(const ImprintLaunch& L) { return L.n_bands > 1 ? (L.views_kernel ? 2 : 1) : 0; }

int env_int(const char* name, int fallback) {
  const char* e = std::getenv(name);
  return e ? std::atoi(e) : fallback;
}

}  // namespace

// Launch classes (a run of consecutive strokes of one class shares a launch) and their shapes, per policy:
//   latency policy      1 : <= 256 active cells, one CTA (128 / 256 threads)
//                       16: <= 4096 cells, cluster 8 x 256 (as fast as 16 x 128 and twice the resident clusters)
//                       17: cluster 16 x 256 (<= 8192 cells) or 16 x 512
//   throughput policy   1 : <= 4096 cells, one CTA of up to 512 threads (r = 30: 8.9 us on ONE SM against 4.0 us on 8)
//                       8 : cluster 8 x 512 (r = 151: 21.8 us on 8 SMs against 12.4 us on 16 — and 16-CTA clusters only
//                           fit 7 times on the 148 SMs, 8-CTA clusters use all of them)
// Two classes per policy, so that a pass of similar strokes stays one launch.
int imprint_cluster_class(int n_active, int policy) {
  if (policy == kShapeThroughput) return n_active <= 4096 ? 1 : 8;
  return n_active <= 256 ? 1 : (n_active <= 4096 ? 16 : 17);
}

double imprint_cost_us(int n_active, int policy) {
  struct P {
    double n, us;
  };
  static const P lat[] = {{146, 3.2}, {1107, 4.0}, {2461, 5.3}, {3958, 6.8}, {5081, 5.4}, {8025, 6.5}, {15200, 6.7}, {20319, 8.2}, {27507, 12.6}};
  static const P thr[] = {{146, 3.2}, {1107, 8.9}, {2461, 15.7}, {3958, 31.7}, {5081, 6.6}, {8025, 8.0}, {15200, 15.2}, {20319, 16.6}, {27507, 21.8}};
  const P* t  = policy == kShapeThroughput ? thr : lat;
  const int m = 9;
  const double x = n_active;
  if (x <= t[0].n) return t[0].us;
  for (int i = 1; i < m; ++i)
    if (x <= t[i].n) {
      // the tables are not continuous where the shape changes (4096 / 8192 cells): no interpolation across a class border
      if (imprint_cluster_class(static_cast<int>(t[i - 1].n), policy) != imprint_cluster_class(static_cast<int>(t[i].n), policy) &&
          imprint_cluster_class(n_active, policy) != imprint_cluster_class(static_cast<int>(t[i].n), policy))
        return t[i - 1].us * x / t[i - 1].n;
      return t[i - 1].us + (t[i].us - t[i - 1].us) * (x - t[i - 1].n) / (t[i].n - t[i - 1].n);
    }
  return t[m - 1].us * x / t[m - 1].n;
}

int imprint_slots(const ImprintLaunch& L) { return std::max(1, L.grid / std::max(1, L.cluster)); }

void imprint_plan(pb_context* ctx, int max_active, ImprintLaunch& L, size_t& smem_bytes) {
  const int cls = imprint_cluster_class(max_active, L.policy);
  // experiment knobs (scratch/imprint_sweep.py): cluster size and block size per class of the latency policy
  int cluster = 1, block = 128;
  if (cls == 1) {
    block = max_active <= 128 ? 128 : (max_active <= 256 ? 256 : 512);
  } else if (cls == 8) {
    cluster = 8, block = 512;
  } else if (cls == 16) {
    cluster = std::min(16, std::max(1, env_int("PB_IMPRINT_CLUSTER16", 8)));
    block   = env_int("PB_IMPRINT_BLOCK16", 256);
  } else {
    cluster = std::min(16, std::max(1, env_int("PB_IMPRINT_CLUSTER17", 16)));
    block   = env_int("PB_IMPRINT_BLOCK17", max_active <= 8192 ? 256 : 512);
  }
  block               = std::min(512, std::max(32, block / 32 * 32));
  const size_t es     = ctx->esize();
  const int per_cta   = (std::max(max_active, 1) + cluster - 1) / cluster;
  const int cpt       = (per_cta + block - 1) / block;  // cells per thread
  const int cta_cells = cpt * block;
  const int chunk     = std::min(cpt, 8);
  const size_t list_bytes = (static_cast<size_t>(2 * chunk) * block * 5 + 15) / 16 * 16;
  const size_t cell_bytes = static_cast<size_t>(cta_cells) * (8 * es + sizeof(float2));
  const size_t budget     = 220 * 1024;  // dynamic shared memory
  if (cell_bytes + list_bytes <= budget) {
    L.cells_in_smem  = 1;
    L.scratch_stride = 0;
    smem_bytes       = cell_bytes + list_bytes;
  } else {  // only reachable for very large footprints (radius > 200 in FP64 mode): cell state in global scratch
    L.cells_in_smem  = 0;
    L.scratch_stride = static_cast<int64_t>((cell_bytes + 255) / 256 * 256);
    smem_bytes       = list_bytes;
  }
  L.cta_cells    = cta_cells;
  L.chunk_cells  = chunk;
  // the incremental ring pass runs on the last warp of every CTA: where a CTA has more threads than cells, those threads own
  // no cells and the pass stays off the critical path (r = 30: 3.7 us per imprint with 32 ring threads, 4.9 with 64)
  L.ring_threads = std::min(block, std::max(32, env_int("PB_RING_THREADS", 32) / 32 * 32));
  L.block        = block;
  L.cluster      = cluster;
  // occupancy queries and attribute changes cost milliseconds: do them once per launch shape
  static std::mutex cache_mutex;  // contexts of different devices may plan from different host threads
  std::lock_guard<std::mutex> lock(cache_mutex);
  static std::map<std::tuple<int, int, int, int, size_t>, int> cache;
  const int mode = launch_mode(L);
  const auto key = std::make_tuple(ctx->device, ctx->precision * 4 + mode, cluster, block, smem_bytes);
  auto it        = cache.find(key);
  if (it != cache.end()) {
    L.grid = it->second;
    return;
  }
  PB_REQUIRE(L.cells_in_smem || (cluster > 1 && block == 512), "footprint too large for this launch shape");
  const void* fn = kernel_ptr(ctx->precision, cluster > 1, block, mode, !L.cells_in_smem);
  PB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(budget)));
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
    PB_REQUIRE(n_clusters >= 1, "imprint kernel: no thread-block cluster of this shape fits on the device");
    L.grid = n_clusters * cluster;
  }
  cache[key] = L.grid;
}

void imprint_launch(pb_context* ctx, const ImprintLaunch& L, size_t smem_bytes, cudaStream_t stream) {
  if (L.n_strokes <= 0) return;
  const void* fn = kernel_ptr(ctx->precision, L.cluster > 1, L.block, launch_mode(L), !L.cells_in_smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim            = dim3(static_cast<unsigned>(L.grid));
  cfg.blockDim           = dim3(static_cast<unsigned>(L.block));
  cfg.dynamicSmemBytes   = smem_bytes;
  cfg.stream             = stream ? stream : ctx->stream;
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
