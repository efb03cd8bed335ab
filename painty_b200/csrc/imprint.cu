// Footprint-brush imprint engine for sm_100a: batched, order-preserving pickup/deposit.
//
// Restates painty/renderer/FootprintBrush.hxx:73-143 (imprint), :278-319 (updateSnapshot),
// :331-340 (blend), :349-384 (pickupPaint), :393-431 (depositPaint) on SoA planes in HBM.
//
// Parallel decomposition (nothing like the reference's serial double loop):
//   * a STROKE (dip -> setRadius -> chain of imprints) is owned by one persistent CTA; strokes are
//     popped from a queue in submission order and wait on completion flags of the earlier strokes
//     whose footprint+snapshot region overlaps theirs (host-built predecessor lists) — a dataflow
//     schedule that keeps the reference's stroke order wherever it is observable;
//   * inside an imprint a thread owns ACTIVE pickup-map cells (footprint height > 0, ~14.5 % of the
//     padded square, compacted once per radius). Its pickup-map state (7 values per cell) lives in shared
//     memory for the whole stroke (global scratch for footprints too large for 227 KB). For each
//     imprint the thread inverts the rotation to find the <= 2 canvas pixels whose rotated+rounded
//     position is its cell, checks each candidate with the reference's exact f64 forward expression, and
//     applies pickup+deposit to them in row-major order — which is exactly the order in which the
//     reference's (row, col) loop hits a shared pickup cell. No two threads ever touch the same cell;
//   * canvas pixels are hit at most once per imprint except at the left/top border, where C++
//     truncation folds column/row (-1,0) onto 0. Those imprints run in <= 4 barrier-separated
//     phases ordered by (row-negative?, col-negative?) which reproduces the row-major order;
//   * per-imprint constants (centre, cos/sin(-theta)) are computed on the host in f64 with the same libm
//     as the reference; all index maths on the device is IEEE f64 without FMA contraction, so every
//     round()/trunc() decision is bit-identical to the CPU's;
//   * canvas and snapshot planes are accessed with L2-only loads/stores (ld/st.global.cg): they are shared
//     between SMs, and the 126 MB L2 keeps the working set of the running strokes resident.
#include <algorithm>

#include "imprint.cuh"

namespace pb {
namespace {

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// FootprintBrush.hxx:331-340
template <typename T>
__device__ __forceinline__ T blend(T va, T a, T vb, T b) {
  const T vt = va + vb;
  return (vt > static_cast<T>(kMinVolume)) ? (va * a + vb * b) / vt : a;
}

struct Hit {
  int px, py, cls;
};

template <typename T>
struct OpCtx {
  T* can[kLayerPlanes];
  T* src[kLayerPlanes];
  T pickup_rate, deposition_rate, cap;
  T paintK[3], paintS[3];
};

// pickupPaint (:349-384) then depositPaint (:393-431) for one (canvas pixel, pickup cell) pair.
template <typename T>
__device__ __forceinline__ void pickup_deposit(const OpCtx<T>& C, int64_t ci, T fh, T* pick, int64_t ps, int cell) {
  // issue every independent load first
  T cK[3], cS[3];
  const T vSrc = __ldcg(C.src[PV] + ci);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    cK[k] = __ldcg(C.can[PK + k] + ci);
    cS[k] = __ldcg(C.can[PS + k] + ci);
  }
  T vCan = __ldcg(C.can[PV] + ci);
  T vP   = pick[PV * ps + cell];
  T pK[3], pS[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    pK[k] = pick[(PK + k) * ps + cell];
    pS[k] = pick[(PS + k) * ps + cell];
  }
  // pickup
  const T leave = C.pickup_rate * vSrc * fh;
  if (leave > static_cast<T>(kMinVolume)) {
    const T remain = vSrc - leave;
    __stcg(C.src[PV] + ci, remain);
    if (C.src[PV] == C.can[PV]) {  // snapshot buffer disabled: pickup source is the canvas itself
      vCan = remain;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        pK[k] = blend(vP, pK[k], leave, cK[k]);
        pS[k] = blend(vP, pS[k], leave, cS[k]);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        pK[k] = blend(vP, pK[k], leave, __ldcg(C.src[PK + k] + ci));
        pS[k] = blend(vP, pS[k], leave, __ldcg(C.src[PS + k] + ci));
      }
    }
    vP = vP + leave;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      pick[(PK + k) * ps + cell] = pK[k];
      pick[(PS + k) * ps + cell] = pS[k];
    }
  }
  // deposit
  const T vFree = fmax(static_cast<T>(0), C.cap - vP);
  T kSrc[3], sSrc[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    kSrc[k] = blend(vP, pK[k], vFree, C.paintK[k]);
    sSrc[k] = blend(vP, pS[k], vFree, C.paintS[k]);
  }
  const T vLeave       = C.deposition_rate * vP * fh;
  pick[PV * ps + cell] = vP - vLeave;
  const T vB           = C.cap * fh;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    __stcg(C.can[PK + k] + ci, blend(vB, kSrc[k], vCan, cK[k]));
    __stcg(C.can[PS + k] + ci, blend(vB, sSrc[k], vCan, cS[k]));
  }
  __stcg(C.can[PV] + ci, vB + vCan);
}

template <typename T>
__global__ void __launch_bounds__(1024, 1) imprint_kernel(const ImprintLaunch L) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ long long s_stroke;
  __shared__ unsigned long long s_active;

  OpCtx<T> C;
#pragma unroll
  for (int k = 0; k < kLayerPlanes; ++k) {
    C.can[k] = static_cast<T*>(L.canvas[k]);
    C.src[k] = static_cast<T*>(L.snapshot[k]);
  }
  C.pickup_rate     = static_cast<T>(L.pickup_rate);
  C.deposition_rate = static_cast<T>(L.deposition_rate);
  C.cap             = static_cast<T>(L.capacity);

  const int tid = threadIdx.x, bd = blockDim.x;
  if (tid == 0) s_active = 0ull;
  unsigned long long my_active = 0;

  for (;;) {
    __syncthreads();
    if (tid == 0) s_stroke = atomicAdd(L.queue, 1);
    __syncthreads();
    const int64_t si = s_stroke;
    if (si >= L.n_strokes) break;
    const DevStroke st = L.strokes[si];

    // dataflow wait: every earlier stroke whose region overlaps ours has completed
    for (int p = st.pred_begin + tid; p < st.pred_end; p += bd) {
      const int* flag = L.done + L.preds[p];
      while (ld_acquire(flag) == 0) __nanosleep(64);
    }
    __syncthreads();

#pragma unroll
    for (int k = 0; k < 3; ++k) {
      C.paintK[k] = static_cast<T>(st.paintK[k]);
      C.paintS[k] = static_cast<T>(st.paintS[k]);
    }
    const int nA       = st.n_active;
    const int wr       = (st.side - 1) / 2;  // == hr (square footprint), FootprintBrush.hxx:75-78
    const T* fhs       = static_cast<const T*>(st.fh);
    const bool in_smem = nA <= L.smem_cells;
    T* pick            = in_smem ? reinterpret_cast<T*>(smem_raw) : static_cast<T*>(L.scratch) + blockIdx.x * L.scratch_stride;
    const int64_t ps   = in_smem ? L.smem_cells : L.scratch_stride / kLayerPlanes;

    // dip() = clean pickup map (:150-166), or continue with the brush's persistent map
    for (int cell = tid; cell < nA; cell += bd) {
      if (st.flags & 1) {
        const uint32_t xy = st.xy[cell];
        const int64_t mi  = static_cast<int64_t>(xy >> 16) * st.size_map + (xy & 0xffffu);
#pragma unroll
        for (int k = 0; k < kLayerPlanes; ++k) pick[k * ps + cell] = static_cast<const T*>(L.pick_dense[k])[mi];
      } else {
#pragma unroll
        for (int k = 0; k < kLayerPlanes; ++k) pick[k * ps + cell] = static_cast<T>(0);
      }
    }

    for (int ii = 0; ii < st.n_imprints; ++ii) {
      const DevImprint im = L.imprints[st.first_imprint + ii];

      if (L.use_snapshot) {
        // updateSnapshot(canvas, centre) (:278-319): copy the ring allowed-box \ open interior
        const int tlx = static_cast<int>(im.cx - wr), tly = static_cast<int>(im.cy - wr);
        const int brx = static_cast<int>(im.cx + wr), bry = static_cast<int>(im.cy + wr);
        const int ax0 = max(static_cast<int>(im.cx - wr - st.radius), 0);
        const int ay0 = max(max(static_cast<int>(im.cy - wr - st.radius), 0), L.store_first);
        const int ax1 = min(static_cast<int>(im.cx + wr + st.radius), L.cols - 1);
        const int ay1 = min(min(static_cast<int>(im.cy + wr + st.radius), L.rows - 1), L.store_first + L.store_rows - 1);
        const int w   = ax1 - ax0 + 1;
        if (w > 0 && ay1 >= ay0) {
          const int total = w * (ay1 - ay0 + 1);
          for (int i = tid; i < total; i += bd) {
            const int row = ay0 + i / w, col = ax0 + i % w;
            if (row > tly && row < bry && col > tlx && col < brx) continue;
            const int64_t ci = static_cast<int64_t>(row - L.store_first) * L.cols + col;
#pragma unroll
            for (int k = 0; k < kLayerPlanes; ++k) __stcg(C.src[k] + ci, __ldcg(C.can[k] + ci));
          }
        }
      }
      // left/top overhang: canvas pixels of column/row 0 can be hit twice (B#11) -> ordered phases
      const bool border = (im.cx - wr < 0.0) || (im.cy - wr < 0.0);
      __syncthreads();

      const int n_phase = border ? 4 : 1;
      for (int ph = 0; ph < n_phase; ++ph) {
        for (int cell = tid; cell < nA; cell += bd) {
          const uint32_t xy = st.xy[cell];
          const int mx = static_cast<int>(xy & 0xffffu), my = static_cast<int>(xy >> 16);
          const double u = mx - wr, v = my - wr;
          // inverse rotation gives the centre of the cell's pre-image; its bounding box has half-width
          // (|c|+|s|)/2 <= 0.7072, so at most 2x2 lattice candidates exist
          const double colf = u * im.c + v * im.s;
          const double rowf = v * im.c - u * im.s;
          const int c_lo = max(static_cast<int>(ceil(colf - 0.7075)), -wr), c_hi = min(static_cast<int>(floor(colf + 0.7075)), wr);
          const int r_lo = max(static_cast<int>(ceil(rowf - 0.7075)), -wr), r_hi = min(static_cast<int>(floor(rowf + 0.7075)), wr);
          for (int row = r_lo; row <= r_hi; ++row) {
            for (int col = c_lo; col <= c_hi; ++col) {
              // the reference's forward map (:95-100), same expression order, no FMA
              const double rc = col * im.c - row * im.s;
              const double rr = col * im.s + row * im.c;
              if (static_cast<int>(round(rc + wr)) != mx || static_cast<int>(round(rr + wr)) != my) continue;
              const double fx = col + im.cx, fy = row + im.cy;
              const int px = static_cast<int>(fx), py = static_cast<int>(fy);  // trunc toward zero (:92-93)
              if (py < 0 || px < 0 || px >= L.cols || py >= L.rows) continue;
              if (border && ((fy >= 0.0 ? 2 : 0) + (fx >= 0.0 ? 1 : 0)) != ph) continue;
              if (py < L.store_first || py >= L.store_first + L.store_rows) continue;  // band canvas
              const int64_t ci = static_cast<int64_t>(py - L.store_first) * L.cols + px;
              pickup_deposit(C, ci, fhs[cell], pick, ps, cell);
              ++my_active;
            }
          }
        }
        __syncthreads();
      }
    }

    if (st.flags & 2) {
      for (int cell = tid; cell < nA; cell += bd) {
        const uint32_t xy = st.xy[cell];
        const int64_t mi  = static_cast<int64_t>(xy >> 16) * st.size_map + (xy & 0xffffu);
#pragma unroll
        for (int k = 0; k < kLayerPlanes; ++k) static_cast<T*>(L.pick_dense[k])[mi] = pick[k * ps + cell];
      }
    }
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      st_release(L.done + si, 1);
    }
  }

  if (my_active) atomicAdd(&s_active, my_active);
  __syncthreads();
  if (tid == 0 && s_active) atomicAdd(L.counters, s_active);
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

}  // namespace

void imprint_plan(pb_context* ctx, int max_active, int& block, int& grid, size_t& smem_bytes, int& smem_cells) {
  block = 128;
  while (block < 1024 && block < max_active) block *= 2;
  const size_t es = ctx->esize();
  // up to 200 KB of dynamic shared memory for the pickup-map state of one stroke
  const size_t budget = 200 * 1024;
  size_t need         = static_cast<size_t>(max_active) * kLayerPlanes * es;
  if (need <= budget) {
    smem_cells = std::max(max_active, 1);
    smem_bytes = static_cast<size_t>(smem_cells) * kLayerPlanes * es;
  } else {
    // the large footprints go to global scratch; keep shared memory for the ones that fit
    smem_cells = static_cast<int>(budget / (kLayerPlanes * es));
    smem_bytes = static_cast<size_t>(smem_cells) * kLayerPlanes * es;
  }
  const void* fn = ctx->precision == PB_F64 ? reinterpret_cast<const void*>(imprint_kernel<double>)
                                             : reinterpret_cast<const void*>(imprint_kernel<float>);
  PB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes)));
  int per_sm = 0;
  PB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, block, smem_bytes));
  PB_REQUIRE(per_sm >= 1, "imprint kernel does not fit on an SM");
  grid = ctx->sm_count * per_sm;
}

void imprint_launch(pb_context* ctx, const ImprintLaunch& L, size_t smem_bytes) {
  if (L.n_strokes <= 0) return;
  if (ctx->precision == PB_F64)
    imprint_kernel<double><<<L.grid, L.block, smem_bytes, ctx->stream>>>(L);
  else
    imprint_kernel<float><<<L.grid, L.block, smem_bytes, ctx->stream>>>(L);
  PB_CUDA(cudaGetLastError());
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
