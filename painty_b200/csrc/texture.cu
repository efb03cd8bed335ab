// Texture-brush stroke deposit for sm_100a (smudge off).
//
// Restates painty/renderer/TextureBrush.hxx:142-204 per pixel: mean-value-coordinate warp of the
// canvas position into stroke-texture space (painty/core/Math.hxx:73-147), BORDER_REFLECT bilinear
// sample of the thickness map (painty/image/Mat.hxx:69-103), volume-weighted K/S blend, V = max.
// The stroke frame (extended spine, polygon, uv, bounding box; TextureBrush.hxx:52-131) is built on
// the host in f64 (texture_host.hpp) — it is per-stroke work.
//
// Decomposition: the reference first collects the covered pixels, then deposits; without smudge the
// two loops fuse exactly because a pixel's result depends only on that pixel — so a pixel only needs the strokes
// that cover it applied in submission order. Work items are (stroke, 64x64 canvas tile) pairs in stroke-major
// order; every canvas tile has a ticket counter, an item runs when all earlier strokes covering its tile are done
// with it (tickets are assigned on the device by one thread per tile walking the stroke list). Persistent CTAs pop
// items from a queue: all SMs stay busy even when consecutive strokes overlap completely.
// The warp and the sample decide discrete outcomes (inside [0,1]^2, Vtex > 0), so they are always
// evaluated in IEEE f64 without FMA contraction, in the reference's operation order; only the final
// blend runs in the context's element type. The polygon (<= kMaxPoly vertices) is staged in shared memory.
#include <cooperative_groups.h>

#include "texture.cuh"

namespace cg = cooperative_groups;

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

// cv::borderInterpolate(BORDER_REFLECT)
__device__ __forceinline__ int reflect(int p, int len) {
  if (static_cast<unsigned>(p) < static_cast<unsigned>(len)) return p;
  if (len == 1) return 0;
  do {
    p = (p < 0) ? (-p - 1) : (len - 1 - (p - len));
  } while (static_cast<unsigned>(p) >= static_cast<unsigned>(len));
  return p;
}

// painty/image/Mat.hxx:69-103
__device__ __forceinline__ double bilinear(const double* __restrict__ m, int rows, int cols, double px, double py) {
  const int x = static_cast<int>(floor(px)), y = static_cast<int>(floor(py));
  const int x0 = reflect(x, cols), x1 = reflect(x + 1, cols);
  const int y0 = reflect(y, rows), y1 = reflect(y + 1, rows);
  const double a = px - static_cast<double>(x), c = py - static_cast<double>(y);
  const double v00 = __ldg(m + static_cast<int64_t>(y0) * cols + x0), v01 = __ldg(m + static_cast<int64_t>(y0) * cols + x1);
  const double v10 = __ldg(m + static_cast<int64_t>(y1) * cols + x0), v11 = __ldg(m + static_cast<int64_t>(y1) * cols + x1);
  return (v00 * (1.0 - a) + v01 * a) * (1.0 - c) + (v10 * (1.0 - a) + v11 * a) * c;
}

struct Edge {
  double r, A, D;
};

// painty/core/Math.hxx:73-147 for n >= 2, single pass: the early-outs are tested in index order before the
// weight of vertex i is accumulated, which returns exactly what the reference's two loops return.
__device__ __forceinline__ void mvc(const double2* __restrict__ poly, const double2* __restrict__ uv, int n, double x,
                                    double y, double& ou, double& ov) {
  const double Eps = 2.220446049250313e-16 * 100.0;
  // r_i = |p_i - x| is needed by edge i - 1 (as r_{i+1}) and by edge i: it is computed once and carried over (the same
  // expression on the same operands, so the value is the one the reference recomputes)
  auto dist = [&](int i) {
    const double2 p = poly[i];
    const double dx = p.x - x, dy = p.y - y;
    return sqrt(dx * dx + dy * dy);
  };
  auto edge = [&](int i, double r_i, double2& si, double2& si1) {
    const double2 p = poly[i], q = poly[i == n - 1 ? 0 : i + 1];
    si  = make_double2(p.x - x, p.y - y);
    si1 = make_double2(q.x - x, q.y - y);
    Edge e;
    e.r = r_i;
    e.A = (si.x * si1.y - si1.x * si.y) / 2.0;
    e.D = si.x * si1.x + si.y * si1.y;
    return e;
  };
  double2 s0, s1;
  const double r_first = dist(0);
  Edge prev = edge(n - 1, dist(n - 1), s0, s1);  // A_{i-1}, r_{i-1}, D_{i-1} for i = 0
  double fu = 0.0, fv = 0.0, W = 0.0;
  double r_cur = r_first;
  for (int i = 0; i < n; ++i) {
    const Edge cur = edge(i, r_cur, s0, s1);
    const double ri1 = (i == n - 1) ? r_first : sqrt(s1.x * s1.x + s1.y * s1.y);  // r_{i+1}
    if (fabs(cur.r - 0.0) < Eps) {  // :102-104
      ou = uv[i].x;
      ov = uv[i].y;
      return;
    }
    if (fabs(cur.A - 0.0) < Eps && cur.D < 0.0) {  // :114-119
      const double2 f1 = uv[i == n - 1 ? 0 : i + 1];
      const double sc  = 1.0 / (cur.r + ri1);
      ou = (ri1 * uv[i].x + cur.r * f1.x) * sc;
      ov = (ri1 * uv[i].y + cur.r * f1.y) * sc;
      return;
    }
    double w = 0.0;
    if (prev.A != 0.0) w = w + (prev.r - prev.D / cur.r) / prev.A;
    if (cur.A != 0.0) w = w + (ri1 - cur.D / cur.r) / cur.A;
    fu = fu + w * uv[i].x;
    fv = fv + w * uv[i].y;
    W  = W + w;
    prev  = cur;
    r_cur = ri1;
  }
  if (!(fabs(W - 0.0) < Eps)) {
    const double sc = 1.0 / W;
    ou = fu * sc;
    ov = fv * sc;
  } else {
    ou = uv[0].x;
    ov = uv[0].y;
  }
}

// One thread per canvas tile walks the strokes in submission order and hands out that tile's tickets.
__global__ void __launch_bounds__(256) texture_ticket_kernel(const TextureLaunch L) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= L.tiles_x * L.tiles_y) return;
  const int tx = t % L.tiles_x, ty = t / L.tiles_x;
  int next = 0;
  for (int64_t s = 0; s < L.n_strokes; ++s) {
    const DevTStroke& st = L.strokes[s];
    if (tx >= st.tx0 && tx <= st.tx1 && ty >= st.ty0 && ty <= st.ty1)
      L.ticket[st.item_begin + static_cast<int64_t>(ty - st.ty0) * (st.tx1 - st.tx0 + 1) + (tx - st.tx0)] = next++;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) texture_kernel(const TextureLaunch L) {
  __shared__ long long s_item, s_stroke;
  __shared__ unsigned long long s_pixels;
  __shared__ double2 s_poly[kMaxPoly];
  __shared__ double2 s_uv[kMaxPoly];
  const int tid = threadIdx.x, bd = blockDim.x;
  if (tid == 0) {
    s_pixels = 0ull;
    s_stroke = -1;
  }
  unsigned long long mine = 0;
  long long staged         = -1;  // stroke whose polygon currently sits in shared memory
  T* can[kLayerPlanes];
#pragma unroll
  for (int k = 0; k < kLayerPlanes; ++k) can[k] = static_cast<T*>(L.canvas[k]);

  for (;;) {
    __syncthreads();
    if (tid == 0) {
      const long long i = static_cast<long long>(atomicAdd(L.queue, 1ull));
      s_item            = i;
      if (i < L.n_items) {  // stroke of item i: last stroke with item_begin <= i
        long long lo = 0, hi = L.n_strokes - 1;
        while (lo < hi) {
          const long long mid = (lo + hi + 1) >> 1;
          if (L.strokes[mid].item_begin <= i)
            lo = mid;
          else
            hi = mid - 1;
        }
        s_stroke = lo;
      }
    }
    __syncthreads();
    const long long item = s_item;
    if (item >= L.n_items) break;
    const long long si  = s_stroke;
    const DevTStroke st = L.strokes[si];
    const int tw        = st.tx1 - st.tx0 + 1;
    const int local     = static_cast<int>(item - st.item_begin);
    const int tx = st.tx0 + local % tw, ty = st.ty0 + local / tw;
    const int tile_id = ty * L.tiles_x + tx;
    if (tid == 0) {
      const int want = L.ticket[item];
      while (ld_acquire(L.tile_done + tile_id) != want) __nanosleep(32);
    }
    if (staged != si) {
      for (int i = tid; i < st.n_poly; i += bd) {
        s_poly[i] = L.poly[st.poly_begin + i];
        s_uv[i]   = L.uv[st.poly_begin + i];
      }
      staged = si;
    }
    __syncthreads();

    const T pK[3] = {static_cast<T>(st.K[0]), static_cast<T>(st.K[1]), static_cast<T>(st.K[2])};
    const T pS[3] = {static_cast<T>(st.S[0]), static_cast<T>(st.S[1]), static_cast<T>(st.S[2])};
    // this tile's part of the reference's (int)boundMin .. (int)boundMax sweep (:142-145)
    const int px0 = max(st.x0, tx * L.tile), px1 = min(st.x1, tx * L.tile + L.tile - 1);
    const int py0 = max(st.y0, ty * L.tile), py1 = min(st.y1, ty * L.tile + L.tile - 1);
    const int w = px1 - px0 + 1, h = py1 - py0 + 1;
    const int total = (st.n_poly >= 2 && w > 0 && h > 0) ? w * h : 0;
    for (int i = tid; i < total; i += bd) {
      const int x = px0 + i % w, y = py0 + i / w;
      if (x < 0 || x >= L.cols || y < 0 || y >= L.rows) continue;
      double u, v;
      mvc(s_poly, s_uv, st.n_poly, static_cast<double>(x), static_cast<double>(y), u, v);
      if (u < 0.0 || u > 1.0 || v < 0.0 || v > 1.0) continue;
      u *= st.map_cols;
      v *= st.map_rows;
      const double Vtex = st.thickness_scale * bilinear(st.map, st.map_rows, st.map_cols, u, v);
      if (!(Vtex > 0.0)) continue;
      const int s = x - st.x0, t = y - st.y0;
      if (!(s >= 0 && t >= 0 && s < st.local_cols && t < st.local_rows)) continue;  // :163-164
      ++mine;
      if (y < L.store_first || y >= L.store_first + L.store_rows) continue;  // band canvas
      const int64_t ci = static_cast<int64_t>(y - L.store_first) * L.cols + x;
      const T vt   = static_cast<T>(Vtex);
      const T vcan = __ldcg(can[PV] + ci);
      const T vsum = vcan + vt;
      if (vsum > static_cast<T>(0)) {  // :189-203
        const T sc = static_cast<T>(1) / vsum;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          __stcg(can[PK + k] + ci, (vcan * __ldcg(can[PK + k] + ci) + vt * pK[k]) * sc);
          __stcg(can[PS + k] + ci, (vcan * __ldcg(can[PS + k] + ci) + vt * pS[k]) * sc);
        }
        __stcg(can[PV] + ci, fmax(vt, vcan));
      }
    }
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      st_release(L.tile_done + tile_id, L.ticket[item] + 1);
    }
  }
  if (mine) atomicAdd(&s_pixels, mine);
  __syncthreads();
  if (tid == 0 && s_pixels) atomicAdd(L.counters, s_pixels);
}

// ---- smudge path -------------------------------------------------------------------------------------------------
// Pass 1 (TextureBrush.hxx:142-173): Vtex for every bounding-box pixel into the local thickness map + its maximum.
__global__ void __launch_bounds__(256) thickness_kernel(const SmudgeLaunch L) {
  __shared__ double2 s_poly[kMaxPoly];
  __shared__ double2 s_uv[kMaxPoly];
  const DevTStroke& st = L.stroke;
  for (int i = threadIdx.x; i < st.n_poly; i += blockDim.x) {
    s_poly[i] = L.poly[st.poly_begin + i];
    s_uv[i]   = L.uv[st.poly_begin + i];
  }
  __syncthreads();
  const int h = st.y1 - st.y0 + 1;
  const int64_t total  = st.n_poly >= 2 ? static_cast<int64_t>(st.x1 - st.x0 + 1) * h : 0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  double vmax = 0.0;
  unsigned long long mine = 0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int x = st.x0 + static_cast<int>(i / h), y = st.y0 + static_cast<int>(i % h);
    if (x < 0 || x >= L.cols || y < 0 || y >= L.rows) continue;
    double u, v;
    mvc(s_poly, s_uv, st.n_poly, static_cast<double>(x), static_cast<double>(y), u, v);
    if (u < 0.0 || u > 1.0 || v < 0.0 || v > 1.0) continue;
    u *= st.map_cols;
    v *= st.map_rows;
    const double Vtex = st.thickness_scale * bilinear(st.map, st.map_rows, st.map_cols, u, v);
    if (!(Vtex > 0.0)) continue;
    const int s = x - st.x0, t = y - st.y0;
    if (!(s >= 0 && t >= 0 && s < st.local_cols && t < st.local_rows)) continue;
    L.tmap[static_cast<int64_t>(t) * st.local_cols + s] = Vtex;
    vmax = fmax(vmax, Vtex);
    ++mine;
  }
  // non-negative doubles order like their bit patterns
  unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(vmax));
  for (int o = 16; o > 0; o >>= 1) {
    bits = max(bits, __shfl_xor_sync(0xffffffffu, bits, o));
    mine += __shfl_xor_sync(0xffffffffu, mine, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (bits) atomicMax(L.max_bits, bits);
    if (mine) atomicAdd(L.counters, mine);
  }
}

// Pass 2 (Smudge.hxx:38-199): the pickup window is dragged along the spine, one dependent step per spine sample.
// Per step a thread first re-orients its window cells (rotate by dtheta + bilinear resample of the previous window,
// ping-pong buffers instead of the reference's copyTo) and then exchanges paint between each cell and the canvas pixel
// under it; both touch only the thread's own cells and pixels, so one cluster barrier per step suffices.
template <typename T>
__global__ void __launch_bounds__(256) smudge_kernel(const SmudgeLaunch L) {
  cg::cluster_group cluster = cg::this_cluster();
  const double maxD = __longlong_as_double(static_cast<long long>(*L.max_bits));
  if (!(maxD > 0.0)) return;  // Smudge.hxx:45-47
  const int n = L.size, cells = n * n;
  const int gt = static_cast<int>(cluster.block_rank()) * blockDim.x + threadIdx.x;
  const int gstride = static_cast<int>(cluster.num_blocks()) * blockDim.x;
  T* can[kLayerPlanes];
#pragma unroll
  for (int k = 0; k < kLayerPlanes; ++k) can[k] = static_cast<T*>(L.canvas[k]);
  const double cw = L.max_size / 2.0, radius = L.max_size * 0.5;
  const T dep = static_cast<T>(L.deposition_rate), pick_rate = static_cast<T>(L.pickup_rate);
  const T Dmax = static_cast<T>(maxD);

  for (int step = 0; step < L.n_steps; ++step) {
    const DevSmudgeStep sp = L.steps[step];
    const int di = (L.first_dst + step + 1) & 1;  // copyTo + overwrite == swap the roles of the two windows
    T* dst[kLayerPlanes];
    const T* src[kLayerPlanes];
#pragma unroll
    for (int k = 0; k < kLayerPlanes; ++k) {
      dst[k] = static_cast<T*>(L.pick[di][k]);
      src[k] = static_cast<const T*>(L.pick[di ^ 1][k]);
    }
    for (int i = gt; i < cells; i += gstride) {
      const int y = i / n, x = i - y * n;
      // updateOrientation :172-197
      const double px = x - cw, py = y - cw;
      const double qx = (px * sp.c - py * sp.s) + cw, qy = (px * sp.s + py * sp.c) + cw;
      T pv[kLayerPlanes];
      if (qx < 0 || qy < 0 || qx >= n || qy >= n) {
#pragma unroll
        for (int k = 0; k < kLayerPlanes; ++k) pv[k] = __ldcg(src[k] + i);
      } else {
        const int fx = static_cast<int>(floor(qx)), fy = static_cast<int>(floor(qy));
        const int x0 = reflect(fx, n), x1 = reflect(fx + 1, n), y0 = reflect(fy, n), y1 = reflect(fy + 1, n);
        const double a = qx - static_cast<double>(fx), c = qy - static_cast<double>(fy);
#pragma unroll
        for (int k = 0; k < kLayerPlanes; ++k) {
          const double v00 = __ldcg(src[k] + y0 * n + x0), v01 = __ldcg(src[k] + y0 * n + x1);
          const double v10 = __ldcg(src[k] + y1 * n + x0), v11 = __ldcg(src[k] + y1 * n + x1);
          pv[k] = static_cast<T>((v00 * (1.0 - a) + v01 * a) * (1.0 - c) + (v10 * (1.0 - a) + v11 * a) * c);
        }
      }
      // exchange with the canvas pixel under the cell (:63-142)
      const int cpx = x + sp.roi_x, cpy = y + sp.roi_y;
      bool active = !(cpx < 0 || cpy < 0 || cpx >= L.cols || cpy >= L.rows);
      double D    = 0.0;
      if (active) {
        const int tpx = static_cast<int>(cpx - L.bmin_x), tpy = static_cast<int>(cpy - L.bmin_y);
        active = !(tpx < 0 || tpy < 0 || tpx >= L.stroke.local_cols || tpy >= L.stroke.local_rows);
        if (active) {
          const double ddx = sp.cx - static_cast<double>(cpx), ddy = sp.cy - static_cast<double>(cpy);
          active = !(sqrt(ddx * ddx + ddy * ddy) > radius);
          if (active) {
            D      = L.tmap[static_cast<int64_t>(tpy) * L.stroke.local_cols + tpx];
            active = D > 0.0 && cpy >= L.store_first && cpy < L.store_first + L.store_rows;
          }
        }
      }
      if (active) {
        const int64_t ci = static_cast<int64_t>(cpy - L.store_first) * L.cols + cpx;
        const T Dt = static_cast<T>(D);
        const T cV = __ldcg(can[PV] + ci), pV = pv[PV];
        const T cVl = cV * dep * Dt / Dmax, cVr = cV - cVl;
        const T pVl = pV * pick_rate * Dt / Dmax, pVr = pV - pVl;
        T cK[3], cS[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          cK[k] = __ldcg(can[PK + k] + ci);
          cS[k] = __ldcg(can[PS + k] + ci);
        }
        const T pVnew = pVr + cVl, cVnew = cVr + pVl;
        if (cVnew > static_cast<T>(kMinVolume)) {  // deposition (uses the window values before the pickup update)
          const T inv = static_cast<T>(1.) / cVnew;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            __stcg(can[PK + k] + ci, inv * (cVr * cK[k] + pVl * pv[PK + k]));
            __stcg(can[PS + k] + ci, inv * (cVr * cS[k] + pVl * pv[PS + k]));
          }
          __stcg(can[PV] + ci, fmax(cVnew, static_cast<T>(0)));
        }
        if (pVnew > static_cast<T>(kMinVolume)) {  // pickup
          const T inv = static_cast<T>(1.0) / pVnew;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            pv[PK + k] = inv * (pVr * pv[PK + k] + cVl * cK[k]);
            pv[PS + k] = inv * (pVr * pv[PS + k] + cVl * cS[k]);
          }
          pv[PV] = fmax(pVnew, static_cast<T>(0));
        }
      }
#pragma unroll
      for (int k = 0; k < kLayerPlanes; ++k) __stcg(dst[k] + i, pv[k]);
    }
    cluster.sync();
  }
}

// Pass 3 (TextureBrush.hxx:179-204): deposit every pixel recorded in the thickness map.
template <typename T>
__global__ void __launch_bounds__(256) deposit_kernel(const SmudgeLaunch L) {
  const DevTStroke& st = L.stroke;
  T* can[kLayerPlanes];
#pragma unroll
  for (int k = 0; k < kLayerPlanes; ++k) can[k] = static_cast<T*>(L.canvas[k]);
  const T pK[3] = {static_cast<T>(st.K[0]), static_cast<T>(st.K[1]), static_cast<T>(st.K[2])};
  const T pS[3] = {static_cast<T>(st.S[0]), static_cast<T>(st.S[1]), static_cast<T>(st.S[2])};
  const int64_t total  = static_cast<int64_t>(st.local_rows) * st.local_cols;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const double Vtex = L.tmap[i];
    if (!(Vtex > 0.0)) continue;
    const int t = static_cast<int>(i / st.local_cols), s = static_cast<int>(i - static_cast<int64_t>(t) * st.local_cols);
    const int x = s + st.x0, y = t + st.y0;
    if (x < 0 || y < 0 || x >= L.cols || y >= L.rows) continue;
    if (y < L.store_first || y >= L.store_first + L.store_rows) continue;
    const int64_t ci = static_cast<int64_t>(y - L.store_first) * L.cols + x;
    const T vt   = static_cast<T>(Vtex);
    const T vcan = __ldcg(can[PV] + ci);
    const T vsum = vcan + vt;
    if (vsum > static_cast<T>(0)) {
      const T sc = static_cast<T>(1) / vsum;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        __stcg(can[PK + k] + ci, (vcan * __ldcg(can[PK + k] + ci) + vt * pK[k]) * sc);
        __stcg(can[PS + k] + ci, (vcan * __ldcg(can[PS + k] + ci) + vt * pS[k]) * sc);
      }
      __stcg(can[PV] + ci, fmax(vt, vcan));
    }
  }
}

}  // namespace

void texture_launch(pb_context* ctx, const TextureLaunch& L) {
  if (L.n_strokes <= 0 || L.n_items <= 0) return;
  const int n_tiles = L.tiles_x * L.tiles_y;
  texture_ticket_kernel<<<(n_tiles + 255) / 256, 256, 0, ctx->stream>>>(L);
  PB_CUDA(cudaGetLastError());
  ctx->launches++;
  const void* fn = ctx->precision == PB_F64 ? reinterpret_cast<const void*>(texture_kernel<double>)
                                             : reinterpret_cast<const void*>(texture_kernel<float>);
  int per_sm = 0;
  PB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 256, 0));
  PB_REQUIRE(per_sm >= 1, "texture kernel does not fit on an SM");
  const int grid = static_cast<int>(std::min<int64_t>(static_cast<int64_t>(ctx->sm_count) * per_sm, L.n_items));
  if (ctx->precision == PB_F64)
    texture_kernel<double><<<grid, 256, 0, ctx->stream>>>(L);
  else
    texture_kernel<float><<<grid, 256, 0, ctx->stream>>>(L);
  PB_CUDA(cudaGetLastError());
  ctx->launches++;
}

void texture_thickness_launch(pb_context* ctx, const SmudgeLaunch& L) {
  const int64_t total = static_cast<int64_t>(L.stroke.x1 - L.stroke.x0 + 1) * (L.stroke.y1 - L.stroke.y0 + 1);
  if (total <= 0 || L.stroke.n_poly < 2) return;
  const int grid = static_cast<int>(std::min<int64_t>((total + 255) / 256, static_cast<int64_t>(ctx->sm_count) * 8));
  thickness_kernel<<<grid, 256, 0, ctx->stream>>>(L);
  PB_CUDA(cudaGetLastError());
  ctx->launches++;
}

void texture_smudge_launch(pb_context* ctx, const SmudgeLaunch& L) {
  if (L.n_steps <= 0 || L.size <= 0) return;
  const void* fn = ctx->precision == PB_F64 ? reinterpret_cast<const void*>(smudge_kernel<double>)
                                             : reinterpret_cast<const void*>(smudge_kernel<float>);
  const int cells   = L.size * L.size;
  const int cluster = cells <= 1024 ? 1 : (cells <= 4096 ? 2 : (cells <= 16384 ? 4 : 8));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim            = dim3(static_cast<unsigned>(cluster));
  cfg.blockDim           = dim3(256);
  cfg.stream             = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id               = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(cluster);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs                = attr;
  cfg.numAttrs             = 1;
  void* args[]             = {const_cast<SmudgeLaunch*>(&L)};
  PB_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
  ctx->launches++;
}

void texture_deposit_launch(pb_context* ctx, const SmudgeLaunch& L) {
  const int64_t total = static_cast<int64_t>(L.stroke.local_rows) * L.stroke.local_cols;
  if (total <= 0) return;
  const int grid = static_cast<int>(std::min<int64_t>((total + 255) / 256, static_cast<int64_t>(ctx->sm_count) * 8));
  if (ctx->precision == PB_F64)
    deposit_kernel<double><<<grid, 256, 0, ctx->stream>>>(L);
  else
    deposit_kernel<float><<<grid, 256, 0, ctx->stream>>>(L);
  PB_CUDA(cudaGetLastError());
  ctx->launches++;
}

}  // namespace pb
