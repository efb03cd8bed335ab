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
#include "texture.cuh"

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
  auto edge = [&](int i, double2& si, double2& si1) {
    const double2 p = poly[i], q = poly[i == n - 1 ? 0 : i + 1];
    si  = make_double2(p.x - x, p.y - y);
    si1 = make_double2(q.x - x, q.y - y);
    Edge e;
    e.r = sqrt(si.x * si.x + si.y * si.y);
    e.A = (si.x * si1.y - si1.x * si.y) / 2.0;
    e.D = si.x * si1.x + si.y * si1.y;
    return e;
  };
  double2 s0, s1;
  Edge prev = edge(n - 1, s0, s1);  // A_{i-1}, r_{i-1}, D_{i-1} for i = 0
  const double r_first = sqrt((poly[0].x - x) * (poly[0].x - x) + (poly[0].y - y) * (poly[0].y - y));
  double fu = 0.0, fv = 0.0, W = 0.0;
  for (int i = 0; i < n; ++i) {
    const Edge cur = edge(i, s0, s1);
    if (fabs(cur.r - 0.0) < Eps) {  // :102-104
      ou = uv[i].x;
      ov = uv[i].y;
      return;
    }
    if (fabs(cur.A - 0.0) < Eps && cur.D < 0.0) {  // :114-119
      const double ri1 = sqrt(s1.x * s1.x + s1.y * s1.y);
      const double2 f1 = uv[i == n - 1 ? 0 : i + 1];
      const double sc  = 1.0 / (cur.r + ri1);
      ou = (ri1 * uv[i].x + cur.r * f1.x) * sc;
      ov = (ri1 * uv[i].y + cur.r * f1.y) * sc;
      return;
    }
    double w = 0.0;
    if (prev.A != 0.0) w = w + (prev.r - prev.D / cur.r) / prev.A;
    if (cur.A != 0.0) {
      const double ri1 = (i == n - 1) ? r_first : sqrt(s1.x * s1.x + s1.y * s1.y);
      w                = w + (ri1 - cur.D / cur.r) / cur.A;
    }
    fu = fu + w * uv[i].x;
    fv = fv + w * uv[i].y;
    W  = W + w;
    prev = cur;
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
      u *= L.map_cols;
      v *= L.map_rows;
      const double Vtex = st.thickness_scale * bilinear(L.map, L.map_rows, L.map_cols, u, v);
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

}  // namespace pb
