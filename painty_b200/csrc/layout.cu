// Device plane storage and the AoS-f64 <-> SoA transposes at the reference boundary.
// The reference hands images over as cv::Mat_<Eigen::Vector3d> (24 B/px AoS f64) and cv::Mat_<double>
// (painty/image/Mat.hxx:40-41); on the device they live as dense SoA planes of float or double.
#include <algorithm>
#include <type_traits>

#include "common.cuh"

namespace pb {
namespace {

struct PlanePtrs7 {
  void* p[kLayerPlanes];
};

constexpr int64_t kChunkPx = int64_t(1) << 23;  // 8 Mpx per staging chunk (192 MB for a vec3 image)

template <typename T>
__global__ void fill_kernel(T* p, int64_t n, T v) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

template <typename T, int CH>
__global__ void aos_to_soa_kernel(const double* __restrict__ src, T* __restrict__ d0, T* __restrict__ d1,
                                  T* __restrict__ d2, int64_t n) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    d0[i] = static_cast<T>(src[i * CH]);
    if (CH > 1) d1[i] = static_cast<T>(src[i * CH + 1]);
    if (CH > 2) d2[i] = static_cast<T>(src[i * CH + 2]);
  }
}

template <typename T, int CH>
__global__ void soa_to_aos_kernel(double* __restrict__ dst, const T* __restrict__ s0, const T* __restrict__ s1,
                                  const T* __restrict__ s2, int64_t n) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    dst[i * CH] = static_cast<double>(s0[i]);
    if (CH > 1) dst[i * CH + 1] = static_cast<double>(s1[i]);
    if (CH > 2) dst[i * CH + 2] = static_cast<double>(s2[i]);
  }
}

// Pixel RECORDS of the imprint engine: 8 elements per pixel (Kr Kg Kb Sr Sg Sb V 0), 32 bytes in FP32 mode — one
// 256-bit load / store and one L2 sector per pixel, where the SoA planes cost 7 scattered accesses (imprint.cu).
// The converters move a rectangle [y0, y1] x [x0, x1] of the stored planes (pitch = cols) to / from the record array.
template <typename T>
__global__ void planes_to_records_kernel(PlanePtrs7 src, T* __restrict__ rec, int cols, int x0, int y0, int w, int64_t n) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t r = i / w;
    const int64_t f = (y0 + r) * cols + x0 + (i - r * w);
    T v[8];
#pragma unroll
    for (int k = 0; k < kLayerPlanes; ++k) v[k] = static_cast<const T*>(src.p[k])[f];
    v[7] = static_cast<T>(0);
    using V = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
    V* out = reinterpret_cast<V*>(rec + f * 8);
    const V* in = reinterpret_cast<const V*>(v);
#pragma unroll
    for (int q = 0; q < static_cast<int>(8 * sizeof(T) / sizeof(V)); ++q) out[q] = in[q];
  }
}
template <typename T>
__global__ void records_to_planes_kernel(PlanePtrs7 dst, const T* __restrict__ rec, int cols, int x0, int y0, int w, int64_t n) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t r = i / w;
    const int64_t f = (y0 + r) * cols + x0 + (i - r * w);
    using V = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
    T v[8];
    V* tmp = reinterpret_cast<V*>(v);
    const V* in = reinterpret_cast<const V*>(rec + f * 8);
#pragma unroll
    for (int q = 0; q < static_cast<int>(8 * sizeof(T) / sizeof(V)); ++q) tmp[q] = in[q];
#pragma unroll
    for (int k = 0; k < kLayerPlanes; ++k) static_cast<T*>(dst.p[k])[f] = v[k];
  }
}

inline unsigned grid_for(pb_context* ctx, int64_t n) {
  const int64_t want = (n + 255) / 256;
  return static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>(want, static_cast<int64_t>(ctx->sm_count) * 16)));
}

template <typename T>
void upload_t(pb_context* ctx, const pb_planes& pl, int p0, int ch, const double* host) {
  const int64_t n = pl.n();
  double* stage   = nullptr;
  const int64_t chunk = std::min(n, kChunkPx);
  PB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&stage), sizeof(double) * chunk * ch, ctx->stream));
  for (int64_t o = 0; o < n; o += chunk) {
    const int64_t m = std::min(chunk, n - o);
    PB_CUDA(cudaMemcpyAsync(stage, host + o * ch, sizeof(double) * m * ch, cudaMemcpyHostToDevice, ctx->stream));
    T* d0 = static_cast<T*>(pl.plane(p0)) + o;
    T* d1 = ch > 1 ? static_cast<T*>(pl.plane(p0 + 1)) + o : nullptr;
    T* d2 = ch > 2 ? static_cast<T*>(pl.plane(p0 + 2)) + o : nullptr;
    if (ch == 3)
      aos_to_soa_kernel<T, 3><<<grid_for(ctx, m), 256, 0, ctx->stream>>>(stage, d0, d1, d2, m);
    else
      aos_to_soa_kernel<T, 1><<<grid_for(ctx, m), 256, 0, ctx->stream>>>(stage, d0, d1, d2, m);
    PB_CUDA(cudaGetLastError());
    ctx->launches++;
  }
  PB_CUDA(cudaFreeAsync(stage, ctx->stream));
  PB_CUDA(cudaStreamSynchronize(ctx->stream));
}

template <typename T>
void download_t(pb_context* ctx, const pb_planes& pl, int p0, int ch, double* host) {
  const int64_t n = pl.n();
  double* stage   = nullptr;
  const int64_t chunk = std::min(n, kChunkPx);
  PB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&stage), sizeof(double) * chunk * ch, ctx->stream));
  for (int64_t o = 0; o < n; o += chunk) {
    const int64_t m = std::min(chunk, n - o);
    const T* s0     = static_cast<const T*>(pl.plane(p0)) + o;
    const T* s1     = ch > 1 ? static_cast<const T*>(pl.plane(p0 + 1)) + o : nullptr;
    const T* s2     = ch > 2 ? static_cast<const T*>(pl.plane(p0 + 2)) + o : nullptr;
    if (ch == 3)
      soa_to_aos_kernel<T, 3><<<grid_for(ctx, m), 256, 0, ctx->stream>>>(stage, s0, s1, s2, m);
    else
      soa_to_aos_kernel<T, 1><<<grid_for(ctx, m), 256, 0, ctx->stream>>>(stage, s0, s1, s2, m);
    PB_CUDA(cudaGetLastError());
    ctx->launches++;
    PB_CUDA(cudaMemcpyAsync(host + o * ch, stage, sizeof(double) * m * ch, cudaMemcpyDeviceToHost, ctx->stream));
  }
  PB_CUDA(cudaFreeAsync(stage, ctx->stream));
  PB_CUDA(cudaStreamSynchronize(ctx->stream));
}

}  // namespace

void planes_alloc(pb_context* ctx, pb_planes& pl, int rows, int cols, int nplanes) {
  PB_REQUIRE(rows >= 0 && cols >= 0, "negative image size");
  pl.ctx     = ctx;
  pl.rows    = rows;
  pl.cols    = cols;
  pl.nplanes = nplanes;
  const size_t bytes = static_cast<size_t>(rows) * cols * ctx->esize();
  pl.stride          = (bytes + 255) / 256 * 256;
  if (pl.stride == 0) pl.stride = 256;
  PB_CUDA(cudaMalloc(&pl.base, pl.stride * nplanes));
}

// stream-ordered temporaries (compose results on their way to the host): no device-wide synchronisation
void planes_alloc_temp(pb_context* ctx, pb_planes& pl, int rows, int cols, int nplanes) {
  pl.ctx     = ctx;
  pl.rows    = rows;
  pl.cols    = cols;
  pl.nplanes = nplanes;
  const size_t bytes = static_cast<size_t>(rows) * cols * ctx->esize();
  pl.stride          = std::max<size_t>((bytes + 255) / 256 * 256, 256);
  PB_CUDA(cudaMallocAsync(&pl.base, pl.stride * nplanes, ctx->stream));
}
void planes_free_temp(pb_planes& pl) {
  if (pl.base) cudaFreeAsync(pl.base, pl.ctx->stream);
  pl.base = nullptr;
}

void planes_free(pb_planes& pl) {
  if (pl.base) cudaFree(pl.base);
  pl.base = nullptr;
}

void fill_plane(pb_context* ctx, void* plane, int64_t n, double value) {
  if (n <= 0) return;
  if (value == 0.0) {
    PB_CUDA(cudaMemsetAsync(plane, 0, static_cast<size_t>(n) * ctx->esize(), ctx->stream));
    return;
  }
  if (ctx->precision == PB_F64)
    fill_kernel<double><<<grid_for(ctx, n), 256, 0, ctx->stream>>>(static_cast<double*>(plane), n, value);
  else
    fill_kernel<float><<<grid_for(ctx, n), 256, 0, ctx->stream>>>(static_cast<float*>(plane), n, static_cast<float>(value));
  PB_CUDA(cudaGetLastError());
  ctx->launches++;
}

void upload_aos(pb_context* ctx, const pb_planes& pl, int p0, int ch, const double* host) {
  if (pl.n() == 0) return;
  if (ctx->precision == PB_F64)
    upload_t<double>(ctx, pl, p0, ch, host);
  else
    upload_t<float>(ctx, pl, p0, ch, host);
}

void download_aos(pb_context* ctx, const pb_planes& pl, int p0, int ch, double* host) {
  if (pl.n() == 0) return;
  if (ctx->precision == PB_F64)
    download_t<double>(ctx, pl, p0, ch, host);
  else
    download_t<float>(ctx, pl, p0, ch, host);
}

void planes_to_records(pb_context* ctx, const pb_planes& pl, void* records, int x0, int y0, int x1, int y1) {
  if (x1 < x0 || y1 < y0) return;
  PlanePtrs7 pp;
  for (int k = 0; k < kLayerPlanes; ++k) pp.p[k] = pl.plane(k);
  const int w = x1 - x0 + 1;
  const int64_t n = static_cast<int64_t>(w) * (y1 - y0 + 1);
  if (ctx->precision == PB_F64)
    planes_to_records_kernel<double><<<grid_for(ctx, n), 256, 0, ctx->stream>>>(pp, static_cast<double*>(records), pl.cols, x0, y0, w, n);
  else
    planes_to_records_kernel<float><<<grid_for(ctx, n), 256, 0, ctx->stream>>>(pp, static_cast<float*>(records), pl.cols, x0, y0, w, n);
  PB_CUDA(cudaGetLastError());
  ctx->launches++;
}

void records_to_planes(pb_context* ctx, const void* records, const pb_planes& pl, int x0, int y0, int x1, int y1) {
  if (x1 < x0 || y1 < y0) return;
  PlanePtrs7 pp;
  for (int k = 0; k < kLayerPlanes; ++k) pp.p[k] = pl.plane(k);
  const int w = x1 - x0 + 1;
  const int64_t n = static_cast<int64_t>(w) * (y1 - y0 + 1);
  if (ctx->precision == PB_F64)
    records_to_planes_kernel<double><<<grid_for(ctx, n), 256, 0, ctx->stream>>>(pp, static_cast<const double*>(records), pl.cols, x0, y0, w, n);
  else
    records_to_planes_kernel<float><<<grid_for(ctx, n), 256, 0, ctx->stream>>>(pp, static_cast<const float*>(records), pl.cols, x0, y0, w, n);
  PB_CUDA(cudaGetLastError());
  ctx->launches++;
}

void copy_planes(pb_context* ctx, const pb_planes& src, pb_planes& dst, int nplanes) {
  PB_REQUIRE(src.rows == dst.rows && src.cols == dst.cols, "copy_planes: size mismatch");
  for (int p = 0; p < nplanes; ++p)
    PB_CUDA(cudaMemcpyAsync(dst.plane(p), src.plane(p), static_cast<size_t>(src.n()) * ctx->esize(),
                            cudaMemcpyDeviceToDevice, ctx->stream));
}

}  // namespace pb
