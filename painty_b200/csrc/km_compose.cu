// Kubelka-Munk layer compose — fused, coalesced, 128-bit vectorised streaming kernels for sm_100a.
//
// Restates (not copies) painty/core/KubelkaMunk.hxx:28-83 + painty/core/Math.hxx:159-169 per pixel
// and the whole-image loops painty/renderer/Renderer.hxx:26-41 (compose), PaintLayer.hxx:81-96
// (composeOnto), Canvas.hxx:105-121 (dryCanvas) over SoA planes in HBM.
//
// Roofline: pure streaming map, no reuse -> HBM bound. Algorithmic bytes FP32: 7 layer planes + 3 R0
// planes read, 3 R planes written = 52 B/px; L stacked layers 28 L + 24; dry 88 B/px; FP64 x2.
// Each thread owns 4 consecutive pixels (float4 / 2 x double2 per plane): 10 independent 16 B loads
// in flight per thread, streaming (evict-first) loads and stores, no shared memory, no tensor cores
// (nothing here is a contraction).
//
// FP32 math is a reformulation that avoids the reference's cancellations and its libm calls:
//   ks = K/S', a = 1+ks, b = sqrt(ks(ks+2))            (= sqrt(a^2-1) without the a^2-1 cancellation)
//   b coth(x) = b + 2b/expm1(2x),  x = b S' d           (expm1: degree-5 polynomial for |2x|<0.25, else
//                                                         ex2.approx; K==0 gives 0/0 = NaN like the reference)
// 5 MUFU ops per channel (rcp, sqrt, ex2, rcp, rcp). The f64 thresholds of the reference are used in
// both precisions (a float instantiation of the reference would early-out at d < 1.19e-3).
// FP64 validation mode follows the reference's operation order exactly; the library is compiled with
// -fmad=false so no FMA contraction happens there, FP32 uses explicit fmaf.
#include <math_constants.h>

#include "common.cuh"

namespace pb {
namespace {

__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ float km_channel(float K, float S_in, float R0, float d) {
  const float S  = (fabsf(S_in) > static_cast<float>(kKmEps)) ? S_in : 1e-11f;
  const float ks = __fdividef(K, S);
  const float a  = 1.0f + ks;
  const float b  = sqrt_approx(fmaxf(fmaf(ks, ks, ks + ks), 0.0f));
  const float y  = 2.0f * b * S * d;
  // expm1(y)
  float p = fmaf(y, 1.0f / 120.0f, 1.0f / 24.0f);
  p       = fmaf(p, y, 1.0f / 6.0f);
  p       = fmaf(p, y, 0.5f);
  p       = fmaf(p, y, 1.0f);
  p       = p * y;
  const float e   = __expf(y) - 1.0f;
  const float em1 = (fabsf(y) < 0.25f) ? p : e;
  const float c   = b + __fdividef(b + b, em1);
  return __fdividef(fmaf(-R0, a - c, 1.0f), (a - R0) + c);
}

// reference order of operations, IEEE f64 (KubelkaMunk.hxx:36-80, Math.hxx:159-169)
__device__ __forceinline__ double km_channel(double K, double S_in, double R0, double d) {
  const double S = (fabs(S_in) > kKmEps) ? S_in : 0.00000000001;
  const double a = 1.0 + K / S;
  const double v = a * a - 1.0;
  const double b = (v < 0.0) ? 0.0 : sqrt(v);
  const double x = b * S * d;
  double coth;
  if (x > 20.0) {
    coth = 1.0;
  } else if (fabs(x) > 0.0) {
    const double r = cosh(x) / sinh(x);
    coth           = isnan(r) ? 1.0 : r;
  } else {
    coth = CUDART_INF;
  }
  const double c = b * coth;
  return (1.0 - R0 * (a - c)) / (a - R0 + c);
}

__device__ __forceinline__ bool is_dry(float v) { return fabsf(v) < static_cast<float>(kKmEps); }
__device__ __forceinline__ bool is_dry(double v) { return fabs(v) < kKmEps; }

template <typename T>
__device__ __forceinline__ void km_pixel(T k0, T k1, T k2, T s0, T s1, T s2, T v, T& r0, T& r1, T& r2) {
  if (is_dry(v)) return;  // KubelkaMunk.hxx:31-34: R = R0
  r0 = km_channel(k0, s0, r0, v);
  r1 = km_channel(k1, s1, r1, v);
  r2 = km_channel(k2, s2, r2, v);
}

template <typename T>
__device__ __forceinline__ void ld4(const T* p, T (&o)[4]);
template <>
__device__ __forceinline__ void ld4<float>(const float* p, float (&o)[4]) {
  const float4 v = __ldcs(reinterpret_cast<const float4*>(p));
  o[0] = v.x, o[1] = v.y, o[2] = v.z, o[3] = v.w;
}
template <>
__device__ __forceinline__ void ld4<double>(const double* p, double (&o)[4]) {
  const double2 a = __ldcs(reinterpret_cast<const double2*>(p));
  const double2 b = __ldcs(reinterpret_cast<const double2*>(p) + 1);
  o[0] = a.x, o[1] = a.y, o[2] = b.x, o[3] = b.y;
}
template <typename T>
__device__ __forceinline__ void st4(T* p, const T (&o)[4]);
template <>
__device__ __forceinline__ void st4<float>(float* p, const float (&o)[4]) {
  __stcs(reinterpret_cast<float4*>(p), make_float4(o[0], o[1], o[2], o[3]));
}
template <>
__device__ __forceinline__ void st4<double>(double* p, const double (&o)[4]) {
  __stcs(reinterpret_cast<double2*>(p), make_double2(o[0], o[1]));
  __stcs(reinterpret_cast<double2*>(p) + 1, make_double2(o[2], o[3]));
}

template <typename T>
struct ComposePtrs {
  const T* K[3];
  const T* S[3];
  const T* V;
  const T* R0[3];
  T* R[3];
};

// One thread = 4 consecutive pixels. All loads are issued before any math (10 x 16 B in flight).
template <typename T>
__global__ void __launch_bounds__(256) km_compose_kernel(ComposePtrs<T> a, int64_t n4, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n4) {
    const int64_t o = i * 4;
    T k[3][4], s[3][4], v[4], r[3][4];
#pragma unroll
    for (int c = 0; c < 3; ++c) ld4(a.K[c] + o, k[c]);
#pragma unroll
    for (int c = 0; c < 3; ++c) ld4(a.S[c] + o, s[c]);
    ld4(a.V + o, v);
#pragma unroll
    for (int c = 0; c < 3; ++c) ld4(a.R0[c] + o, r[c]);
#pragma unroll
    for (int j = 0; j < 4; ++j) km_pixel(k[0][j], k[1][j], k[2][j], s[0][j], s[1][j], s[2][j], v[j], r[0][j], r[1][j], r[2][j]);
#pragma unroll
    for (int c = 0; c < 3; ++c) st4(a.R[c] + o, r[c]);
  } else if (i == n4) {  // scalar tail (n % 4 pixels)
    for (int64_t o = n4 * 4; o < n; ++o) {
      T r0 = a.R0[0][o], r1 = a.R0[1][o], r2 = a.R0[2][o];
      km_pixel(a.K[0][o], a.K[1][o], a.K[2][o], a.S[0][o], a.S[1][o], a.S[2][o], a.V[o], r0, r1, r2);
      a.R[0][o] = r0, a.R[1][o] = r1, a.R[2][o] = r2;
    }
  }
}

// Compose + band gather (multi GPU): the band's reflectance rows are stored straight into the assembled image of up to
// kGatherMax ranks — the rank's own HBM and CUDA-IPC peer mappings reached over NVLink — instead of composing into a local
// buffer and running a collective afterwards. dst[d][c] already points at this band's first row in destination d.
constexpr int kGatherMax = 8;
template <typename T>
struct GatherPtrs {
  T* dst[kGatherMax][3];
  int n_dst;
};
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) km_compose_gather_kernel(ComposePtrs<T> a, GatherPtrs<T> g, int64_t n4, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (VEC && i < n4) {
    const int64_t o = i * 4;
    T k[3][4], s[3][4], v[4], r[3][4];
#pragma unroll
    for (int c = 0; c < 3; ++c) ld4(a.K[c] + o, k[c]);
#pragma unroll
    for (int c = 0; c < 3; ++c) ld4(a.S[c] + o, s[c]);
    ld4(a.V + o, v);
#pragma unroll
    for (int c = 0; c < 3; ++c) ld4(a.R0[c] + o, r[c]);
#pragma unroll
    for (int j = 0; j < 4; ++j) km_pixel(k[0][j], k[1][j], k[2][j], s[0][j], s[1][j], s[2][j], v[j], r[0][j], r[1][j], r[2][j]);
    for (int d = 0; d < g.n_dst; ++d) {
#pragma unroll
      for (int c = 0; c < 3; ++c) st4(g.dst[d][c] + o, r[c]);
    }
  } else if (VEC ? i == n4 : i < n) {  // scalar tail (n % 4 pixels), or everything when a destination is not 16 B aligned
    const int64_t o0 = VEC ? n4 * 4 : i, o1 = VEC ? n : i + 1;
    for (int64_t o = o0; o < o1; ++o) {
      T r0 = a.R0[0][o], r1 = a.R0[1][o], r2 = a.R0[2][o];
      km_pixel(a.K[0][o], a.K[1][o], a.K[2][o], a.S[0][o], a.S[1][o], a.S[2][o], a.V[o], r0, r1, r2);
      for (int d = 0; d < g.n_dst; ++d) g.dst[d][0][o] = r0, g.dst[d][1][o] = r1, g.dst[d][2][o] = r2;
    }
  }
}

// fallback for caller-provided planes that are not 16 B aligned: one pixel per thread
template <typename T>
__global__ void __launch_bounds__(256) km_compose_scalar_kernel(ComposePtrs<T> a, int64_t n) {
  const int64_t o = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (o < n) {
    T r0 = a.R0[0][o], r1 = a.R0[1][o], r2 = a.R0[2][o];
    km_pixel(a.K[0][o], a.K[1][o], a.K[2][o], a.S[0][o], a.S[1][o], a.S[2][o], a.V[o], r0, r1, r2);
    a.R[0][o] = r0, a.R[1][o] = r1, a.R[2][o] = r2;
  }
}

template <typename T>
struct StackPtrs {
  const T* K[kMaxStack][3];
  const T* S[kMaxStack][3];
  const T* V[kMaxStack];
  const T* R0[3];
  T* R[3];
  int n_layers;
};

// L layers bottom-up, R stays in registers between layers: (28 L + 24) B/px instead of 52 L.
template <typename T>
__global__ void __launch_bounds__(256) km_stacked_kernel(StackPtrs<T> a, int64_t n4, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n4) {
    const int64_t o = i * 4;
    T r[3][4];
#pragma unroll
    for (int c = 0; c < 3; ++c) ld4(a.R0[c] + o, r[c]);
    for (int l = 0; l < a.n_layers; ++l) {
      T k[3][4], s[3][4], v[4];
#pragma unroll
      for (int c = 0; c < 3; ++c) ld4(a.K[l][c] + o, k[c]);
#pragma unroll
      for (int c = 0; c < 3; ++c) ld4(a.S[l][c] + o, s[c]);
      ld4(a.V[l] + o, v);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        km_pixel(k[0][j], k[1][j], k[2][j], s[0][j], s[1][j], s[2][j], v[j], r[0][j], r[1][j], r[2][j]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) st4(a.R[c] + o, r[c]);
  } else if (i == n4) {
    for (int64_t o = n4 * 4; o < n; ++o) {
      T r0 = a.R0[0][o], r1 = a.R0[1][o], r2 = a.R0[2][o];
      for (int l = 0; l < a.n_layers; ++l)
        km_pixel(a.K[l][0][o], a.K[l][1][o], a.K[l][2][o], a.S[l][0][o], a.S[l][1][o], a.S[l][2][o], a.V[l][o], r0, r1, r2);
      a.R[0][o] = r0, a.R[1][o] = r1, a.R[2][o] = r2;
    }
  }
}

// ColorConverter::rgb2srgb (core/Color.hxx:198-206)
__device__ __forceinline__ float to_srgb(float l) {
  return l <= 0.00313066844250063f ? l * 12.92f : fmaf(1.055f, powf(l, 1.0f / 2.4f), -0.055f);
}
__device__ __forceinline__ double to_srgb(double l) {
  return l <= 0.00313066844250063 ? l * 12.92 : 1.055 * pow(l, 1.0 / 2.4) - 0.055;
}

// compose + display epilogue, one pixel per thread (the output is 3-6 B/px against 40 B/px of input)
template <typename T>
__global__ void __launch_bounds__(256) km_compose_display_kernel(ComposePtrs<T> a, int64_t n, int mode, bool srgb, void* out) {
  const int64_t o = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (o >= n) return;
  T r[3] = {__ldcs(a.R0[0] + o), __ldcs(a.R0[1] + o), __ldcs(a.R0[2] + o)};
  km_pixel(__ldcs(a.K[0] + o), __ldcs(a.K[1] + o), __ldcs(a.K[2] + o), __ldcs(a.S[0] + o), __ldcs(a.S[1] + o), __ldcs(a.S[2] + o),
           __ldcs(a.V + o), r[0], r[1], r[2]);
  if (srgb) {
#pragma unroll
    for (int c = 0; c < 3; ++c) r[c] = to_srgb(r[c]);
  }
  if (mode == 0) {  // qRgb(static_cast<uint8_t>(v * 255.0)): truncation; out-of-range values are clamped here
    unsigned q[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const T v = r[c] * static_cast<T>(255.0);
      q[c]      = v >= static_cast<T>(255.0) ? 255u : (v > static_cast<T>(0) ? static_cast<unsigned>(v) : 0u);
    }
    static_cast<unsigned*>(out)[o] = 0xff000000u | (q[0] << 16) | (q[1] << 8) | q[2];
  } else {  // cv::Mat::convertTo: saturate_cast(cvRound(v * scale)), then RGB -> BGR
    const T scale = mode == 1 ? static_cast<T>(255.0) : static_cast<T>(65535.0);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const T v   = rint(r[c] * scale);  // round half to even like cvRound
      const T cl  = v < static_cast<T>(0) ? static_cast<T>(0) : (v > scale ? scale : v);
      const int q = isnan(v) ? 0 : static_cast<int>(cl);
      if (mode == 1)
        static_cast<unsigned char*>(out)[3 * o + (2 - c)] = static_cast<unsigned char>(q);
      else
        static_cast<unsigned short*>(out)[3 * o + (2 - c)] = static_cast<unsigned short>(q);
    }
  }
}

// ---- planner read-back prep: compose -> CIELab -> LANCZOS4 down-scale --------------------------
// ColorConverter::rgb2lab (core/Color.hxx:248-252): rgb2xyz with the sRGB D65 matrix (:169-176), xyz2lab (:89-93) with
// f(t) = t > (6/29)^3 ? t^(1/3) : (1/3)(29/6)^2 t + 4/29 (:786-790), white point D65 (:48-50). FP64 keeps the reference's
// operation order and pow(t, 1/3); FP32 uses cbrtf.
__device__ __forceinline__ double lab_f(double t) {
  return t > 0.008856451679035631 ? pow(t, 1. / 3.) : 7.787037037037035 * t + (4. / 29.);
}
__device__ __forceinline__ float lab_f(float t) { return t > 0.008856451679035631f ? cbrtf(t) : fmaf(7.787037037037035f, t, 4.f / 29.f); }
template <typename T>
__device__ __forceinline__ void rgb_to_lab(const T rgb[3], T lab[3]) {
  const T X = static_cast<T>(0.4124564) * rgb[0] + static_cast<T>(0.3575761) * rgb[1] + static_cast<T>(0.1804375) * rgb[2];
  const T Y = static_cast<T>(0.2126729) * rgb[0] + static_cast<T>(0.7151522) * rgb[1] + static_cast<T>(0.0721750) * rgb[2];
  const T Z = static_cast<T>(0.0193339) * rgb[0] + static_cast<T>(0.1191920) * rgb[1] + static_cast<T>(0.9503041) * rgb[2];
  const T fx = lab_f(X / static_cast<T>(0.95047)), fy = lab_f(Y / static_cast<T>(1.00000)), fz = lab_f(Z / static_cast<T>(1.08883));
  lab[0] = static_cast<T>(116.) * fy - static_cast<T>(16.);
  lab[1] = static_cast<T>(500.) * (fx - fy);
  lab[2] = static_cast<T>(200.) * (fy - fz);
}

// compose + Lab, one pixel per thread; the Lab image stays on the device (3 planes of T)
template <typename T>
__global__ void __launch_bounds__(256) km_compose_lab_kernel(ComposePtrs<T> a, int64_t n) {
  const int64_t o = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (o >= n) return;
  T r[3] = {__ldcs(a.R0[0] + o), __ldcs(a.R0[1] + o), __ldcs(a.R0[2] + o)};
  km_pixel(__ldcs(a.K[0] + o), __ldcs(a.K[1] + o), __ldcs(a.K[2] + o), __ldcs(a.S[0] + o), __ldcs(a.S[1] + o), __ldcs(a.S[2] + o),
           __ldcs(a.V + o), r[0], r[1], r[2]);
  T lab[3];
  rgb_to_lab(r, lab);
#pragma unroll
  for (int c = 0; c < 3; ++c) a.R[c][o] = lab[c];
}

// cv::resize(INTER_LANCZOS4) as OpenCV evaluates it for a CV_64F image (the reference's ScaledMat, image/Mat.hxx:141-147):
// separable 8-tap passes, horizontal first; tap j of output column dx reads source column clamp(xofs[dx] - 3 + j) with the
// single-precision weight alpha[dx][j]; products and the left-to-right sums are double. Offsets and weights come from the
// host (lanczos4_taps in capi.cu). Horizontal pass: Lab planes (T) -> tmp[c][row][dx] (double).
template <typename T>
__global__ void __launch_bounds__(256) lanczos4_h_kernel(const T* l0, const T* l1, const T* l2, int rows, int cols, int ocols,
                                                         const int* __restrict__ xofs, const float* __restrict__ alpha, double* tmp) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(rows) * ocols) return;
  const int row = static_cast<int>(i / ocols), dx = static_cast<int>(i - static_cast<int64_t>(row) * ocols);
  const int sx = xofs[dx] - 3;
  const T* src[3] = {l0, l1, l2};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const T* S = src[c] + static_cast<int64_t>(row) * cols;
    double v   = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int x    = min(max(sx + j, 0), cols - 1);
      const double p = static_cast<double>(S[x]) * static_cast<double>(alpha[dx * 8 + j]);
      v              = j == 0 ? p : v + p;
    }
    tmp[(static_cast<int64_t>(c) * rows + row) * ocols + dx] = v;
  }
}
// Vertical pass: out[dy][dx][c] (AoS double, the layout of the Mat3d the planner receives)
__global__ void __launch_bounds__(256) lanczos4_v_kernel(const double* tmp, int rows, int ocols, int orows, const int* __restrict__ yofs,
                                                         const float* __restrict__ beta, double* out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(orows) * ocols) return;
  const int dy = static_cast<int>(i / ocols), dx = static_cast<int>(i - static_cast<int64_t>(dy) * ocols);
  const int sy = yofs[dy] - 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int y    = min(max(sy + k, 0), rows - 1);
      const double p = tmp[(static_cast<int64_t>(c) * rows + y) * ocols + dx] * static_cast<double>(beta[dy * 8 + k]);
      v              = k == 0 ? p : v + p;
    }
    out[i * 3 + c] = v;
  }
}
template <typename T>
__global__ void __launch_bounds__(256) planes_to_aos_f64_kernel(const T* l0, const T* l1, const T* l2, int64_t n, double* out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[3 * i]     = static_cast<double>(l0[i]);
  out[3 * i + 1] = static_cast<double>(l1[i]);
  out[3 * i + 2] = static_cast<double>(l2[i]);
}

// Renderer::render (renderer/Renderer.hxx:60-156) per pixel, after the fused compose. Operation order follows the
// reference (3-vector reductions left to right). T = float uses the single-precision libm-equivalents.
template <typename T>
struct V3 {
  T x, y, z;
};
template <typename T>
__device__ __forceinline__ T dot3(const V3<T>& a, const V3<T>& b) {
  return a.x * b.x + a.y * b.y + a.z * b.z;
}
template <typename T>
__device__ __forceinline__ V3<T> normalized(const V3<T>& a) {
  const T n = sqrt(dot3(a, a));
  return {a.x / n, a.y / n, a.z / n};
}
__device__ __forceinline__ int reflect_idx(int p, int len) {  // cv::borderInterpolate(BORDER_REFLECT)
  if (static_cast<unsigned>(p) < static_cast<unsigned>(len)) return p;
  if (len == 1) return 0;
  do {
    p = (p < 0) ? (-p - 1) : (len - 1 - (p - len));
  } while (static_cast<unsigned>(p) >= static_cast<unsigned>(len));
  return p;
}

template <typename T>
__global__ void __launch_bounds__(256) km_render_kernel(ComposePtrs<T> a, int rows, int cols) {
  const int64_t o = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (o >= static_cast<int64_t>(rows) * cols) return;
  const int i = static_cast<int>(o / cols), j = static_cast<int>(o - static_cast<int64_t>(i) * cols);
  // compose (Kd)
  T kd[3] = {a.R0[0][o], a.R0[1][o], a.R0[2][o]};
  const T s11 = a.V[o];
  km_pixel(a.K[0][o], a.K[1][o], a.K[2][o], a.S[0][o], a.S[1][o], a.S[2][o], s11, kd[0], kd[1], kd[2]);
  // Interpolate() at integer offsets degenerates to the reflected neighbour (:99-107)
  auto H = [&](int y, int x) {
    const int xx = reflect_idx(x, cols), yy = reflect_idx(y, rows);
    // bilinear with zero fractional part: (v00*(1-0) + v01*0)*(1-0) + (v10*(1-0) + v11*0)*0
    const int x1 = reflect_idx(x + 1, cols), y1 = reflect_idx(y + 1, rows);
    const T v00 = a.V[static_cast<int64_t>(yy) * cols + xx], v01 = a.V[static_cast<int64_t>(yy) * cols + x1];
    const T v10 = a.V[static_cast<int64_t>(y1) * cols + xx], v11 = a.V[static_cast<int64_t>(y1) * cols + x1];
    const T z = static_cast<T>(0), one = static_cast<T>(1);
    return (v00 * (one - z) + v01 * z) * (one - z) + (v10 * (one - z) + v11 * z) * z;
  };
  const T s01 = H(i, j - 1), s21 = H(i, j + 1), s10 = H(i - 1, j), s12 = H(i + 1, j);
  const T two = static_cast<T>(2.0), zero = static_cast<T>(0.0), one = static_cast<T>(1.0);
  const V3<T> va = normalized(V3<T>{two, zero, s21 - s01});
  const V3<T> vb = normalized(V3<T>{zero, two, s12 - s10});
  V3<T> n = normalized(V3<T>{va.y * vb.z - va.z * vb.y, va.z * vb.x - va.x * vb.z, va.x * vb.y - va.y * vb.x});
  n.z *= static_cast<T>(-1.);
  const V3<T> pix = {static_cast<T>(j), static_cast<T>(i), s11};
  const V3<T> light_dir = normalized(V3<T>{static_cast<T>(-200) - pix.x, static_cast<T>(-1500) - pix.y, static_cast<T>(-2000.) - pix.z});
  const V3<T> l = normalized(light_dir);
  const V3<T> v = normalized(V3<T>{static_cast<T>(cols / 2.0) - pix.x, static_cast<T>(rows / 2.0) - pix.y, static_cast<T>(-100.) - pix.z});
  const V3<T> h = normalized(V3<T>{v.x + l.x, v.y + l.y, v.z + l.z});
  const T NdotH = fmax(zero, dot3(n, h)), VdotH = fmax(zero, dot3(v, h));
  const T NdotV = fmax(zero, dot3(n, v)), NdotL = fmax(zero, dot3(n, l));
  const T m = static_cast<T>(0.5), sfrac = static_cast<T>(0.2), pi = static_cast<T>(3.141592653589793238462643383279502884);
  T spec = zero;
  if (NdotL > zero && NdotV > zero) {
    const T A  = one / (pow(m, two) + pow(NdotH, static_cast<T>(4.0)) * pi);
    const T B  = exp(-pow(tan(acos(NdotH)), two) / pow(m, two));
    const T G1 = two * NdotH * NdotV / VdotH, G2 = two * NdotH * NdotL / VdotH;
    const T G  = fmin(one, fmin(G1, G2));
    // R_F = Ks + (1 - Ks) * pow(1 - VdotH, 5) with Ks = 1
    const T rf = one + (one - one) * pow(one - VdotH, static_cast<T>(5.0));
    spec       = (A * B * G * rf) / (NdotL * NdotV);
  }
  const T ln   = sqrt(dot3(light_dir, light_dir));
  const T beta = static_cast<T>(15.0) * (one / (static_cast<T>(4.0) * pi * pow(ln, two)));
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const T ambient = kd[c] * sfrac;
    const T r       = (beta * NdotL) * ((one - sfrac) * kd[c] + sfrac * spec) + ambient * kd[c];
    a.R[c][o]       = fmin(fmax(r, zero), one);
  }
}

template <typename T>
struct DryPtrs {
  T* p[kCanvasPlanes];
};

// Canvas::dryCanvas (Canvas.hxx:105-121): h += V; R0 = KM(K,S,R0,V); K = S = V = 0. 88 B/px FP32.
template <typename T>
__global__ void __launch_bounds__(256) km_dry_kernel(DryPtrs<T> a, int64_t n4, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n4) {
    const int64_t o = i * 4;
    T k[3][4], s[3][4], v[4], r[3][4], h[4];
#pragma unroll
    for (int c = 0; c < 3; ++c) ld4(a.p[PK + c] + o, k[c]);
#pragma unroll
    for (int c = 0; c < 3; ++c) ld4(a.p[PS + c] + o, s[c]);
    ld4(a.p[PV] + o, v);
#pragma unroll
    for (int c = 0; c < 3; ++c) ld4(a.p[PR + c] + o, r[c]);
    ld4(a.p[PH] + o, h);
    const T z[4] = {T(0), T(0), T(0), T(0)};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      h[j] += v[j];
      km_pixel(k[0][j], k[1][j], k[2][j], s[0][j], s[1][j], s[2][j], v[j], r[0][j], r[1][j], r[2][j]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) st4(a.p[PR + c] + o, r[c]);
    st4(a.p[PH] + o, h);
#pragma unroll
    for (int c = 0; c < 7; ++c) st4(a.p[c] + o, z);
  } else if (i == n4) {
    for (int64_t o = n4 * 4; o < n; ++o) {
      T r0 = a.p[PR][o], r1 = a.p[PR + 1][o], r2 = a.p[PR + 2][o];
      const T v = a.p[PV][o];
      a.p[PH][o] += v;
      km_pixel(a.p[0][o], a.p[1][o], a.p[2][o], a.p[3][o], a.p[4][o], a.p[5][o], v, r0, r1, r2);
      a.p[PR][o] = r0, a.p[PR + 1][o] = r1, a.p[PR + 2][o] = r2;
      for (int c = 0; c < 7; ++c) a.p[c][o] = T(0);
    }
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T>
void compose_t(pb_context* ctx, int64_t n, const ComposeArgs& a) {
  ComposePtrs<T> p;
  bool al = aligned16(a.V);
  for (int c = 0; c < 3; ++c) {
    p.K[c]  = static_cast<const T*>(a.K[c]);
    p.S[c]  = static_cast<const T*>(a.S[c]);
    p.R0[c] = static_cast<const T*>(a.R0[c]);
    p.R[c]  = static_cast<T*>(a.R[c]);
    al      = al && aligned16(a.K[c]) && aligned16(a.S[c]) && aligned16(a.R0[c]) && aligned16(a.R[c]);
  }
  p.V = static_cast<const T*>(a.V);
  if (!al) {
    km_compose_scalar_kernel<T><<<static_cast<unsigned>((n + 255) / 256), 256, 0, ctx->stream>>>(p, n);
    PB_CUDA(cudaGetLastError());
    ctx->launches++;
    return;
  }
  // 8-byte elements keep 16 B alignment for 4-pixel groups too (32 B per group)
  const int64_t n4     = n / 4;
  const int64_t blocks = (n4 + 1 + 255) / 256;
  km_compose_kernel<T><<<static_cast<unsigned>(blocks), 256, 0, ctx->stream>>>(p, n4, n);
  PB_CUDA(cudaGetLastError());
  ctx->launches++;
}

template <typename T>
void stacked_t(pb_context* ctx, int64_t n, const StackArgs& a) {
  StackPtrs<T> p;
  bool al = true;
  for (int l = 0; l < a.n_layers; ++l) {
    for (int c = 0; c < 3; ++c) {
      p.K[l][c] = static_cast<const T*>(a.K[l][c]);
      p.S[l][c] = static_cast<const T*>(a.S[l][c]);
      al        = al && aligned16(a.K[l][c]) && aligned16(a.S[l][c]);
    }
    p.V[l] = static_cast<const T*>(a.V[l]);
    al     = al && aligned16(a.V[l]);
  }
  for (int c = 0; c < 3; ++c) {
    p.R0[c] = static_cast<const T*>(a.R0[c]);
    p.R[c]  = static_cast<T*>(a.R[c]);
    al      = al && aligned16(a.R0[c]) && aligned16(a.R[c]);
  }
  p.n_layers           = a.n_layers;
  const int64_t n4     = al ? n / 4 : 0;
  const int64_t blocks = (n4 + 1 + 255) / 256;
  km_stacked_kernel<T><<<static_cast<unsigned>(blocks), 256, 0, ctx->stream>>>(p, n4, n);
  PB_CUDA(cudaGetLastError());
  ctx->launches++;
}

template <typename T>
void dry_t(pb_context* ctx, int64_t n, void* const planes[11]) {
  DryPtrs<T> p;
  bool al = true;
  for (int c = 0; c < kCanvasPlanes; ++c) {
    p.p[c] = static_cast<T*>(planes[c]);
    al     = al && aligned16(planes[c]);
  }
  const int64_t n4     = al ? n / 4 : 0;
  const int64_t blocks = (n4 + 1 + 255) / 256;
  km_dry_kernel<T><<<static_cast<unsigned>(blocks), 256, 0, ctx->stream>>>(p, n4, n);
  PB_CUDA(cudaGetLastError());
  ctx->launches++;
}

}  // namespace

void km_compose(pb_context* ctx, int64_t n, const ComposeArgs& a) {
  if (n <= 0) return;
  if (ctx->precision == PB_F64)
    compose_t<double>(ctx, n, a);
  else
    compose_t<float>(ctx, n, a);
}

void km_compose_display(pb_context* ctx, int64_t n, const ComposeArgs& a, int mode, bool srgb, void* out) {
  if (n <= 0) return;
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  if (ctx->precision == PB_F64) {
    ComposePtrs<double> p;
    for (int c = 0; c < 3; ++c) {
      p.K[c]  = static_cast<const double*>(a.K[c]);
      p.S[c]  = static_cast<const double*>(a.S[c]);
      p.R0[c] = static_cast<const double*>(a.R0[c]);
      p.R[c]  = nullptr;
    }
    p.V = static_cast<const double*>(a.V);
    km_compose_display_kernel<double><<<blocks, 256, 0, ctx->stream>>>(p, n, mode, srgb, out);
  } else {
    ComposePtrs<float> p;
    for (int c = 0; c < 3; ++c) {
      p.K[c]  = static_cast<const float*>(a.K[c]);
      p.S[c]  = static_cast<const float*>(a.S[c]);
      p.R0[c] = static_cast<const float*>(a.R0[c]);
      p.R[c]  = nullptr;
    }
    p.V = static_cast<const float*>(a.V);
    km_compose_display_kernel<float><<<blocks, 256, 0, ctx->stream>>>(p, n, mode, srgb, out);
  }
  PB_CUDA(cudaGetLastError());
  ctx->launches++;
}

template <typename T>
static void compose_gather_t(pb_context* ctx, int64_t n, const ComposeArgs& a, int n_dst, void* const (*dst)[3]) {
  ComposePtrs<T> p;
  GatherPtrs<T> g;
  bool aligned = true;
  auto al      = [&](const void* q) { aligned = aligned && (reinterpret_cast<uintptr_t>(q) % 16 == 0); };
  for (int c = 0; c < 3; ++c) {
    p.K[c]  = static_cast<const T*>(a.K[c]);
    p.S[c]  = static_cast<const T*>(a.S[c]);
    p.R0[c] = static_cast<const T*>(a.R0[c]);
    p.R[c]  = nullptr;
    al(a.K[c]), al(a.S[c]), al(a.R0[c]);
  }
  p.V = static_cast<const T*>(a.V);
  al(a.V);
  g.n_dst = n_dst;
  for (int d = 0; d < n_dst; ++d)
    for (int c = 0; c < 3; ++c) {
      g.dst[d][c] = static_cast<T*>(dst[d][c]);
      al(dst[d][c]);
    }
  if (aligned) {
    const int64_t n4 = n / 4;
    km_compose_gather_kernel<T, true><<<static_cast<unsigned>((n4 + 1 + 255) / 256), 256, 0, ctx->stream>>>(p, g, n4, n);
  } else {
    km_compose_gather_kernel<T, false><<<static_cast<unsigned>((n + 255) / 256), 256, 0, ctx->stream>>>(p, g, 0, n);
  }
  PB_CUDA(cudaGetLastError());
  ctx->launches++;
}
void km_compose_gather(pb_context* ctx, int64_t n, const ComposeArgs& a, int n_dst, void* const (*dst)[3]) {
  PB_REQUIRE(n_dst >= 1 && n_dst <= kGatherMax, "compose_gather: 1..8 destinations");
  if (n <= 0) return;
  if (ctx->precision == PB_F64)
    compose_gather_t<double>(ctx, n, a, n_dst, dst);
  else
    compose_gather_t<float>(ctx, n, a, n_dst, dst);
}

template <typename T>
static void compose_lab_t(pb_context* ctx, int64_t n, const ComposeArgs& a) {
  ComposePtrs<T> p;
  for (int c = 0; c < 3; ++c) {
    p.K[c]  = static_cast<const T*>(a.K[c]);
    p.S[c]  = static_cast<const T*>(a.S[c]);
    p.R0[c] = static_cast<const T*>(a.R0[c]);
    p.R[c]  = static_cast<T*>(a.R[c]);
  }
  p.V = static_cast<const T*>(a.V);
  km_compose_lab_kernel<T><<<static_cast<unsigned>((n + 255) / 256), 256, 0, ctx->stream>>>(p, n);
  PB_CUDA(cudaGetLastError());
  ctx->launches++;
}
void km_compose_lab(pb_context* ctx, int64_t n, const ComposeArgs& a) {
  if (n <= 0) return;
  if (ctx->precision == PB_F64)
    compose_lab_t<double>(ctx, n, a);
  else
    compose_lab_t<float>(ctx, n, a);
}
void lab_resize_lanczos4(pb_context* ctx, void* const lab[3], int rows, int cols, int orows, int ocols, const int* d_xofs,
                         const float* d_alpha, const int* d_yofs, const float* d_beta, double* d_tmp, double* d_out) {
  const int64_t nh = static_cast<int64_t>(rows) * ocols, nv = static_cast<int64_t>(orows) * ocols;
  if (ctx->precision == PB_F64)
    lanczos4_h_kernel<double><<<static_cast<unsigned>((nh + 255) / 256), 256, 0, ctx->stream>>>(
        static_cast<const double*>(lab[0]), static_cast<const double*>(lab[1]), static_cast<const double*>(lab[2]), rows, cols, ocols,
        d_xofs, d_alpha, d_tmp);
  else
    lanczos4_h_kernel<float><<<static_cast<unsigned>((nh + 255) / 256), 256, 0, ctx->stream>>>(
        static_cast<const float*>(lab[0]), static_cast<const float*>(lab[1]), static_cast<const float*>(lab[2]), rows, cols, ocols, d_xofs,
        d_alpha, d_tmp);
  PB_CUDA(cudaGetLastError());
  lanczos4_v_kernel<<<static_cast<unsigned>((nv + 255) / 256), 256, 0, ctx->stream>>>(d_tmp, rows, ocols, orows, d_yofs, d_beta, d_out);
  PB_CUDA(cudaGetLastError());
  ctx->launches += 2;
}
void lab_planes_to_aos(pb_context* ctx, void* const lab[3], int64_t n, double* d_out) {
  if (n <= 0) return;
  if (ctx->precision == PB_F64)
    planes_to_aos_f64_kernel<double><<<static_cast<unsigned>((n + 255) / 256), 256, 0, ctx->stream>>>(
        static_cast<const double*>(lab[0]), static_cast<const double*>(lab[1]), static_cast<const double*>(lab[2]), n, d_out);
  else
    planes_to_aos_f64_kernel<float><<<static_cast<unsigned>((n + 255) / 256), 256, 0, ctx->stream>>>(
        static_cast<const float*>(lab[0]), static_cast<const float*>(lab[1]), static_cast<const float*>(lab[2]), n, d_out);
  PB_CUDA(cudaGetLastError());
  ctx->launches++;
}

template <typename T>
static void render_t(pb_context* ctx, int rows, int cols, const ComposeArgs& a) {
  ComposePtrs<T> p;
  for (int c = 0; c < 3; ++c) {
    p.K[c]  = static_cast<const T*>(a.K[c]);
    p.S[c]  = static_cast<const T*>(a.S[c]);
    p.R0[c] = static_cast<const T*>(a.R0[c]);
    p.R[c]  = static_cast<T*>(a.R[c]);
  }
  p.V             = static_cast<const T*>(a.V);
  const int64_t n = static_cast<int64_t>(rows) * cols;
  km_render_kernel<T><<<static_cast<unsigned>((n + 255) / 256), 256, 0, ctx->stream>>>(p, rows, cols);
  PB_CUDA(cudaGetLastError());
  ctx->launches++;
}

void km_render(pb_context* ctx, int rows, int cols, const ComposeArgs& a) {
  if (rows <= 0 || cols <= 0) return;
  if (ctx->precision == PB_F64)
    render_t<double>(ctx, rows, cols, a);
  else
    render_t<float>(ctx, rows, cols, a);
}

void km_compose_stacked(pb_context* ctx, int64_t n, const StackArgs& a) {
  PB_REQUIRE(a.n_layers >= 1 && a.n_layers <= kMaxStack, "compose_stacked: 1..8 layers supported per pass");
  if (n <= 0) return;
  if (ctx->precision == PB_F64)
    stacked_t<double>(ctx, n, a);
  else
    stacked_t<float>(ctx, n, a);
}

void km_dry(pb_context* ctx, int64_t n, void* const planes[11]) {
  if (n <= 0) return;
  if (ctx->precision == PB_F64)
    dry_t<double>(ctx, n, planes);
  else
    dry_t<float>(ctx, n, planes);
}

}  // namespace pb
