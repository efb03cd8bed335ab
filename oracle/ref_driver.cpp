// TEST INFRASTRUCTURE ONLY — never linked into or called from the product path.
//
// oracle/_ref/libpainty_ref.so: the *unmodified* reference hot-path headers from
// /root/reference (KubelkaMunk.hxx, Math.hxx, Spline.hxx, image/Mat.hxx,
// renderer/{PaintLayer,Canvas,FootprintBrush,TextureBrush,Renderer,Smudge,BrushStrokeSample}.hxx,
// renderer/src/BrushStrokeSample.cxx, image/src/TextureWarp.cxx) compiled where they lie against
// the stand-ins in oracle/shim/ (Eigen/Dense, opencv2/imgproc.hpp), behind a flat C API so that
// tests and bench.py's cpu_baseline can drive the real reference code through ctypes.
// Built by oracle/Makefile with `-O3 -ffp-contract=off` and no -march (the reference's CI build has
// no FMA contraction). Nothing from /root/reference is copied: the
// compiler reads the headers in place via -I/root/reference.
//
// Substitutions (both are *inputs* of the hot path):
//   * painty::io::imRead(path, Mat<double>&, bool)  -> returns the f64 image registered under
//     `path` by ref_register_image (baked with Python cv2, oracle/bake_assets.py);
//   * cv::resize                                     -> returns the f64 image registered for that
//     output size by ref_register_resize (cv2.resize INTER_LANCZOS4 on the f64 input).
#include <algorithm>
#include <array>
#include <chrono>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "painty/core/KubelkaMunk.hxx"
#include "painty/core/Spline.hxx"
#include "painty/image/Mat.hxx"
#include "painty/io/ImageIO.hxx"
#include "painty/renderer/Canvas.hxx"
#include "painty/renderer/FootprintBrush.hxx"
#include "painty/renderer/PaintLayer.hxx"
#include "painty/renderer/Renderer.hxx"
#include "painty/renderer/TextureBrush.hxx"
// out-of-line parts of the reference that the headers above need at link time
#include "painty/image/src/TextureWarp.cxx"
#include "painty/renderer/src/BrushStrokeSample.cxx"

namespace {
struct Image {
  int rows, cols;
  std::vector<double> data;
};
std::map<std::string, Image>& images() {
  static std::map<std::string, Image> m;
  return m;
}
std::map<std::pair<int, int>, Image>& resized() {
  static std::map<std::pair<int, int>, Image> m;
  return m;
}
using Canvas    = painty::Canvas<painty::vec3>;
using Layer     = painty::PaintLayer<painty::vec3>;
using FBrush    = painty::FootprintBrush<painty::vec3>;
using TBrush    = painty::TextureBrush<painty::vec3>;
using painty::vec2;
using painty::vec3;

void layer_to_aos(const Layer& l, double* K, double* S, double* V) {
  const auto n = static_cast<int32_t>(l.getK_buffer().total());
  for (int32_t i = 0; i < n; ++i) {
    for (int c = 0; c < 3; ++c) {
      if (K) K[3 * i + c] = l.getK_buffer()(i)[static_cast<size_t>(c)];
      if (S) S[3 * i + c] = l.getS_buffer()(i)[static_cast<size_t>(c)];
    }
    if (V) V[i] = l.getV_buffer()(i);
  }
}
void aos_to_layer(Layer& l, const double* K, const double* S, const double* V) {
  const auto n = static_cast<int32_t>(l.getK_buffer().total());
  for (int32_t i = 0; i < n; ++i) {
    for (int c = 0; c < 3; ++c) {
      l.getK_buffer()(i)[static_cast<size_t>(c)] = K[3 * i + c];
      l.getS_buffer()(i)[static_cast<size_t>(c)] = S[3 * i + c];
    }
    l.getV_buffer()(i) = V[i];
  }
}
}  // namespace

// ---- substituted inputs -------------------------------------------------------------------------
extern "C" bool oracle_shim_resize_f64(int, int, int out_rows, int out_cols, double* out) {
  auto it = resized().find({out_rows, out_cols});
  if (it == resized().end()) return false;
  std::memcpy(out, it->second.data.data(), sizeof(double) * it->second.data.size());
  return true;
}

void painty::io::imRead(const std::string& filename, Mat<double>& gray, bool) {
  auto it = images().find(filename);
  if (it == images().end()) {
    throw std::ios_base::failure(filename);
  }
  gray = Mat<double>(it->second.rows, it->second.cols);
  std::memcpy(gray.data, it->second.data.data(), sizeof(double) * it->second.data.size());
}

extern "C" {

void ref_register_image(const char* path, int rows, int cols, const double* data) {
  images()[path] = Image{rows, cols, std::vector<double>(data, data + static_cast<size_t>(rows) * cols)};
}
void ref_register_resize(int out_rows, int out_cols, const double* data) {
  resized()[{out_rows, out_cols}] =
    Image{out_rows, out_cols, std::vector<double>(data, data + static_cast<size_t>(out_rows) * out_cols)};
}

// ---- scalar functions ---------------------------------------------------------------------------
void ref_compute_reflectance(const double* K, const double* S, const double* R0, double d, double* out) {
  const vec3 r = painty::ComputeReflectance<double, 3>(vec3(K[0], K[1], K[2]), vec3(S[0], S[1], S[2]),
                                                       vec3(R0[0], R0[1], R0[2]), d);
  out[0] = r[0];
  out[1] = r[1];
  out[2] = r[2];
}
void ref_compute_reflectance_f32(const float* K, const float* S, const float* R0, float d, float* out) {
  using v3f     = painty::vec<float, 3>;
  const v3f r = painty::ComputeReflectance<float, 3>(v3f(K[0], K[1], K[2]), v3f(S[0], S[1], S[2]),
                                                     v3f(R0[0], R0[1], R0[2]), d);
  out[0] = r[0];
  out[1] = r[1];
  out[2] = r[2];
}
double ref_coth(double x) { return painty::coth(x); }
double ref_acoth(double x) { return painty::acoth(x); }
// returns 0 on success, 1 when the reference throws std::invalid_argument
int ref_compute_scattering_absorption(const double* Rb, const double* Rw, double* K, double* S) {
  vec3 k, s;
  try {
    painty::ComputeScatteringAndAbsorption<double, 3>(vec3(Rb[0], Rb[1], Rb[2]), vec3(Rw[0], Rw[1], Rw[2]), k, s);
  } catch (const std::invalid_argument&) {
    return 1;
  }
  for (size_t c = 0; c < 3; ++c) {
    K[c] = k[c];
    S[c] = s[c];
  }
  return 0;
}
void ref_catmull_rom(const double* p_1, const double* p0, const double* p1, const double* p2, double t, double* out) {
  const vec2 r = painty::CatmullRom(vec2(p_1[0], p_1[1]), vec2(p0[0], p0[1]), vec2(p1[0], p1[1]), vec2(p2[0], p2[1]), t);
  out[0] = r[0];
  out[1] = r[1];
}
void ref_catmull_rom_d1(const double* p_1, const double* p0, const double* p1, const double* p2, double t, double* out) {
  const vec2 r =
    painty::CatmullRomDerivativeFirst(vec2(p_1[0], p_1[1]), vec2(p0[0], p0[1]), vec2(p1[0], p1[1]), vec2(p2[0], p2[1]), t);
  out[0] = r[0];
  out[1] = r[1];
}
double ref_catmull_rom_scalar(double p_1, double p0, double p1, double p2, double t) {
  return painty::CatmullRom(p_1, p0, p1, p2, t);
}
double ref_cubic_scalar(double p_1, double p0, double p1, double p2, double t) {
  return painty::Cubic(p_1, p0, p1, p2, t);
}
// kind: 0 catmullRom, 1 catmullRomDerivativeFirst, 2 cubic
void ref_spline_eval(int n, const double* xy, double u, int kind, double* out) {
  std::vector<vec2> pts;
  for (int i = 0; i < n; ++i) pts.emplace_back(xy[2 * i], xy[2 * i + 1]);
  painty::SplineEval<std::vector<vec2>::const_iterator> sp(pts.cbegin(), pts.cend());
  const vec2 r = kind == 0 ? sp.catmullRom(u) : (kind == 1 ? sp.catmullRomDerivativeFirst(u) : sp.cubic(u));
  out[0] = r[0];
  out[1] = r[1];
}
// returns 0 ok, 1 invalid_argument
int ref_mvc_interpolate(int n, const double* polygon_xy, int nv, const double* values_uv, double px, double py,
                        double* out) {
  std::vector<vec2> poly, vals;
  for (int i = 0; i < n; ++i) poly.emplace_back(polygon_xy[2 * i], polygon_xy[2 * i + 1]);
  for (int i = 0; i < nv; ++i) vals.emplace_back(values_uv[2 * i], values_uv[2 * i + 1]);
  try {
    const vec2 r = painty::generalizedBarycentricCoordinatesInterpolate(poly, vec2(px, py), vals);
    out[0] = r[0];
    out[1] = r[1];
  } catch (const std::invalid_argument&) {
    return 1;
  }
  return 0;
}
double ref_interpolate_bilinear(const double* data, int rows, int cols, double x, double y) {
  painty::Mat<double> m(rows, cols);
  std::memcpy(m.data, data, sizeof(double) * static_cast<size_t>(rows) * cols);
  return painty::Interpolate(m, vec2(x, y));
}

// ---- whole-image compose (Renderer.hxx:26-41, PaintLayer.hxx:81-96) -----------------------------
// All image arguments are AoS f64: K,S,R0,out = rows*cols*3, V = rows*cols.
void ref_compose(int rows, int cols, const double* K, const double* S, const double* V, const double* R0, double* out) {
  Layer layer(rows, cols);
  aos_to_layer(layer, K, S, V);
  painty::Mat<vec3> r0(rows, cols);
  for (int32_t i = 0; i < rows * cols; ++i) r0(i) = vec3(R0[3 * i], R0[3 * i + 1], R0[3 * i + 2]);
  const painty::Mat<vec3> r1 = painty::Renderer<vec3>().compose(layer, r0);
  for (int32_t i = 0; i < rows * cols; ++i)
    for (size_t c = 0; c < 3; ++c) out[3 * i + static_cast<int32_t>(c)] = r1(i)[c];
}
// in place: R0 <- KM(K,S,R0,V)  (the stacked-layer unit)
void ref_compose_onto(int rows, int cols, const double* K, const double* S, const double* V, double* R0) {
  Layer layer(rows, cols);
  aos_to_layer(layer, K, S, V);
  painty::Mat<vec3> r0(rows, cols);
  for (int32_t i = 0; i < rows * cols; ++i) r0(i) = vec3(R0[3 * i], R0[3 * i + 1], R0[3 * i + 2]);
  layer.composeOnto(r0);
  for (int32_t i = 0; i < rows * cols; ++i)
    for (size_t c = 0; c < 3; ++c) R0[3 * i + static_cast<int32_t>(c)] = r0(i)[c];
}
// Timed kernel for the CPU baseline: the reference's per-pixel function over `n` pixels, row-split over
// `threads` std::threads (the reference itself is single threaded; threads>1 is our fair ceiling).
// Returns seconds. Layout AoS f64 like above.
double ref_compose_timed(int64_t n, const double* K, const double* S, const double* V, const double* R0, double* out,
                         int threads) {
  auto work = [&](int64_t b, int64_t e) {
    for (int64_t i = b; i < e; ++i) {
      const vec3 r = painty::ComputeReflectance<double, 3>(vec3(K[3 * i], K[3 * i + 1], K[3 * i + 2]),
                                                           vec3(S[3 * i], S[3 * i + 1], S[3 * i + 2]),
                                                           vec3(R0[3 * i], R0[3 * i + 1], R0[3 * i + 2]), V[i]);
      out[3 * i]     = r[0];
      out[3 * i + 1] = r[1];
      out[3 * i + 2] = r[2];
    }
  };
  const auto t0 = std::chrono::steady_clock::now();
  if (threads <= 1) {
    work(0, n);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(work, n * t / threads, n * (t + 1) / threads);
    for (auto& th : pool) th.join();
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ---- display epilogue: the reference's ColorConverter::rgb2srgb + the GUI's qRgb cast ---------------
// (apps/painty_gui/DigitalCanvas.cxx:164-177; clamped outside [0,255] where the reference's cast is UB)
void ref_qrgb32(int64_t n, const double* rgb, uint32_t* out) {
  painty::ColorConverter<double> converter;
  for (int64_t i = 0; i < n; ++i) {
    vec3 v(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
    converter.rgb2srgb(v, v);
    uint32_t q[3];
    for (size_t c = 0; c < 3; ++c) {
      const double s = v[c] * 255.0;
      q[c]           = s >= 255.0 ? 255u : (s > 0.0 ? static_cast<uint32_t>(static_cast<uint8_t>(s)) : 0u);
    }
    out[i] = 0xff000000u | (q[0] << 16) | (q[1] << 8) | q[2];
  }
}

// The planner's read-back prep, colour part: convertColor(rgb, rgb_2_CIELab) = ColorConverter<double>::rgb2lab per pixel
// (painty/sbr/src/PictureTargetSbrPainter.cxx:338-340, painty/core/Color.hxx:248-252); AoS f64 in and out.
void ref_rgb2lab(int64_t n, const double* rgb, double* lab) {
  painty::ColorConverter<double> converter;
  for (int64_t i = 0; i < n; ++i) {
    vec3 v(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]), o;
    converter.rgb2lab(v, o);
    for (size_t c = 0; c < 3; ++c) lab[3 * i + static_cast<int64_t>(c)] = o[c];
  }
}

// ---- Canvas -------------------------------------------------------------------------------------
void* ref_canvas_create(int rows, int cols) {
  auto* c = new Canvas(rows, cols);
  c->setDryingTime(std::chrono::milliseconds(0));  // disable the wall-clock branch
  return c;
}
void ref_canvas_destroy(void* c) { delete static_cast<Canvas*>(c); }
void ref_canvas_clear(void* c) { static_cast<Canvas*>(c)->clear(); }
void ref_canvas_set_background(void* cv, const double* R0) {
  auto* c = static_cast<Canvas*>(cv);
  painty::Mat<vec3> bg(c->getPaintLayer().getRows(), c->getPaintLayer().getCols());
  for (int32_t i = 0; i < static_cast<int32_t>(bg.total()); ++i) bg(i) = vec3(R0[3 * i], R0[3 * i + 1], R0[3 * i + 2]);
  c->setBackground(bg);
}
void ref_canvas_dry(void* c) { static_cast<Canvas*>(c)->dryCanvas(); }
void ref_canvas_set_layer(void* cv, const double* K, const double* S, const double* V) {
  aos_to_layer(static_cast<Canvas*>(cv)->getPaintLayer(), K, S, V);
}
// any pointer may be null
void ref_canvas_get(void* cv, double* K, double* S, double* V, double* R0, double* h) {
  auto* c = static_cast<Canvas*>(cv);
  layer_to_aos(c->getPaintLayer(), K, S, V);
  const auto n = static_cast<int32_t>(c->getR0().total());
  for (int32_t i = 0; i < n; ++i) {
    if (R0)
      for (size_t ch = 0; ch < 3; ++ch) R0[3 * i + static_cast<int32_t>(ch)] = c->getR0()(i)[ch];
    if (h) h[i] = c->get_h()(i);
  }
}
void ref_canvas_compose(void* cv, double* out) {
  auto* c                    = static_cast<Canvas*>(cv);
  const painty::Mat<vec3> r1 = painty::Renderer<vec3>().compose(*c);
  for (int32_t i = 0; i < static_cast<int32_t>(r1.total()); ++i)
    for (size_t ch = 0; ch < 3; ++ch) out[3 * i + static_cast<int32_t>(ch)] = r1(i)[ch];
}

void ref_canvas_render(void* cv, double* out) {  // Renderer::render (Renderer.hxx:60-156)
  auto* c                    = static_cast<Canvas*>(cv);
  const painty::Mat<vec3> r1 = painty::Renderer<vec3>().render(*c);
  for (int32_t i = 0; i < static_cast<int32_t>(r1.total()); ++i)
    for (size_t ch = 0; ch < 3; ++ch) out[3 * i + static_cast<int32_t>(ch)] = r1(i)[ch];
}

// ---- FootprintBrush -----------------------------------------------------------------------------
// The footprint for ceil(radius) must have been registered with ref_register_resize(width,width)
// and "./data/footprint/footprint.png" with ref_register_image (any content; it only feeds resize).
void* ref_fbrush_create(double radius) {
  try {
    return new FBrush(radius);
  } catch (...) {
    return nullptr;
  }
}
void ref_fbrush_destroy(void* b) { delete static_cast<FBrush*>(b); }
int ref_fbrush_set_radius(void* b, double radius) {
  try {
    static_cast<FBrush*>(b)->setRadius(radius);
  } catch (...) {
    return 1;
  }
  return 0;
}
void ref_fbrush_dip(void* b, const double* K, const double* S) {
  static_cast<FBrush*>(b)->dip({vec3(K[0], K[1], K[2]), vec3(S[0], S[1], S[2])});
}
void ref_fbrush_set_rates(void* b, double pickup, double deposition) {
  static_cast<FBrush*>(b)->setPickupRate(pickup);
  static_cast<FBrush*>(b)->setDepositionRate(deposition);
}
void ref_fbrush_set_use_snapshot(void* b, int use) { static_cast<FBrush*>(b)->setUseSnapshotBuffer(use != 0); }
int ref_fbrush_size_map(void* b) { return static_cast<FBrush*>(b)->getPickupMap().getRows(); }
int ref_fbrush_footprint_size(void* b) { return static_cast<FBrush*>(b)->getFootprint().rows; }
void ref_fbrush_get_footprint(void* b, double* out) {
  const auto& f = static_cast<FBrush*>(b)->getFootprint();
  std::memcpy(out, f.data, sizeof(double) * f.total());
}
void ref_fbrush_get_pickup_map(void* b, double* K, double* S, double* V) {
  layer_to_aos(static_cast<FBrush*>(b)->getPickupMap(), K, S, V);
}
// n imprints (cx, cy, theta) applied in order (FootprintBrush.hxx:73-143). Returns seconds spent.
double ref_fbrush_imprint_batch(void* b, void* canvas, int n, const double* cx, const double* cy, const double* theta) {
  auto* br      = static_cast<FBrush*>(b);
  auto* c       = static_cast<Canvas*>(canvas);
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < n; ++i) br->imprint(vec2(cx[i], cy[i]), theta[i], *c);
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ---- TextureBrush -------------------------------------------------------------------------------
// `<sample_dir>/thickness_map.png` must have been registered with ref_register_image.
void* ref_tbrush_create(const char* sample_dir, int enable_smudge) {
  try {
    auto* b = new TBrush(sample_dir);
    b->enableSmudge(enable_smudge != 0);
    return b;
  } catch (...) {
    return nullptr;
  }
}
void ref_tbrush_destroy(void* b) { delete static_cast<TBrush*>(b); }
void ref_tbrush_set_radius(void* b, double r) { static_cast<TBrush*>(b)->setRadius(r); }
void ref_tbrush_enable_smudge(void* b, int enable) { static_cast<TBrush*>(b)->enableSmudge(enable != 0); }
void ref_tbrush_dip(void* b, const double* K, const double* S) {
  static_cast<TBrush*>(b)->dip({vec3(K[0], K[1], K[2]), vec3(S[0], S[1], S[2])});
}
void ref_tbrush_set_thickness_scale(void* b, double s) { static_cast<TBrush*>(b)->setThicknessScale(s); }
double ref_tbrush_paint_stroke(void* b, void* canvas, int n, const double* xy) {
  std::vector<vec2> path;
  for (int i = 0; i < n; ++i) path.emplace_back(xy[2 * i], xy[2 * i + 1]);
  const auto t0 = std::chrono::steady_clock::now();
  static_cast<TBrush*>(b)->paintStroke(path, *static_cast<Canvas*>(canvas));
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
// BrushStrokeSample::getSampleAt on the registered thickness map (BrushStrokeSampleTest.cxx:14-21)
double ref_stroke_sample_at(const char* sample_dir, double x, double y) {
  painty::BrushStrokeSample s(sample_dir);
  return s.getSampleAt(vec2(x, y));
}
void ref_stroke_sample_dims(const char* sample_dir, int* rows, int* cols) {
  painty::BrushStrokeSample s(sample_dir);
  *rows = s.getThicknessMap().rows;
  *cols = s.getThicknessMap().cols;
}

}  // extern "C"
