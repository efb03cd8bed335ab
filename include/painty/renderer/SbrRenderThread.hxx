/**
 * Drop-in for painty/renderer/SbrRenderThread.hxx (reference lines 19-74, src/SbrRenderThread.cxx:14-98): the class
 * sbr_painter's PictureTargetSbrPainter drives (sbr/src/PictureTargetSbrPainter.cxx:33,169,226,307,335,355,362-397).
 * Same name, same constructor signature, same methods; put this repository's include/ first on the include path and the
 * planner compiles against it unchanged. SURVEY.md §8f #3.
 *
 * What stands behind it:
 *  - the GpuTaskQueue argument is accepted and ignored: there is no GL context to own, device work is ordered on the
 *    CUDA context's stream (GpuTaskQueue is only forward-declared here, a null pointer is fine);
 *  - render() records the stroke (dip -> setRadius -> paintStroke, src/SbrRenderThread.cxx:64-73); the recorded strokes are
 *    flushed as ONE batched device call by the next read-back / dryCanvas / destructor, so independent strokes run
 *    concurrently and overlapping ones keep their submission order. Futures are returned ready;
 *  - like TextureBrushGpu (src/TextureBrushGpu.cxx:238) every stroke samples the brush texture the dictionary picks for
 *    (path, 2 * radius) from <data>/textures (src/TextureBrushDictionary.cxx:25-79; the random draw among the candidates of
 *    the chosen class uses std::mt19937 seeded from std::random_device like the reference, or from
 *    SbrRenderOptions::seed for reproducible runs), and like CanvasGpu::clear (src/CanvasGpu.cxx:27-40) the substrate is
 *    <data>/canvas_patterns/0.png scaled to the canvas. Both are optional (SbrRenderOptions) — without them the brush
 *    samples data/sample_0 and the canvas is white, which is the CPU TextureBrush / Canvas behaviour;
 *  - numerics are those of the reference's CPU TextureBrush + Canvas::dryCanvas (what BASELINE.json pins), not of the GL
 *    shaders; there are no wall-clock timers (the GL path's periodic dryStep makes its output time dependent);
 *  - getLabImageScaled(rows, cols) is the planner's read-back prep done on the device
 *    (PictureTargetSbrPainter.cxx:334-341: ScaledMat(convertColor(getLinearRgbImage().get(), rgb_2_CIELab), rows, cols)):
 *    only the down-scaled Lab image crosses PCIe.
 */
#pragma once

#include <algorithm>
#include <array>
#include <cstdlib>
#include <filesystem>
#include <future>
#include <map>
#include <memory>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "painty/b200/Device.hxx"
#include "painty/core/Types.hxx"
#include "painty/core/Vec.hxx"
#include "painty/image/Mat.hxx"
#include "painty/io/ImageIO.hxx"
#include "painty/renderer/BrushStrokeSample.hxx"

namespace painty {

class GpuTaskQueue;  // painty/gpu/GpuTaskQueue.hxx (GL); only its shared_ptr appears in the constructor

struct SbrRenderOptions {
  bool useTextureDictionary = true;   // <dataDir>/textures/<size>_<lengthClass>_<nn>.png
  bool useCanvasPattern     = true;   // <dataDir>/canvas_patterns/0.png
  std::string dataDir       = "data"; // the reference hard-codes cwd-relative "data/..." paths
  std::string sampleDir     = "data/sample_0";
  uint64_t seed             = 0;      // 0: std::random_device like the reference
};

class SbrRenderThread final {
 public:
  SbrRenderThread(const std::shared_ptr<GpuTaskQueue>& /*gpuTaskQueue*/, const Size& canvasSize,
                  const SbrRenderOptions& options = SbrRenderOptions())
      : _canvasSize(canvasSize), _sample(options.sampleDir), _gen(options.seed ? options.seed : std::random_device{}()) {
    const int rows = static_cast<int>(canvasSize.height), cols = static_cast<int>(canvasSize.width);
    b200::check(pb_canvas_create(b200::context(), rows, cols, &_canvas));
    const Mat<double>& m = _sample.getThicknessMap();
    b200::check(pb_tbrush_create(b200::context(), m.rows, m.cols, reinterpret_cast<const double*>(m.data), &_brush));
    if (options.useCanvasPattern) loadCanvasPattern(options.dataDir + "/canvas_patterns/0.png");
    if (options.useTextureDictionary) loadDictionary(options.dataDir + "/textures");
  }
  SbrRenderThread(const SbrRenderThread&) = delete;
  SbrRenderThread& operator=(const SbrRenderThread&) = delete;
  ~SbrRenderThread() {
    try {
      flush();
    } catch (...) {
    }
    if (_dict) pb_texdict_destroy(_dict);
    if (_brush) pb_tbrush_destroy(_brush);
    if (_canvas) pb_canvas_destroy(_canvas);
  }

  auto getSize() const -> Size { return _canvasSize; }
  auto getBrushThicknessScale() const -> double { return _thicknessScale; }

  /** Records dip(ks); setRadius(radius); paintStroke(path) (SbrRenderThread.cxx:64-73). */
  auto render(const std::vector<vec2>& path, const double radius, const std::array<vec3, 2UL>& ks) -> std::future<void> {
    pb_tstroke s{};
    s.radius = radius;
    for (size_t i = 0; i < 3; ++i) {
      s.K[i] = ks[0U][i];
      s.S[i] = ks[1U][i];
    }
    s.thickness_scale = _thicknessScale;
    s.first_vertex    = static_cast<int64_t>(_xy.size() / 2);
    s.n_vertices      = static_cast<int32_t>(path.size());
    const size_t at   = _xy.size();
    for (const auto& p : path) {
      _xy.push_back(p[0U]);
      _xy.push_back(p[1U]);
    }
    s.texture_id = 0;
    if (_dict && path.size() >= 2U) {  // TextureBrushGpu.cxx:238: lookup(vertices, 2.0 * _radius)
      int32_t n = 0;
      b200::check(pb_texdict_lookup(_dict, static_cast<int>(path.size()), _xy.data() + at, 2.0 * radius, nullptr, nullptr,
                                    static_cast<int>(_candidates.size()), _candidates.data(), &n));
      std::uniform_int_distribution<std::size_t> dis(0UL, static_cast<std::size_t>(n) - 1UL);  // TextureBrushDictionary.cxx:61-66
      s.texture_id = _textureIds[static_cast<size_t>(_candidates[dis(_gen)])];
    }
    _pending.push_back(s);
    return ready();
  }

  /** Flushes the recorded strokes and composes (CanvasGpu::getCompositionLinearRgb's role, SbrRenderThread.cxx:75-80). */
  auto getLinearRgbImage() -> std::future<Mat3d> {
    flush();
    Mat3d rgb(static_cast<int>(_canvasSize.height), static_cast<int>(_canvasSize.width));
    b200::check(pb_canvas_compose(_canvas, reinterpret_cast<double*>(rgb.data)));
    std::promise<Mat3d> p;
    p.set_value(rgb);
    return p.get_future();
  }

  /** Device-side read-back prep of the planner: CIELab of the composition, LANCZOS4-scaled to rows x cols. */
  auto getLabImageScaled(const int32_t rows, const int32_t cols) -> std::future<Mat3d> {
    flush();
    Mat3d lab(rows, cols);
    b200::check(pb_canvas_compose_lab_scaled(_canvas, rows, cols, reinterpret_cast<double*>(lab.data)));
    std::promise<Mat3d> p;
    p.set_value(lab);
    return p.get_future();
  }

  void setBrushThicknessScale(const double scale) { _thicknessScale = scale; }

  void enableSmudge(bool enable) {
    flush();  // the switch applies to the strokes submitted after it
    b200::check(pb_tbrush_enable_smudge(_brush, enable ? 1 : 0));
  }

  /** dryStep(1.0): everything wet is composed into the substrate (SbrRenderThread.cxx:94-98, Canvas::dryCanvas). */
  auto dryCanvas() -> std::future<void> {
    flush();
    b200::check(pb_canvas_dry(_canvas));
    return ready();
  }

  /** Device canvas handle (e.g. for pb_canvas_compose_qrgb32 previews). */
  pb_canvas* canvas() {
    flush();
    return _canvas;
  }
  /** Number of brush textures in the device atlas (0 without a dictionary). */
  auto getTextureCount() const -> size_t { return _textureIds.size(); }

 private:
  static std::future<void> ready() {
    std::promise<void> p;
    p.set_value();
    return p.get_future();
  }
  void flush() {
    if (_pending.empty()) return;
    b200::check(pb_tbrush_stroke_batch(_brush, _canvas, static_cast<int64_t>(_pending.size()), _pending.data(),
                                       static_cast<int64_t>(_xy.size() / 2), _xy.data()));
    _pending.clear();
    _xy.clear();
  }

  // CanvasGpu::clear (src/CanvasGpu.cxx:27-40): linear RGB pattern, float32, LANCZOS4-scaled to the canvas, as R0.
  void loadCanvasPattern(const std::string& file) {
    Mat3d pattern = {};
    io::imRead(file, pattern, true);
    Mat4f patternF(pattern.size());
    for (auto i = 0; i < static_cast<int32_t>(pattern.total()); i++) {
      patternF(i) = {static_cast<float>(pattern(i)[0U]), static_cast<float>(pattern(i)[1U]), static_cast<float>(pattern(i)[2U]), 1.0F};
    }
    patternF = ScaledMat(patternF, _canvasSize);
    std::vector<double> r0(static_cast<size_t>(patternF.total()) * 3U);
    for (auto i = 0; i < static_cast<int32_t>(patternF.total()); i++) {
      for (size_t c = 0; c < 3U; ++c) r0[3U * static_cast<size_t>(i) + c] = static_cast<double>(patternF(i)[c]);
    }
    b200::check(pb_canvas_set_background(_canvas, r0.data()));
  }

  // TextureBrushDictionary::createBrushTexturesFromFolder (src/TextureBrushDictionary.cxx:81-118) + loadHeightMap (:71-79)
  void loadDictionary(const std::string& folder) {
    auto split = [](const std::string& input, const char delim) {
      std::vector<std::string> elems;
      std::stringstream ss(input);
      std::string item;
      while (std::getline(ss, item, delim)) elems.push_back(item);
      return elems;
    };
    std::vector<std::string> files;
    for (const auto& p : std::filesystem::directory_iterator(folder)) files.push_back(p.path().string());
    std::sort(files.begin(), files.end());  // directory order is unspecified; fixed here so that seeded runs repeat
    std::vector<int32_t> sizeKey, lengthKey, rows, cols;
    for (const auto& filepath : files) {
      const std::string filename      = split(split(filepath, '.').front(), '/').back();
      const std::vector<std::string> t = split(filename, '_');
      Mat1d gray;
      io::imRead(filepath, gray, false);
      // cv::normalize(gray, gray, 0.0, 1.0, cv::NORM_MINMAX): dst = src * scale + shift, scale = 1 / (max - min)
      double lo = gray(0), hi = gray(0);
      for (const auto& v : gray) {
        lo = std::min(lo, v);
        hi = std::max(hi, v);
      }
      const double scale = (hi - lo) > 2.220446049250313e-16 ? 1. / (hi - lo) : 0.;
      const double shift = 0.0 - lo * scale;
      for (auto& v : gray) v = v * scale + shift;
      int id = 0;
      b200::check(pb_tbrush_add_texture(_brush, gray.rows, gray.cols, reinterpret_cast<const double*>(gray.data), &id));
      _textureIds.push_back(id);
      sizeKey.push_back(std::stoi(t[0]));
      lengthKey.push_back(std::stoi(t[1]));
      rows.push_back(gray.rows);
      cols.push_back(gray.cols);
    }
    if (files.empty()) return;
    b200::check(pb_texdict_create(static_cast<int>(files.size()), sizeKey.data(), lengthKey.data(), rows.data(), cols.data(), &_dict));
    _candidates.resize(files.size());
  }

  Size _canvasSize;
  BrushStrokeSample _sample;
  pb_canvas* _canvas  = nullptr;
  pb_tbrush* _brush   = nullptr;
  pb_texdict* _dict   = nullptr;
  double _thicknessScale = 1.0;
  std::vector<pb_tstroke> _pending;
  std::vector<double> _xy;
  std::vector<int32_t> _textureIds, _candidates;
  std::mt19937 _gen;
};

}  // namespace painty
