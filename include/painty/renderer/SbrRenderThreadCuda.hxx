/**
 * CUDA back end with the public interface of painty::SbrRenderThread (reference renderer/SbrRenderThread.hxx:19-74,
 * src/SbrRenderThread.cxx:64-98) — the class sbr_painter's PictureTargetSbrPainter drives. SURVEY.md §8f #3.
 *
 * The reference marshals every call onto one GL worker thread and answers with std::future; render() is fire and
 * forget, read-backs block. Here render() only records the stroke (dip -> setRadius -> paintStroke, in submission
 * order); the recorded strokes are flushed as ONE batched device call by getLinearRgbImage() / dryCanvas() / the
 * destructor, which lets independent strokes run concurrently while overlapping ones keep their order. Futures are
 * returned ready: the work is stream ordered behind them.
 *
 * Semantics are those of the reference's CPU TextureBrush (what BASELINE.json pins), not of the GL shaders; there are no
 * wall-clock timers (the GL path's periodic dryStep makes its output time dependent).
 * The constructor takes no GpuTaskQueue (there is no GL context to own).
 */
#pragma once

#include <array>
#include <future>
#include <memory>
#include <string>
#include <vector>

#include "painty/b200/Device.hxx"
#include "painty/core/Types.hxx"
#include "painty/core/Vec.hxx"
#include "painty/image/Mat.hxx"
#include "painty/renderer/BrushStrokeSample.hxx"

namespace painty {

class SbrRenderThreadCuda final {
 public:
  SbrRenderThreadCuda(const Size& canvasSize, const std::string& sampleDir = "data/sample_0")
      : _canvasSize(canvasSize), _sample(sampleDir) {
    b200::check(pb_canvas_create(b200::context(), static_cast<int>(canvasSize.height), static_cast<int>(canvasSize.width), &_canvas));
    const Mat<double>& m = _sample.getThicknessMap();
    b200::check(pb_tbrush_create(b200::context(), m.rows, m.cols, m.data, &_brush));
  }
  SbrRenderThreadCuda(const SbrRenderThreadCuda&) = delete;
  SbrRenderThreadCuda& operator=(const SbrRenderThreadCuda&) = delete;
  ~SbrRenderThreadCuda() {
    try {
      flush();
    } catch (...) {
    }
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
    for (const auto& p : path) {
      _xy.push_back(p[0U]);
      _xy.push_back(p[1U]);
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

  Size _canvasSize;
  BrushStrokeSample _sample;
  pb_canvas* _canvas = nullptr;
  pb_tbrush* _brush  = nullptr;
  double _thicknessScale = 1.0;
  std::vector<pb_tstroke> _pending;
  std::vector<double> _xy;
};

}  // namespace painty
