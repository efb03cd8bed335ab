/**
 * Round-1 name of the CUDA render thread, kept for source compatibility: the drop-in now carries the reference's own
 * name and constructor (include/painty/renderer/SbrRenderThread.hxx). SbrRenderThreadCuda(size[, sampleDir]) is that class
 * without a GpuTaskQueue argument, a white canvas and the single data/sample_0 brush texture (the CPU TextureBrush setup).
 */
#pragma once

#include "painty/renderer/SbrRenderThread.hxx"

namespace painty {

class SbrRenderThreadCuda final {
 public:
  explicit SbrRenderThreadCuda(const Size& canvasSize, const std::string& sampleDir = "data/sample_0")
      : _impl(nullptr, canvasSize, plain(sampleDir)) {}
  auto getSize() const -> Size { return _impl.getSize(); }
  auto getBrushThicknessScale() const -> double { return _impl.getBrushThicknessScale(); }
  auto render(const std::vector<vec2>& path, const double radius, const std::array<vec3, 2UL>& ks) -> std::future<void> {
    return _impl.render(path, radius, ks);
  }
  auto getLinearRgbImage() -> std::future<Mat3d> { return _impl.getLinearRgbImage(); }
  auto getLabImageScaled(const int32_t rows, const int32_t cols) -> std::future<Mat3d> { return _impl.getLabImageScaled(rows, cols); }
  void setBrushThicknessScale(const double scale) { _impl.setBrushThicknessScale(scale); }
  void enableSmudge(bool enable) { _impl.enableSmudge(enable); }
  auto dryCanvas() -> std::future<void> { return _impl.dryCanvas(); }
  pb_canvas* canvas() { return _impl.canvas(); }

 private:
  static SbrRenderOptions plain(const std::string& sampleDir) {
    SbrRenderOptions o;
    o.useTextureDictionary = false;
    o.useCanvasPattern     = false;
    o.sampleDir            = sampleDir;
    return o;
  }
  SbrRenderThread _impl;
};

}  // namespace painty
