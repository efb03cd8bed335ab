/**
 * Drop-in for painty/renderer/TextureBrush.hxx (reference lines 19-237). The stroke-texture sample is still loaded
 * by painty's own BrushStrokeSample (host); the warp + sample + deposit and the Smudge walk (renderer/Smudge.hxx,
 * default-on like the reference, off in sbr_painter's config) run on the device.
 */
#pragma once

#include <array>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "painty/b200/Device.hxx"
#include "painty/renderer/BrushBase.hxx"
#include "painty/renderer/BrushStrokeSample.hxx"
#include "painty/renderer/Canvas.hxx"

namespace painty {
template <class vector_type>
class TextureBrush final : public BrushBase<vector_type> {
  using T                 = typename BrushBase<vector_type>::T;
  static constexpr auto N = BrushBase<vector_type>::N;

  struct Handle {
    pb_tbrush* h = nullptr;
    ~Handle() {
      if (h) pb_tbrush_destroy(h);
    }
  };

 public:
  TextureBrush(const std::string& sampleDir) : _brushStrokeSample(sampleDir), _h(std::make_shared<Handle>()) {
    const Mat<double>& m = _brushStrokeSample.getThicknessMap();
    b200::check(pb_tbrush_create(b200::context(), m.rows, m.cols, m.data, &_h->h));
    b200::check(pb_tbrush_enable_smudge(_h->h, 1));  // TextureBrush.hxx:236: _useSmudge = true
  }

  void setRadius(const double radius) override { b200::check(pb_tbrush_set_radius(_h->h, radius)); }  // reference :33-41

  void dip(const std::array<vector_type, 2UL>& paint) override {  // reference :48-50
    const double K[3] = {paint[0U][0U], paint[0U][1U], paint[0U][2U]}, S[3] = {paint[1U][0U], paint[1U][1U], paint[1U][2U]};
    pb_tbrush_dip(_h->h, K, S);
  }

  void paintStroke(const std::vector<vec2>& verticesArg, Canvas<vector_type>& canvas) override {  // reference :52-205
    if (verticesArg.size() < 2UL) return;
    std::vector<double> xy(2 * verticesArg.size());
    for (size_t i = 0; i < verticesArg.size(); ++i) xy[2 * i] = verticesArg[i][0U], xy[2 * i + 1] = verticesArg[i][1U];
    b200::check(pb_tbrush_set_thickness_scale(_h->h, BrushBase<vector_type>::getThicknessScale()));
    b200::check(pb_tbrush_paint_stroke(_h->h, canvas.device(), static_cast<int>(verticesArg.size()), xy.data()));
    canvas.deviceWritten();
  }

  void enableSmudge(const bool enable) { b200::check(pb_tbrush_enable_smudge(_h->h, enable ? 1 : 0)); }

 private:
  BrushStrokeSample _brushStrokeSample;
  std::shared_ptr<Handle> _h;
};
}  // namespace painty
