/**
 * Drop-in for painty/renderer/FootprintBrush.hxx (reference lines 22-503). The footprint image is still loaded,
 * LANCZOS-scaled and padded by painty's own host code (io::imRead, ScaledMat, PaddedMat — reference :49-58); the
 * pickup map, snapshot buffer and the imprint itself live on the device (painty_b200 imprint engine).
 *
 * Differences, all documented in DESIGN.md §2: paintStroke uses p_pre = path[0] on the first segment (the
 * reference reads path[-1], UB); out-of-range footprint reads are height 0; _paintIntrinsic starts as zero.
 */
#pragma once

#include <array>
#include <memory>
#include <vector>

#include "painty/b200/Device.hxx"
#include "painty/core/Spline.hxx"
#include "painty/io/ImageIO.hxx"
#include "painty/renderer/BrushBase.hxx"
#include "painty/renderer/Canvas.hxx"
#include "painty/renderer/PaintLayer.hxx"

namespace painty {
template <class vector_type>
class FootprintBrush final : public BrushBase<vector_type> {
  using T                 = typename BrushBase<vector_type>::T;
  static constexpr auto N = BrushBase<vector_type>::N;

  struct Handle {
    pb_fbrush* h = nullptr;
    Handle() { b200::check(pb_fbrush_create(b200::context(), &h)); }
    ~Handle() {
      if (h) pb_fbrush_destroy(h);
    }
  };

 public:
  FootprintBrush(const double radius) : _h(std::make_shared<Handle>()), _footprint(0, 0) { setRadius(radius); }
  ~FootprintBrush() override = default;

  void setRadius(const double radius) override {  // reference :46-63
    int acted = 0;
    b200::check(pb_fbrush_set_radius(_h->h, radius, 0, nullptr, &acted));
    if (!acted) return;
    io::imRead("./data/footprint/footprint.png", _footprintFullSize, true);
    const auto width   = static_cast<int32_t>(2.0 * std::ceil(radius) + 1.0);
    const auto sizeMap = static_cast<int32_t>(std::ceil(std::sqrt(2.0) * width));
    const auto pad     = (sizeMap - width) / 2;
    _footprint         = PaddedMat(ScaledMat(_footprintFullSize, width, width), pad, pad, pad, pad, 0.0);
    b200::check(pb_fbrush_set_radius(_h->h, radius, _footprint.rows, _footprint.data, &acted));
    _pickupMap.reset();
  }

  /** One imprint (reference :73-143). */
  void imprint(const vec2& center, const double theta, Canvas<vector_type>& canvas) {
    const double cx = center[0U], cy = center[1U];
    b200::check(pb_fbrush_imprint_batch(_h->h, canvas.device(), 1, &cx, &cy, &theta));
    touched(canvas);
  }
  /** n imprints in order as one device-side chain (the GUI's per-mouse-move loop, DigitalCanvas.cxx:117-122). */
  void imprint(const std::vector<vec2>& centers, const std::vector<double>& thetas, Canvas<vector_type>& canvas) {
    std::vector<double> cx(centers.size()), cy(centers.size());
    for (size_t i = 0; i < centers.size(); ++i) cx[i] = centers[i][0U], cy[i] = centers[i][1U];
    b200::check(pb_fbrush_imprint_batch(_h->h, canvas.device(), static_cast<int64_t>(cx.size()), cx.data(), cy.data(), thetas.data()));
    touched(canvas);
  }

  void dip(const std::array<vector_type, 2UL>& paint) override {  // reference :150-154
    const double K[3] = {paint[0U][0U], paint[0U][1U], paint[0U][2U]}, S[3] = {paint[1U][0U], paint[1U][1U], paint[1U][2U]};
    b200::check(pb_fbrush_dip(_h->h, K, S));
    if (_pickupMap) _pickupMap->deviceWritten();
  }
  void clean() {  // reference :160-166
    b200::check(pb_fbrush_clean(_h->h));
    if (_pickupMap) _pickupMap->deviceWritten();
  }
  void updateSnapshot(const Canvas<vector_type>& canvas) { b200::check(pb_fbrush_update_snapshot(_h->h, canvas.device())); }

  const PaintLayer<vector_type>& getPickupMap() const {  // reference :174
    if (!_pickupMap) {
      pb_layer* view = nullptr;
      b200::check(pb_fbrush_pickup_layer(_h->h, &view));
      _pickupMap = std::make_unique<PaintLayer<vector_type>>(view);
    }
    return *_pickupMap;
  }
  const Mat<double>& getFootprint() const { return _footprint; }

  void setPickupRate(const T rate) { pb_fbrush_set_pickup_rate(_h->h, rate); }
  void setDepositionRate(const T rate) { pb_fbrush_set_deposition_rate(_h->h, rate); }
  T getPickupRate() const { return pb_fbrush_get_pickup_rate(_h->h); }
  T getDepositionRate() const { return pb_fbrush_get_deposition_rate(_h->h); }
  bool getUseSnapshotBuffer() const { return pb_fbrush_get_use_snapshot(_h->h) != 0; }
  void setUseSnapshotBuffer(const bool use) { pb_fbrush_set_use_snapshot(_h->h, use ? 1 : 0); }

  void paintStroke(const std::vector<vec2>& path, Canvas<vector_type>& canvas) override {  // reference :206-268
    if (path.size() < 2UL) return;
    std::vector<double> xy(2 * path.size());
    for (size_t i = 0; i < path.size(); ++i) xy[2 * i] = path[i][0U], xy[2 * i + 1] = path[i][1U];
    int64_t n = 0;
    b200::check(pb_expand_stroke(0, static_cast<int>(path.size()), xy.data(), 0, nullptr, nullptr, nullptr, &n));
    std::vector<double> cx(static_cast<size_t>(n)), cy(cx.size()), th(cx.size());
    b200::check(pb_expand_stroke(0, static_cast<int>(path.size()), xy.data(), n, cx.data(), cy.data(), th.data(), &n));
    b200::check(pb_fbrush_imprint_batch(_h->h, canvas.device(), n, cx.data(), cy.data(), th.data()));
    touched(canvas);
  }

  pb_fbrush* device() const { return _h->h; }

 private:
  void touched(Canvas<vector_type>& canvas) {
    canvas.deviceWritten();
    if (_pickupMap) _pickupMap->deviceWritten();
  }

  std::shared_ptr<Handle> _h;
  Mat<double> _footprint;
  Mat<double> _footprintFullSize;
  mutable std::unique_ptr<PaintLayer<vector_type>> _pickupMap;
};
}  // namespace painty
