/** Abstract brush contract of the B200 façade.
 *
 * Keeps the member names, argument types and defaults of the reference's interface
 * (painty/renderer/BrushBase.hxx:15-47), so code written against BrushBase<vec3> compiles unchanged; the two concrete
 * brushes of this façade (FootprintBrush, TextureBrush) forward to the C ABI in painty_b200.h. */
#pragma once

#include <array>
#include <vector>

#include "painty/renderer/Canvas.hxx"

namespace painty {

template <class vector_type>
class BrushBase {
  using Traits = DataType<vector_type>;

 public:
  using T                 = typename Traits::channel_type;
  static constexpr auto N = Traits::dim;
  using Paint             = std::array<vector_type, 2UL>;  ///< {absorption K, scattering S}
  using Path              = std::vector<vec2>;              ///< canvas coordinates, x = column

  BrushBase()          = default;
  virtual ~BrushBase() = default;

  /// Brush radius in canvas pixels.
  virtual void setRadius(double radius) = 0;
  /// Load the brush with a paint.
  virtual void dip(const Paint& paint) = 0;
  /// Paint one stroke along `path`.
  virtual void paintStroke(const Path& path, Canvas<vector_type>& canvas) = 0;

  /// Factor applied to the thickness of the deposited layer; 1 unless set.
  auto getThicknessScale() const -> T { return thickness_scale_; }
  void setThicknessScale(const T scale) { thickness_scale_ = scale; }

 private:
  T thickness_scale_ = static_cast<T>(1);
};

}  // namespace painty
