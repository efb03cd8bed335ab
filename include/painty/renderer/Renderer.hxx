/**
 * Drop-in for painty/renderer/Renderer.hxx (reference lines 15-158): compose() runs the fused streaming
 * Kubelka-Munk kernel, render() the same kernel with the directional-light relighting (reference :60-156) fused
 * behind it; both return a host Mat by value like the reference.
 */
#pragma once

#include "painty/b200/Device.hxx"
#include "painty/image/Mat.hxx"
#include "painty/renderer/Canvas.hxx"
#include "painty/renderer/PaintLayer.hxx"

namespace painty {
template <class vector_type>
class Renderer final {
  using T                 = typename DataType<vector_type>::channel_type;
  static constexpr auto N = DataType<vector_type>::dim;

 public:
  /** Compose wet layer onto substrate (reference :26-41). */
  Mat<vector_type> compose(const PaintLayer<vector_type>& paintLayer, const Mat<vector_type>& R0_buffer) const {
    Mat<vector_type> R1(R0_buffer.rows, R0_buffer.cols);
    b200::check(pb_layer_compose(paintLayer.device(), reinterpret_cast<const double*>(R0_buffer.data),
                                 reinterpret_cast<double*>(R1.data)));
    return R1;
  }

  /** Compose current wet layer of canvas onto substrate (reference :48-53). */
  Mat<vector_type> compose(const Canvas<vector_type>& canvas) const {
    Mat<vector_type> R1(canvas.getPaintLayer().getRows(), canvas.getPaintLayer().getCols());
    b200::check(pb_canvas_compose(canvas.device(), reinterpret_cast<double*>(R1.data)));
    return R1;
  }

  /** Render the canvas with directional light (reference :60-156). */
  Mat<vector_type> render(const Canvas<vector_type>& canvas) const {
    Mat<vector_type> rgb(canvas.getPaintLayer().getRows(), canvas.getPaintLayer().getCols());
    b200::check(pb_canvas_render(canvas.device(), reinterpret_cast<double*>(rgb.data)));
    return rgb;
  }
};
}  // namespace painty
