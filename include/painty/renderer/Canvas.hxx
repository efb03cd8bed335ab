/**
 * Drop-in for painty/renderer/Canvas.hxx (reference lines 20-198): wet PaintLayer + dry substrate R0 + height
 * h live in 11 device SoA planes; getR0()/get_h() are lazily synchronised host mirrors.
 *
 * Drying clock: the reference keeps a wall-clock time point per pixel and dries inside checkDry(). The device
 * path keeps the API (getTimeMap, get/setDryingTime, checkDry) on the host; brush kernels do not consult the
 * clock (with the reference's default of ~4.17 h, or 0, checkDry only stamps times — SURVEY.md §5).
 */
#pragma once

#include <algorithm>
#include <chrono>
#include <cstring>
#include <memory>
#include <type_traits>
#include <vector>

#include "painty/b200/Device.hxx"
#include "painty/core/KubelkaMunk.hxx"
#include "painty/core/Vec.hxx"
#include "painty/image/Mat.hxx"
#include "painty/renderer/PaintLayer.hxx"

namespace painty {
template <class vector_type>
class Canvas final {
  using T                 = typename DataType<vector_type>::channel_type;
  static constexpr auto N = DataType<vector_type>::dim;

  struct State {
    pb_canvas* handle = nullptr;
    Mat<vector_type> R0;
    Mat<T> h;
    bool host_valid = false, device_valid = true;
    bool host_access = false;    // a mutable reference to R0 / h has been handed out (see PaintLayer.hxx, "Held references")
    std::vector<double> shadow;  // R0 | h as of the last synchronisation
    ~State() {
      if (handle) pb_canvas_destroy(handle);
    }
  };
  static std::shared_ptr<State> make(int32_t rows, int32_t cols) {
    auto s = std::make_shared<State>();
    b200::check(pb_canvas_create(b200::context(), rows, cols, &s->handle));
    return s;
  }
  static pb_layer* layerView(pb_canvas* c) {
    pb_layer* l = nullptr;
    b200::check(pb_canvas_paint_layer(c, &l));
    return l;
  }

 public:
  Canvas(const int32_t rows, const int32_t cols)
      : _s(make(rows, cols)),
        _paintLayer(layerView(_s->handle)),
        _timeMap(static_cast<size_t>(rows * cols), std::chrono::system_clock::now()),
        _dryingTime(static_cast<uint32_t>(0.25 * 60 * 1000000)) {}

  void clear() {  // reference :37-58
    b200::check(pb_canvas_clear(_s->handle));
    deviceWritten();
    std::fill(_timeMap.begin(), _timeMap.end(), std::chrono::system_clock::now());
  }

  const Mat<vector_type>& getR0() const { return toHost(), _s->R0; }
  const Mat<T>& get_h() const { return toHost(), _s->h; }
  Mat<vector_type>& getR0() { return hostAccess(), _s->R0; }
  Mat<T>& get_h() { return hostAccess(), _s->h; }
  /** Leave host-access mode of the canvas and its wet layer (see PaintLayer.hxx, "Held references"). */
  void endHostAccess() {
    device();
    _paintLayer.endHostAccess();
    _s->host_access = false;
    _s->shadow.clear();
    _s->shadow.shrink_to_fit();
  }
  Mat<vector_type> getReflectanceLayerDry() const { return getR0().clone(); }

  void setBackground(const Mat<vector_type>& background) {  // reference :80-87
    b200::check(pb_canvas_set_background(_s->handle, reinterpret_cast<const double*>(background.data)));
    deviceWritten();
  }

  const PaintLayer<vector_type>& getPaintLayer() const { return _paintLayer; }
  PaintLayer<vector_type>& getPaintLayer() { return _paintLayer; }

  const std::vector<std::chrono::system_clock::time_point>& getTimeMap() const { return _timeMap; }
  std::vector<std::chrono::system_clock::time_point>& getTimeMap() { return _timeMap; }

  void dryCanvas() {  // reference :105-121, one fused kernel
    b200::check(pb_canvas_dry(device()));
    deviceWritten();
    std::fill(_timeMap.begin(), _timeMap.end(), std::chrono::system_clock::now());
  }

  /** Reference :123-155. Host-side on the mirrors; only stamps the time map unless a drying time elapsed. */
  void checkDry(int32_t x, int32_t y, const std::chrono::system_clock::time_point& timePoint) {
    const size_t idx = static_cast<size_t>(y * pb_canvas_cols(_s->handle) + x);
    if (_dryingTime.count() > 0U) {
      T v = static_cast<const PaintLayer<vector_type>&>(_paintLayer).getV_buffer()(y, x);
      if (v > 0.001) {
        auto dur = std::chrono::duration_cast<std::chrono::milliseconds>(timePoint - _timeMap[idx]);
        if (dur >= _dryingTime) {
          get_h()(y, x) += v;
          getR0()(y, x) = ComputeReflectance(_paintLayer.getK_buffer()(y, x), _paintLayer.getS_buffer()(y, x), getR0()(y, x), v);
          _paintLayer.getV_buffer()(y, x) = 0.0;
          _paintLayer.getK_buffer()(y, x).fill(0.0);
          _paintLayer.getS_buffer()(y, x).fill(0.0);
        } else {
          const T rate = static_cast<T>(dur.count() / _dryingTime.count());
          if (rate > 0.01) {
            T vl = rate * v;
            get_h()(y, x) += vl;
            getR0()(y, x) = ComputeReflectance(_paintLayer.getK_buffer()(y, x), _paintLayer.getS_buffer()(y, x), getR0()(y, x), vl);
            _paintLayer.getV_buffer()(y, x) = v - vl;
          }
        }
      }
    }
    _timeMap[idx] = timePoint;
  }

  std::chrono::milliseconds getDryingTime() { return _dryingTime; }
  void setDryingTime(std::chrono::milliseconds msecs) { _dryingTime = msecs; }

  // ---- façade plumbing -------------------------------------------------------------------------------
  /** Device handle with pending host edits (wet layer, R0, h) uploaded. */
  pb_canvas* device() const {
    _paintLayer.device();
    if (_s->host_access && _s->host_valid && differsFromShadow()) _s->device_valid = false;  // written through a held reference
    if (!_s->device_valid) {
      b200::check(pb_canvas_upload_substrate(_s->handle, reinterpret_cast<const double*>(_s->R0.data),
                                             reinterpret_cast<const double*>(_s->h.data)));
      _s->device_valid = true;
      if (_s->host_access) snapshotShadow();
    }
    return _s->handle;
  }
  void deviceWritten() const {
    _paintLayer.deviceWritten();
    _s->host_valid   = false;
    _s->device_valid = true;
    if (_s->host_access) {
      toHost();
      snapshotShadow();
    }
  }

 private:
  void hostAccess() {
    toHost();
    if (!_s->host_access) {
      _s->host_access = true;
      snapshotShadow();
    }
  }
  size_t pixels() const { return static_cast<size_t>(pb_canvas_rows(_s->handle)) * static_cast<size_t>(pb_canvas_cols(_s->handle)); }
  void snapshotShadow() const {
    const size_t n = pixels();
    _s->shadow.resize(4 * n);
    std::memcpy(_s->shadow.data(), _s->R0.data, 3 * n * sizeof(double));
    std::memcpy(_s->shadow.data() + 3 * n, _s->h.data, n * sizeof(double));
  }
  bool differsFromShadow() const {
    const size_t n = pixels();
    if (_s->shadow.size() != 4 * n) return true;
    return std::memcmp(_s->shadow.data(), _s->R0.data, 3 * n * sizeof(double)) != 0 ||
           std::memcmp(_s->shadow.data() + 3 * n, _s->h.data, n * sizeof(double)) != 0;
  }
  void toHost() const {
    if (_s->host_valid) return;
    const int32_t r = pb_canvas_rows(_s->handle), c = pb_canvas_cols(_s->handle);
    if (_s->R0.rows != r || _s->R0.cols != c) {
      _s->R0 = Mat<vector_type>(r, c);
      _s->h  = Mat<T>(r, c);
    }
    b200::check(pb_canvas_download(_s->handle, nullptr, nullptr, nullptr, reinterpret_cast<double*>(_s->R0.data),
                                   reinterpret_cast<double*>(_s->h.data)));
    _s->host_valid = true;
  }

  std::shared_ptr<State> _s;
  PaintLayer<vector_type> _paintLayer;
  std::vector<std::chrono::system_clock::time_point> _timeMap;
  std::chrono::milliseconds _dryingTime;
};
}  // namespace painty
