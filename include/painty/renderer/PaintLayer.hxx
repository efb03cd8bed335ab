/**
 * Drop-in for painty/renderer/PaintLayer.hxx (reference lines 22-145): same class, same members, storage moved
 * to device-resident SoA planes behind the painty_b200 C ABI.
 *
 * getK/S/V_buffer() still hand out references to host cv::Mat_ matrices like the reference; they are lazily
 * synchronised mirrors (download on access, upload before the next device operation after a host write).
 * Copies are shallow like the reference's cv::Mat_ members (they share state); copyTo() is the deep copy.
 *
 * Held references. The reference's callers keep `auto& vBuffer = layer.getV_buffer()` across operations. Once a MUTABLE
 * getter has handed out a reference the layer is in "host access" mode: the mirrors keep their buffers, are refreshed
 * right after every device operation (reads through a held reference stay current), and are compared with a shadow copy
 * before every device operation (writes through a held reference are uploaded, nothing is silently dropped). That costs a
 * download + a memcmp per operation; endHostAccess() leaves the mode again. Read-only access (const getters) never enters it.
 */
#pragma once

#include <cstring>
#include <memory>
#include <type_traits>
#include <vector>

#include "painty/b200/Device.hxx"
#include "painty/core/KubelkaMunk.hxx"
#include "painty/image/Mat.hxx"

namespace painty {
template <class vector_type>
class PaintLayer final {
  using T                 = typename DataType<vector_type>::channel_type;
  static constexpr auto N = DataType<vector_type>::dim;
  static_assert(std::is_same<T, double>::value && N == 3, "painty_b200 implements the vec3 (double x 3) instantiation");

  struct State {
    pb_layer* handle = nullptr;
    Mat<vector_type> K, S;
    Mat<T> V;
    bool host_valid = false, device_valid = true;
    bool host_access = false;        // a mutable reference to the mirrors has been handed out
    std::vector<double> shadow;      // mirrors as of the last synchronisation (K | S | V), host-access mode only
    ~State() {
      if (handle) pb_layer_destroy(handle);
    }
  };

 public:
  PaintLayer(int32_t rows, int32_t cols) : _s(std::make_shared<State>()) {
    b200::check(pb_layer_create(b200::context(), rows, cols, &_s->handle));
  }
  // adopt a view handed out by the C ABI (canvas wet layer, brush pickup map)
  explicit PaintLayer(pb_layer* view) : _s(std::make_shared<State>()) { _s->handle = view; }

  const Mat<vector_type>& getK_buffer() const { return toHost(), _s->K; }
  const Mat<vector_type>& getS_buffer() const { return toHost(), _s->S; }
  const Mat<T>& getV_buffer() const { return toHost(), _s->V; }
  Mat<vector_type>& getK_buffer() { return hostAccess(), _s->K; }
  Mat<vector_type>& getS_buffer() { return hostAccess(), _s->S; }
  Mat<T>& getV_buffer() { return hostAccess(), _s->V; }
  /** Leave host-access mode: references handed out by the mutable getters must no longer be written through. */
  void endHostAccess() {
    device();
    _s->host_access = false;
    _s->shadow.clear();
    _s->shadow.shrink_to_fit();
  }

  int32_t getCols() const { return pb_layer_cols(_s->handle); }
  int32_t getRows() const { return pb_layer_rows(_s->handle); }

  /** Set all values to zero (reference :68-74). */
  void clear() {
    b200::check(pb_layer_clear(_s->handle));
    deviceWritten();
  }

  /** Compose this layer onto a substrate, in place (reference :81-96). */
  void composeOnto(Mat<vector_type>& R0) const {
    if ((R0.rows != getRows()) || (R0.cols != getCols())) {
      R0 = Mat<vector_type>(getRows(), getCols());
      for (auto& v : R0) v.fill(1.0);
    }
    b200::check(pb_layer_compose_onto(device(), reinterpret_cast<double*>(R0.data)));
  }

  /** Deep copy (reference :103-111). */
  void copyTo(PaintLayer& other) const {
    b200::check(pb_layer_copy(device(), other._s->handle));
    other.deviceWritten();
  }

  /** Update a cell (reference :122-127): i = row, j = col. */
  void set(int32_t i, int32_t j, const vector_type& k, const vector_type& s, const T v) {
    getK_buffer()(i, j) = k;
    getS_buffer()(i, j) = s;
    getV_buffer()(i, j) = v;
  }

  // ---- façade plumbing -------------------------------------------------------------------------------
  /** Device handle with pending host edits uploaded. */
  pb_layer* device() const {
    if (_s->host_access && _s->host_valid && differsFromShadow()) _s->device_valid = false;  // written through a held reference
    if (!_s->device_valid) {
      b200::check(pb_layer_upload(_s->handle, reinterpret_cast<const double*>(_s->K.data),
                                  reinterpret_cast<const double*>(_s->S.data), reinterpret_cast<const double*>(_s->V.data)));
      _s->device_valid = true;
      if (_s->host_access) snapshotShadow();
    }
    return _s->handle;
  }
  /** A kernel wrote the planes: host mirrors are stale (refreshed at once while references to them are held). */
  void deviceWritten() const {
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
  size_t pixels() const { return static_cast<size_t>(getRows()) * static_cast<size_t>(getCols()); }
  void snapshotShadow() const {
    const size_t n = pixels();
    _s->shadow.resize(7 * n);
    std::memcpy(_s->shadow.data(), _s->K.data, 3 * n * sizeof(double));
    std::memcpy(_s->shadow.data() + 3 * n, _s->S.data, 3 * n * sizeof(double));
    std::memcpy(_s->shadow.data() + 6 * n, _s->V.data, n * sizeof(double));
  }
  bool differsFromShadow() const {
    const size_t n = pixels();
    if (_s->shadow.size() != 7 * n) return true;
    return std::memcmp(_s->shadow.data(), _s->K.data, 3 * n * sizeof(double)) != 0 ||
           std::memcmp(_s->shadow.data() + 3 * n, _s->S.data, 3 * n * sizeof(double)) != 0 ||
           std::memcmp(_s->shadow.data() + 6 * n, _s->V.data, n * sizeof(double)) != 0;
  }
  void toHost() const {
    if (_s->host_valid) return;
    const int32_t r = getRows(), c = getCols();
    if (_s->K.rows != r || _s->K.cols != c) {
      _s->K = Mat<vector_type>(r, c);
      _s->S = Mat<vector_type>(r, c);
      _s->V = Mat<T>(r, c);
    }
    static_assert(sizeof(vector_type) == 3 * sizeof(double), "vec3 must be 3 packed doubles");
    b200::check(pb_layer_download(_s->handle, reinterpret_cast<double*>(_s->K.data), reinterpret_cast<double*>(_s->S.data),
                                  reinterpret_cast<double*>(_s->V.data)));
    _s->host_valid = true;
  }

  std::shared_ptr<State> _s;
};
}  // namespace painty
