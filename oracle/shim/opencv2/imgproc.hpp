// TEST INFRASTRUCTURE ONLY (oracle/). Minimal stand-in for the subset of OpenCV core/imgproc
// that the painty hot-path headers use: cv::Mat_<T> (ref-counted shallow copies, clone()),
// cv::Size, cv::DataType/DataDepth traits, cv::borderInterpolate(BORDER_REFLECT) and a
// cv::resize that is *substituted*: it asks the oracle driver for a pre-baked result
// (Python cv2 4.13 INTER_LANCZOS4 on the same f64 input, see oracle/bake_assets.py), because
// resize results are *inputs* of the hot path. Written from scratch; NOT OpenCV.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <vector>

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) (((depth)&7) + (((cn)-1) << CV_CN_SHIFT))

// Provided by the oracle driver translation unit: fill `out` (rows*cols doubles) with the baked
// resize of an (in_rows x in_cols) f64 image; returns false when nothing was registered.
extern "C" bool oracle_shim_resize_f64(int in_rows, int in_cols, int out_rows, int out_cols,
                                       double* out);

namespace cv {

enum BorderTypes { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_REFLECT_101 = 4 };
enum InterpolationFlags { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3, INTER_LANCZOS4 = 4 };

struct Size {
  int width = 0, height = 0;
  Size() = default;
  Size(int w, int h) : width(w), height(h) {}
};

template <class T>
struct DataDepth {
  enum { value = CV_64F, fmt = 'd' };
};
template <>
struct DataDepth<uint8_t> {
  enum { value = CV_8U, fmt = 'u' };
};
template <>
struct DataDepth<int32_t> {
  enum { value = CV_32S, fmt = 'i' };
};
template <>
struct DataDepth<uint32_t> {
  enum { value = CV_32S, fmt = 'i' };
};
template <>
struct DataDepth<float> {
  enum { value = CV_32F, fmt = 'f' };
};
template <>
struct DataDepth<double> {
  enum { value = CV_64F, fmt = 'd' };
};

template <class T>
class DataType {
 public:
  typedef T value_type;
  typedef T channel_type;
  enum { generic_type = 1, depth = DataDepth<T>::value, channels = 1, type = CV_MAKETYPE(depth, 1) };
};

// BORDER_REFLECT: fedcba|abcdefgh|hgfedcb
inline int borderInterpolate(int p, int len, int borderType) {
  if (static_cast<unsigned>(p) < static_cast<unsigned>(len)) return p;
  if (borderType == BORDER_REPLICATE) return p < 0 ? 0 : len - 1;
  if (borderType == BORDER_REFLECT || borderType == BORDER_REFLECT_101) {
    const int delta = borderType == BORDER_REFLECT_101;
    if (len == 1) return 0;
    do {
      if (p < 0)
        p = -p - 1 + delta;
      else
        p = len - 1 - (p - len) - delta;
    } while (static_cast<unsigned>(p) >= static_cast<unsigned>(len));
    return p;
  }
  throw std::invalid_argument("shim borderInterpolate: unsupported border type");
}

template <class T>
class Mat_ {
 public:
  typedef T value_type;
  typedef T* iterator;
  typedef const T* const_iterator;

  int rows = 0, cols = 0;
  T* data = nullptr;

  Mat_() = default;
  Mat_(int r, int c) { create(r, c); }
  explicit Mat_(Size s) { create(s.height, s.width); }

  void create(int r, int c) {
    rows = r;
    cols = c;
    const std::size_t n = static_cast<std::size_t>(r < 0 ? 0 : r) * static_cast<std::size_t>(c < 0 ? 0 : c);
    _buf = std::make_shared<std::vector<T>>(n);
    data = _buf->data();
  }

  std::size_t total() const { return static_cast<std::size_t>(rows) * static_cast<std::size_t>(cols); }
  bool empty() const { return total() == 0 || data == nullptr; }
  Size size() const { return Size(cols, rows); }

  T& operator()(int i) { return data[i]; }
  const T& operator()(int i) const { return data[i]; }
  T& operator()(int i, int j) { return data[static_cast<std::size_t>(i) * cols + j]; }
  const T& operator()(int i, int j) const { return data[static_cast<std::size_t>(i) * cols + j]; }

  iterator begin() { return data; }
  iterator end() { return data + total(); }
  const_iterator begin() const { return data; }
  const_iterator end() const { return data + total(); }

  Mat_ clone() const {
    Mat_ m(rows, cols);
    for (std::size_t i = 0; i < total(); ++i) m.data[i] = data[i];
    return m;
  }

 private:
  std::shared_ptr<std::vector<T>> _buf;
};

template <class T>
inline void resize(const Mat_<T>& in, Mat_<T>& out, Size dsize, double = 0.0, double = 0.0,
                   int /*flag*/ = INTER_LANCZOS4) {
  out = Mat_<T>(dsize.height, dsize.width);
  if constexpr (std::is_same<T, double>::value) {
    if (!oracle_shim_resize_f64(in.rows, in.cols, dsize.height, dsize.width, out.data)) {
      throw std::runtime_error("shim cv::resize: no baked result registered for this size");
    }
  } else {
    throw std::runtime_error("shim cv::resize: only f64 single-channel images are supported");
  }
}

}  // namespace cv
