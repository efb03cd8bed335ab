// Device-side interface of the texture-brush deposit engine (texture.cu).
#pragma once
#include <algorithm>

#include "common.cuh"

namespace pb {

constexpr int kMaxPoly = 1024;  // polygon vertices staged in shared memory (2 * (path vertices + 2))

struct DevTStroke {
  double K[3], S[3];
  double thickness_scale;
  int32_t x0, x1, y0, y1;          // (int)boundMin .. (int)boundMax, TextureBrush.hxx:142-145
  int32_t local_rows, local_cols;  // size of the reference's local thicknessMap (:135-136)
  int32_t poly_begin, n_poly;
  int32_t pred_begin, pred_end;
};

struct TextureLaunch {
  void* canvas[kLayerPlanes];
  int rows, cols, store_first, store_rows;
  const double* map;  // thickness map, f64 row-major
  int map_rows, map_cols;
  const DevTStroke* strokes;
  int64_t n_strokes;
  const double2* poly;
  const double2* uv;
  const int32_t* preds;
  int* done;
  int* queue;
  unsigned long long* counters;  // [0] deposited stroke-pixels
};

void texture_launch(pb_context* ctx, const TextureLaunch& L);

}  // namespace pb
