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
  int32_t tx0, ty0, tx1, ty1;  // canvas tiles covered by the bounding box (inclusive); empty if tx1 < tx0
  int64_t item_begin;          // first work item (stroke, tile) of this stroke
  // the stroke's thickness texture (f64 row-major): the brush's own sample map or an entry of its texture atlas
  // (TextureBrushDictionary.cxx:25-79 picks one per stroke)
  const double* map;
  int32_t map_rows, map_cols;
};

struct TextureLaunch {
  void* canvas[kLayerPlanes];
  int rows, cols, store_first, store_rows;
  const DevTStroke* strokes;
  int64_t n_strokes;
  const double2* poly;
  const double2* uv;
  // tile-ticket dataflow: item i = (stroke, canvas tile); an item may run when `tile_done[tile] == ticket[i]`,
  // i.e. when every earlier stroke that covers the tile has finished with it
  int tile, tiles_x, tiles_y;
  int64_t n_items;
  int32_t* ticket;               // per item, filled on the device by texture_ticket_kernel
  int* tile_done;                // per canvas tile (zeroed)
  unsigned long long* queue;     // item counter (zeroed)
  unsigned long long* counters;  // [0] deposited stroke-pixels
};

constexpr int kTextureTile = 64;

void texture_launch(pb_context* ctx, const TextureLaunch& L);

// ---- smudge path (Smudge.hxx): a stroke = thickness pass -> serial smudge walk -> deposit pass -----------------
struct DevSmudgeStep {
  double cx, cy, c, s;
  int32_t roi_x, roi_y;
};
struct SmudgeLaunch {
  void* canvas[kLayerPlanes];
  int rows, cols, store_first, store_rows;
  DevTStroke stroke;  // one stroke at a time: the smudge state chains strokes serially
  const double2* poly;
  const double2* uv;
  double* tmap;  // the reference's local thicknessMap (local_rows x local_cols, f64, zeroed)
  unsigned long long* max_bits;  // max over tmap as the bit pattern of a non-negative double (zeroed)
  unsigned long long* counters;
  // smudge state
  void* pick[2][kLayerPlanes];  // ping-pong pickup windows (size x size)
  int size, max_size, first_dst;
  const DevSmudgeStep* steps;
  int n_steps;
  double bmin_x, bmin_y, pickup_rate, deposition_rate;
};
void texture_thickness_launch(pb_context* ctx, const SmudgeLaunch& L);
void texture_smudge_launch(pb_context* ctx, const SmudgeLaunch& L);
void texture_deposit_launch(pb_context* ctx, const SmudgeLaunch& L);

}  // namespace pb
