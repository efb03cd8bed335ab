// Shared declarations of libpainty_b200.so (device-resident paint renderer for sm_100a).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/painty_b200.h"

namespace pb {

void set_error(const std::string& msg);

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define PB_CUDA(call)                                                                                   \
  do {                                                                                                  \
    cudaError_t _e = (call);                                                                            \
    if (_e != cudaSuccess) {                                                                            \
      throw pb::Error(std::string(#call) + " failed: " + cudaGetErrorString(_e) + " (" __FILE__ ":" + \
                      std::to_string(__LINE__) + ")");                                                  \
    }                                                                                                   \
  } while (0)

#define PB_REQUIRE(cond, msg)                 \
  do {                                        \
    if (!(cond)) throw pb::Error(msg);        \
  } while (0)

constexpr int kLayerPlanes  = 7;   // Kr Kg Kb Sr Sg Sb V
constexpr int kCanvasPlanes = 11;  // + R0r R0g R0b h
constexpr int PK = 0, PS = 3, PV = 6, PR = 7, PH = 10;

// thresholds of the f64 reference, used in both precisions
constexpr double kKmEps     = 2.220446049250313e-16 * 10000.0;  // KubelkaMunk.hxx:31-32,39
constexpr double kMinVolume = 0.001;                            // FootprintBrush.hxx:26

}  // namespace pb

struct pb_context {
  int device       = 0;
  int precision    = PB_F32;
  cudaStream_t stream = nullptr;
  // second stream + fork / join events for the straddling strokes of a multi-GPU run (created on first use)
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int64_t launches = 0;
  int sm_count     = 148;
  size_t esize() const { return precision == PB_F64 ? 8 : 4; }
};

struct pb_canvas;
// 7 (layer) or 11 (canvas) dense planes carved out of one allocation; plane p at base + p*stride bytes.
struct pb_planes {
  pb_context* ctx = nullptr;
  int rows = 0, cols = 0;  // stored extent
  int nplanes   = 0;
  void* base    = nullptr;
  size_t stride = 0;  // bytes between planes (256 B aligned)
  int64_t n() const { return static_cast<int64_t>(rows) * cols; }
  void* plane(int p) const { return static_cast<char*>(base) + static_cast<size_t>(p) * stride; }
};

struct pb_layer {
  pb_planes pl;
  bool owns         = true;     // false: a view of a canvas' wet planes or of a brush's pickup map
  pb_canvas* canvas = nullptr;  // set for canvas views (writes bump the canvas version)
};

struct pb_canvas {
  pb_planes pl;       // 11 planes over the stored rows
  int rows = 0;       // logical (global) canvas height
  int cols = 0;
  int row_begin = 0, row_end = 0;  // owned band
  int halo      = 0;
  int store_first = 0;  // first stored global row
  uint64_t id      = 0;  // unique per canvas
  uint64_t version = 1;  // bumped by every operation that modifies the wet layer
};

namespace pb {

// ---- layout.cu ----------------------------------------------------------------------------------
void planes_alloc(pb_context* ctx, pb_planes& pl, int rows, int cols, int nplanes);
void planes_free(pb_planes& pl);
void planes_alloc_temp(pb_context* ctx, pb_planes& pl, int rows, int cols, int nplanes);
void planes_free_temp(pb_planes& pl);
void fill_plane(pb_context* ctx, void* plane, int64_t n, double value);
// host AoS f64 (n*ch) <-> `ch` consecutive device planes starting at plane index p0
void upload_aos(pb_context* ctx, const pb_planes& pl, int p0, int ch, const double* host);
void download_aos(pb_context* ctx, const pb_planes& pl, int p0, int ch, double* host);
void copy_planes(pb_context* ctx, const pb_planes& src, pb_planes& dst, int nplanes);
// The imprint engine's pixel records (8 elements per pixel: Kr Kg Kb Sr Sg Sb V 0; record index = stored row * cols +
// column): convert the rectangle [y0, y1] x [x0, x1] (stored rows) of the 7 layer planes to / from a record array.
constexpr int kRecord = 8;
void planes_to_records(pb_context* ctx, const pb_planes& pl, void* records, int x0, int y0, int x1, int y1);
void records_to_planes(pb_context* ctx, const void* records, const pb_planes& pl, int x0, int y0, int x1, int y1);

// ---- km_compose.cu ------------------------------------------------------------------------------
struct ComposeArgs {
  const void* K[3];
  const void* S[3];
  const void* V;
  const void* R0[3];
  void* R[3];
};
void km_compose(pb_context* ctx, int64_t n, const ComposeArgs& a);
// compose with a gather epilogue: R is stored into n_dst (<= 8) destination images (own HBM or NVLink peer mappings)
void km_compose_gather(pb_context* ctx, int64_t n, const ComposeArgs& a, int n_dst, void* const (*dst)[3]);
constexpr int kMaxStack = 8;
struct StackArgs {
  const void* K[kMaxStack][3];
  const void* S[kMaxStack][3];
  const void* V[kMaxStack];
  const void* R0[3];
  void* R[3];
  int n_layers;
};
void km_compose_stacked(pb_context* ctx, int64_t n, const StackArgs& a);
// compose + rgb2srgb + quantise into `out` (device). mode 0: uint32 0xffRRGGBB (truncating cast); mode 1 / 2:
// BGR interleaved uint8 / uint16 with round-half-even + saturation; srgb = false skips the transfer curve.
void km_compose_display(pb_context* ctx, int64_t n, const ComposeArgs& a, int mode, bool srgb, void* out);
// Renderer::render: R = light(KM(K,S,V over R0), height = V), rows x cols image, R planes out
void km_render(pb_context* ctx, int rows, int cols, const ComposeArgs& a);
// planner read-back prep: compose + CIELab into a.R (3 planes of the element type), then OpenCV's LANCZOS4 resize
void km_compose_lab(pb_context* ctx, int64_t n, const ComposeArgs& a);
void lab_resize_lanczos4(pb_context* ctx, void* const lab[3], int rows, int cols, int orows, int ocols, const int* d_xofs,
                         const float* d_alpha, const int* d_yofs, const float* d_beta, double* d_tmp, double* d_out);
void lab_planes_to_aos(pb_context* ctx, void* const lab[3], int64_t n, double* d_out);
// Canvas::dryCanvas: h += V; R0 = KM(K,S,R0,V); K=S=V=0
void km_dry(pb_context* ctx, int64_t n, void* const planes[11]);

}  // namespace pb
