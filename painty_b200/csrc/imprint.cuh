// Device-side interface of the footprint-brush imprint engine (imprint.cu).
#pragma once
#include "common.cuh"
#include "imprint_geom.hpp"

namespace pb {

// One footprint geometry (= one ceil(radius)): the compacted list of active cells (height > 0) of the
// padded footprint. Cell i sits at map position (xy[i] & 0xffff, xy[i] >> 16) and has height fh[i].
struct FootprintGeom {
  int width    = 0;  // 2*ceil(r)+1
  int size_map = 0;  // ceil(sqrt(2)*width)
  int side     = 0;  // width + 2*pad (footprint rows == cols)
  int n_active = 0;
  // max over the active cells of |(|u| + 0.5, |v| + 0.5)|, (u, v) = cell - map centre: no pixel farther than this from
  // the imprint centre can map to an active cell. compact = it is <= wr - 2, i.e. every pixel an imprint touches lies in
  // the open interior of its footprint box, disjoint from the snapshot ring of the same imprint.
  double reach = 0.0;
  bool compact = false;
  uint32_t* d_xy = nullptr;  // device
  void* d_fh     = nullptr;  // device, context element type
};

constexpr int kMaxBands = 8;  // GPUs of one NVSwitch node
constexpr int kTraceImprints = 256, kTraceStamps = 8;
constexpr int kStrokeDone = 0x7fffffff;  // progress value of a finished stroke

// DevStroke::flags
constexpr int kStrokeLoadPick  = 1;   // load the pickup state from the brush's dense map (continue without dip)
constexpr int kStrokeStorePick = 2;   // store it back after the stroke
constexpr int kStrokeWindows   = 8;   // straddling stroke (multi GPU): every dataflow segment works on a local staging window
constexpr int kStrokeTwoPhase  = 16;  // footprint is not compact: ring pass and main pass are separated by a barrier,
                                      // every ring pass scans the whole ring, hits are decided in f64 only

struct alignas(16) DevStroke {
  int64_t first_imprint;
  int32_t n_imprints;
  int32_t n_active;
  const uint32_t* xy;
  const void* fh;
  int32_t size_map, side;
  double radius;  // FootprintBrush::_radius as used by updateSnapshot (:298-305)
  double paintK[3], paintS[3];
  int32_t seg_begin;  // first entry of this stroke in seg_off[] / windows[] (one entry per segment)
  int32_t seg_len;    // imprints per dataflow segment (>= 1)
  int32_t flags;      // kStroke*
  float eps;          // half width of the undecided band of the single-precision hit test (imprint_geom.hpp)
  int32_t flag_index;        // progress word of this stroke in the executor's flag array (its number among the rank's strokes)
  // straddling strokes (multi GPU): the frame of the stroke's staging window = its whole region on the canvas
  // (x0 and cols multiples of 4); the segments' rectangles (DevWindow) lie inside it
  int32_t win_x0, win_y0, win_cols, win_rows;
  int32_t pad[3];            // 144 bytes: the kernel copies the record into shared memory in 16-byte pieces
};
static_assert(sizeof(DevStroke) == 144, "DevStroke layout");

// Multi-GPU staging of a straddling stroke. The stroke owns a local WINDOW that mirrors its whole region of the canvas
// (DevStroke::win_*: rows of the executor's own band and of its neighbours alike) and gives the imprint chain one uniform
// view. Per dataflow segment (DevWindow = the segment's region: union of the allowed boxes of its imprints, clipped to
// the canvas, x0 and cols multiples of 4) the window is brought up to date when the segment starts and the touched pixels
// are written back when it ends. Bringing up to date = pulling the whole segment region after a wait on other strokes
// (they may have written into it), otherwise only the part that was not in the previous segment's region — consecutive
// segments overlap by ~90 %, and without an intervening writer the window is the truth for everything it already holds.
struct DevWindow {
  int32_t x0, y0, cols, rows;
};

struct ImprintLaunch {
  // The engine works on pixel RECORDS (8 elements per pixel: Kr Kg Kb Sr Sg Sb V 0 — 32 bytes in FP32 mode, one
  // 256-bit access and one L2 sector per pixel): a record copy of the canvas' wet layer (converted from / to the SoA
  // planes around a batch) and the snapshot buffer, per row band. Single GPU: one band (n_bands = 1) holding the rows
  // [store_first, store_first + store_rows). Multi GPU (n_bands > 1): band b holds the rows
  // [b*rows_per_band, min((b+1)*rows_per_band, rows)) and lives in GPU b's HBM; the pointers of the other bands are
  // peer mappings (CUDA IPC) reached through NVLink.
  void* canvas[kMaxBands];          // canvas records
  void* snapshot[kMaxBands];        // snapshot records; == canvas when the snapshot buffer is disabled
  unsigned char* dirty[kMaxBands];  // 1 byte per pixel (flat, same index as the records): snapshot may differ from canvas
  // the executor's own band once more (== index my_band above): strokes that stay inside it use these fields
  void* own_canvas;
  void* own_snapshot;
  unsigned char* own_dirty;
  int n_bands, rows_per_band, my_band;
  int use_snapshot;
  int rows, cols;               // logical canvas size (bounds checks)
  int store_first, store_rows;  // stored row window (single band only)
  // brush constants
  double pickup_rate, deposition_rate, capacity;
  // dense pickup map of the brush (7 planes of size_map^2), used by strokes with flags
  void* pick_dense[kLayerPlanes];
  // work
  const DevStroke* strokes;
  int64_t n_strokes;
  const DevImprint* imprints;
  // Dataflow: segment j of the launch waits for preds[seg_off[j] .. seg_off[j+1]) = (stroke, segments needed);
  // stroke = flag index (single GPU) or (rank << 27) | flag index on that rank (multi GPU).
  const int2* preds;
  const int32_t* seg_off;
  const DevWindow* windows;      // per segment (multi GPU, strokes with kStrokeWindows); may be null otherwise
  long long* done[kMaxBands];    // per-rank progress words ([my_band] is local): (epoch << 32) | segments completed
  int epoch;
  // Multi GPU: a run of strokes is split over TWO concurrent launches — the strokes that stay inside the executor's band
  // (kernel variant without the band-view chain: exactly the single-GPU code plus system-scope progress words) and the
  // straddling strokes (variant with the view chain, on a second stream). Compiled into one kernel, the view chain's
  // register pressure slowed EVERY stroke by 40-50 % (r = 151: 20.6 instead of 13.8 us per imprint). views_kernel selects.
  int views_kernel;
  unsigned long long watchdog_ns;  // longest legitimate dataflow wait (0 = no watchdog), see seg_wait
  int* queue;                    // single counter (zeroed): tickets
  const int32_t* order;          // ticket -> stroke of this launch (host-planned claim order); nullptr = identity
  unsigned long long* counters;  // [0] active stroke-pixels
  // diagnostics (pb_fbrush_enable_trace): SM cycle stamps of the first kTraceImprints imprints of stroke 0 of the launch,
  // taken by the first and the last thread of the cluster's rank-0 CTA: [imprint][thread][kTraceStamps]; nullptr = off
  unsigned long long* trace;
  unsigned char* win_scratch;    // staging windows, win_stride bytes per stroke slot
  int64_t win_stride;
  // per-CTA cell state in global memory for footprints that do not fit shared memory
  void* scratch;
  int64_t scratch_stride;  // bytes per CTA
  int cta_cells;           // cell slots per CTA (a multiple of block)
  int cells_in_smem;       // 1: pickup state, heights and cell coordinates live in dynamic shared memory
  int chunk_cells;         // cells per thread whose interactions are listed at a time (list capacity 2 * chunk_cells)
  int ring_threads;        // threads per CTA (the last ones) that run the incremental snapshot-ring pass
  int block;               // threads per CTA
  int cluster;             // CTAs per thread-block cluster
  int grid;                // CTAs (multiple of cluster)
  int policy;              // kShapeLatency / kShapeThroughput (input of imprint_plan)
};

// Launch-shape policies. A stroke is a chain of dependent imprints, so its speed is a latency; a batch, however, is bound
// either by its dependency graph (critical path: spread every stroke over many SMs to cut the latency per imprint) or by
// the number of strokes that can run at once (give every stroke as few SMs as it needs, run many). The host planner
// simulates the batch under both policies with the measured cost curves and takes the faster one (capi.cu).
constexpr int kShapeLatency = 0, kShapeThroughput = 1;
// measured single-stroke latency of the policy's shape, us per imprint (scratch/imprint_sweep.py, brush at 45 degrees)
double imprint_cost_us(int n_active, int policy);

// Launch shape for a run of strokes whose largest footprint has max_active cells: a stroke is owned by a
// thread-block cluster of `cluster` CTAs x `block` threads.
void imprint_plan(pb_context* ctx, int max_active, ImprintLaunch& L, size_t& smem_bytes);  // reads L.n_bands, L.policy
int imprint_cluster_class(int n_active, int policy);
void imprint_launch(pb_context* ctx, const ImprintLaunch& L, size_t smem_bytes, cudaStream_t stream = nullptr);  // nullptr: ctx->stream
// concurrent strokes of a launch shape (resident clusters), for the host's claim-order model
int imprint_slots(const ImprintLaunch& L);

// visited stroke-pixel count (the reference's `counter`, FootprintBrush.hxx:119): one pass over all
// (2hr+1)(2wr+1) cells of every imprint, both bounds checks evaluated exactly in f64.
void imprint_count_visited(pb_context* ctx, const DevStroke* strokes, int64_t n_strokes, const DevImprint* imprints,
                           int rows, int cols, unsigned long long* counter);

}  // namespace pb
