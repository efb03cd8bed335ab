// Device-side interface of the footprint-brush imprint engine (imprint.cu).
#pragma once
#include "common.cuh"

namespace pb {

// One footprint geometry (= one ceil(radius)): the compacted list of active cells (height > 0) of the
// padded footprint. Cell i sits at map position (xy[i] & 0xffff, xy[i] >> 16) and has height fh[i].
struct FootprintGeom {
  int width    = 0;  // 2*ceil(r)+1
  int size_map = 0;  // ceil(sqrt(2)*width)
  int side     = 0;  // width + 2*pad (footprint rows == cols)
  int n_active = 0;
  uint32_t* d_xy = nullptr;  // device
  void* d_fh     = nullptr;  // device, context element type
};

struct DevImprint {  // per-imprint constants, computed on the host in f64 (FootprintBrush.hxx:95-96)
  double cx, cy, c, s;  // c = cos(-theta), s = sin(-theta)
};

constexpr int kMaxBands = 8;  // GPUs of one NVSwitch node
constexpr int kStrokeDone = 0x7fffffff;  // progress value of a finished stroke

struct DevStroke {
  int64_t first_imprint;
  int32_t n_imprints;
  int32_t n_active;
  const uint32_t* xy;
  const void* fh;
  int32_t size_map, side;
  double radius;  // FootprintBrush::_radius as used by updateSnapshot (:298-305)
  double paintK[3], paintS[3];
  int32_t seg_begin;             // first entry of this stroke in seg_off[] (one entry per segment, + 1)
  int32_t seg_len;               // imprints per dataflow segment (>= 1)
  int32_t flags;                 // bit0: load pick state from the dense map, bit1: store it back,
                                 // bit2: rows of another GPU are accessed directly through NVLink (system-scope
                                 //       fences at every barrier), bit3: they are staged in local windows instead
  int32_t pad;
  // Multi-GPU staging windows: the part of the stroke's region that lies in a neighbour's band is pulled into local
  // scratch before the stroke and the touched pixels are pushed back afterwards (one bulk NVLink transfer each way
  // instead of remote round trips on every imprint). Window w covers band-local rows [row0, row0+rows) and canvas
  // columns [win_ox, win_ox + win_cols) of band win_band[w]; win_band[w] < 0 = unused.
  int32_t win_band[2], win_row0[2], win_rows[2];
  int32_t win_ox, win_cols;  // multiples of 4 (the dirty map is scanned in 32-bit words)
};

struct ImprintLaunch {
  // Canvas / snapshot / dirty planes per row band. Single GPU: one band (n_bands = 1) holding the rows
  // [store_first, store_first + store_rows). Multi GPU (n_bands > 1): band b holds the rows
  // [b*rows_per_band, min((b+1)*rows_per_band, rows)) and lives in GPU b's HBM; the pointers of the other bands are
  // peer mappings (CUDA IPC) reached through NVLink.
  void* canvas[kMaxBands][kLayerPlanes];
  void* snapshot[kMaxBands][kLayerPlanes];  // == canvas planes when the snapshot buffer is disabled
  unsigned char* dirty[kMaxBands];          // 1 byte per pixel: snapshot(p) may differ from canvas(p)
  // the executor's own band once more (== index my_band above): strokes that stay inside it use these fields
  void* own_canvas[kLayerPlanes];
  void* own_snapshot[kLayerPlanes];
  unsigned char* own_dirty;
  int n_bands, rows_per_band, my_band;
  int dirty_pitch;
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
  long long* done[kMaxBands];    // per-rank progress words ([my_band] is local): (epoch << 32) | segments completed
  int epoch;
  int flag_offset;               // flag index of this launch's stroke 0 (strokes of earlier launches come first)
  int* queue;                    // single counter (zeroed): tickets
  const int32_t* order;          // ticket -> stroke of this launch (host-planned claim order); nullptr = identity
  unsigned long long* counters;  // [0] active stroke-pixels
  unsigned char* win_scratch;  // staging windows, win_stride bytes per stroke slot (two halves, one per window)
  int64_t win_stride;
  // per-CTA pick scratch in global memory for footprints that do not fit shared memory
  void* scratch;
  int64_t scratch_stride;  // elements per CTA
  int smem_cells;          // cells per CTA that fit in dynamic shared memory
  int block;               // threads per CTA
  int cluster;             // CTAs per thread-block cluster
  int group;               // clusters cooperating on one stroke (software barrier on top of the hardware one)
  unsigned* group_bar;     // per group: monotonic arrival counter of the inter-cluster barrier (zeroed)
  long long* group_stroke; // per group: stroke index popped by the group leader
  int grid;                // CTAs (multiple of cluster)
};

// Launch shape for a run of strokes whose largest footprint has max_active cells: a stroke is owned by a
// thread-block cluster of `cluster` CTAs x `block` threads (~2 active cells per thread).
void imprint_plan(pb_context* ctx, int max_active, ImprintLaunch& L, size_t& smem_bytes);
int imprint_cluster_class(int n_active);
void imprint_launch(pb_context* ctx, const ImprintLaunch& L, size_t smem_bytes);

// visited stroke-pixel count (the reference's `counter`, FootprintBrush.hxx:119): one pass over all
// (2hr+1)(2wr+1) cells of every imprint, both bounds checks evaluated exactly in f64.
void imprint_count_visited(pb_context* ctx, const DevStroke* strokes, int64_t n_strokes, const DevImprint* imprints,
                           int rows, int cols, unsigned long long* counter);

}  // namespace pb
