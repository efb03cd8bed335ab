// C ABI of libpainty_b200.so (include/painty_b200.h): contexts, device-resident canvas / layer /
// brush state, host-side stroke planning, kernel launches. No CPU fallback anywhere: every entry point
// that computes pixels needs a CUDA device and fails loudly otherwise.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "host_math.hpp"
#include "imprint.cuh"
#include "schedule.hpp"
#include "texture.cuh"
#include "texture_host.hpp"

namespace pb {
namespace {
thread_local std::string g_error;
}
void set_error(const std::string& msg) { g_error = msg; }
}  // namespace pb

#define PB_API_BEGIN try {
#define PB_API_END                           \
  return 0;                                  \
  }                                          \
  catch (const std::exception& e) {          \
    pb::set_error(e.what());                 \
    return 1;                                \
  }                                          \
  catch (...) {                              \
    pb::set_error("unknown C++ exception");  \
    return 1;                                \
  }

// entry points that only touch host state: reject a null handle with an error code instead of crashing
#define PB_CHECK_HANDLE(h, name)                      \
  if ((h) == nullptr) {                               \
    pb::set_error(name ": null handle");              \
    return 1;                                         \
  }

using namespace pb;

namespace {

struct DeviceGuard {
  explicit DeviceGuard(const pb_context* ctx) { PB_CUDA(cudaSetDevice(ctx->device)); }
};

template <typename T>
struct DevBuf {  // stream-ordered temporary
  T* p = nullptr;
  cudaStream_t s;
  DevBuf(pb_context* ctx, size_t n) : s(ctx->stream) {
    if (n) PB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&p), n * sizeof(T), s));
  }
  ~DevBuf() {
    if (p) cudaFreeAsync(p, s);
  }
  void upload(const T* h, size_t n) {
    if (n) PB_CUDA(cudaMemcpyAsync(p, h, n * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void zero(size_t n) {
    if (n) PB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
  }
};

uint64_t fnv1a(const void* data, size_t bytes) {
  const unsigned char* p = static_cast<const unsigned char*>(data);
  uint64_t h             = 1469598103934665603ull;
  for (size_t i = 0; i < bytes; ++i) h = (h ^ p[i]) * 1099511628211ull;
  return h;
}

void canvas_clear(pb_canvas* c) {  // Canvas.hxx:37-58: wet layer 0, R0 = background (1), h = 0
  pb_context* ctx = c->pl.ctx;
  const int64_t n = c->pl.n();
  for (int p = 0; p < kLayerPlanes; ++p) fill_plane(ctx, c->pl.plane(p), n, 0.0);
  for (int p = 0; p < 3; ++p) fill_plane(ctx, c->pl.plane(PR + p), n, 1.0);
  fill_plane(ctx, c->pl.plane(PH), n, 0.0);
}

ComposeArgs compose_args(const pb_planes& layer, const pb_planes& r0, int r0_first, void* const out[3], int64_t elem_off,
                         size_t esize) {
  ComposeArgs a;
  for (int k = 0; k < 3; ++k) {
    a.K[k]  = static_cast<char*>(layer.plane(PK + k)) + elem_off * esize;
    a.S[k]  = static_cast<char*>(layer.plane(PS + k)) + elem_off * esize;
    a.R0[k] = static_cast<char*>(r0.plane(r0_first + k)) + elem_off * esize;
    a.R[k]  = out[k];
  }
  a.V = static_cast<char*>(layer.plane(PV)) + elem_off * esize;
  return a;
}

}  // namespace

struct pb_fbrush {
  pb_context* ctx = nullptr;
  double radius   = 0.0;  // FootprintBrush::_radius
  std::map<std::pair<int, uint64_t>, FootprintGeom> geoms;  // (side, content hash) -> compacted footprint
  std::map<int, std::pair<int, uint64_t>> by_width;         // width -> key of the footprint registered for it
  const FootprintGeom* cur = nullptr;
  pb_planes pick;      // dense pickup map, 7 planes of size_map^2
  // The engine's pixel records (8 elements per pixel, imprint.cuh), both canvas sized: the snapshot buffer, allocated
  // at the first imprint (:281-284), and the working copy of the canvas' wet layer the kernels run on (converted from
  // the SoA planes when a batch starts and back when it ends).
  void* snap_rec = nullptr;
  void* work_rec = nullptr;
  int snap_rows = 0, snap_cols = 0, work_rows = 0, work_cols = 0;
  // 1 byte per canvas pixel: snapshot may differ from canvas there (touched by an imprint since its last ring copy)
  unsigned char* dirty = nullptr;  // flat: byte index == pixel index of the stored planes (+ kDirtyPad bytes)
  uint64_t snap_canvas_id = 0, snap_canvas_version = 0;  // canvas state the dirty map is valid for
  // multi-GPU: 64-bit progress words other GPUs poll (kDistFlagCapacity words + 1024 queue counters), batch epoch
  long long* dist_flags = nullptr;
  int dist_epoch        = 0;
  int dist_queue_slot   = 0;
  bool use_snapshot = true;
  double pickup_rate = 0.9, deposition_rate = 0.05, capacity = 1.0;  // :477-495
  double paintK[3] = {0, 0, 0}, paintS[3] = {0, 0, 0};                // zero-initialised
  unsigned long long* d_counters = nullptr;                          // [0] active [1] visited
  unsigned long long* d_trace    = nullptr;                          // diagnostics, pb_fbrush_enable_trace
  bool count_visited             = false;
  // host-side figures of the last stroke / imprint batch (pb_fbrush_batch_stats)
  double stats[PB_BATCH_STATS] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
};

struct pb_tbrush {
  pb_context* ctx = nullptr;
  double radius   = 0.0;
  double thickness_scale = 1.0;                                        // BrushBase.hxx:45
  double paintK[3] = {0.1, 0.1, 0.1}, paintS[3] = {0.1, 0.1, 0.1};      // TextureBrush.hxx:29-31
  // thickness textures: [0] = the stroke sample handed to the constructor (BrushStrokeSample::getThicknessMap), the
  // others were added with pb_tbrush_add_texture (the brush-texture dictionary of TextureBrushDictionary.cxx)
  struct Texture {
    int rows = 0, cols = 0;
    double* d_map = nullptr;
  };
  std::vector<Texture> textures;
  int current_texture = 0;  // used by pb_tbrush_paint_stroke (BrushStrokeSample::setThicknessMap's role)
  unsigned long long* d_counters = nullptr;  // [0] deposited stroke-pixels, [1] scratch: max thickness bits
  // Smudge (renderer/Smudge.hxx): two ping-pong pickup windows of size x size, state carried across strokes
  bool use_smudge       = false;  // the reference defaults to true (TextureBrush.hxx:236), see INTEGRATION.md
  int smudge_size       = 0;      // Smudge(0) as built by TextureBrush's ctor (:27): empty maps, _maxSize = 1
  int smudge_max_size   = 1;
  int smudge_dst        = 0;      // which window currently is _pickupMapDst
  double smudge_rotation = 0.0;   // Smudge::_currentRotation
  pb_planes smudge_map[2];
};

namespace {

const FootprintGeom* register_footprint(pb_fbrush* b, double radius, int side, const double* fp) {
  pb_context* ctx = b->ctx;
  const int width = static_cast<int32_t>(2.0 * std::ceil(radius) + 1.0);  // FootprintBrush.hxx:54-55
  const int size_map = static_cast<int32_t>(std::ceil(std::sqrt(2.0) * width));
  PB_REQUIRE(side >= 0 && side <= 0xffff, "footprint side out of range");
  const auto key = std::make_pair(side, fnv1a(fp, sizeof(double) * side * side) ^ static_cast<uint64_t>(size_map));
  auto it        = b->geoms.find(key);
  if (it == b->geoms.end()) {
    FootprintGeom g;
    g.width    = width;
    g.size_map = size_map;
    g.side     = side;
    std::vector<uint32_t> xy;
    std::vector<double> fh;
    // reads at map index >= side are out of range in the reference (B#2) -> height 0; cells whose
    // index is >= size_map are rejected by the reference's map bounds check
    const int lim = std::min(side, size_map);
    for (int my = 0; my < lim; ++my)
      for (int mx = 0; mx < lim; ++mx) {
        const double h = fp[static_cast<size_t>(my) * side + mx];
        if (h > 0.0) {
          xy.push_back(static_cast<uint32_t>(my) << 16 | static_cast<uint32_t>(mx));
          fh.push_back(h);
        }
      }
    g.n_active = static_cast<int>(xy.size());
    {  // how far from the imprint centre a touched pixel can lie (imprint.cuh: FootprintGeom::reach)
      const int wr = (side - 1) / 2;
      double reach = 0.0;
      for (uint32_t c : xy) {
        const double u = std::fabs(static_cast<double>(static_cast<int>(c & 0xffffu) - wr)) + 0.5;
        const double v = std::fabs(static_cast<double>(static_cast<int>(c >> 16) - wr)) + 0.5;
        reach          = std::max(reach, std::sqrt(u * u + v * v));
      }
      g.reach   = reach;
      g.compact = reach <= static_cast<double>(wr - 2) && std::getenv("PB_IMPRINT_TWO_PHASE") == nullptr;
    }
    if (std::getenv("PB_CELL_ORDER") == nullptr || std::atoi(std::getenv("PB_CELL_ORDER")) != 0) {
      // Order the compacted cells by 8x4 tiles of the pickup map: a warp (32 consecutive cells) then covers a compact
      // 2-D patch whose rotated image touches far fewer 32 B sectors of the SoA canvas planes than a 32-cell row
      // segment does at oblique angles. Cells are thread-private, so their order is free.
      std::vector<size_t> order(xy.size());
      for (size_t i = 0; i < order.size(); ++i) order[i] = i;
      auto key = [&](size_t i) {
        const uint64_t mx = xy[i] & 0xffffu, my = xy[i] >> 16;
        return ((my >> 2) << 40) | ((mx >> 3) << 20) | ((my & 3) << 3) | (mx & 7);
      };
      std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return key(a) < key(b); });
      std::vector<uint32_t> xy2(xy.size());
      std::vector<double> fh2(fh.size());
      for (size_t i = 0; i < order.size(); ++i) {
        xy2[i] = xy[order[i]];
        fh2[i] = fh[order[i]];
      }
      xy.swap(xy2);
      fh.swap(fh2);
    }
    if (g.n_active) {
      PB_CUDA(cudaMalloc(reinterpret_cast<void**>(&g.d_xy), sizeof(uint32_t) * xy.size()));
      PB_CUDA(cudaMemcpyAsync(g.d_xy, xy.data(), sizeof(uint32_t) * xy.size(), cudaMemcpyHostToDevice, ctx->stream));
      PB_CUDA(cudaMalloc(&g.d_fh, ctx->esize() * fh.size()));
      if (ctx->precision == PB_F64) {
        PB_CUDA(cudaMemcpyAsync(g.d_fh, fh.data(), sizeof(double) * fh.size(), cudaMemcpyHostToDevice, ctx->stream));
      } else {
        std::vector<float> f32(fh.begin(), fh.end());
        PB_CUDA(cudaMemcpyAsync(g.d_fh, f32.data(), sizeof(float) * f32.size(), cudaMemcpyHostToDevice, ctx->stream));
      }
      PB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    it = b->geoms.emplace(key, g).first;
  }
  b->by_width[width] = key;
  return &it->second;
}

void brush_set_geometry(pb_fbrush* b, double radius, const FootprintGeom* g) {
  b->radius = radius;
  b->cur    = g;
  if (b->pick.rows != g->size_map) {
    planes_free(b->pick);
    planes_alloc(b->ctx, b->pick, g->size_map, g->size_map, kLayerPlanes);
  }
  for (int p = 0; p < kLayerPlanes; ++p) fill_plane(b->ctx, b->pick.plane(p), b->pick.n(), 0.0);
}

// the ring pass reads the dirty map in aligned 32-bit words that may extend a few bytes past the last pixel
constexpr size_t kDirtyPad = 64;
size_t dirty_bytes(const pb_canvas* c) { return static_cast<size_t>(c->pl.n()) + kDirtyPad; }

size_t record_bytes(const pb_canvas* c) { return std::max<size_t>(static_cast<size_t>(c->pl.n()) * kRecord * c->pl.ctx->esize(), 256); }

void ensure_work(pb_fbrush* b, pb_canvas* c) {
  if (b->work_rec != nullptr && b->work_rows == c->pl.rows && b->work_cols == c->pl.cols) return;
  PB_CUDA(cudaStreamSynchronize(b->ctx->stream));
  if (b->work_rec) cudaFree(b->work_rec);
  b->work_rec = nullptr;
  PB_CUDA(cudaMalloc(&b->work_rec, record_bytes(c)));
  b->work_rows = c->pl.rows;
  b->work_cols = c->pl.cols;
}

bool snapshot_matches(const pb_fbrush* b, const pb_canvas* c) {
  return b->snap_rec != nullptr && b->snap_rows == c->pl.rows && b->snap_cols == c->pl.cols;
}

void ensure_snapshot(pb_fbrush* b, pb_canvas* c) {  // FootprintBrush.hxx:281-284
  pb_context* ctx = b->ctx;
  if (!snapshot_matches(b, c)) {
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (b->snap_rec) cudaFree(b->snap_rec);
    if (b->dirty) cudaFree(b->dirty);
    b->snap_rec = nullptr;
    b->dirty    = nullptr;
    PB_CUDA(cudaMalloc(&b->snap_rec, record_bytes(c)));
    b->snap_rows = c->pl.rows;
    b->snap_cols = c->pl.cols;
    planes_to_records(ctx, c->pl, b->snap_rec, 0, 0, c->pl.cols - 1, c->pl.rows - 1);  // canvas.copyTo(snapshot)
    const size_t bytes = dirty_bytes(c);
    PB_CUDA(cudaMalloc(reinterpret_cast<void**>(&b->dirty), bytes));
    PB_CUDA(cudaMemsetAsync(b->dirty, 0, bytes, ctx->stream));
  } else if (b->snap_canvas_id != c->id || b->snap_canvas_version != c->version) {
    // the canvas changed behind this brush's back (clear / dry / upload / another brush / another canvas of
    // the same size): every pixel may now differ from the snapshot
    PB_CUDA(cudaMemsetAsync(b->dirty, 1, dirty_bytes(c), ctx->stream));
  }
  b->snap_canvas_id      = c->id;
  b->snap_canvas_version = c->version;
}

// Granularity of the dependency planner's tile tables in pixels (PB_PLAN_TILE, default 32): regions are rounded outward to
// tiles, so smaller tiles mean fewer false dependencies between neighbouring strokes and more planning work (4K bench:
// 3.17 s per step and 68 ms of planning at 64 px, 3.08 s / 108 ms at 32 px, 3.09 s / 253 ms at 16 px).
int plan_tile() {
  const char* e = std::getenv("PB_PLAN_TILE");
  const int t   = e ? std::atoi(e) : 32;
  return std::min(512, std::max(8, t));
}

struct HostStroke {
  const FootprintGeom* g;
  double radius;
  double K[3], S[3];
  int64_t first, n;
  int flags;
};

// Multi-GPU view of one logical canvas: band b (rows [b*rows_per_band, ...)) lives on GPU b; the base pointers of
// the other ranks' allocations are CUDA-IPC peer mappings.
struct DistInfo {
  int world = 1, rank = 0, rows_per_band = 0;
  void* canvas_base[kMaxBands]   = {};
  int64_t canvas_stride[kMaxBands] = {};
  void* snapshot_base[kMaxBands] = {};
  int64_t snapshot_stride[kMaxBands] = {};
  unsigned char* dirty_base[kMaxBands] = {};
  long long* flags_base[kMaxBands] = {};
};
constexpr int64_t kDistFlagCapacity = int64_t(1) << 22;

// Host-side plan of a stroke batch: everything run_plan needs that does not depend on device state — the dataflow graph,
// the claim order, per-run stroke records / wait lists / staging windows and the per-imprint constants of the strokes this
// rank executes. Planning touches no stream, so the plan of the next batch can be made (on another host thread) while the
// device still executes the current one.
struct RunPlan {
  size_t begin = 0, end = 0;  // range in `mine`
  std::vector<DevStroke> ds;
  std::vector<int2> preds;
  std::vector<int32_t> seg_off;
  std::vector<DevWindow> windows;  // one entry per segment of the run, parallel to seg_off
  std::vector<int32_t> order;      // queue ticket -> stroke of this run (empty = submission order)
  int max_active    = 1;
  size_t max_window = 0;
  // multi GPU: a run is executed by two concurrent launches — `views` = the one for the straddling strokes; both carry the
  // same group number. share = this launch's part of the run's imprints (splits the resident clusters between the two).
  bool views   = false;
  int group    = 0;
  double share = 1.0;
};
}  // namespace
struct pb_batch_plan {
  bool multi = false;
  int policy = kShapeLatency;  // launch-shape policy the runs were cut for
  int world = 1, rank = 0, rows_per_band = 0;
  int rows = 0, cols = 0, store_first = 0, store_rows = 0;  // the canvas the plan was made for
  bool use_snapshot = true;
  size_t n_mine = 0;
  std::vector<RunPlan> runs;
  std::vector<DevImprint> im;
  Region batch{0, 0, -1, -1};
  double stats[PB_BATCH_STATS] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  // brush state the plan assumes before / leaves after the batch (stroke batches; imprint batches keep the state)
  bool sets_state = false;
  double radius_before = 0.0, radius_after = 0.0;
  const FootprintGeom* geom_after = nullptr;
  double K_after[3] = {0, 0, 0}, S_after[3] = {0, 0, 0};
};
namespace {

// Plans the persistent imprint kernel launches for a submission-ordered stroke list: one launch per run of
// consecutive (local) strokes that share a launch class; dependencies are tracked across runs and — with `dist` —
// across GPUs (every rank plans the same global list and executes the strokes whose first imprint lies in its band).
void plan_imprints(pb_fbrush* b, pb_canvas* c, const std::vector<HostStroke>& hs, int64_t n_imprints, const double* cx,
                   const double* cy, const double* theta, const DistInfo* dist, pb_batch_plan& P) {
  pb_context* ctx = b->ctx;
  PB_REQUIRE(c->pl.ctx == ctx, "canvas and brush belong to different contexts");
  const bool multi = dist != nullptr && dist->world > 1;
  P.multi          = multi;
  P.world          = multi ? dist->world : 1;
  P.rank           = multi ? dist->rank : 0;
  P.rows_per_band  = multi ? dist->rows_per_band : std::max(c->pl.rows, 1);
  P.rows = c->rows, P.cols = c->cols, P.store_first = c->store_first, P.store_rows = c->pl.rows;
  P.use_snapshot = b->use_snapshot;
  if (hs.empty()) return;
  PB_REQUIRE(!multi || b->use_snapshot, "distributed strokes need the snapshot buffer enabled");
  (void)n_imprints;

  // Global plan: executor rank, local numbering, and the dataflow graph at segment granularity (schedule.hpp).
  const size_t n = hs.size();
  static const int kSegmentLength = [] {
    const char* e = std::getenv("PB_IMPRINT_SEGMENT");  // imprints per dataflow segment; 0 = whole strokes
    return e ? std::max(0, std::atoi(e)) : 64;  // 4K bench step, round 2: 16 / 32 / 64 / 128 -> 3.26 / 3.11 / 3.09 / 3.18 s
  }();
  std::vector<int32_t> executor(n, 0), local_index(n, -1);
  std::vector<int32_t> counts(kMaxBands, 0);
  std::vector<char> remote(n, 0);
  std::vector<Region> allowed(multi ? n : 0);
  auto span_of = [&](size_t s) {
    const HostStroke& h = hs[s];
    return StrokeSpan{h.first, h.n, (h.g->side - 1) / 2, h.radius, false};
  };
  const auto t_plan0 = std::chrono::steady_clock::now();
  double model_makespan = 0.0;
  Region batch{c->cols, c->rows, -1, -1};
  const SegmentPlan plan = plan_segments(c->rows, c->cols, n, span_of, cx, cy, kSegmentLength, b->use_snapshot,
                                         [&](size_t s, const Region&, const Region& r) {
    const HostStroke& h = hs[s];
    if (r.x1 >= r.x0 && r.y1 >= r.y0) {  // everything the batch reads or writes
      batch.x0 = std::min(batch.x0, r.x0), batch.y0 = std::min(batch.y0, r.y0);
      batch.x1 = std::max(batch.x1, r.x1), batch.y1 = std::max(batch.y1, r.y1);
    }
    if (multi && h.n > 0) {
      const int y0 = std::min(std::max(static_cast<int>(cy[h.first]), 0), c->rows - 1);
      executor[s]  = std::min(y0 / dist->rows_per_band, dist->world - 1);
      const int b0 = executor[s] * dist->rows_per_band, b1 = std::min(b0 + dist->rows_per_band, c->rows) - 1;
      remote[s]    = (r.y1 >= r.y0) && (r.y0 < b0 || r.y1 > b1);
      allowed[s]   = r;
    }
    local_index[s] = counts[executor[s]]++;
  }, plan_tile());
  const int my_rank = multi ? dist->rank : 0, world = multi ? dist->world : 1;
  std::vector<std::vector<size_t>> locals(static_cast<size_t>(world));  // per rank: its strokes in submission order
  for (size_t s = 0; s < n; ++s) locals[executor[s]].push_back(s);
  const std::vector<size_t>& mine = locals[my_rank];
  // a run = consecutive local strokes of one launch class = one kernel launch
  // (one launch per class: sharing a launch between the two cluster classes, sized for the larger footprints, was
  // measured slower — 7.43 vs 7.22 s on the 10k-stroke workload — although it removes a kernel boundary)
  int policy = kShapeLatency;  // launch-shape policy of the batch (imprint.cuh), chosen below
  auto launch_class = [&](size_t s) { return imprint_cluster_class(hs[s].g->n_active, policy); };
  auto split_runs = [&](const std::vector<size_t>& list) {
    std::vector<std::pair<size_t, size_t>> runs;
    for (size_t a = 0; a < list.size();) {
      const int cls = launch_class(list[a]);
      size_t e      = a + 1;
      while (e < list.size() && launch_class(list[e]) == cls) ++e;
      runs.emplace_back(a, e);
      a = e;
    }
    return runs;
  };
  auto run_max_active = [&](const std::vector<size_t>& list, const std::pair<size_t, size_t>& run) {
    int m = 1;
    for (size_t k = run.first; k < run.second; ++k) m = std::max(m, hs[list[k]].g->n_active);
    return m;
  };

  // Claim order of the device queues (schedule.hpp): list-schedule the whole batch, all ranks included, with the measured
  // cost curve of the launch shapes (imprint_cost_us) and pop the strokes of every launch in the order the model started
  // them. The same simulation picks the launch-shape policy: the batch is planned under the latency shapes and under the
  // throughput shapes, and the policy with the shorter model makespan runs (FP32 only: the throughput shapes keep the
  // whole cell state of a stroke in one CTA's shared memory). PB_IMPRINT_POLICY = latency | throughput forces one.
  static const bool kReorder = [] {
    const char* e = std::getenv("PB_IMPRINT_REORDER");
    return e == nullptr || std::atoi(e) != 0;
  }();
  static const int kForcedPolicy = [] {
    const char* e = std::getenv("PB_IMPRINT_POLICY");
    if (e == nullptr) return -1;
    return std::strcmp(e, "throughput") == 0 ? kShapeThroughput : (std::strcmp(e, "latency") == 0 ? kShapeLatency : -1);
  }();
  std::vector<int32_t> claim_pos;  // global stroke -> position in the claim sequence (empty = submission order)
  std::vector<int64_t> counts64(n);
  for (size_t s = 0; s < n; ++s) counts64[s] = hs[s].n;
  constexpr double kViewsCost = 1.75;
  auto simulate = [&](int pol, std::vector<int32_t>& pos) -> double {  // model makespan of the batch under policy `pol`
    policy = pol;
    std::vector<ClaimSpec> spec(n);
    std::vector<std::vector<int>> slots(static_cast<size_t>(world));
    for (int r = 0; r < world; ++r) {
      const auto runs = split_runs(locals[r]);
      for (size_t j = 0; j < runs.size(); ++j) {
        ImprintLaunch L{};
        L.n_bands   = world;
        L.policy    = pol;
        size_t smem = 0;
        imprint_plan(ctx, run_max_active(locals[r], runs[j]), L, smem);
        const int64_t n_run = static_cast<int64_t>(runs[j].second - runs[j].first);
        slots[r].push_back(static_cast<int>(std::min<int64_t>(imprint_slots(L), n_run)));
        for (size_t k = runs[j].first; k < runs[j].second; ++k) {
          const size_t s = locals[r][k];
          // straddling strokes run in the kernel variant with the band-view chain: ~1.75x the latency per imprint
          // (scratch/dist_micro.py: r = 151 along a band boundary 23.2 us against 12.4 us inside the band)
          spec[s] = ClaimSpec{r, static_cast<int32_t>(j), imprint_cost_us(hs[s].g->n_active, pol) * ((multi && remote[s]) ? kViewsCost : 1.0)};
        }
      }
    }
    double makespan = 0.0;
    const std::vector<int32_t> seq = plan_claim_order(plan, counts64, spec, slots, 64, &makespan);
    pos.assign(n, -1);
    for (size_t q = 0; q < seq.size(); ++q) pos[seq[q]] = static_cast<int32_t>(q);
    // the device relies on the order being topological; fall back to submission order otherwise
    bool ok = seq.size() == n;
    for (size_t s = 0; s < n && ok; ++s) {
      ok = pos[s] >= 0;
      for (int32_t p = plan.seg_off[plan.seg_first[s]]; p < plan.seg_off[plan.seg_first[s + 1]] && ok; ++p)
        ok = pos[plan.pred_stroke[p]] < pos[s];
    }
    if (!ok) pos.clear();
    return makespan;
  };
  const bool may_choose = ctx->precision == PB_F32 && kForcedPolicy < 0;
  if (kReorder && n > 1) {
    model_makespan = simulate(kShapeLatency, claim_pos);
    if (may_choose) {
      std::vector<int32_t> pos_t;
      const double mk_t = simulate(kShapeThroughput, pos_t);
      if (mk_t < model_makespan) {
        model_makespan = mk_t;
        claim_pos.swap(pos_t);
      } else {
        policy = kShapeLatency;
      }
    }
  }
  if (kForcedPolicy >= 0 && ctx->precision == PB_F32 && policy != kForcedPolicy) {
    if (kReorder && n > 1)
      model_makespan = simulate(kForcedPolicy, claim_pos);
    else
      policy = kForcedPolicy;
  }
  P.policy = policy;
  const std::vector<std::pair<size_t, size_t>> my_runs = split_runs(mine);
  PB_REQUIRE(static_cast<int64_t>(mine.size()) <= kDistFlagCapacity, "too many strokes in one batch");

  const auto t_plan1 = std::chrono::steady_clock::now();

  // per-imprint constants of the strokes this rank executes, compacted in execution order
  std::vector<DevImprint> im;
  std::vector<int64_t> local_first(mine.size());
  {
    size_t total = 0;
    for (size_t k = 0; k < mine.size(); ++k) total += static_cast<size_t>(hs[mine[k]].n);
    im.resize(total);
    size_t o = 0;
    for (size_t k = 0; k < mine.size(); ++k) {
      local_first[k] = static_cast<int64_t>(o);
      o += static_cast<size_t>(hs[mine[k]].n);
    }
    // cos / sin of every imprint in f64 with the host's libm (the reference's values): independent per imprint, so the
    // strokes are dealt out to a few host threads (PB_HOST_THREADS, default min(cores, 8)); same results for any count
    static const unsigned kHostThreads = [] {
      const char* e = std::getenv("PB_HOST_THREADS");
      const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
      return e ? static_cast<unsigned>(std::max(1, std::atoi(e))) : std::min(hw, 8u);
    }();
    const unsigned n_threads = total < (1u << 16) ? 1u : kHostThreads;
    auto work = [&](unsigned t) {
      for (size_t k = t; k < mine.size(); k += n_threads) {
        const HostStroke& h = hs[mine[k]];
        const int wr        = (h.g->side - 1) / 2;
        DevImprint* out     = im.data() + local_first[k];
        for (int64_t i = 0; i < h.n; ++i) out[i] = make_imprint(cx[h.first + i], cy[h.first + i], theta[h.first + i], wr);  // :95-96
      }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < n_threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
  }
  const auto t_prep1 = std::chrono::steady_clock::now();
  {
    auto ms = [](auto a, auto b2) { return std::chrono::duration<double, std::milli>(b2 - a).count(); };
    P.stats[0] = ms(t_plan0, t_plan1);                                   // dataflow planning (segments + claim order)
    P.stats[1] = ms(t_plan1, t_prep1);                                   // per-imprint constants (cos / sin)
    P.stats[2] = static_cast<double>(n);                                 // strokes planned (all ranks)
    P.stats[3] = static_cast<double>(plan.seg_first[n]);                 // dataflow segments
    P.stats[4] = static_cast<double>(plan.pred_stroke.size());           // wait entries
    P.stats[5] = model_makespan * 1e-3;                                  // the planner's model of the batch, ms
    P.stats[6] = static_cast<double>(mine.size());                       // strokes this rank executes
    P.stats[7] = static_cast<double>(my_runs.size());                    // kernel launches of this rank
  }
  P.im.swap(im);
  P.batch  = batch;
  P.n_mine = mine.size();

  for (const auto& my_run : my_runs) {
    const size_t run_begin = my_run.first, run_end = my_run.second;
    const size_t n_run = run_end - run_begin;
    P.runs.emplace_back();
    RunPlan& RP = P.runs.back();
    RP.begin = run_begin, RP.end = run_end;
    std::vector<DevStroke>& ds = RP.ds;
    ds.resize(n_run);
    std::vector<int2>& run_preds = RP.preds;
    std::vector<int32_t>& run_seg_off = RP.seg_off;
    run_seg_off.assign(1, 0);
    std::vector<DevWindow>& run_windows = RP.windows;
    int& max_active = RP.max_active;
    size_t& max_window = RP.max_window;
    for (size_t k = 0; k < n_run; ++k) {
      const size_t s      = mine[run_begin + k];
      const HostStroke& h = hs[s];
      DevStroke& d        = ds[k];
      d.first_imprint     = local_first[run_begin + k];
      d.n_imprints        = static_cast<int32_t>(h.n);
      d.n_active          = h.g->n_active;
      d.xy                = h.g->d_xy;
      d.fh                = h.g->d_fh;
      d.size_map          = h.g->size_map;
      d.side              = h.g->side;
      d.radius            = h.radius;
      for (int q = 0; q < 3; ++q) {
        d.paintK[q] = h.K[q];
        d.paintS[q] = h.S[q];
      }
      d.flags    = h.flags;
      if (!h.g->compact) d.flags |= kStrokeTwoPhase;
      // half width of the undecided band of the single-precision hit test: 3x the error bound of imprint_geom.hpp
      d.eps      = static_cast<float>(1e-6 * ((h.g->side - 1) / 2) + 2e-5);
      d.flag_index = static_cast<int32_t>(run_begin + k);
      const int nseg = plan.seg_first[s + 1] - plan.seg_first[s];
      const size_t win_first = run_windows.size();
      run_windows.resize(win_first + static_cast<size_t>(nseg), DevWindow{0, 0, 0, 0});
      d.win_x0 = d.win_y0 = d.win_cols = d.win_rows = 0;
      if (multi && remote[s]) {
        // The stroke's staging window mirrors its whole region (own rows and neighbour rows); per dataflow segment the
        // kernel brings the segment's region inside it up to date (imprint.cuh: DevWindow).
        const Region& r = allowed[s];
        const int half  = (h.g->side - 1) / 2;
        if (r.x1 >= r.x0 && r.y1 >= r.y0) {
          d.win_x0   = r.x0 & ~3;
          d.win_cols = ((r.x1 - d.win_x0 + 1) + 3) & ~3;
          d.win_y0   = r.y0;
          d.win_rows = r.y1 - r.y0 + 1;
        }
        for (int k2 = 0; k2 < nseg; ++k2) {
          Region sbox, sall;
          const int64_t len = plan.seg_len[s];
          imprint_regions(h.first + k2 * len, std::min<int64_t>(len, h.n - k2 * len), half, h.radius, cx, cy, c->rows, c->cols, sbox,
                          sall);
          DevWindow& w = run_windows[win_first + static_cast<size_t>(k2)];
          if (sall.y1 < sall.y0 || sall.x1 < sall.x0) continue;  // nothing on the canvas: an empty rectangle
          w.x0   = sall.x0 & ~3;
          w.cols = ((sall.x1 - w.x0 + 1) + 3) & ~3;
          w.y0   = sall.y0;
          w.rows = sall.y1 - sall.y0 + 1;
        }
        const size_t bytes = static_cast<size_t>(d.win_rows) * d.win_cols * (2 * kRecord * ctx->esize() + 2) + 64;
        PB_REQUIRE(bytes <= (size_t(2) << 30), "stroke too large for a multi-GPU staging window");
        d.flags |= kStrokeWindows;
        max_window = std::max(max_window, (bytes + 255) / 256 * 256);
      }
      max_active = std::max(max_active, d.n_active);
      d.seg_begin = static_cast<int32_t>(run_seg_off.size()) - 1;
      d.seg_len   = plan.seg_len[s];
      for (int32_t gseg = plan.seg_first[s]; gseg < plan.seg_first[s + 1]; ++gseg) {
        for (int32_t p = plan.seg_off[gseg]; p < plan.seg_off[gseg + 1]; ++p) {
          const int32_t g = plan.pred_stroke[p];
          run_preds.push_back(make_int2(multi ? ((executor[g] << 27) | local_index[g]) : local_index[g], plan.pred_need[p]));
        }
        run_seg_off.push_back(static_cast<int32_t>(run_preds.size()));
      }
    }

    {  // resident clusters of this run's launch shape (diagnostics)
      ImprintLaunch Lq{};
      Lq.n_bands   = multi ? dist->world : 1;
      Lq.policy    = policy;
      size_t smemq = 0;
      imprint_plan(ctx, max_active, Lq, smemq);
      P.stats[max_active <= 256 ? 8 : (max_active <= 4096 ? 9 : 10)] = static_cast<double>(imprint_slots(Lq));
      P.stats[11] = std::max(P.stats[11], static_cast<double>(Lq.block) * Lq.cluster);
    }
    if (!claim_pos.empty()) {
      RP.order.resize(n_run);
      for (size_t k = 0; k < n_run; ++k) RP.order[k] = static_cast<int32_t>(k);
      std::sort(RP.order.begin(), RP.order.end(),
                [&](int32_t a, int32_t b2) { return claim_pos[mine[run_begin + a]] < claim_pos[mine[run_begin + b2]]; });
    }
    RP.group = static_cast<int>(P.runs.size());
    if (multi) {
      // Split the run: strokes inside the band -> the launch without the view chain, straddling strokes -> the launch
      // with it. Each part keeps the run's claim order among its own strokes.
      auto is_views = [&](size_t k) { return (RP.ds[k].flags & kStrokeWindows) != 0; };
      size_t n_views = 0;
      double imprints_all = 0.0, imprints_views = 0.0;  // work = imprints x modelled latency of the footprint
      for (size_t k = 0; k < n_run; ++k) {
        const double w = RP.ds[k].n_imprints * imprint_cost_us(RP.ds[k].n_active, policy);
        imprints_all += w;
        if (is_views(k)) {
          ++n_views;
          imprints_views += w;
        }
      }
      if (n_views == n_run) {
        RP.views = true;
      } else if (n_views > 0) {
        auto subset = [&](bool want_views) {
          RunPlan S;
          S.begin = RP.begin, S.end = RP.end, S.max_active = RP.max_active, S.group = RP.group, S.views = want_views;
          S.max_window = want_views ? RP.max_window : 0;
          S.seg_off.assign(1, 0);
          std::vector<int32_t> new_index(n_run, -1);
          for (size_t k = 0; k < n_run; ++k) {
            if (is_views(k) != want_views) continue;
            new_index[k]     = static_cast<int32_t>(S.ds.size());
            DevStroke d      = RP.ds[k];
            const int nseg   = (k + 1 < n_run ? RP.ds[k + 1].seg_begin : static_cast<int32_t>(RP.seg_off.size()) - 1) - d.seg_begin;
            const int old_sb = d.seg_begin;
            d.seg_begin      = static_cast<int32_t>(S.seg_off.size()) - 1;
            for (int g = 0; g < nseg; ++g) {
              for (int32_t p = RP.seg_off[old_sb + g]; p < RP.seg_off[old_sb + g + 1]; ++p) S.preds.push_back(RP.preds[p]);
              S.seg_off.push_back(static_cast<int32_t>(S.preds.size()));
              S.windows.push_back(RP.windows[static_cast<size_t>(old_sb + g)]);
            }
            S.ds.push_back(d);
          }
          for (int32_t k : RP.order)
            if (new_index[static_cast<size_t>(k)] >= 0) S.order.push_back(new_index[static_cast<size_t>(k)]);
          // share of the run's work: imprints weighted with the view chain's cost factor
          const double w_views = 1.75 * imprints_views, w_main = imprints_all - imprints_views;
          S.share = (want_views ? w_views : w_main) / std::max(w_views + w_main, 1.0);
          return S;
        };
        RunPlan main_part = subset(false), views_part = subset(true);
        P.runs.back() = std::move(main_part);
        P.runs.push_back(std::move(views_part));
      }
    }
  }
}

// Uploads a plan and launches its kernels on the context's stream.
void run_plan(pb_fbrush* b, pb_canvas* c, const pb_batch_plan& P, const DistInfo* dist) {
  pb_context* ctx = b->ctx;
  PB_REQUIRE(c->pl.ctx == ctx, "canvas and brush belong to different contexts");
  const bool multi = dist != nullptr && dist->world > 1;
  PB_REQUIRE(multi == P.multi && (!multi || (dist->world == P.world && dist->rank == P.rank && dist->rows_per_band == P.rows_per_band)),
             "batch plan was made for another band layout");
  PB_REQUIRE(P.rows == c->rows && P.cols == c->cols && P.store_first == c->store_first && P.store_rows == c->pl.rows,
             "batch plan was made for another canvas shape");
  PB_REQUIRE(P.use_snapshot == b->use_snapshot, "batch plan was made with another snapshot-buffer setting");
  for (int i = 0; i < PB_BATCH_STATS; ++i) b->stats[i] = P.stats[i];
  if (b->use_snapshot || snapshot_matches(b, c)) ensure_snapshot(b, c);
  ensure_work(b, c);
  const bool have_dirty = b->dirty != nullptr && snapshot_matches(b, c);
  PB_REQUIRE(!multi || (have_dirty && b->use_snapshot), "distributed strokes need the snapshot buffer enabled");
  const std::vector<DevImprint>& im = P.im;
  const Region& batch               = P.batch;
  DevBuf<DevImprint> d_im(ctx, im.size());
  d_im.upload(im.data(), im.size());

  // completion flags: a per-batch buffer on one GPU, the brush's exported buffer + a fresh epoch across GPUs
  DevBuf<long long> d_flags(ctx, multi ? 0 : P.n_mine + 1);
  int epoch = 1;
  if (multi) {
    epoch = ++b->dist_epoch;
  } else {
    d_flags.zero(P.n_mine + 1);
  }

  // Single GPU: the kernels run on the record copy of the wet layer — convert the batch's region (stored rows, columns
  // rounded to 4) on the way in and back on the way out. Multi GPU: the whole band is converted by pb_fbrush_dist_begin /
  // _end around the batch (peers read each other's records, so every rank must be converted before any kernel starts).
  Region conv{0, 0, -1, -1};
  if (!multi && batch.x1 >= batch.x0 && batch.y1 >= batch.y0) {
    conv.x0 = batch.x0 & ~3;
    conv.x1 = std::min(c->cols - 1, batch.x1 | 3);
    conv.y0 = std::max(batch.y0, c->store_first) - c->store_first;
    conv.y1 = std::min(batch.y1, c->store_first + c->pl.rows - 1) - c->store_first;
    planes_to_records(ctx, c->pl, b->work_rec, conv.x0, conv.y0, conv.x1, conv.y1);
  }

  // device copies of one launch's arrays; they live until the launch (and, for a pair of concurrent launches, the join of
  // the two streams) has been enqueued — the stream-ordered frees then run behind it
  struct LaunchBuffers {
    DevBuf<DevStroke> strokes;
    DevBuf<int2> preds;
    DevBuf<int32_t> seg_off, order;
    DevBuf<char> scratch;
    DevBuf<unsigned char> windows;
    DevBuf<DevWindow> win_desc;
    LaunchBuffers(pb_context* cx_, const RunPlan& RP, size_t scratch_bytes, size_t window_bytes, bool multi_)
        : strokes(cx_, RP.ds.size()), preds(cx_, RP.preds.size()), seg_off(cx_, RP.seg_off.size()), order(cx_, RP.order.size()),
          scratch(cx_, scratch_bytes), windows(cx_, window_bytes), win_desc(cx_, multi_ ? RP.windows.size() : 0) {}
  };
  struct Prepared {
    ImprintLaunch L{};
    size_t smem = 0;
    std::unique_ptr<LaunchBuffers> B;
  };
  // Everything of a launch except the launch itself (uploads on the main stream). clusters: upper bound of resident clusters
  // the launch may take (0 = all that fit).
  auto prepare_run = [&](const RunPlan& RP, int clusters) -> Prepared {
    Prepared out;
    const size_t n_run = RP.ds.size();
    ImprintLaunch& L = out.L;
    L.n_bands      = multi ? dist->world : 1;
    L.policy       = P.policy;
    L.views_kernel = (multi && RP.views) ? 1 : 0;
    static const unsigned long long kWatchdogNs = [] {
      const char* e = std::getenv("PB_IMPRINT_WATCHDOG_S");
      return static_cast<unsigned long long>((e ? std::atof(e) : 60.0) * 1e9);
    }();
    L.watchdog_ns  = kWatchdogNs;
    size_t& smem   = out.smem;
    imprint_plan(ctx, RP.max_active, L, smem);
    if (clusters > 0) L.grid = std::min(L.grid, clusters * L.cluster);
    L.grid = static_cast<int>(std::min<int64_t>(L.grid, static_cast<int64_t>(n_run) * L.cluster));
    if (multi) {
      for (int r = 0; r < dist->world; ++r) {
        L.canvas[r]   = dist->canvas_base[r];
        L.snapshot[r] = dist->snapshot_base[r];
        L.dirty[r]    = dist->dirty_base[r];
        L.done[r]     = dist->flags_base[r];
      }
      L.rows_per_band = dist->rows_per_band;
      L.my_band       = dist->rank;
      L.queue         = reinterpret_cast<int*>(b->dist_flags + kDistFlagCapacity + (b->dist_queue_slot++ % 1024));
      PB_CUDA(cudaMemsetAsync(L.queue, 0, sizeof(int), ctx->stream));
    } else {
      L.canvas[0]     = b->work_rec;
      L.snapshot[0]   = b->use_snapshot ? b->snap_rec : b->work_rec;
      L.dirty[0]      = have_dirty ? b->dirty : nullptr;
      L.done[0]       = d_flags.p;
      L.rows_per_band = std::max(c->pl.rows, 1);
      L.my_band       = 0;
      L.queue         = reinterpret_cast<int*>(d_flags.p + P.n_mine);
      if (RP.begin > 0) PB_CUDA(cudaMemsetAsync(L.queue, 0, sizeof(int), ctx->stream));
    }
    L.own_canvas   = L.canvas[L.my_band];
    L.own_snapshot = L.snapshot[L.my_band];
    for (int p = 0; p < kLayerPlanes; ++p) L.pick_dense[p] = b->pick.base ? b->pick.plane(p) : nullptr;
    L.own_dirty       = L.dirty[L.my_band];
    L.epoch           = epoch;
    L.use_snapshot    = b->use_snapshot ? 1 : 0;
    L.rows            = c->rows;
    L.cols            = c->cols;
    L.store_first     = c->store_first;
    L.store_rows      = c->pl.rows;
    L.pickup_rate     = b->pickup_rate;
    L.deposition_rate = b->deposition_rate;
    L.capacity        = b->capacity;
    L.n_strokes       = static_cast<int64_t>(n_run);

    const size_t n_slots = static_cast<size_t>(imprint_slots(L));
    auto B = std::make_unique<LaunchBuffers>(ctx, RP, static_cast<size_t>(L.scratch_stride) * L.grid, RP.max_window * n_slots, multi);
    B->order.upload(RP.order.data(), RP.order.size());
    B->strokes.upload(RP.ds.data(), RP.ds.size());
    B->preds.upload(RP.preds.data(), RP.preds.size());
    B->seg_off.upload(RP.seg_off.data(), RP.seg_off.size());
    if (multi) B->win_desc.upload(RP.windows.data(), RP.windows.size());
    L.order       = RP.order.empty() ? nullptr : B->order.p;
    L.windows     = multi ? B->win_desc.p : nullptr;
    L.win_scratch = B->windows.p;
    L.win_stride  = static_cast<int64_t>(RP.max_window);
    L.scratch     = B->scratch.p;
    L.strokes     = B->strokes.p;
    L.imprints    = d_im.p;
    L.preds       = B->preds.p;
    L.seg_off     = B->seg_off.p;
    L.counters    = b->d_counters;
    L.trace       = b->d_trace;
    out.B         = std::move(B);
    return out;
  };
  auto count_visited = [&](const Prepared& pr) {
    if (b->count_visited) imprint_count_visited(ctx, pr.B->strokes.p, pr.L.n_strokes, d_im.p, c->rows, c->cols, b->d_counters + 1);
  };
  for (size_t i = 0; i < P.runs.size(); ++i) {
    const RunPlan& RP = P.runs[i];
    const bool paired = multi && i + 1 < P.runs.size() && P.runs[i + 1].group == RP.group;
    if (!paired) {
      const Prepared pr = prepare_run(RP, 0);
      imprint_launch(ctx, pr.L, pr.smem, nullptr);
      count_visited(pr);
      continue;
    }
    // the run's two launches run side by side: the straddling strokes on the second stream with their share of the
    // resident clusters, the strokes inside the band on the main stream with the rest
    const RunPlan& VP = P.runs[i + 1];
    if (ctx->aux_stream == nullptr) {
      PB_CUDA(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
      PB_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
      PB_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    }
    ImprintLaunch Lq{};
    Lq.n_bands      = dist->world;
    Lq.policy       = P.policy;
    Lq.views_kernel = 1;
    size_t smemq    = 0;
    imprint_plan(ctx, VP.max_active, Lq, smemq);
    const int total   = imprint_slots(Lq);
    PB_REQUIRE(total >= 2, "multi-GPU stroke batches need room for two resident thread-block clusters of the launch shape");
    const int n_views = std::min<int>(std::max(1, static_cast<int>(std::lround(total * VP.share))), std::max(1, total - 1));
    // The two kernels wait for each other's strokes: nothing that can block the host (a first-use module load, a
    // synchronous copy) may come between their launches — prepare both, then launch back to back.
    const Prepared pv = prepare_run(VP, n_views);
    const Prepared pm = prepare_run(RP, std::max(1, total - n_views));
    PB_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));  // uploads and memsets of both launches precede the forked one
    PB_CUDA(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
    imprint_launch(ctx, pv.L, pv.smem, ctx->aux_stream);
    imprint_launch(ctx, pm.L, pm.smem, nullptr);
    PB_CUDA(cudaEventRecord(ctx->ev_join, ctx->aux_stream));
    PB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    count_visited(pv);
    count_visited(pm);
    ++i;
  }
  if (!multi) records_to_planes(ctx, b->work_rec, c->pl, conv.x0, conv.y0, conv.x1, conv.y1);
  c->version++;
  if (have_dirty) {
    b->snap_canvas_id      = c->id;
    b->snap_canvas_version = c->version;
  }
}

void run_imprints(pb_fbrush* b, pb_canvas* c, const std::vector<HostStroke>& hs, int64_t n_imprints, const double* cx,
                  const double* cy, const double* theta, const DistInfo* dist = nullptr) {
  if (hs.empty()) return;
  pb_batch_plan P;
  plan_imprints(b, c, hs, n_imprints, cx, cy, theta, dist, P);
  run_plan(b, c, P, dist);
}

}  // namespace

extern "C" {

const char* pb_last_error(void) { return g_error.c_str(); }
int pb_version(void) { return 100; }

// ---- context -------------------------------------------------------------------------------------
int pb_context_create(int device, int precision, pb_context** out) {
  PB_API_BEGIN
  // Persistent kernels wait for each other across streams and GPUs; a lazily loaded kernel module would be loaded at its
  // first launch or attribute query, which synchronises the context — behind kernels that may be waiting for exactly that
  // launch. Load everything up front (no effect once the CUDA runtime of this process is initialised).
  setenv("CUDA_MODULE_LOADING", "EAGER", 0);
  PB_REQUIRE(out != nullptr, "pb_context_create: out is null");
  PB_REQUIRE(precision == PB_F32 || precision == PB_F64, "precision must be PB_F32 or PB_F64");
  int count = 0;
  PB_CUDA(cudaGetDeviceCount(&count));
  PB_REQUIRE(device >= 0 && device < count, "no such CUDA device (painty_b200 has no CPU fallback)");
  PB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  PB_CUDA(cudaGetDeviceProperties(&prop, device));
  PB_REQUIRE(prop.major == 10, std::string("painty_b200 is built for sm_100a only; device is sm_") +
                                 std::to_string(prop.major) + std::to_string(prop.minor));
  auto ctx       = std::make_unique<pb_context>();
  ctx->device    = device;
  ctx->precision = precision;
  ctx->sm_count  = prop.multiProcessorCount;
  PB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  *out = ctx.release();
  PB_API_END
}
int pb_context_destroy(pb_context* ctx) {
  PB_API_BEGIN
  if (ctx) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamDestroy(ctx->stream);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    delete ctx;
  }
  PB_API_END
}
int pb_context_synchronize(pb_context* ctx) {
  PB_API_BEGIN
  PB_REQUIRE(ctx != nullptr, "pb_context_synchronize: null handle");
  DeviceGuard g(ctx);
  PB_CUDA(cudaStreamSynchronize(ctx->stream));
  PB_API_END
}
int pb_context_precision(const pb_context* ctx) { return ctx->precision; }
void* pb_context_stream(pb_context* ctx) { return ctx->stream; }
int64_t pb_context_launch_count(const pb_context* ctx) { return ctx->launches; }

// ---- host scalars ----------------------------------------------------------------------------------
int pb_compute_reflectance(const double K[3], const double S[3], const double R0[3], double d, double out[3]) {
  PB_API_BEGIN
  host::compute_reflectance(K, S, R0, d, out);
  PB_API_END
}
int pb_compute_scattering_absorption(const double Rb[3], const double Rw[3], double K[3], double S[3]) {
  PB_API_BEGIN
  PB_REQUIRE(host::compute_scattering_absorption(Rb, Rw, K, S),
             "invalid_argument: on black or white inputs violate one of the following conditions: 0 < black < white < 1");
  PB_API_END
}
int pb_paint_mixed(const double K1[3], const double S1[3], double v1, const double K2[3], const double S2[3], double v2,
                   double K[3], double S[3]) {
  PB_API_BEGIN
  const double inv = 1.0 / (v1 + v2);  // PaintMixer.cxx:539-545
  for (int i = 0; i < 3; ++i) {
    K[i] = ((v1 * K1[i]) + (v2 * K2[i])) * inv;
    S[i] = ((v1 * S1[i]) + (v2 * S2[i])) * inv;
  }
  PB_API_END
}
int pb_paint_mix_single(int n, const double* baseK, const double* baseS, int n_weights, const double* w, double K[3],
                        double S[3]) {
  PB_API_BEGIN
  PB_REQUIRE(n == n_weights, "invalid_argument: Palette size does not match underlying size.");
  double norm = 1.0, sum = 0.0;  // PaintMixer.cxx:332-340
  for (int i = 0; i < n; ++i) sum += w[i];
  if (!(std::fabs(sum - 1.0) < 0.00001)) norm = 1. / sum;
  for (int i = 0; i < 3; ++i) K[i] = S[i] = 0.0;
  for (int l = 0; l < n; ++l)
    for (int i = 0; i < 3; ++i) {
      K[i] += norm * w[l] * baseK[3 * l + i];
      S[i] += norm * w[l] * baseS[3 * l + i];
    }
  PB_API_END
}
int pb_expand_stroke(int mode, int n, const double* path_xy, int64_t capacity, double* cx, double* cy, double* theta,
                     int64_t* n_imprints) {
  PB_API_BEGIN
  PB_REQUIRE(mode == 0 || mode == 1, "pb_expand_stroke: mode must be 0 (library) or 1 (GUI)");
  std::vector<host::Imprint> out;
  host::expand_stroke(mode, reinterpret_cast<const host::V2*>(path_xy), n, out);
  if (n_imprints) *n_imprints = static_cast<int64_t>(out.size());
  for (int64_t i = 0; i < std::min<int64_t>(capacity, static_cast<int64_t>(out.size())); ++i) {
    cx[i]    = out[i].cx;
    cy[i]    = out[i].cy;
    theta[i] = out[i].theta;
  }
  PB_API_END
}

int pb_expand_stroke_batch(int mode, int64_t n_strokes, const int64_t* first_vertex, const int32_t* n_vertices, const double* path_xy,
                           int64_t capacity, double* cx, double* cy, double* theta, int64_t* first_imprint, int64_t* n_imprints,
                           int64_t* total) {
  PB_API_BEGIN
  PB_REQUIRE(mode == 0 || mode == 1, "pb_expand_stroke_batch: mode must be 0 (library) or 1 (GUI)");
  PB_REQUIRE(n_strokes >= 0 && (n_strokes == 0 || (first_vertex && n_vertices && path_xy)), "pb_expand_stroke_batch: null arrays");
  int64_t at = 0;
  std::vector<host::Imprint> out;
  for (int64_t s = 0; s < n_strokes; ++s) {
    out.clear();
    host::expand_stroke(mode, reinterpret_cast<const host::V2*>(path_xy) + first_vertex[s], n_vertices[s], out);
    if (first_imprint) first_imprint[s] = at;
    if (n_imprints) n_imprints[s] = static_cast<int64_t>(out.size());
    for (size_t i = 0; i < out.size(); ++i, ++at) {
      if (at < capacity) {
        cx[at]    = out[i].cx;
        cy[at]    = out[i].cy;
        theta[at] = out[i].theta;
      }
    }
  }
  if (total) *total = at;
  PB_API_END
}

int pb_plan_dependencies(int rows, int cols, int64_t n, const int32_t* box, const int32_t* allowed, int64_t* offsets,
                         int64_t capacity, int32_t* preds, int64_t* n_preds) {
  PB_API_BEGIN
  PB_REQUIRE(rows > 0 && cols > 0 && n >= 0 && n < (int64_t(1) << 31), "pb_plan_dependencies: bad sizes");
  DataflowPlanner planner(rows, cols);
  std::vector<int32_t> all;
  for (int64_t s = 0; s < n; ++s) {
    auto clip = [&](const int32_t* r) {
      return Region{std::max(r[0], 0), std::max(r[1], 0), std::min(r[2], cols - 1), std::min(r[3], rows - 1)};
    };
    int32_t b = 0, e = 0;
    planner.add_footprint(static_cast<int32_t>(s), clip(box + 4 * s), clip(allowed + 4 * s), all, b, e);
    offsets[s]     = b;
    offsets[s + 1] = e;
  }
  if (n == 0) offsets[0] = 0;
  for (int64_t i = 0; i < std::min<int64_t>(capacity, static_cast<int64_t>(all.size())); ++i) preds[i] = all[static_cast<size_t>(i)];
  if (n_preds) *n_preds = static_cast<int64_t>(all.size());
  PB_API_END
}

int pb_plan_segments(int rows, int cols, int64_t n, const int64_t* first, const int64_t* count, const int32_t* side,
                     const double* radius, const unsigned char* single, const double* cx, const double* cy, int segment_length,
                     int use_snapshot,
                     int32_t* seg_first, int32_t* seg_len, int64_t seg_capacity, int32_t* seg_off, int64_t pred_capacity,
                     int32_t* pred_stroke, int32_t* pred_need, int64_t* n_preds) {
  PB_API_BEGIN
  PB_REQUIRE(rows > 0 && cols > 0 && n >= 0 && n < (int64_t(1) << 31), "pb_plan_segments: bad sizes");
  const SegmentPlan plan = plan_segments(
      rows, cols, static_cast<size_t>(n),
      [&](size_t s) { return StrokeSpan{first[s], count[s], (side[s] - 1) / 2, radius[s], single != nullptr && single[s] != 0}; },
      cx, cy, segment_length, use_snapshot != 0, [](size_t, const Region&, const Region&) {}, plan_tile());
  for (int64_t s = 0; s < n; ++s) {
    seg_first[s] = plan.seg_first[static_cast<size_t>(s)];
    seg_len[s]   = plan.seg_len[static_cast<size_t>(s)];
  }
  seg_first[n] = plan.seg_first[static_cast<size_t>(n)];
  for (int64_t i = 0; i < std::min<int64_t>(seg_capacity + 1, static_cast<int64_t>(plan.seg_off.size())); ++i)
    seg_off[i] = plan.seg_off[static_cast<size_t>(i)];
  const int64_t np = static_cast<int64_t>(plan.pred_stroke.size());
  for (int64_t i = 0; i < std::min(pred_capacity, np); ++i) {
    pred_stroke[i] = plan.pred_stroke[static_cast<size_t>(i)];
    pred_need[i]   = plan.pred_need[static_cast<size_t>(i)];
  }
  if (n_preds) *n_preds = np;
  PB_API_END
}

int pb_plan_claim_order(int rows, int cols, int64_t n, const int64_t* first, const int64_t* count, const int32_t* side,
                        const double* radius, const unsigned char* single, const double* cx, const double* cy,
                        int segment_length, int use_snapshot, const int32_t* pool, const int32_t* run, const double* cost,
                        int n_pools, const int32_t* runs_per_pool, const int32_t* slots, int32_t* order, double* makespan) {
  PB_API_BEGIN
  PB_REQUIRE(rows > 0 && cols > 0 && n >= 0 && n < (int64_t(1) << 31) && n_pools >= 1, "pb_plan_claim_order: bad sizes");
  const SegmentPlan plan = plan_segments(
      rows, cols, static_cast<size_t>(n),
      [&](size_t s) { return StrokeSpan{first[s], count[s], (side[s] - 1) / 2, radius[s], single != nullptr && single[s] != 0}; },
      cx, cy, segment_length, use_snapshot != 0, [](size_t, const Region&, const Region&) {}, plan_tile());
  std::vector<std::vector<int>> sl(static_cast<size_t>(n_pools));
  for (int p = 0, o = 0; p < n_pools; ++p)
    for (int j = 0; j < runs_per_pool[p]; ++j) sl[p].push_back(slots[o++]);
  std::vector<ClaimSpec> spec(static_cast<size_t>(n));
  std::vector<int64_t> counts(static_cast<size_t>(n));
  for (int64_t s = 0; s < n; ++s) {
    PB_REQUIRE(pool[s] >= 0 && pool[s] < n_pools && run[s] >= 0 && run[s] < runs_per_pool[pool[s]], "pb_plan_claim_order: bad pool/run");
    spec[s]   = ClaimSpec{pool[s], run[s], cost[s]};
    counts[s] = count[s];
  }
  const std::vector<int32_t> seq = plan_claim_order(plan, counts, spec, sl, 64, makespan);
  for (int64_t s = 0; s < n; ++s) order[s] = seq[static_cast<size_t>(s)];
  PB_API_END
}

int pb_ring_rects(const int32_t box[4], const int32_t allowed[4], const int32_t* prev_box, const int32_t* prev_allowed, int pitch,
                  int64_t capacity, int32_t* rows, int32_t* words, int64_t* n_words) {
  PB_API_BEGIN
  PB_REQUIRE(box != nullptr && allowed != nullptr && n_words != nullptr && pitch > 0, "pb_ring_rects: bad argument");
  PB_REQUIRE((prev_box == nullptr) == (prev_allowed == nullptr), "pb_ring_rects: previous box and allowed go together");
  const RingGeom g{box[0], box[1], box[2], box[3], allowed[0], allowed[1], allowed[2], allowed[3]};
  RingGeom prev{};
  if (prev_box) prev = RingGeom{prev_box[0], prev_box[1], prev_box[2], prev_box[3], prev_allowed[0], prev_allowed[1], prev_allowed[2], prev_allowed[3]};
  int64_t n = 0;
  // the device's enumeration (imprint.cu: ring_scan -> imprint_geom.hpp: ring_list / ring_list_item), item by item
  RingList rl;
  ring_list(g, prev_box ? &prev : nullptr, rl);
  for (int t = 0; t < rl.total; ++t, ++n) {
    int row = 0, x0 = 0, j = 0;
    ring_list_item(rl, t, row, x0, j);
    if (n < capacity) {
      rows[n]  = row;
      words[n] = ((row * pitch + x0) >> 2) + j;
    }
  }
  *n_words = n;
  PB_API_END
}

int pb_imprint_hits(double cx, double cy, double theta, int half_side, int rows, int cols, int64_t n_cells, const int32_t* mx,
                    const int32_t* my, int mode, double eps, int phase, int32_t* n_hits, int32_t* px, int32_t* py) {
  PB_API_BEGIN
  PB_REQUIRE(n_cells >= 0 && (n_cells == 0 || (mx && my && n_hits && px && py)), "pb_imprint_hits: null argument");
  PB_REQUIRE(mode == 0 || mode == 1, "pb_imprint_hits: mode must be 0 (exact) or 1 (single precision)");
  const DevImprint im = make_imprint(cx, cy, theta, half_side);
  const float lo = 0.5f - static_cast<float>(eps), hi = 0.5f + static_cast<float>(eps);
  for (int64_t i = 0; i < n_cells; ++i) {
    PixelHits h;
    if (mode == 0) {
      hits_exact(im, half_side, mx[i], my[i], rows, cols, phase, h);
    } else {
      hits_fast(im.fc, im.fs, im.ix, im.iy, im.flags, static_cast<float>(mx[i] - half_side), static_cast<float>(my[i] - half_side), lo, hi,
                rows, cols, phase, h);
    }
    n_hits[i] = h.n;
    for (int j = 0; j < 2; ++j) {
      px[2 * i + j] = h.px[j];
      py[2 * i + j] = h.py[j];
    }
  }
  PB_API_END
}

// ---- PaintLayer --------------------------------------------------------------------------------------
int pb_layer_create(pb_context* ctx, int rows, int cols, pb_layer** out) {
  PB_API_BEGIN
  PB_REQUIRE(ctx != nullptr, "pb_layer_create: null handle");
  DeviceGuard g(ctx);
  auto l = std::make_unique<pb_layer>();
  planes_alloc(ctx, l->pl, rows, cols, kLayerPlanes);
  // the reference leaves a fresh PaintLayer's memory value-initialised by cv::Mat_; we zero it
  for (int p = 0; p < kLayerPlanes; ++p) fill_plane(ctx, l->pl.plane(p), l->pl.n(), 0.0);
  *out = l.release();
  PB_API_END
}
int pb_layer_destroy(pb_layer* l) {
  PB_API_BEGIN
  if (l) {
    DeviceGuard g(l->pl.ctx);
    if (l->owns) {
      PB_CUDA(cudaStreamSynchronize(l->pl.ctx->stream));
      planes_free(l->pl);
    }
    delete l;
  }
  PB_API_END
}
int pb_layer_rows(const pb_layer* l) { return l->pl.rows; }
int pb_layer_cols(const pb_layer* l) { return l->pl.cols; }
int pb_layer_clear(pb_layer* l) {
  PB_API_BEGIN
  PB_REQUIRE(l != nullptr, "pb_layer_clear: null handle");
  DeviceGuard g(l->pl.ctx);
  if (l->canvas) l->canvas->version++;
  for (int p = 0; p < kLayerPlanes; ++p) fill_plane(l->pl.ctx, l->pl.plane(p), l->pl.n(), 0.0);
  PB_API_END
}
int pb_layer_upload(pb_layer* l, const double* K, const double* S, const double* V) {
  PB_API_BEGIN
  PB_REQUIRE(l != nullptr, "pb_layer_upload: null handle");
  DeviceGuard g(l->pl.ctx);
  if (l->canvas) l->canvas->version++;
  if (K) upload_aos(l->pl.ctx, l->pl, PK, 3, K);
  if (S) upload_aos(l->pl.ctx, l->pl, PS, 3, S);
  if (V) upload_aos(l->pl.ctx, l->pl, PV, 1, V);
  PB_API_END
}
int pb_layer_download(pb_layer* l, double* K, double* S, double* V) {
  PB_API_BEGIN
  PB_REQUIRE(l != nullptr, "pb_layer_download: null handle");
  DeviceGuard g(l->pl.ctx);
  if (K) download_aos(l->pl.ctx, l->pl, PK, 3, K);
  if (S) download_aos(l->pl.ctx, l->pl, PS, 3, S);
  if (V) download_aos(l->pl.ctx, l->pl, PV, 1, V);
  PB_API_END
}
int pb_layer_copy(const pb_layer* src, pb_layer* dst) {
  PB_API_BEGIN
  PB_REQUIRE(src != nullptr, "pb_layer_copy: null handle");
  PB_REQUIRE(dst != nullptr, "pb_layer_copy: null handle");
  DeviceGuard g(src->pl.ctx);
  if (dst->canvas) dst->canvas->version++;
  if (dst->pl.rows != src->pl.rows || dst->pl.cols != src->pl.cols) {  // PaintLayer.hxx:104-107
    PB_REQUIRE(dst->owns, "copyTo: cannot resize a layer view");
    PB_CUDA(cudaStreamSynchronize(dst->pl.ctx->stream));
    planes_free(dst->pl);
    planes_alloc(src->pl.ctx, dst->pl, src->pl.rows, src->pl.cols, kLayerPlanes);
  }
  copy_planes(src->pl.ctx, src->pl, dst->pl, kLayerPlanes);
  PB_API_END
}
int pb_layer_compose(pb_layer* l, const double* R0, double* out) {
  PB_API_BEGIN
  PB_REQUIRE(l != nullptr, "pb_layer_compose: null handle");
  pb_context* ctx = l->pl.ctx;
  DeviceGuard g(ctx);
  pb_planes r;
  planes_alloc_temp(ctx, r, l->pl.rows, l->pl.cols, 3);
  try {
    upload_aos(ctx, r, 0, 3, R0);
    void* o[3] = {r.plane(0), r.plane(1), r.plane(2)};
    km_compose(ctx, l->pl.n(), compose_args(l->pl, r, 0, o, 0, ctx->esize()));
    download_aos(ctx, r, 0, 3, out);
  } catch (...) {
    planes_free_temp(r);
    throw;
  }
  planes_free_temp(r);
  PB_API_END
}
int pb_layer_compose_onto(pb_layer* l, double* R0) { return pb_layer_compose(l, R0, R0); }

// ---- Canvas ------------------------------------------------------------------------------------------
int pb_canvas_create_band(pb_context* ctx, int rows, int cols, int row_begin, int row_end, int halo, pb_canvas** out) {
  PB_API_BEGIN
  PB_REQUIRE(ctx != nullptr, "pb_canvas_create_band: null handle");
  DeviceGuard g(ctx);
  PB_REQUIRE(rows >= 0 && cols >= 0, "canvas size must be non-negative");
  PB_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= rows && halo >= 0, "invalid band");
  PB_REQUIRE(static_cast<int64_t>(rows) * cols < (int64_t(1) << 31), "canvas too large (int32 pixel index like the reference)");
  auto c         = std::make_unique<pb_canvas>();
  c->rows        = rows;
  c->cols        = cols;
  c->row_begin   = row_begin;
  c->row_end     = row_end;
  c->halo        = halo;
  static std::atomic<uint64_t> next_id{1};
  c->id          = next_id.fetch_add(1);
  c->store_first = std::max(0, row_begin - halo);
  const int last = std::min(rows, row_end + halo);
  planes_alloc(ctx, c->pl, last - c->store_first, cols, kCanvasPlanes);
  canvas_clear(c.get());
  *out = c.release();
  PB_API_END
}
int pb_canvas_create(pb_context* ctx, int rows, int cols, pb_canvas** out) {
  return pb_canvas_create_band(ctx, rows, cols, 0, rows, 0, out);
}
int pb_canvas_destroy(pb_canvas* c) {
  PB_API_BEGIN
  if (c) {
    DeviceGuard g(c->pl.ctx);
    PB_CUDA(cudaStreamSynchronize(c->pl.ctx->stream));
    planes_free(c->pl);
    delete c;
  }
  PB_API_END
}
int pb_canvas_rows(const pb_canvas* c) { return c->rows; }
int pb_canvas_cols(const pb_canvas* c) { return c->cols; }
int pb_canvas_stored_rows(const pb_canvas* c, int* first_row, int* n_rows) {
  PB_CHECK_HANDLE(c, "pb_canvas_stored_rows");
  if (first_row) *first_row = c->store_first;
  if (n_rows) *n_rows = c->pl.rows;
  return 0;
}
int pb_canvas_clear(pb_canvas* c) {
  PB_API_BEGIN
  PB_REQUIRE(c != nullptr, "pb_canvas_clear: null handle");
  DeviceGuard g(c->pl.ctx);
  c->version++;
  canvas_clear(c);
  PB_API_END
}
int pb_canvas_set_background(pb_canvas* c, const double* R0) {
  PB_API_BEGIN
  PB_REQUIRE(c != nullptr, "pb_canvas_set_background: null handle");
  DeviceGuard g(c->pl.ctx);
  c->version++;
  canvas_clear(c);
  upload_aos(c->pl.ctx, c->pl, PR, 3, R0);
  PB_API_END
}
int pb_canvas_dry(pb_canvas* c) {
  PB_API_BEGIN
  PB_REQUIRE(c != nullptr, "pb_canvas_dry: null handle");
  DeviceGuard g(c->pl.ctx);
  c->version++;
  void* planes[kCanvasPlanes];
  for (int p = 0; p < kCanvasPlanes; ++p) planes[p] = c->pl.plane(p);
  km_dry(c->pl.ctx, c->pl.n(), planes);
  PB_API_END
}
int pb_canvas_upload_layer(pb_canvas* c, const double* K, const double* S, const double* V) {
  PB_API_BEGIN
  PB_REQUIRE(c != nullptr, "pb_canvas_upload_layer: null handle");
  DeviceGuard g(c->pl.ctx);
  c->version++;
  if (K) upload_aos(c->pl.ctx, c->pl, PK, 3, K);
  if (S) upload_aos(c->pl.ctx, c->pl, PS, 3, S);
  if (V) upload_aos(c->pl.ctx, c->pl, PV, 1, V);
  PB_API_END
}
int pb_canvas_download(pb_canvas* c, double* K, double* S, double* V, double* R0, double* h) {
  PB_API_BEGIN
  PB_REQUIRE(c != nullptr, "pb_canvas_download: null handle");
  DeviceGuard g(c->pl.ctx);
  if (K) download_aos(c->pl.ctx, c->pl, PK, 3, K);
  if (S) download_aos(c->pl.ctx, c->pl, PS, 3, S);
  if (V) download_aos(c->pl.ctx, c->pl, PV, 1, V);
  if (R0) download_aos(c->pl.ctx, c->pl, PR, 3, R0);
  if (h) download_aos(c->pl.ctx, c->pl, PH, 1, h);
  PB_API_END
}
int pb_canvas_compose_device(pb_canvas* c, void* d_out, int64_t plane_stride) {
  PB_API_BEGIN
  PB_REQUIRE(c != nullptr, "pb_canvas_compose_device: null handle");
  pb_context* ctx = c->pl.ctx;
  DeviceGuard g(ctx);
  void* o[3];
  for (int k = 0; k < 3; ++k) o[k] = static_cast<char*>(d_out) + static_cast<size_t>(k) * plane_stride * ctx->esize();
  km_compose(ctx, c->pl.n(), compose_args(c->pl, c->pl, PR, o, 0, ctx->esize()));
  PB_API_END
}
int pb_canvas_compose_band_device(pb_canvas* c, void* d_out, int64_t plane_stride) {
  PB_API_BEGIN
  PB_REQUIRE(c != nullptr, "pb_canvas_compose_band_device: null handle");
  pb_context* ctx = c->pl.ctx;
  DeviceGuard g(ctx);
  void* o[3];
  for (int k = 0; k < 3; ++k) o[k] = static_cast<char*>(d_out) + static_cast<size_t>(k) * plane_stride * ctx->esize();
  const int64_t off = static_cast<int64_t>(c->row_begin - c->store_first) * c->cols;
  const int64_t n   = static_cast<int64_t>(c->row_end - c->row_begin) * c->cols;
  km_compose(ctx, n, compose_args(c->pl, c->pl, PR, o, off, ctx->esize()));
  PB_API_END
}
// ---- assembled reflectance image of a band-sharded canvas ------------------------------------------------------------
struct pb_band_image {
  pb_planes pl;  // 3 planes (r, g, b) of rows x cols elements of the context's type, one cudaMalloc (IPC exportable)
};
int pb_band_image_create(pb_context* ctx, int rows, int cols, pb_band_image** out) {
  PB_API_BEGIN
  PB_REQUIRE(ctx != nullptr && out != nullptr, "pb_band_image_create: null argument");
  PB_REQUIRE(rows > 0 && cols > 0 && static_cast<int64_t>(rows) * cols < (1ll << 31), "pb_band_image_create: bad size");
  DeviceGuard g(ctx);
  auto im = std::make_unique<pb_band_image>();
  planes_alloc(ctx, im->pl, rows, cols, 3);
  *out = im.release();
  PB_API_END
}
int pb_band_image_destroy(pb_band_image* im) {
  PB_API_BEGIN
  if (im) {
    DeviceGuard g(im->pl.ctx);
    PB_CUDA(cudaStreamSynchronize(im->pl.ctx->stream));
    planes_free(im->pl);
    delete im;
  }
  PB_API_END
}
int pb_band_image_device(pb_band_image* im, void** base, int64_t* plane_stride_bytes) {
  PB_CHECK_HANDLE(im, "pb_band_image_device");
  if (base) *base = im->pl.base;
  if (plane_stride_bytes) *plane_stride_bytes = static_cast<int64_t>(im->pl.stride);
  return 0;
}
int pb_band_image_download(pb_band_image* im, double* out) {
  PB_API_BEGIN
  PB_REQUIRE(im != nullptr && out != nullptr, "pb_band_image_download: null argument");
  DeviceGuard g(im->pl.ctx);
  download_aos(im->pl.ctx, im->pl, 0, 3, out);
  PB_API_END
}
int pb_canvas_compose_gather(pb_canvas* c, int n_dst, void* const* dst_base, int64_t plane_stride_bytes) {
  PB_API_BEGIN
  PB_REQUIRE(c != nullptr && dst_base != nullptr, "pb_canvas_compose_gather: null argument");
  PB_REQUIRE(n_dst >= 1 && n_dst <= PB_MAX_BANDS, "pb_canvas_compose_gather: 1..8 destinations");
  pb_context* ctx = c->pl.ctx;
  DeviceGuard g(ctx);
  const size_t es   = ctx->esize();
  const int64_t off = static_cast<int64_t>(c->row_begin - c->store_first) * c->cols;  // first owned pixel in the stored planes
  const int64_t at  = static_cast<int64_t>(c->row_begin) * c->cols;                   // ... and in the assembled image
  const int64_t n   = static_cast<int64_t>(c->row_end - c->row_begin) * c->cols;
  void* dst[PB_MAX_BANDS][3];
  for (int d = 0; d < n_dst; ++d) {
    PB_REQUIRE(dst_base[d] != nullptr, "pb_canvas_compose_gather: null destination");
    for (int k = 0; k < 3; ++k)
      dst[d][k] = static_cast<char*>(dst_base[d]) + static_cast<size_t>(k) * static_cast<size_t>(plane_stride_bytes) + static_cast<size_t>(at) * es;
  }
  void* none[3] = {nullptr, nullptr, nullptr};
  km_compose_gather(ctx, n, compose_args(c->pl, c->pl, PR, none, off, es), n_dst, dst);
  PB_API_END
}
int pb_canvas_compose(pb_canvas* c, double* out) {
  PB_API_BEGIN
  PB_REQUIRE(c != nullptr, "pb_canvas_compose: null handle");
  pb_context* ctx = c->pl.ctx;
  DeviceGuard g(ctx);
  pb_planes r;
  planes_alloc_temp(ctx, r, c->pl.rows, c->pl.cols, 3);
  try {
    void* o[3] = {r.plane(0), r.plane(1), r.plane(2)};
    km_compose(ctx, c->pl.n(), compose_args(c->pl, c->pl, PR, o, 0, ctx->esize()));
    download_aos(ctx, r, 0, 3, out);
  } catch (...) {
    planes_free_temp(r);
    throw;
  }
  planes_free_temp(r);
  PB_API_END
}
namespace {
void compose_display(pb_canvas* c, int mode, bool srgb, size_t bytes_per_px, void* host_out) {
  pb_context* ctx = c->pl.ctx;
  DeviceGuard g(ctx);
  const int64_t n = c->pl.n();
  if (n == 0) return;
  DevBuf<unsigned char> d(ctx, static_cast<size_t>(n) * bytes_per_px);
  void* none[3] = {nullptr, nullptr, nullptr};
  km_compose_display(ctx, n, compose_args(c->pl, c->pl, PR, none, 0, ctx->esize()), mode, srgb, d.p);
  PB_CUDA(cudaMemcpyAsync(host_out, d.p, static_cast<size_t>(n) * bytes_per_px, cudaMemcpyDeviceToHost, ctx->stream));
  PB_CUDA(cudaStreamSynchronize(ctx->stream));
}
}  // namespace
int pb_canvas_render(pb_canvas* c, double* out) {
  PB_API_BEGIN
  PB_REQUIRE(c != nullptr, "pb_canvas_render: null handle");
  pb_context* ctx = c->pl.ctx;
  DeviceGuard g(ctx);
  PB_REQUIRE(c->store_first == 0 && c->pl.rows == c->rows, "pb_canvas_render needs a full canvas (not a band)");
  pb_planes r;
  planes_alloc_temp(ctx, r, c->pl.rows, c->pl.cols, 3);
  try {
    void* o[3] = {r.plane(0), r.plane(1), r.plane(2)};
    km_render(ctx, c->pl.rows, c->pl.cols, compose_args(c->pl, c->pl, PR, o, 0, ctx->esize()));
    download_aos(ctx, r, 0, 3, out);
  } catch (...) {
    planes_free_temp(r);
    throw;
  }
  planes_free_temp(r);
  PB_API_END
}
namespace {
// OpenCV's INTER_LANCZOS4 taps for one axis of a (src -> dst) resize, exactly as cv::resize prepares them for a
// floating-point image (modules/imgproc/src/resize.cpp: fx = (float)((d + 0.5) * scale - 0.5), s = floor(fx), and
// interpolateLanczos4 — single-precision intermediate x + 3 - i, double sin/cos of the first angle rotated by 45 degrees per
// tap, weights normalised in float). tests/test_oracle.py pins this against cv2.resize bit for bit.
void lanczos4_taps(int src, int dst, std::vector<int>& ofs, std::vector<float>& w) {
  static const double s45    = 0.70710678118654752440084436210485;
  static const double cs[][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45}, {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
  const double pi    = 3.1415926535897932384626433832795;
  const double scale = static_cast<double>(src) / dst;
  ofs.resize(static_cast<size_t>(dst));
  w.resize(static_cast<size_t>(dst) * 8);
  for (int d = 0; d < dst; ++d) {
    float fx     = static_cast<float>((d + 0.5) * scale - 0.5);
    const int sx = static_cast<int>(std::floor(fx));
    fx -= static_cast<float>(sx);
    ofs[static_cast<size_t>(d)] = sx;
    float* c        = &w[static_cast<size_t>(d) * 8];
    const float x3  = fx + 3.0f;
    const double y0 = -x3 * pi * 0.25, s0 = std::sin(y0), c0 = std::cos(y0);
    float sum       = 0.f;
    for (int i = 0; i < 8; ++i) {
      const float y0_ = x3 - static_cast<float>(i);
      if (std::fabs(y0_) >= 1e-6f) {
        const double y = -y0_ * pi * 0.25;
        c[i]           = static_cast<float>((cs[i][0] * s0 + cs[i][1] * c0) / (y * y));
      } else {
        c[i] = 1e30f;
      }
      sum += c[i];
    }
    sum = 1.f / sum;
    for (int i = 0; i < 8; ++i) c[i] *= sum;
  }
}
}  // namespace
int pb_lanczos4_taps(int src, int dst, int32_t* offsets, float* weights) {
  PB_API_BEGIN
  PB_REQUIRE(src > 0 && dst > 0 && offsets && weights, "pb_lanczos4_taps: bad arguments");
  std::vector<int> o;
  std::vector<float> w;
  lanczos4_taps(src, dst, o, w);
  std::copy(o.begin(), o.end(), offsets);
  std::copy(w.begin(), w.end(), weights);
  PB_API_END
}
int pb_canvas_compose_lab_scaled(pb_canvas* c, int out_rows, int out_cols, double* out) {
  PB_API_BEGIN
  PB_REQUIRE(c != nullptr && out != nullptr, "pb_canvas_compose_lab_scaled: null argument");
  PB_REQUIRE(out_rows > 0 && out_cols > 0, "pb_canvas_compose_lab_scaled: empty output");
  pb_context* ctx = c->pl.ctx;
  DeviceGuard g(ctx);
  PB_REQUIRE(c->store_first == 0 && c->pl.rows == c->rows, "pb_canvas_compose_lab_scaled needs a full canvas (not a band)");
  const int rows = c->pl.rows, cols = c->pl.cols;
  pb_planes lab;
  planes_alloc_temp(ctx, lab, rows, cols, 3);
  try {
    void* o[3] = {lab.plane(0), lab.plane(1), lab.plane(2)};
    km_compose_lab(ctx, c->pl.n(), compose_args(c->pl, c->pl, PR, o, 0, ctx->esize()));
    const size_t n_out = static_cast<size_t>(out_rows) * out_cols * 3;
    DevBuf<double> d_out(ctx, n_out);
    if (out_rows == rows && out_cols == cols) {  // cv::resize copies when the size does not change
      lab_planes_to_aos(ctx, o, c->pl.n(), d_out.p);
    } else {
      std::vector<int> xofs, yofs;
      std::vector<float> alpha, beta;
      lanczos4_taps(cols, out_cols, xofs, alpha);
      lanczos4_taps(rows, out_rows, yofs, beta);
      DevBuf<int> d_x(ctx, xofs.size()), d_y(ctx, yofs.size());
      DevBuf<float> d_a(ctx, alpha.size()), d_b(ctx, beta.size());
      DevBuf<double> d_tmp(ctx, static_cast<size_t>(rows) * out_cols * 3);
      d_x.upload(xofs.data(), xofs.size());
      d_y.upload(yofs.data(), yofs.size());
      d_a.upload(alpha.data(), alpha.size());
      d_b.upload(beta.data(), beta.size());
      lab_resize_lanczos4(ctx, o, rows, cols, out_rows, out_cols, d_x.p, d_a.p, d_y.p, d_b.p, d_tmp.p, d_out.p);
      PB_CUDA(cudaStreamSynchronize(ctx->stream));  // the host tap vectors were uploaded asynchronously
    }
    PB_CUDA(cudaMemcpyAsync(out, d_out.p, n_out * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
  } catch (...) {
    planes_free_temp(lab);
    throw;
  }
  planes_free_temp(lab);
  PB_API_END
}
int pb_canvas_compose_qrgb32(pb_canvas* c, uint32_t* out) {
  PB_API_BEGIN
  PB_REQUIRE(c != nullptr, "pb_canvas_compose_qrgb32: null handle");
  compose_display(c, 0, true, 4, out);
  PB_API_END
}
int pb_canvas_compose_bgr(pb_canvas* c, int bits, int srgb, void* out) {
  PB_API_BEGIN
  PB_REQUIRE(c != nullptr, "pb_canvas_compose_bgr: null handle");
  PB_REQUIRE(bits == 8 || bits == 16, "pb_canvas_compose_bgr: bits must be 8 or 16");
  compose_display(c, bits == 8 ? 1 : 2, srgb != 0, bits == 8 ? 3 : 6, out);
  PB_API_END
}
int pb_canvas_paint_layer(pb_canvas* c, pb_layer** out) {
  PB_API_BEGIN
  PB_REQUIRE(c != nullptr, "pb_canvas_paint_layer: null handle");
  auto l        = std::make_unique<pb_layer>();
  l->pl         = c->pl;
  l->pl.nplanes = kLayerPlanes;
  l->owns       = false;
  l->canvas     = c;
  *out          = l.release();
  PB_API_END
}
int pb_canvas_upload_substrate(pb_canvas* c, const double* R0, const double* h) {
  PB_API_BEGIN
  PB_REQUIRE(c != nullptr, "pb_canvas_upload_substrate: null handle");
  DeviceGuard g(c->pl.ctx);
  if (R0) upload_aos(c->pl.ctx, c->pl, PR, 3, R0);
  if (h) upload_aos(c->pl.ctx, c->pl, PH, 1, h);
  PB_API_END
}
int pb_canvas_device_planes(pb_canvas* c, void* planes[11], int64_t* elems_per_plane) {
  PB_CHECK_HANDLE(c, "pb_canvas_device_planes");
  for (int p = 0; p < kCanvasPlanes; ++p) planes[p] = c->pl.plane(p);
  if (elems_per_plane) *elems_per_plane = c->pl.n();
  c->version++;  // the pointers are writable: brushes must assume the wet layer changed
  return 0;
}
int pb_canvas_mark_modified(pb_canvas* c) {
  PB_CHECK_HANDLE(c, "pb_canvas_mark_modified");
  c->version++;
  return 0;
}

// ---- raw compose ---------------------------------------------------------------------------------------
int pb_km_compose_planes(pb_context* ctx, int64_t n, const void* const K[3], const void* const S[3], const void* V,
                         const void* const R0[3], void* const R[3]) {
  PB_API_BEGIN
  PB_REQUIRE(ctx != nullptr, "pb_km_compose_planes: null handle");
  DeviceGuard g(ctx);
  ComposeArgs a;
  for (int k = 0; k < 3; ++k) {
    a.K[k]  = K[k];
    a.S[k]  = S[k];
    a.R0[k] = R0[k];
    a.R[k]  = R[k];
  }
  a.V = V;
  km_compose(ctx, n, a);
  PB_API_END
}
int pb_km_compose_stacked_planes(pb_context* ctx, int64_t n, int n_layers, const void* const* K, const void* const* S,
                                 const void* const* V, const void* const R0[3], void* const R[3]) {
  PB_API_BEGIN
  PB_REQUIRE(ctx != nullptr, "pb_km_compose_stacked_planes: null handle");
  DeviceGuard g(ctx);
  PB_REQUIRE(n_layers >= 1, "compose_stacked: need at least one layer");
  // more than kMaxStack layers: chain passes of <= kMaxStack, intermediate R stays on the device in R
  int done = 0;
  while (done < n_layers) {
    const int m = std::min(kMaxStack, n_layers - done);
    StackArgs a;
    a.n_layers = m;
    for (int l = 0; l < m; ++l) {
      for (int k = 0; k < 3; ++k) {
        a.K[l][k] = K[3 * (done + l) + k];
        a.S[l][k] = S[3 * (done + l) + k];
      }
      a.V[l] = V[done + l];
    }
    for (int k = 0; k < 3; ++k) {
      a.R0[k] = done == 0 ? R0[k] : R[k];
      a.R[k]  = R[k];
    }
    km_compose_stacked(ctx, n, a);
    done += m;
  }
  PB_API_END
}

// ---- FootprintBrush --------------------------------------------------------------------------------------
int pb_fbrush_create(pb_context* ctx, pb_fbrush** out) {
  PB_API_BEGIN
  PB_REQUIRE(ctx != nullptr, "pb_fbrush_create: null handle");
  DeviceGuard g(ctx);
  auto b = std::make_unique<pb_fbrush>();
  b->ctx = ctx;
  PB_CUDA(cudaMalloc(reinterpret_cast<void**>(&b->d_counters), 2 * sizeof(unsigned long long)));
  PB_CUDA(cudaMemsetAsync(b->d_counters, 0, 2 * sizeof(unsigned long long), ctx->stream));
  *out = b.release();
  PB_API_END
}
int pb_fbrush_destroy(pb_fbrush* b) {
  PB_API_BEGIN
  if (b) {
    DeviceGuard g(b->ctx);
    PB_CUDA(cudaStreamSynchronize(b->ctx->stream));
    for (auto& kv : b->geoms) {
      if (kv.second.d_xy) cudaFree(kv.second.d_xy);
      if (kv.second.d_fh) cudaFree(kv.second.d_fh);
    }
    planes_free(b->pick);
    if (b->snap_rec) cudaFree(b->snap_rec);
    if (b->work_rec) cudaFree(b->work_rec);
    if (b->dirty) cudaFree(b->dirty);
    if (b->dist_flags) cudaFree(b->dist_flags);
    cudaFree(b->d_counters);
    if (b->d_trace) cudaFree(b->d_trace);
    delete b;
  }
  PB_API_END
}
int pb_fbrush_register_footprint(pb_fbrush* b, double radius, int side, const double* footprint) {
  PB_API_BEGIN
  PB_REQUIRE(b != nullptr, "pb_fbrush_register_footprint: null handle");
  DeviceGuard g(b->ctx);
  PB_REQUIRE(footprint != nullptr, "footprint is null");
  register_footprint(b, radius, side, footprint);
  PB_API_END
}
int pb_fbrush_set_radius(pb_fbrush* b, double radius, int side, const double* footprint, int* acted) {
  PB_API_BEGIN
  PB_REQUIRE(b != nullptr, "pb_fbrush_set_radius: null handle");
  DeviceGuard g(b->ctx);
  const bool act = !(std::fabs(b->radius - radius) < 0.5);  // fuzzyCompare, FootprintBrush.hxx:47-48
  if (acted) *acted = act ? 1 : 0;
  if (act && footprint != nullptr) brush_set_geometry(b, radius, register_footprint(b, radius, side, footprint));
  PB_API_END
}
int pb_fbrush_clean(pb_fbrush* b) {
  PB_API_BEGIN
  PB_REQUIRE(b != nullptr, "pb_fbrush_clean: null handle");
  DeviceGuard g(b->ctx);
  if (b->pick.base)
    for (int p = 0; p < kLayerPlanes; ++p) fill_plane(b->ctx, b->pick.plane(p), b->pick.n(), 0.0);
  PB_API_END
}
int pb_fbrush_dip(pb_fbrush* b, const double K[3], const double S[3]) {
  PB_CHECK_HANDLE(b, "pb_fbrush_dip");
  const int rc = pb_fbrush_clean(b);
  if (rc) return rc;
  for (int i = 0; i < 3; ++i) {
    b->paintK[i] = K[i];
    b->paintS[i] = S[i];
  }
  return 0;
}
int pb_fbrush_set_pickup_rate(pb_fbrush* b, double rate) {
  PB_CHECK_HANDLE(b, "pb_fbrush_set_pickup_rate");
  b->pickup_rate = rate;
  return 0;
}
int pb_fbrush_set_deposition_rate(pb_fbrush* b, double rate) {
  PB_CHECK_HANDLE(b, "pb_fbrush_set_deposition_rate");
  b->deposition_rate = rate;
  return 0;
}
double pb_fbrush_get_pickup_rate(const pb_fbrush* b) { return b->pickup_rate; }
double pb_fbrush_get_deposition_rate(const pb_fbrush* b) { return b->deposition_rate; }
int pb_fbrush_set_use_snapshot(pb_fbrush* b, int use) {
  PB_CHECK_HANDLE(b, "pb_fbrush_set_use_snapshot");
  b->use_snapshot = use != 0;
  return 0;
}
int pb_fbrush_get_use_snapshot(const pb_fbrush* b) { return b->use_snapshot ? 1 : 0; }
int pb_fbrush_size_map(const pb_fbrush* b) { return b->cur ? b->cur->size_map : 0; }
int pb_fbrush_pickup_map(pb_fbrush* b, double* K, double* S, double* V) {
  PB_API_BEGIN
  PB_REQUIRE(b != nullptr, "pb_fbrush_pickup_map: null handle");
  DeviceGuard g(b->ctx);
  PB_REQUIRE(b->pick.base != nullptr, "brush has no pickup map yet (setRadius was never applied)");
  if (K) download_aos(b->ctx, b->pick, PK, 3, K);
  if (S) download_aos(b->ctx, b->pick, PS, 3, S);
  if (V) download_aos(b->ctx, b->pick, PV, 1, V);
  PB_API_END
}
int pb_fbrush_pickup_layer(pb_fbrush* b, pb_layer** out) {
  PB_API_BEGIN
  PB_REQUIRE(b != nullptr, "pb_fbrush_pickup_layer: null handle");
  PB_REQUIRE(b->pick.base != nullptr, "brush has no pickup map yet (setRadius was never applied)");
  auto l  = std::make_unique<pb_layer>();
  l->pl   = b->pick;
  l->owns = false;
  *out    = l.release();
  PB_API_END
}
int pb_fbrush_update_snapshot(pb_fbrush* b, pb_canvas* c) {
  PB_API_BEGIN
  PB_REQUIRE(b != nullptr, "pb_fbrush_update_snapshot: null handle");
  PB_REQUIRE(c != nullptr, "pb_fbrush_update_snapshot: null handle");
  DeviceGuard g(b->ctx);
  if (!snapshot_matches(b, c)) {
    ensure_snapshot(b, c);
  } else {
    planes_to_records(b->ctx, c->pl, b->snap_rec, 0, 0, c->pl.cols - 1, c->pl.rows - 1);
    PB_CUDA(cudaMemsetAsync(b->dirty, 0, dirty_bytes(c), b->ctx->stream));
    b->snap_canvas_id      = c->id;
    b->snap_canvas_version = c->version;
  }
  PB_API_END
}
int pb_fbrush_snapshot_download(pb_fbrush* b, double* K, double* S, double* V) {
  PB_API_BEGIN
  PB_REQUIRE(b != nullptr, "pb_fbrush_snapshot_download: null handle");
  DeviceGuard g(b->ctx);
  PB_REQUIRE(b->snap_rec != nullptr, "brush has no snapshot buffer yet");
  pb_planes tmp;
  planes_alloc_temp(b->ctx, tmp, b->snap_rows, b->snap_cols, kLayerPlanes);
  try {
    records_to_planes(b->ctx, b->snap_rec, tmp, 0, 0, b->snap_cols - 1, b->snap_rows - 1);
    if (K) download_aos(b->ctx, tmp, PK, 3, K);
    if (S) download_aos(b->ctx, tmp, PS, 3, S);
    if (V) download_aos(b->ctx, tmp, PV, 1, V);
  } catch (...) {
    planes_free_temp(tmp);
    throw;
  }
  planes_free_temp(tmp);
  PB_API_END
}
int pb_fbrush_imprint_batch(pb_fbrush* b, pb_canvas* c, int64_t n, const double* cx, const double* cy,
                            const double* theta) {
  PB_API_BEGIN
  PB_REQUIRE(b != nullptr, "pb_fbrush_imprint_batch: null handle");
  PB_REQUIRE(c != nullptr, "pb_fbrush_imprint_batch: null handle");
  DeviceGuard g(b->ctx);
  PB_REQUIRE(b->cur != nullptr, "imprint before setRadius: the brush has no footprint");
  if (n <= 0) return 0;
  PB_REQUIRE(cx != nullptr && cy != nullptr && theta != nullptr, "imprint_batch: null imprint arrays");
  PB_REQUIRE(n < (int64_t(1) << 31), "too many imprints in one batch");
  HostStroke h;
  h.g      = b->cur;
  h.radius = b->radius;
  for (int i = 0; i < 3; ++i) {
    h.K[i] = b->paintK[i];
    h.S[i] = b->paintS[i];
  }
  h.first = 0;
  h.n     = n;
  h.flags = 3;  // continue with, and write back, the brush's persistent pickup map
  run_imprints(b, c, {h}, n, cx, cy, theta);
  PB_API_END
}
namespace {
DistInfo dist_info(const pb_canvas* c, const pb_dist_desc* d) {
  PB_REQUIRE(d != nullptr && d->world >= 1 && d->world <= kMaxBands && d->rank >= 0 && d->rank < d->world, "invalid pb_dist_desc");
  PB_REQUIRE(d->rows_per_band > 0 && c->halo == 0 && c->row_begin == d->rank * d->rows_per_band &&
               c->row_end == std::min(c->rows, (d->rank + 1) * d->rows_per_band),
             "canvas is not this rank's band of a rows_per_band partition");
  DistInfo di;
  di.world         = d->world;
  di.rank          = d->rank;
  di.rows_per_band = d->rows_per_band;
  for (int r = 0; r < d->world; ++r) {
    di.canvas_base[r]     = d->canvas_base[r];
    di.canvas_stride[r]   = d->canvas_stride[r];
    di.snapshot_base[r]   = d->snapshot_base[r];
    di.snapshot_stride[r] = d->snapshot_stride[r];
    di.dirty_base[r]      = static_cast<unsigned char*>(d->dirty_base[r]);
    di.flags_base[r]      = static_cast<long long*>(d->flags_base[r]);
  }
  return di;
}

// dip -> setRadius -> paintStroke per stroke (SbrRenderThread.cxx:68-72), planned from the brush's current state
void plan_stroke_batch(pb_fbrush* b, pb_canvas* c, int64_t n_strokes, const pb_stroke* strokes, int64_t n_imprints, const double* cx,
                       const double* cy, const double* theta, const DistInfo* dist, pb_batch_plan& P) {
  PB_REQUIRE(n_strokes >= 0 && n_strokes < (int64_t(1) << 27), "too many strokes in one batch");
  PB_REQUIRE(n_strokes == 0 || (strokes != nullptr && n_imprints >= 0 && (n_imprints == 0 || (cx != nullptr && cy != nullptr && theta != nullptr))),
             "stroke_batch: null stroke or imprint arrays");
  std::vector<HostStroke> hs(static_cast<size_t>(n_strokes));
  double radius            = b->radius;
  const FootprintGeom* cur = b->cur;
  P.radius_before          = b->radius;
  for (int64_t s = 0; s < n_strokes; ++s) {
    const pb_stroke& in = strokes[s];
    PB_REQUIRE(in.first_imprint >= 0 && in.n_imprints >= 0 && in.first_imprint + in.n_imprints <= n_imprints,
               "stroke imprint range out of bounds");
    PB_REQUIRE(in.n_imprints < (int64_t(1) << 31), "too many imprints in one stroke");
    // setRadius only acts on a change >= 0.5
    if (!(std::fabs(radius - in.radius) < 0.5)) {
      radius          = in.radius;
      const int width = static_cast<int32_t>(2.0 * std::ceil(radius) + 1.0);
      auto it         = b->by_width.find(width);
      PB_REQUIRE(it != b->by_width.end(), "stroke_batch: no footprint registered for radius " + std::to_string(radius));
      cur = &b->geoms.at(it->second);
    }
    PB_REQUIRE(cur != nullptr, "stroke_batch: the brush has no footprint");
    HostStroke& h = hs[static_cast<size_t>(s)];
    h.g           = cur;
    h.radius      = radius;
    for (int i = 0; i < 3; ++i) {
      h.K[i] = in.K[i];
      h.S[i] = in.S[i];
    }
    h.first = in.first_imprint;
    h.n     = in.n_imprints;
    h.flags = 0;  // dip(): clean pickup map
  }
  if (n_strokes > 0) {
    hs.back().flags = 2;
    // brush state after the batch = state after the last stroke
    P.sets_state   = true;
    P.radius_after = radius;
    P.geom_after   = cur;
    for (int i = 0; i < 3; ++i) {
      P.K_after[i] = strokes[n_strokes - 1].K[i];
      P.S_after[i] = strokes[n_strokes - 1].S[i];
    }
  }
  plan_imprints(b, c, hs, n_imprints, cx, cy, theta, dist, P);
}

void run_stroke_plan(pb_fbrush* b, pb_canvas* c, const pb_batch_plan& P, const DistInfo* dist) {
  if (!P.sets_state) return;  // empty batch
  // the stroke records depend on the radius the brush had when the plan was made (the 0.5 rule of setRadius)
  PB_REQUIRE(b->radius == P.radius_before, "batch plan is stale: the brush radius changed since it was planned");
  brush_set_geometry(b, P.radius_after, P.geom_after);
  for (int i = 0; i < 3; ++i) {
    b->paintK[i] = P.K_after[i];
    b->paintS[i] = P.S_after[i];
  }
  run_plan(b, c, P, dist);
}
}  // namespace

int pb_fbrush_stroke_batch(pb_fbrush* b, pb_canvas* c, int64_t n_strokes, const pb_stroke* strokes, int64_t n_imprints,
                           const double* cx, const double* cy, const double* theta) {
  PB_API_BEGIN
  PB_REQUIRE(b != nullptr, "pb_fbrush_stroke_batch: null handle");
  PB_REQUIRE(c != nullptr, "pb_fbrush_stroke_batch: null handle");
  DeviceGuard g(b->ctx);
  if (n_strokes <= 0) return 0;
  pb_batch_plan P;
  plan_stroke_batch(b, c, n_strokes, strokes, n_imprints, cx, cy, theta, nullptr, P);
  run_stroke_plan(b, c, P, nullptr);
  PB_API_END
}
int pb_fbrush_plan_stroke_batch(pb_fbrush* b, pb_canvas* c, const pb_dist_desc* dist, int64_t n_strokes, const pb_stroke* strokes,
                                int64_t n_imprints, const double* cx, const double* cy, const double* theta, pb_batch_plan** out) {
  PB_API_BEGIN
  PB_REQUIRE(b != nullptr && c != nullptr && out != nullptr, "pb_fbrush_plan_stroke_batch: null argument");
  DeviceGuard g(b->ctx);
  auto P = std::make_unique<pb_batch_plan>();
  if (dist != nullptr) {
    const DistInfo di = dist_info(c, dist);
    plan_stroke_batch(b, c, n_strokes, strokes, n_imprints, cx, cy, theta, &di, *P);
  } else {
    plan_stroke_batch(b, c, n_strokes, strokes, n_imprints, cx, cy, theta, nullptr, *P);
  }
  *out = P.release();
  PB_API_END
}
int pb_fbrush_run_batch_plan(pb_fbrush* b, pb_canvas* c, const pb_dist_desc* dist, const pb_batch_plan* plan) {
  PB_API_BEGIN
  PB_REQUIRE(b != nullptr && c != nullptr && plan != nullptr, "pb_fbrush_run_batch_plan: null argument");
  DeviceGuard g(b->ctx);
  if (dist != nullptr) {
    PB_REQUIRE(b->dist_flags != nullptr, "call pb_fbrush_dist_storage first");
    const DistInfo di = dist_info(c, dist);
    run_stroke_plan(b, c, *plan, &di);
  } else {
    run_stroke_plan(b, c, *plan, nullptr);
  }
  PB_API_END
}
int pb_batch_plan_destroy(pb_batch_plan* plan) {
  delete plan;
  return 0;
}
int pb_batch_plan_stats(const pb_batch_plan* plan, double out[PB_BATCH_STATS]) {
  PB_CHECK_HANDLE(plan, "pb_batch_plan_stats");
  PB_CHECK_HANDLE(out, "pb_batch_plan_stats");
  for (int i = 0; i < PB_BATCH_STATS; ++i) out[i] = plan->stats[i];
  return 0;
}

// ---- multi-GPU ---------------------------------------------------------------------------------------------
int pb_ipc_export(pb_context* ctx, void* dev_ptr, unsigned char handle[PB_IPC_HANDLE_BYTES]) {
  PB_API_BEGIN
  PB_REQUIRE(ctx != nullptr, "pb_ipc_export: null handle");
  DeviceGuard g(ctx);
  static_assert(sizeof(cudaIpcMemHandle_t) == PB_IPC_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  PB_CUDA(cudaIpcGetMemHandle(&h, dev_ptr));
  std::memcpy(handle, &h, sizeof(h));
  PB_API_END
}
int pb_ipc_import(pb_context* ctx, const unsigned char handle[PB_IPC_HANDLE_BYTES], void** dev_ptr) {
  PB_API_BEGIN
  PB_REQUIRE(ctx != nullptr, "pb_ipc_import: null handle");
  DeviceGuard g(ctx);
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  PB_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  PB_API_END
}
int pb_ipc_close(pb_context* ctx, void* dev_ptr) {
  PB_API_BEGIN
  PB_REQUIRE(ctx != nullptr, "pb_ipc_close: null handle");
  DeviceGuard g(ctx);
  PB_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  PB_API_END
}
int pb_canvas_storage(pb_canvas* c, void** base, int64_t* plane_stride_bytes) {
  PB_CHECK_HANDLE(c, "pb_canvas_storage");
  if (base) *base = c->pl.base;
  if (plane_stride_bytes) *plane_stride_bytes = static_cast<int64_t>(c->pl.stride);
  c->version++;  // writable pointer handed out
  return 0;
}
int pb_fbrush_dist_storage(pb_fbrush* b, pb_canvas* c, void** canvas_records, void** snapshot_records, void** dirty_base,
                           void** flags_base) {
  PB_API_BEGIN
  PB_REQUIRE(b != nullptr, "pb_fbrush_dist_storage: null handle");
  PB_REQUIRE(c != nullptr, "pb_fbrush_dist_storage: null handle");
  DeviceGuard g(b->ctx);
  ensure_snapshot(b, c);
  ensure_work(b, c);
  if (b->dist_flags == nullptr) {
    PB_CUDA(cudaMalloc(reinterpret_cast<void**>(&b->dist_flags), sizeof(long long) * (kDistFlagCapacity + 1024)));
    PB_CUDA(cudaMemsetAsync(b->dist_flags, 0, sizeof(long long) * (kDistFlagCapacity + 1024), b->ctx->stream));
  }
  PB_CUDA(cudaStreamSynchronize(b->ctx->stream));
  if (canvas_records) *canvas_records = b->work_rec;
  if (snapshot_records) *snapshot_records = b->snap_rec;
  if (dirty_base) *dirty_base = b->dirty;
  if (flags_base) *flags_base = b->dist_flags;
  PB_API_END
}
int pb_fbrush_dist_begin(pb_fbrush* b, pb_canvas* c) {
  PB_API_BEGIN
  PB_REQUIRE(b != nullptr && c != nullptr, "pb_fbrush_dist_begin: null handle");
  DeviceGuard g(b->ctx);
  PB_REQUIRE(b->work_rec != nullptr && b->work_rows == c->pl.rows && b->work_cols == c->pl.cols && snapshot_matches(b, c),
             "call pb_fbrush_dist_storage first");
  ensure_snapshot(b, c);  // marks everything dirty when the canvas changed behind the brush's back
  planes_to_records(b->ctx, c->pl, b->work_rec, 0, 0, c->pl.cols - 1, c->pl.rows - 1);
  PB_CUDA(cudaStreamSynchronize(b->ctx->stream));
  PB_API_END
}
int pb_fbrush_dist_end(pb_fbrush* b, pb_canvas* c) {
  PB_API_BEGIN
  PB_REQUIRE(b != nullptr && c != nullptr, "pb_fbrush_dist_end: null handle");
  DeviceGuard g(b->ctx);
  PB_REQUIRE(b->work_rec != nullptr && b->work_rows == c->pl.rows && b->work_cols == c->pl.cols, "call pb_fbrush_dist_storage first");
  records_to_planes(b->ctx, b->work_rec, c->pl, 0, 0, c->pl.cols - 1, c->pl.rows - 1);
  // the brush's own writes: its dirty map stays valid for the new canvas state
  c->version++;
  b->snap_canvas_id      = c->id;
  b->snap_canvas_version = c->version;
  PB_API_END
}
int pb_fbrush_stroke_batch_dist(pb_fbrush* b, pb_canvas* c, const pb_dist_desc* d, int64_t n_strokes, const pb_stroke* strokes,
                                int64_t n_imprints, const double* cx, const double* cy, const double* theta) {
  PB_API_BEGIN
  PB_REQUIRE(b != nullptr, "pb_fbrush_stroke_batch_dist: null handle");
  PB_REQUIRE(c != nullptr, "pb_fbrush_stroke_batch_dist: null handle");
  DeviceGuard g(b->ctx);
  PB_REQUIRE(b->dist_flags != nullptr, "call pb_fbrush_dist_storage first");
  const DistInfo di = dist_info(c, d);
  if (n_strokes <= 0) return 0;
  pb_batch_plan P;
  plan_stroke_batch(b, c, n_strokes, strokes, n_imprints, cx, cy, theta, &di, P);
  run_stroke_plan(b, c, P, &di);
  PB_API_END
}
int pb_fbrush_enable_trace(pb_fbrush* b, int enable) {
  PB_API_BEGIN
  PB_REQUIRE(b != nullptr, "pb_fbrush_enable_trace: null handle");
  DeviceGuard g(b->ctx);
  const size_t bytes = sizeof(unsigned long long) * kTraceImprints * 2 * kTraceStamps;
  PB_CUDA(cudaStreamSynchronize(b->ctx->stream));
  if (enable && b->d_trace == nullptr) PB_CUDA(cudaMalloc(reinterpret_cast<void**>(&b->d_trace), bytes));
  if (!enable && b->d_trace != nullptr) {
    cudaFree(b->d_trace);
    b->d_trace = nullptr;
  }
  if (b->d_trace) PB_CUDA(cudaMemset(b->d_trace, 0, bytes));
  PB_API_END
}
int pb_fbrush_read_trace(pb_fbrush* b, uint64_t* out) {
  PB_API_BEGIN
  PB_REQUIRE(b != nullptr && out != nullptr && b->d_trace != nullptr, "pb_fbrush_read_trace: tracing is not enabled");
  DeviceGuard g(b->ctx);
  PB_CUDA(cudaStreamSynchronize(b->ctx->stream));
  PB_CUDA(cudaMemcpy(out, b->d_trace, sizeof(unsigned long long) * kTraceImprints * 2 * kTraceStamps, cudaMemcpyDeviceToHost));
  PB_API_END
}
int pb_fbrush_enable_visited_count(pb_fbrush* b, int enable) {
  PB_CHECK_HANDLE(b, "pb_fbrush_enable_visited_count");
  b->count_visited = enable != 0;
  return 0;
}
int pb_fbrush_batch_stats(const pb_fbrush* b, double out[PB_BATCH_STATS]) {
  PB_CHECK_HANDLE(b, "pb_fbrush_batch_stats");
  PB_CHECK_HANDLE(out, "pb_fbrush_batch_stats");
  for (int i = 0; i < PB_BATCH_STATS; ++i) out[i] = b->stats[i];
  return 0;
}
int pb_fbrush_counters(pb_fbrush* b, uint64_t* visited, uint64_t* active) {
  PB_API_BEGIN
  PB_REQUIRE(b != nullptr, "pb_fbrush_counters: null handle");
  DeviceGuard g(b->ctx);
  unsigned long long h[2];
  PB_CUDA(cudaMemcpyAsync(h, b->d_counters, sizeof(h), cudaMemcpyDeviceToHost, b->ctx->stream));
  PB_CUDA(cudaStreamSynchronize(b->ctx->stream));
  if (active) *active = h[0];
  if (visited) *visited = h[1];
  PB_API_END
}

// ---- TextureBrush -----------------------------------------------------------------------------------------
namespace {
// TextureBrush::paintStroke with _useSmudge (TextureBrush.hxx:52-205 + Smudge.hxx). The smudge windows and their
// orientation carry over from stroke to stroke, so strokes are a serial chain: per stroke a parallel thickness pass,
// the serial walk along the spine (one cluster), and a parallel deposit pass, all ordered by the stream.
void smudge_strokes(pb_tbrush* b, pb_canvas* c, int64_t n_strokes, const pb_tstroke* strokes, int64_t n_vertices,
                    const double* path_xy) {
  pb_context* ctx = b->ctx;
  double radius   = b->radius;
  for (int64_t s = 0; s < n_strokes; ++s) {
    const pb_tstroke& in = strokes[s];
    PB_REQUIRE(in.first_vertex >= 0 && in.n_vertices >= 0 && in.first_vertex + in.n_vertices <= n_vertices,
               "stroke vertex range out of bounds");
    if (!(std::fabs(radius - in.radius) < 0.5)) {
      PB_REQUIRE(pb_tbrush_set_radius(b, in.radius) == 0, pb_last_error());
      radius = b->radius;
    }
    for (int i = 0; i < 3; ++i) {
      b->paintK[i] = in.K[i];
      b->paintS[i] = in.S[i];
    }
    const host::TextureFrame f = host::build_texture_frame(reinterpret_cast<const host::V2*>(path_xy) + in.first_vertex,
                                                           in.n_vertices, radius, c->rows, c->cols);
    if (!f.valid) continue;
    PB_REQUIRE(f.poly.size() <= static_cast<size_t>(kMaxPoly), "stroke has too many vertices (max 510 per stroke)");
    SmudgeLaunch L{};
    for (int p = 0; p < kLayerPlanes; ++p) L.canvas[p] = c->pl.plane(p);
    L.rows        = c->rows;
    L.cols        = c->cols;
    L.store_first = c->store_first;
    L.store_rows  = c->pl.rows;
    DevTStroke& d = L.stroke;
    PB_REQUIRE(in.texture_id >= 0 && static_cast<size_t>(in.texture_id) < b->textures.size(), "stroke texture id out of range");
    d.map      = b->textures[static_cast<size_t>(in.texture_id)].d_map;
    d.map_rows = b->textures[static_cast<size_t>(in.texture_id)].rows;
    d.map_cols = b->textures[static_cast<size_t>(in.texture_id)].cols;
    for (int i = 0; i < 3; ++i) {
      d.K[i] = in.K[i];
      d.S[i] = in.S[i];
    }
    d.thickness_scale = in.thickness_scale;
    d.x0 = f.x0, d.x1 = f.x1, d.y0 = f.y0, d.y1 = f.y1;
    d.local_rows = std::max(f.local_rows, 0);
    d.local_cols = std::max(f.local_cols, 0);
    d.poly_begin = 0;
    d.n_poly     = static_cast<int32_t>(f.poly.size());
    const size_t n_local = static_cast<size_t>(d.local_rows) * d.local_cols;
    DevBuf<double2> d_poly(ctx, f.poly.size()), d_uv(ctx, f.uv.size());
    DevBuf<double> d_tmap(ctx, n_local);
    d_poly.upload(reinterpret_cast<const double2*>(f.poly.data()), f.poly.size());
    d_uv.upload(reinterpret_cast<const double2*>(f.uv.data()), f.uv.size());
    d_tmap.zero(n_local);
    PB_CUDA(cudaMemsetAsync(b->d_counters + 1, 0, sizeof(unsigned long long), ctx->stream));
    L.poly     = d_poly.p;
    L.uv       = d_uv.p;
    L.tmap     = d_tmap.p;
    L.max_bits = b->d_counters + 1;
    L.counters = b->d_counters;
    c->version++;
    texture_thickness_launch(ctx, L);
    // Smudge::smudge returns before touching its state when the stroke deposits nothing (:41-47)
    unsigned long long max_bits = 0;
    PB_CUDA(cudaMemcpyAsync(&max_bits, b->d_counters + 1, sizeof(max_bits), cudaMemcpyDeviceToHost, ctx->stream));
    PB_CUDA(cudaStreamSynchronize(ctx->stream));
    double maxD;
    std::memcpy(&maxD, &max_bits, sizeof(maxD));
    std::vector<host::SmudgeStep> steps;
    if (maxD > 0.0) host::build_smudge_steps(f, b->smudge_size, b->smudge_rotation, steps);
    static_assert(sizeof(host::SmudgeStep) == sizeof(DevSmudgeStep), "SmudgeStep layout");
    DevBuf<DevSmudgeStep> d_steps(ctx, steps.size());
    d_steps.upload(reinterpret_cast<const DevSmudgeStep*>(steps.data()), steps.size());
    if (!steps.empty() && b->smudge_size > 0) {
      for (int w = 0; w < 2; ++w)
        for (int p = 0; p < kLayerPlanes; ++p) L.pick[w][p] = b->smudge_map[w].plane(p);
      L.size            = b->smudge_size;
      L.max_size        = b->smudge_max_size;
      L.first_dst       = b->smudge_dst;
      L.steps           = d_steps.p;
      L.n_steps         = static_cast<int>(steps.size());
      L.bmin_x          = f.bound_min.x;
      L.bmin_y          = f.bound_min.y;
      L.pickup_rate     = 0.1;  // Smudge.hxx:156-158
      L.deposition_rate = 0.1;
      texture_smudge_launch(ctx, L);
      b->smudge_dst = (b->smudge_dst + L.n_steps) & 1;
    }
    texture_deposit_launch(ctx, L);
  }
}
}  // namespace

namespace {
pb_tbrush::Texture upload_texture(pb_context* ctx, int rows, int cols, const double* map) {
  pb_tbrush::Texture t;
  t.rows = rows;
  t.cols = cols;
  const size_t bytes = sizeof(double) * static_cast<size_t>(rows) * cols;
  PB_CUDA(cudaMalloc(reinterpret_cast<void**>(&t.d_map), bytes));
  PB_CUDA(cudaMemcpyAsync(t.d_map, map, bytes, cudaMemcpyHostToDevice, ctx->stream));
  PB_CUDA(cudaStreamSynchronize(ctx->stream));  // the host buffer is the caller's
  return t;
}
}  // namespace

int pb_tbrush_create(pb_context* ctx, int map_rows, int map_cols, const double* thickness_map, pb_tbrush** out) {
  PB_API_BEGIN
  PB_REQUIRE(ctx != nullptr, "pb_tbrush_create: null handle");
  DeviceGuard g(ctx);
  PB_REQUIRE(map_rows > 0 && map_cols > 0 && thickness_map != nullptr, "texture brush needs a thickness map");
  auto b      = std::make_unique<pb_tbrush>();
  b->ctx      = ctx;
  b->textures.push_back(upload_texture(ctx, map_rows, map_cols, thickness_map));
  PB_CUDA(cudaMalloc(reinterpret_cast<void**>(&b->d_counters), 2 * sizeof(unsigned long long)));
  PB_CUDA(cudaMemsetAsync(b->d_counters, 0, 2 * sizeof(unsigned long long), ctx->stream));
  PB_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = b.release();
  PB_API_END
}
int pb_tbrush_destroy(pb_tbrush* b) {
  PB_API_BEGIN
  if (b) {
    DeviceGuard g(b->ctx);
    PB_CUDA(cudaStreamSynchronize(b->ctx->stream));
    for (auto& t : b->textures) cudaFree(t.d_map);
    cudaFree(b->d_counters);
    planes_free(b->smudge_map[0]);
    planes_free(b->smudge_map[1]);
    delete b;
  }
  PB_API_END
}
int pb_tbrush_add_texture(pb_tbrush* b, int map_rows, int map_cols, const double* thickness_map, int* texture_id) {
  PB_API_BEGIN
  PB_REQUIRE(b != nullptr, "pb_tbrush_add_texture: null handle");
  PB_REQUIRE(map_rows > 0 && map_cols > 0 && thickness_map != nullptr, "pb_tbrush_add_texture: empty texture");
  DeviceGuard g(b->ctx);
  b->textures.push_back(upload_texture(b->ctx, map_rows, map_cols, thickness_map));
  if (texture_id) *texture_id = static_cast<int>(b->textures.size()) - 1;
  PB_API_END
}
int pb_tbrush_select_texture(pb_tbrush* b, int texture_id) {
  PB_API_BEGIN
  PB_REQUIRE(b != nullptr, "pb_tbrush_select_texture: null handle");
  PB_REQUIRE(texture_id >= 0 && static_cast<size_t>(texture_id) < b->textures.size(), "texture id out of range");
  b->current_texture = texture_id;
  PB_API_END
}
int pb_tbrush_texture_count(const pb_tbrush* b) { return b ? static_cast<int>(b->textures.size()) : 0; }
int pb_tbrush_set_radius(pb_tbrush* b, double radius) {
  PB_API_BEGIN
  PB_REQUIRE(b != nullptr, "pb_tbrush_set_radius: null handle");
  if (!(std::fabs(b->radius - radius) < 0.5)) {  // TextureBrush.hxx:33-41
    b->radius = radius;
    if (b->use_smudge) {  // _smudge = Smudge(int(2 * radius)): fresh, clean windows
      DeviceGuard g(b->ctx);
      const int size = static_cast<int32_t>(2.0 * radius);
      PB_CUDA(cudaStreamSynchronize(b->ctx->stream));
      for (auto& m : b->smudge_map) {
        planes_free(m);
        if (size > 0) {
          planes_alloc(b->ctx, m, size, size, kLayerPlanes);
          for (int p = 0; p < kLayerPlanes; ++p) fill_plane(b->ctx, m.plane(p), m.n(), 0.0);
        }
      }
      b->smudge_size     = size;
      b->smudge_max_size = (size % 2 == 0) ? size + 1 : size;  // Smudge.hxx:25-27
      b->smudge_dst      = 0;
      b->smudge_rotation = 0.0;
    }
  }
  PB_API_END
}
int pb_tbrush_enable_smudge(pb_tbrush* b, int enable) {
  PB_CHECK_HANDLE(b, "pb_tbrush_enable_smudge");
  b->use_smudge = enable != 0;
  return 0;
}
int pb_tbrush_dip(pb_tbrush* b, const double K[3], const double S[3]) {
  PB_CHECK_HANDLE(b, "pb_tbrush_dip");
  for (int i = 0; i < 3; ++i) {
    b->paintK[i] = K[i];
    b->paintS[i] = S[i];
  }
  return 0;
}
int pb_tbrush_set_thickness_scale(pb_tbrush* b, double scale) {
  PB_CHECK_HANDLE(b, "pb_tbrush_set_thickness_scale");
  b->thickness_scale = scale;
  return 0;
}
int pb_tbrush_stroke_batch(pb_tbrush* b, pb_canvas* c, int64_t n_strokes, const pb_tstroke* strokes, int64_t n_vertices,
                           const double* path_xy) {
  PB_API_BEGIN
  PB_REQUIRE(b != nullptr, "pb_tbrush_stroke_batch: null handle");
  PB_REQUIRE(c != nullptr, "pb_tbrush_stroke_batch: null handle");
  pb_context* ctx = b->ctx;
  DeviceGuard g(ctx);
  PB_REQUIRE(c->pl.ctx == ctx, "canvas and brush belong to different contexts");
  if (n_strokes <= 0) return 0;
  if (b->use_smudge) {
    smudge_strokes(b, c, n_strokes, strokes, n_vertices, path_xy);
    return 0;
  }
  std::vector<DevTStroke> ds;
  ds.reserve(static_cast<size_t>(n_strokes));
  std::vector<host::V2> poly, uv;
  const int tile = kTextureTile, tiles_x = (c->cols + tile - 1) / tile, tiles_y = (c->rows + tile - 1) / tile;
  int64_t n_items = 0;
  double radius   = b->radius;
  for (int64_t s = 0; s < n_strokes; ++s) {
    const pb_tstroke& in = strokes[s];
    PB_REQUIRE(in.first_vertex >= 0 && in.n_vertices >= 0 && in.first_vertex + in.n_vertices <= n_vertices,
               "stroke vertex range out of bounds");
    if (!(std::fabs(radius - in.radius) < 0.5)) radius = in.radius;
    const host::TextureFrame f = host::build_texture_frame(reinterpret_cast<const host::V2*>(path_xy) + in.first_vertex,
                                                           in.n_vertices, radius, c->rows, c->cols);
    DevTStroke d{};
    for (int i = 0; i < 3; ++i) {
      d.K[i] = in.K[i];
      d.S[i] = in.S[i];
    }
    d.thickness_scale = in.thickness_scale;
    PB_REQUIRE(in.texture_id >= 0 && static_cast<size_t>(in.texture_id) < b->textures.size(), "stroke texture id out of range");
    d.map             = b->textures[static_cast<size_t>(in.texture_id)].d_map;
    d.map_rows        = b->textures[static_cast<size_t>(in.texture_id)].rows;
    d.map_cols        = b->textures[static_cast<size_t>(in.texture_id)].cols;
    d.poly_begin      = static_cast<int32_t>(poly.size());
    d.item_begin      = n_items;
    d.tx0 = d.ty0 = 0;
    d.tx1 = d.ty1 = -1;
    if (f.valid) {
      PB_REQUIRE(f.poly.size() <= static_cast<size_t>(kMaxPoly), "stroke has too many vertices (max 510 per stroke)");
      d.x0 = f.x0, d.x1 = f.x1, d.y0 = f.y0, d.y1 = f.y1;
      d.local_rows = f.local_rows;
      d.local_cols = f.local_cols;
      d.n_poly     = static_cast<int32_t>(f.poly.size());
      poly.insert(poly.end(), f.poly.begin(), f.poly.end());
      uv.insert(uv.end(), f.uv.begin(), f.uv.end());
      // tiles of the part of the bounding box that lies on the canvas and in the stored rows
      const int bx0 = std::max(f.x0, 0), bx1 = std::min(f.x1, c->cols - 1);
      const int by0 = std::max(f.y0, c->store_first), by1 = std::min(f.y1, c->store_first + c->pl.rows - 1);
      if (bx1 >= bx0 && by1 >= by0) {
        d.tx0 = bx0 / tile, d.tx1 = bx1 / tile, d.ty0 = by0 / tile, d.ty1 = by1 / tile;
        n_items += static_cast<int64_t>(d.tx1 - d.tx0 + 1) * (d.ty1 - d.ty0 + 1);
      }
    } else {
      d.x0 = d.y0 = 0;
      d.x1 = d.y1 = -1;
      d.n_poly    = 0;
    }
    ds.push_back(d);
  }
  b->radius = radius;
  for (int i = 0; i < 3; ++i) {
    b->paintK[i] = strokes[n_strokes - 1].K[i];
    b->paintS[i] = strokes[n_strokes - 1].S[i];
  }
  static_assert(sizeof(host::V2) == sizeof(double2), "V2 must match double2");
  DevBuf<DevTStroke> d_strokes(ctx, ds.size());
  DevBuf<double2> d_poly(ctx, poly.size()), d_uv(ctx, uv.size());
  DevBuf<int32_t> d_ticket(ctx, static_cast<size_t>(n_items));
  DevBuf<int> d_tiles(ctx, static_cast<size_t>(tiles_x) * tiles_y + 2);
  d_strokes.upload(ds.data(), ds.size());
  d_poly.upload(reinterpret_cast<const double2*>(poly.data()), poly.size());
  d_uv.upload(reinterpret_cast<const double2*>(uv.data()), uv.size());
  d_tiles.zero(static_cast<size_t>(tiles_x) * tiles_y + 2);
  TextureLaunch L{};
  for (int p = 0; p < kLayerPlanes; ++p) L.canvas[p] = c->pl.plane(p);
  L.rows        = c->rows;
  L.cols        = c->cols;
  L.store_first = c->store_first;
  L.store_rows  = c->pl.rows;
  L.strokes     = d_strokes.p;
  L.n_strokes   = n_strokes;
  L.poly        = d_poly.p;
  L.uv          = d_uv.p;
  L.tile        = tile;
  L.tiles_x     = tiles_x;
  L.tiles_y     = tiles_y;
  L.n_items     = n_items;
  L.ticket      = d_ticket.p;
  L.tile_done   = d_tiles.p;
  L.queue       = reinterpret_cast<unsigned long long*>(d_tiles.p + static_cast<size_t>(tiles_x) * tiles_y + (((tiles_x * tiles_y) & 1) ? 1 : 0));
  L.counters    = b->d_counters;
  c->version++;
  texture_launch(ctx, L);
  PB_API_END
}
int pb_tbrush_paint_stroke(pb_tbrush* b, pb_canvas* c, int n, const double* path_xy) {
  PB_CHECK_HANDLE(b, "pb_tbrush_paint_stroke");
  pb_tstroke s{};
  s.radius = b->radius;
  for (int i = 0; i < 3; ++i) {
    s.K[i] = b->paintK[i];
    s.S[i] = b->paintS[i];
  }
  s.thickness_scale = b->thickness_scale;
  s.first_vertex    = 0;
  s.n_vertices      = n;
  s.texture_id      = b->current_texture;
  return pb_tbrush_stroke_batch(b, c, 1, &s, n, path_xy);
}
int pb_tbrush_counters(pb_tbrush* b, uint64_t* pixels) {
  PB_API_BEGIN
  PB_REQUIRE(b != nullptr, "pb_tbrush_counters: null handle");
  DeviceGuard g(b->ctx);
  unsigned long long h = 0;
  PB_CUDA(cudaMemcpyAsync(&h, b->d_counters, sizeof(h), cudaMemcpyDeviceToHost, b->ctx->stream));
  PB_CUDA(cudaStreamSynchronize(b->ctx->stream));
  if (pixels) *pixels = h;
  PB_API_END
}

// ---- TextureBrushDictionary (host only) ----------------------------------------------------------------------------
struct pb_texdict {
  // groups[i0][i1] = entries, size keys and length keys ascending (std::map order, TextureBrushDictionary.cxx:92-118)
  std::vector<std::vector<std::vector<int32_t>>> groups;
  std::vector<double> avg_sizes;                // :120-137
  std::vector<std::vector<double>> avg_length;  // :139-164
};
int pb_texdict_create(int n, const int32_t* size_key, const int32_t* length_key, const int32_t* rows, const int32_t* cols,
                      pb_texdict** out) {
  PB_API_BEGIN
  PB_REQUIRE(n > 0 && size_key && length_key && rows && cols && out, "pb_texdict_create: bad arguments");
  std::map<uint32_t, std::map<uint32_t, std::vector<int32_t>>> m;
  for (int i = 0; i < n; ++i) m[static_cast<uint32_t>(size_key[i])][static_cast<uint32_t>(length_key[i])].push_back(i);
  auto d = std::make_unique<pb_texdict>();
  for (const auto& e : m) {
    d->groups.emplace_back();
    for (const auto& a : e.second) d->groups.back().push_back(a.second);
  }
  d->avg_sizes.assign(d->groups.size(), 0.0);
  d->avg_length.resize(d->groups.size());
  for (size_t i = 0; i < d->groups.size(); ++i) {
    uint32_t count = 0;
    d->avg_length[i].assign(d->groups[i].size(), 0.0);
    for (size_t j = 0; j < d->groups[i].size(); ++j) {
      uint32_t cj = 0;
      for (const int32_t t : d->groups[i][j]) {
        d->avg_sizes[i] += static_cast<double>(rows[t]);
        d->avg_length[i][j] += static_cast<double>(cols[t]);
        ++count;
        ++cj;
      }
      d->avg_length[i][j] *= (1.0 / static_cast<double>(cj));
    }
    d->avg_sizes[i] *= (1.0 / static_cast<double>(count));
  }
  *out = d.release();
  PB_API_END
}
int pb_texdict_destroy(pb_texdict* d) {
  delete d;
  return 0;
}
int pb_texdict_lookup(const pb_texdict* d, int n, const double* path_xy, double brush_size, int32_t* size_group,
                      int32_t* length_group, int capacity, int32_t* candidates, int32_t* n_candidates) {
  PB_API_BEGIN
  PB_REQUIRE(d != nullptr && n >= 1 && path_xy != nullptr, "pb_texdict_lookup: bad arguments");
  double length = 0.0;  // :27-30
  for (int i = 0; i + 1 < n; ++i) {
    const double dx = path_xy[2 * i] - path_xy[2 * i + 2], dy = path_xy[2 * i + 1] - path_xy[2 * i + 3];
    length += std::sqrt(dx * dx + dy * dy);
  }
  uint32_t i0 = 0, i1 = 1;  // :32-33
  double mr = d->avg_sizes[0];
  for (uint32_t i = 0; i < d->avg_sizes.size(); ++i) {  // :36-43
    const double dd = std::abs(d->avg_sizes[i] - brush_size);
    if (dd < mr) {
      mr = dd;
      i0 = i;
    }
  }
  double ml = d->avg_length[i0][0];
  const size_t n_len = std::min(d->avg_sizes.size(), d->avg_length[i0].size());  // the reference reads out of range beyond
  for (uint32_t i = 0; i < n_len; ++i) {                                         // :46-53
    const double dd = std::abs(d->avg_length[i0][i] - length);
    if (dd < ml) {
      ml = dd;
      i1 = i;
    }
  }
  PB_REQUIRE(i1 < d->groups[i0].size() && !d->groups[i0][i1].empty(), "no candidate found");  // :57-59
  const auto& cand = d->groups[i0][i1];
  if (size_group) *size_group = static_cast<int32_t>(i0);
  if (length_group) *length_group = static_cast<int32_t>(i1);
  if (n_candidates) *n_candidates = static_cast<int32_t>(cand.size());
  for (int k = 0; k < capacity && k < static_cast<int>(cand.size()) && candidates; ++k) candidates[k] = cand[static_cast<size_t>(k)];
  PB_API_END
}

}  // extern "C"
