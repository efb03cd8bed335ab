// A C++ host driving the band-sharded (multi-GPU) path through the C ABI alone — no Python, no torch, no NCCL:
//   one process per GPU (fork), CUDA-IPC handles exchanged through a shared-memory page, a process barrier built on it,
//   footprint strokes that cross the band boundary (pb_fbrush_stroke_batch_dist: peer memory over NVLink, cross-GPU stroke
//   dependencies), and the final image assembled by the compose kernel's peer stores (pb_canvas_compose_gather) into rank
//   0's pb_band_image. Rank 0 then renders the same stroke list on ONE GPU and checks that the two images are identical.
//
// Build:  g++ -std=c++17 -O2 -I include examples/band_two_process.cpp painty_b200/libpainty_b200.so \
//             -Wl,-rpath,'$ORIGIN/../painty_b200' -o examples/_build/band_two_process
// Run:    examples/_build/band_two_process [world=2]      (needs `world` B200s on one node)
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "painty_b200.h"

namespace {

constexpr int kMaxWorld = PB_MAX_BANDS;

struct Shared {  // one page shared by all ranks (mmap before fork)
  std::atomic<int> arrived[2];
  std::atomic<int> sense;
  std::atomic<int> failed;
  unsigned char canvas[kMaxWorld][PB_IPC_HANDLE_BYTES], snapshot[kMaxWorld][PB_IPC_HANDLE_BYTES];
  unsigned char dirty[kMaxWorld][PB_IPC_HANDLE_BYTES], flags[kMaxWorld][PB_IPC_HANDLE_BYTES];
  unsigned char image[PB_IPC_HANDLE_BYTES];
  int64_t image_stride;
};

void barrier(Shared* sh, int world, int& phase) {  // sense-reversing barrier over the shared page
  const int p = phase & 1;
  if (sh->arrived[p].fetch_add(1) + 1 == world) {
    sh->arrived[p].store(0);
    sh->sense.store(phase + 1);
  } else {
    while (sh->sense.load() <= phase && !sh->failed.load()) usleep(50);
  }
  ++phase;
}

#define CHECK(call)                                                                      \
  do {                                                                                   \
    if ((call) != 0) {                                                                   \
      std::fprintf(stderr, "rank %d: %s failed: %s\n", rank, #call, pb_last_error());    \
      sh->failed.store(1);                                                               \
      return 1;                                                                          \
    }                                                                                    \
  } while (0)

// A synthetic padded footprint with the geometry FootprintBrush::setRadius produces (FootprintBrush.hxx:46-63); painty's own
// imRead + ScaledMat + PaddedMat supply the real one in an integration.
std::vector<double> footprint(double radius, int& side) {
  const int width = static_cast<int>(2.0 * std::ceil(radius) + 1.0);
  const int size_map = static_cast<int>(std::ceil(std::sqrt(2.0) * width));
  const int pad = (size_map - width) / 2;
  side = width + 2 * pad;
  std::vector<double> fp(static_cast<size_t>(side) * side, 0.0);
  const double c = (side - 1) / 2.0;
  for (int y = 0; y < side; ++y)
    for (int x = 0; x < side; ++x) {
      const double d = std::hypot(x - c, y - c) / radius;
      if (d < 0.8 && (x + 2 * y) % 3 != 0) fp[static_cast<size_t>(y) * side + x] = (1.0 - d) * (0.5 + 0.5 * std::cos(0.7 * x));
    }
  return fp;
}

struct Strokes {
  std::vector<pb_stroke> rec;
  std::vector<double> cx, cy, th;
};
Strokes make_strokes(int rows, int cols, const double* radii, int n_radii) {
  Strokes s;
  unsigned seed = 12345;
  auto rnd = [&]() { return (seed = seed * 1664525u + 1013904223u) / 4294967296.0; };
  for (int i = 0; i < 40; ++i) {
    pb_stroke r{};
    r.radius = radii[i % n_radii];
    for (int k = 0; k < 3; ++k) r.K[k] = 0.05 + rnd(), r.S[k] = 0.05 + 0.8 * rnd();
    double path[8];
    double x = rnd() * cols, y = rows * (0.25 + 0.5 * rnd()), a = 6.28 * rnd();  // around the band boundaries
    for (int p = 0; p < 4; ++p) {
      path[2 * p] = x, path[2 * p + 1] = y;
      a += rnd() - 0.5;
      x += 40 * std::cos(a), y += 40 * std::sin(a);
    }
    int64_t n = 0;
    pb_expand_stroke(0, 4, path, 0, nullptr, nullptr, nullptr, &n);
    r.first_imprint = static_cast<int64_t>(s.cx.size());
    r.n_imprints    = n;
    s.cx.resize(s.cx.size() + n), s.cy.resize(s.cy.size() + n), s.th.resize(s.th.size() + n);
    pb_expand_stroke(0, 4, path, n, s.cx.data() + r.first_imprint, s.cy.data() + r.first_imprint, s.th.data() + r.first_imprint, &n);
    s.rec.push_back(r);
  }
  return s;
}

int run_rank(int rank, int world, Shared* sh) {
  const int rows = 600 * world, cols = 1000, rpb = 600;
  int phase = 0;
  pb_context* ctx = nullptr;
  CHECK(pb_context_create(rank, PB_F32, &ctx));
  pb_canvas* band = nullptr;
  CHECK(pb_canvas_create_band(ctx, rows, cols, rank * rpb, std::min(rows, (rank + 1) * rpb), 0, &band));
  pb_fbrush* brush = nullptr;
  CHECK(pb_fbrush_create(ctx, &brush));
  const double radii[3] = {30.0, 40.0, 64.0};
  for (double r : radii) {
    int side = 0;
    const std::vector<double> fp = footprint(r, side);
    CHECK(pb_fbrush_register_footprint(brush, r, side, fp.data()));
  }
  const Strokes st = make_strokes(rows, cols, radii, 3);  // every rank builds the same global list

  // exchange the peer mappings of canvas records, snapshot records, dirty map and progress flags
  void *wrec = nullptr, *srec = nullptr, *dirty = nullptr, *flags = nullptr;
  CHECK(pb_fbrush_dist_storage(brush, band, &wrec, &srec, &dirty, &flags));
  CHECK(pb_ipc_export(ctx, wrec, sh->canvas[rank]));
  CHECK(pb_ipc_export(ctx, srec, sh->snapshot[rank]));
  CHECK(pb_ipc_export(ctx, dirty, sh->dirty[rank]));
  CHECK(pb_ipc_export(ctx, flags, sh->flags[rank]));
  pb_band_image* image = nullptr;
  void* image_base     = nullptr;
  if (rank == 0) {  // gather to rank 0
    CHECK(pb_band_image_create(ctx, rows, cols, &image));
    CHECK(pb_band_image_device(image, &image_base, &sh->image_stride));
    CHECK(pb_ipc_export(ctx, image_base, sh->image));
  }
  barrier(sh, world, phase);
  pb_dist_desc d{};
  d.world = world, d.rank = rank, d.rows_per_band = rpb;
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      d.canvas_base[r] = wrec, d.snapshot_base[r] = srec, d.dirty_base[r] = dirty, d.flags_base[r] = flags;
    } else {
      CHECK(pb_ipc_import(ctx, sh->canvas[r], &d.canvas_base[r]));
      CHECK(pb_ipc_import(ctx, sh->snapshot[r], &d.snapshot_base[r]));
      CHECK(pb_ipc_import(ctx, sh->dirty[r], &d.dirty_base[r]));
      CHECK(pb_ipc_import(ctx, sh->flags[r], &d.flags_base[r]));
    }
  }
  if (rank != 0) CHECK(pb_ipc_import(ctx, sh->image, &image_base));

  // the batch: planes -> records, barrier, kernels (peer memory + cross-GPU flags), barrier, records -> planes
  CHECK(pb_fbrush_dist_begin(brush, band));
  barrier(sh, world, phase);
  CHECK(pb_fbrush_stroke_batch_dist(brush, band, &d, static_cast<int64_t>(st.rec.size()), st.rec.data(),
                                    static_cast<int64_t>(st.cx.size()), st.cx.data(), st.cy.data(), st.th.data()));
  CHECK(pb_context_synchronize(ctx));
  barrier(sh, world, phase);
  CHECK(pb_fbrush_dist_end(brush, band));

  // compose + gather: every rank stores its band's reflectance rows into rank 0's image
  void* dst[1] = {image_base};
  CHECK(pb_canvas_compose_gather(band, 1, dst, sh->image_stride));
  CHECK(pb_context_synchronize(ctx));
  barrier(sh, world, phase);

  int rc = 0;
  if (rank == 0) {
    std::vector<double> got(static_cast<size_t>(rows) * cols * 3), want(got.size());
    CHECK(pb_band_image_download(image, got.data()));
    pb_canvas* full = nullptr;
    pb_fbrush* b1   = nullptr;
    CHECK(pb_canvas_create(ctx, rows, cols, &full));
    CHECK(pb_fbrush_create(ctx, &b1));
    for (double r : radii) {
      int side = 0;
      const std::vector<double> fp = footprint(r, side);
      CHECK(pb_fbrush_register_footprint(b1, r, side, fp.data()));
    }
    CHECK(pb_fbrush_stroke_batch(b1, full, static_cast<int64_t>(st.rec.size()), st.rec.data(), static_cast<int64_t>(st.cx.size()),
                                 st.cx.data(), st.cy.data(), st.th.data()));
    CHECK(pb_canvas_compose(full, want.data()));
    size_t diff = 0, painted = 0;
    for (size_t i = 0; i < got.size(); ++i) {
      diff += got[i] != want[i];
      painted += want[i] != 1.0;
    }
    std::printf("band_two_process: world %d, %zu strokes, %zu imprints, %zu painted values, %zu differing values -> %s\n", world,
                st.rec.size(), st.cx.size(), painted, diff, (diff == 0 && painted > 0) ? "OK" : "MISMATCH");
    rc = (diff == 0 && painted > 0) ? 0 : 1;
    pb_fbrush_destroy(b1);
    pb_canvas_destroy(full);
  }
  barrier(sh, world, phase);  // nobody unmaps while a peer may still read
  for (int r = 0; r < world; ++r) {
    if (r == rank) continue;
    pb_ipc_close(ctx, d.canvas_base[r]), pb_ipc_close(ctx, d.snapshot_base[r]);
    pb_ipc_close(ctx, d.dirty_base[r]), pb_ipc_close(ctx, d.flags_base[r]);
  }
  if (rank != 0) pb_ipc_close(ctx, image_base);
  barrier(sh, world, phase);
  if (image) pb_band_image_destroy(image);
  pb_fbrush_destroy(brush);
  pb_canvas_destroy(band);
  pb_context_destroy(ctx);
  return rc;
}

}  // namespace

int main(int argc, char** argv) {
  const int world = argc > 1 ? std::atoi(argv[1]) : 2;
  if (world < 1 || world > kMaxWorld) return 2;
  auto* sh = static_cast<Shared*>(mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0));
  if (sh == MAP_FAILED) return 2;
  new (sh) Shared();
  std::vector<pid_t> kids;
  for (int r = 1; r < world; ++r) {  // fork BEFORE any CUDA call: every process creates its own CUDA context
    const pid_t pid = fork();
    if (pid == 0) _exit(run_rank(r, world, sh));
    kids.push_back(pid);
  }
  int rc = run_rank(0, world, sh);
  for (pid_t pid : kids) {
    int status = 0;
    waitpid(pid, &status, 0);
    if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) rc = 1;
  }
  return rc;
}
