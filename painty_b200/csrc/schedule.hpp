// Host-side dataflow planner shared by both brushes: turns a submission-ordered list of stroke
// regions into predecessor lists so that the device can run independent strokes concurrently while
// every pair of overlapping strokes keeps its submission order (the reference renders strictly one
// stroke after the other on a single thread, SbrRenderThread.cxx:64-73).
//
// The canvas is cut into coarse tiles; for every tile we remember the last stroke that touched it.
// A stroke depends on the distinct "last strokes" of the tiles its region covers. Waiting for those is
// sufficient: each of them in turn waited for the previous toucher of the shared tile (induction).
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace pb {

struct Region {  // inclusive pixel rectangle, already clipped to the canvas; empty if x1 < x0 or y1 < y0
  int x0, y0, x1, y1;
};

class DataflowPlanner {
 public:
  DataflowPlanner(int rows, int cols, int tile = 64)
      : tile_(tile), tx_((cols + tile - 1) / tile), ty_((rows + tile - 1) / tile), last_(static_cast<size_t>(tx_) * ty_, -1) {}

  // Appends the predecessors of stroke `index` (region r) to `preds` and returns [begin,end).
  void add(int32_t index, const Region& r, std::vector<int32_t>& preds, int32_t& begin, int32_t& end) {
    begin = static_cast<int32_t>(preds.size());
    if (r.x1 >= r.x0 && r.y1 >= r.y0) {
      const int tx0 = r.x0 / tile_, tx1 = std::min(r.x1 / tile_, tx_ - 1);
      const int ty0 = r.y0 / tile_, ty1 = std::min(r.y1 / tile_, ty_ - 1);
      for (int ty = ty0; ty <= ty1; ++ty) {
        for (int tx = tx0; tx <= tx1; ++tx) {
          int32_t& l = last_[static_cast<size_t>(ty) * tx_ + tx];
          if (l >= 0 && l != index) {
            bool seen = false;
            for (int32_t k = static_cast<int32_t>(preds.size()) - 1; k >= begin && !seen; --k) seen = preds[k] == l;
            if (!seen) preds.push_back(l);
          }
          l = index;
        }
      }
    }
    end = static_cast<int32_t>(preds.size());
  }

  // Footprint strokes have two footprints on the canvas: the BOX they modify (canvas, snapshot, dirty flags under
  // the brush) and the larger ALLOWED region in which their snapshot ring only refreshes pixels that are already
  // dirty (an idempotent copy of values the stroke itself never changes). Two strokes must keep their order iff the
  // box of one meets the allowed region of the other; two rings that merely overlap each other commute.
  // Per tile we keep the last stroke whose box touched it and the strokes whose ring touched it since then.
  void add_footprint(int32_t index, const Region& box, const Region& allowed, std::vector<int32_t>& preds, int32_t& begin,
                     int32_t& end) {
    begin = static_cast<int32_t>(preds.size());
    if (ring_.empty()) ring_.resize(last_.size());
    auto push = [&](int32_t l) {
      if (l < 0 || l == index) return;
      for (int32_t k = static_cast<int32_t>(preds.size()) - 1; k >= begin; --k)
        if (preds[k] == l) return;
      preds.push_back(l);
    };
    auto tiles = [&](const Region& r, auto&& fn) {
      if (r.x1 < r.x0 || r.y1 < r.y0) return;
      const int tx0 = r.x0 / tile_, tx1 = std::min(r.x1 / tile_, tx_ - 1);
      const int ty0 = r.y0 / tile_, ty1 = std::min(r.y1 / tile_, ty_ - 1);
      for (int ty = ty0; ty <= ty1; ++ty)
        for (int tx = tx0; tx <= tx1; ++tx) fn(static_cast<size_t>(ty) * tx_ + tx);
    };
    tiles(allowed, [&](size_t t) { push(last_[t]); });  // earlier boxes under our ring or box
    tiles(box, [&](size_t t) {                          // earlier rings over our box
      for (int32_t l : ring_[t]) push(l);
    });
    tiles(allowed, [&](size_t t) { ring_[t].push_back(index); });
    tiles(box, [&](size_t t) {
      last_[t] = index;
      ring_[t].clear();
    });
    end = static_cast<int32_t>(preds.size());
  }

 private:
  int tile_, tx_, ty_;
  std::vector<int32_t> last_;
  std::vector<std::vector<int32_t>> ring_;
};

}  // namespace pb
