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
#include <cmath>
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

// The BOX a run of footprint imprints modifies (footprint square of half side `half_side` around every centre) and
// everything it reads or writes (the union of the snapshot "allowed" boxes, FootprintBrush.hxx:298-305: the box
// grown by the radius). Both padded by 2 px, clipped to the canvas.
inline void imprint_regions(int64_t first, int64_t count, int half_side, double radius, const double* cx, const double* cy,
                            int rows, int cols, Region& box, Region& allowed) {
  box = allowed = Region{1, 1, 0, 0};
  if (count <= 0) return;
  double lx = cx[first], hx = lx, ly = cy[first], hy = ly;
  for (int64_t i = first; i < first + count; ++i) {
    lx = std::min(lx, cx[i]);
    hx = std::max(hx, cx[i]);
    ly = std::min(ly, cy[i]);
    hy = std::max(hy, cy[i]);
  }
  auto grow = [&](double margin) {
    const double m = half_side + margin + 2.0;
    Region r;
    r.x0 = static_cast<int>(std::max(0.0, std::floor(lx - m)));
    r.y0 = static_cast<int>(std::max(0.0, std::floor(ly - m)));
    r.x1 = static_cast<int>(std::min<double>(cols - 1, std::ceil(hx + m)));
    r.y1 = static_cast<int>(std::min<double>(rows - 1, std::ceil(hy + m)));
    return r;
  };
  box     = grow(0.0);
  allowed = grow(radius);
}

// Dataflow graph of a footprint stroke list at SEGMENT granularity. A stroke is cut into segments of seg_len
// consecutive imprints; segment k of a stroke may start once every EARLIER stroke g listed for it has completed
// `need` segments. The executor publishes a stroke's progress at segment boundaries, so a later stroke starts (or
// continues) as soon as the earlier ones have moved past the part of the canvas its next segment needs — the
// critical path of densely overlapping stroke lists shrinks accordingly.
// The planner sees the segments in submission order. A dependency on segment m of stroke g becomes "g has
// completed m + 1 segments" (largest requirement per g kept); dependencies on the own stroke are dropped: its
// segments run in order anyway, and each of them registered the predecessors it displaced in the tile tables, so
// the induction argument of DataflowPlanner still holds.
struct StrokeSpan {
  int64_t first, count;  // imprints [first, first + count) of the cx/cy arrays
  int half_side;         // (footprint side - 1) / 2
  double radius;
  bool single;           // keep the stroke in one segment (it reads its whole region up front)
};
struct SegmentPlan {
  std::vector<int32_t> seg_first;  // [n + 1] first segment of each stroke
  std::vector<int32_t> seg_len;    // [n] imprints per segment
  std::vector<int32_t> seg_off;    // [segments + 1] CSR offsets into pred_stroke / pred_need
  std::vector<int32_t> pred_stroke, pred_need;
};
constexpr int kMaxSegmentsPerStroke = 4000;

template <typename SpanFn, typename VisitFn>
SegmentPlan plan_segments(int rows, int cols, size_t n, SpanFn&& span_of, const double* cx, const double* cy, int segment_length,
                          bool use_snapshot, VisitFn&& visit_stroke) {
  SegmentPlan plan;
  plan.seg_first.assign(n + 1, 0);
  plan.seg_len.assign(n, 1);
  plan.seg_off.assign(1, 0);
  std::vector<int32_t> owner, raw;
  DataflowPlanner planner(rows, cols);
  for (size_t s = 0; s < n; ++s) {
    const StrokeSpan sp = span_of(s);
    Region box, allowed;
    imprint_regions(sp.first, sp.count, sp.half_side, sp.radius, cx, cy, rows, cols, box, allowed);
    visit_stroke(s, box, allowed);
    const int64_t whole = std::max<int64_t>(sp.count, 1);
    int64_t len = (segment_length <= 0 || sp.single)
                      ? whole
                      : std::max<int64_t>(segment_length, (whole + kMaxSegmentsPerStroke - 1) / kMaxSegmentsPerStroke);
    len                  = std::min(len, whole);
    const int nseg       = static_cast<int>((whole + len - 1) / len);
    plan.seg_len[s]      = static_cast<int32_t>(len);
    plan.seg_first[s + 1] = plan.seg_first[s] + nseg;
    for (int k = 0; k < nseg; ++k) {
      const int32_t gseg = plan.seg_first[s] + k;
      Region sbox = box, sall = allowed;
      if (nseg > 1)
        imprint_regions(sp.first + k * len, std::min<int64_t>(len, sp.count - k * len), sp.half_side, sp.radius, cx, cy, rows,
                        cols, sbox, sall);
      raw.clear();
      int32_t rb = 0, re = 0;
      if (use_snapshot) {
        planner.add_footprint(gseg, sbox, sall, raw, rb, re);
      } else {
        planner.add(gseg, sbox, raw, rb, re);
      }
      const size_t begin = plan.pred_stroke.size();
      for (int32_t p : raw) {
        const int32_t g = owner[p];
        if (g == static_cast<int32_t>(s)) continue;
        const int32_t need = p - plan.seg_first[g] + 1;
        bool merged        = false;
        for (size_t q = begin; q < plan.pred_stroke.size() && !merged; ++q) {
          if (plan.pred_stroke[q] == g) {
            plan.pred_need[q] = std::max(plan.pred_need[q], need);
            merged            = true;
          }
        }
        if (!merged) {
          plan.pred_stroke.push_back(g);
          plan.pred_need.push_back(need);
        }
      }
      owner.push_back(static_cast<int32_t>(s));
      plan.seg_off.push_back(static_cast<int32_t>(plan.pred_stroke.size()));
    }
  }
  return plan;
}

}  // namespace pb
