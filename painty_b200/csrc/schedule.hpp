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
#include <functional>
#include <utility>
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
                          bool use_snapshot, VisitFn&& visit_stroke, int tile = 64) {
  SegmentPlan plan;
  plan.seg_first.assign(n + 1, 0);
  plan.seg_len.assign(n, 1);
  plan.seg_off.assign(1, 0);
  // Same tile tables and rules as DataflowPlanner (last box per tile, rings since then), specialised for segments:
  // entries are global segment ids, consecutive ring entries of one stroke collapse into the latest (a dependency
  // on a later segment implies the earlier ones), and the per-segment lists are deduplicated per STROKE with stamps.
  const int tx = (cols + tile - 1) / tile, ty = (rows + tile - 1) / tile;
  std::vector<int32_t> last(static_cast<size_t>(tx) * ty, -1), owner;
  std::vector<std::vector<int32_t>> ring(use_snapshot ? last.size() : 0);
  std::vector<int32_t> stamp(n, -1), slot(n, 0);
  auto tiles = [&](const Region& r, auto&& fn) {
    if (r.x1 < r.x0 || r.y1 < r.y0) return;
    const int tx0 = r.x0 / tile, tx1 = std::min(r.x1 / tile, tx - 1);
    const int ty0 = r.y0 / tile, ty1 = std::min(r.y1 / tile, ty - 1);
    for (int y = ty0; y <= ty1; ++y)
      for (int x = tx0; x <= tx1; ++x) fn(static_cast<size_t>(y) * tx + x);
  };
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
    const int32_t self   = static_cast<int32_t>(s);
    for (int k = 0; k < nseg; ++k) {
      const int32_t gseg = plan.seg_first[s] + k;
      Region sbox = box, sall = allowed;
      if (nseg > 1)
        imprint_regions(sp.first + k * len, std::min<int64_t>(len, sp.count - k * len), sp.half_side, sp.radius, cx, cy, rows,
                        cols, sbox, sall);
      auto note = [&](int32_t l) {  // segment l of an earlier stroke must be complete
        if (l < 0) return;
        const int32_t g = owner[l];
        if (g == self) return;
        const int32_t need = l - plan.seg_first[g] + 1;
        if (stamp[g] != gseg) {
          stamp[g] = gseg;
          slot[g]  = static_cast<int32_t>(plan.pred_stroke.size());
          plan.pred_stroke.push_back(g);
          plan.pred_need.push_back(need);
        } else if (plan.pred_need[slot[g]] < need) {
          plan.pred_need[slot[g]] = need;
        }
      };
      if (use_snapshot) {
        tiles(sall, [&](size_t t) { note(last[t]); });  // earlier boxes under our ring or box
        tiles(sbox, [&](size_t t) {                     // earlier rings over our box
          for (int32_t l : ring[t]) note(l);
        });
        tiles(sall, [&](size_t t) {
          std::vector<int32_t>& r = ring[t];
          if (!r.empty() && owner[r.back()] == self) {
            r.back() = gseg;
          } else {
            r.push_back(gseg);
          }
        });
        tiles(sbox, [&](size_t t) {
          last[t] = gseg;
          ring[t].clear();
        });
      } else {
        tiles(sbox, [&](size_t t) {
          note(last[t]);
          last[t] = gseg;
        });
      }
      owner.push_back(self);
      plan.seg_off.push_back(static_cast<int32_t>(plan.pred_stroke.size()));
    }
  }
  return plan;
}

// Claim order of the device queues. The imprint kernel's clusters pop strokes from a queue strictly in order and
// block on the dataflow waits, so a stroke that is not ready yet holds a cluster while later strokes that could
// run stay in the queue (head-of-line blocking). The host therefore list-schedules the batch once with a cost
// model and hands the device the resulting CLAIM ORDER: the device still pops in order, but the order is the one
// in which an ideal ready-first scheduler would have started the strokes.
//
// Model: every stroke belongs to a slot POOL (the GPU that executes it) and to a RUN (kernel launch) of that pool;
// a pool works on one run at a time, with slots[pool][run] concurrent strokes. A free slot claims the first
// stroke among the next `window` unclaimed ones of the current run (submission order) whose predecessor strokes
// are all CLAIMED and whose first segment is ready. Claiming only after all predecessors makes the global claim
// sequence T a topological order of the dependency graph, and per-pool orders are T restricted to the pool —
// which is what keeps in-order popping deadlock free: the T-earliest unfinished stroke is always claimed and all
// its predecessors are finished. Deterministic: every rank computes the same T from the same inputs.
struct ClaimSpec {
  int32_t pool, run;
  double cost;  // estimated duration of one imprint (any consistent unit)
};

inline std::vector<int32_t> plan_claim_order(const SegmentPlan& plan, const std::vector<int64_t>& count,
                                             const std::vector<ClaimSpec>& spec, const std::vector<std::vector<int>>& slots,
                                             int window = 64, double* makespan = nullptr) {
  const int32_t n = static_cast<int32_t>(count.size());
  std::vector<int32_t> order;
  order.reserve(static_cast<size_t>(n));
  if (n == 0) return order;
  // stroke-level predecessor sets (unique) and their transpose
  std::vector<int32_t> unclaimed_preds(static_cast<size_t>(n), 0), succ_off(static_cast<size_t>(n) + 1, 0), succ;
  {
    std::vector<std::pair<int32_t, int32_t>> edges;  // (pred, stroke)
    std::vector<int32_t> tmp;
    for (int32_t s = 0; s < n; ++s) {
      tmp.assign(plan.pred_stroke.begin() + plan.seg_off[plan.seg_first[s]], plan.pred_stroke.begin() + plan.seg_off[plan.seg_first[s + 1]]);
      std::sort(tmp.begin(), tmp.end());
      tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
      unclaimed_preds[s] = static_cast<int32_t>(tmp.size());
      for (int32_t g : tmp) {
        edges.emplace_back(g, s);
        ++succ_off[static_cast<size_t>(g) + 1];
      }
    }
    for (int32_t g = 0; g < n; ++g) succ_off[g + 1] += succ_off[g];
    succ.resize(edges.size());
    std::vector<int32_t> fill(succ_off.begin(), succ_off.end() - 1);
    for (const auto& e : edges) succ[fill[e.first]++] = e.second;
  }
  struct Pool {
    std::vector<std::vector<int32_t>> pending;  // per run: strokes in submission order
    std::vector<size_t> head;                   // per run: first possibly unclaimed entry
    std::vector<int32_t> left;                  // per run: strokes not finished yet
    int run = 0;
    std::vector<int32_t> slot_stroke, slot_seg;  // -1 = idle
    std::vector<char> slot_running;
  };
  std::vector<Pool> pools(slots.size());
  for (size_t p = 0; p < pools.size(); ++p) {
    pools[p].pending.resize(slots[p].size());
    pools[p].head.assign(slots[p].size(), 0);
    pools[p].left.assign(slots[p].size(), 0);
  }
  for (int32_t s = 0; s < n; ++s) {
    pools[spec[s].pool].pending[spec[s].run].push_back(s);
    ++pools[spec[s].pool].left[spec[s].run];
  }
  auto open_run = [&](Pool& P) {
    while (P.run < static_cast<int>(P.pending.size()) && P.left[P.run] == 0) ++P.run;
    if (P.run < static_cast<int>(P.pending.size())) {
      const int k = std::max(1, slots[&P - pools.data()][P.run]);
      P.slot_stroke.assign(k, -1);
      P.slot_seg.assign(k, 0);
      P.slot_running.assign(k, 0);
    } else {
      P.slot_stroke.clear();
    }
  };
  for (Pool& P : pools) open_run(P);

  std::vector<int32_t> progress(static_cast<size_t>(n), 0);
  std::vector<char> claimed(static_cast<size_t>(n), 0);
  // wake lists: pools to rescan / (pool, slot) to recheck when a stroke makes progress
  std::vector<std::vector<int32_t>> wake_pool(static_cast<size_t>(n)), wake_slot(static_cast<size_t>(n));
  struct Event {
    double t;
    int32_t pool, slot;
    bool operator>(const Event& o) const { return t != o.t ? t > o.t : (pool != o.pool ? pool > o.pool : slot > o.slot); }
  };
  std::vector<Event> heap;
  auto push = [&](Event e) {
    heap.push_back(e);
    std::push_heap(heap.begin(), heap.end(), std::greater<Event>());
  };
  double now = 0.0;
  // first unsatisfied predecessor of segment k of stroke s, or -1
  auto blocker = [&](int32_t s, int32_t k) {
    const int32_t g = plan.seg_first[s] + k;
    for (int32_t i = plan.seg_off[g]; i < plan.seg_off[g + 1]; ++i)
      if (progress[plan.pred_stroke[i]] < plan.pred_need[i]) return plan.pred_stroke[i];
    return -1;
  };
  auto seg_count = [&](int32_t s) { return plan.seg_first[s + 1] - plan.seg_first[s]; };
  auto start_segment = [&](int32_t p, int32_t i) {  // slot holds a stroke that is not running: run its segment if ready
    Pool& P          = pools[p];
    const int32_t s = P.slot_stroke[i], k = P.slot_seg[i];
    const int32_t b = blocker(s, k);
    if (b >= 0) {
      wake_slot[b].push_back(p);
      wake_slot[b].push_back(i);
      return;
    }
    const int64_t m   = std::max<int64_t>(0, std::min<int64_t>(plan.seg_len[s], count[s] - static_cast<int64_t>(k) * plan.seg_len[s]));
    P.slot_running[i] = 1;
    push(Event{now + static_cast<double>(m) * spec[s].cost, p, i});
  };
  auto scan_pool = [&](int32_t p) {
    Pool& P = pools[p];
    if (P.run >= static_cast<int>(P.pending.size())) return;
    std::vector<int32_t>& pend = P.pending[P.run];
    size_t& head               = P.head[P.run];
    for (size_t i = 0; i < P.slot_stroke.size(); ++i) {
      if (P.slot_stroke[i] >= 0) continue;
      while (head < pend.size() && claimed[pend[head]]) ++head;
      int32_t pick = -1;
      int seen     = 0;
      for (size_t q = head; q < pend.size() && seen < window; ++q) {
        const int32_t s = pend[q];
        if (claimed[s]) continue;
        ++seen;
        if (unclaimed_preds[s] > 0) continue;  // wakes up through the claim of its predecessor (same scan loop)
        const int32_t b = blocker(s, 0);
        if (b < 0) {
          pick = s;
          break;
        }
        if (wake_pool[b].empty() || wake_pool[b].back() != p) wake_pool[b].push_back(p);
      }
      if (pick < 0) return;  // no candidate for this slot => none for the other idle slots either
      claimed[pick] = 1;
      order.push_back(pick);
      for (int32_t j = succ_off[pick]; j < succ_off[pick + 1]; ++j) --unclaimed_preds[succ[j]];
      P.slot_stroke[i] = pick;
      P.slot_seg[i]    = 0;
      start_segment(p, static_cast<int32_t>(i));
    }
  };
  auto scan_all = [&]() {
    // a claim can unblock candidates of other pools (their predecessors are now all claimed): iterate to a fixpoint
    size_t before;
    do {
      before = order.size();
      for (size_t p = 0; p < pools.size(); ++p) scan_pool(static_cast<int32_t>(p));
    } while (order.size() != before);
  };
  scan_all();
  std::vector<int32_t> wp, ws;
  while (!heap.empty()) {
    std::pop_heap(heap.begin(), heap.end(), std::greater<Event>());
    const Event e = heap.back();
    heap.pop_back();
    now       = e.t;
    Pool& P   = pools[e.pool];
    const int32_t s = P.slot_stroke[e.slot];
    const int32_t k = P.slot_seg[e.slot] + 1;
    P.slot_running[e.slot] = 0;
    bool freed = false;
    if (k >= seg_count(s)) {
      progress[s]           = 0x7fffffff;
      P.slot_stroke[e.slot] = -1;
      freed                 = true;
      if (--P.left[P.run] == 0) open_run(P);
    } else {
      progress[s]        = k;
      P.slot_seg[e.slot] = k;
    }
    wp.swap(wake_pool[s]);
    ws.swap(wake_slot[s]);
    wake_pool[s].clear();
    wake_slot[s].clear();
    if (!freed) start_segment(e.pool, e.slot);
    for (size_t i = 0; i + 1 < ws.size(); i += 2) {
      Pool& Q = pools[ws[i]];
      const int32_t sl = ws[i + 1];
      if (sl < static_cast<int32_t>(Q.slot_stroke.size()) && Q.slot_stroke[sl] >= 0 && !Q.slot_running[sl]) start_segment(ws[i], sl);
    }
    const size_t before = order.size();
    std::sort(wp.begin(), wp.end());
    wp.erase(std::unique(wp.begin(), wp.end()), wp.end());
    if (freed && !std::binary_search(wp.begin(), wp.end(), e.pool)) scan_pool(e.pool);
    for (int32_t p : wp) scan_pool(p);
    if (order.size() != before) scan_all();
    wp.clear();
    ws.clear();
  }
  if (makespan) *makespan = now;  // the model's completion time of the batch, in units of ClaimSpec::cost
  // Anything the model left unclaimed (cannot happen for a consistent plan) keeps its submission order.
  if (order.size() != static_cast<size_t>(n)) {
    order.clear();
    for (int32_t s = 0; s < n; ++s) order.push_back(s);
  }
  return order;
}

}  // namespace pb
