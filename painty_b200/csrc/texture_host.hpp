// Host-side frame construction of a texture-brush stroke (per-stroke f64 work that stays on the CPU):
// restates painty/renderer/TextureBrush.hxx:52-136 — end extension by the radius, bounding box,
// Catmull-Rom frames, polygon = reverse(right side) ++ left side with uv = reverse((u,1)) ++ ((u,0)).
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

#include "host_math.hpp"

namespace pb {
namespace host {

struct TextureFrame {
  bool valid = false;
  std::vector<V2> poly, uv;  // 2*(n+2) each
  int x0 = 0, x1 = -1, y0 = 0, y1 = -1;
  int local_rows = 0, local_cols = 0;
  V2 bound_min   = {0.0, 0.0};  // boundMin as handed to Smudge::smudge
  double length  = 0.0;         // polyline length of the extended vertices (:86-89)
  std::vector<V2> spine;        // extended vertices (control points of the spine spline)
};

// One step of Smudge::smudge's walk along the spine (Smudge.hxx:48-60, 161-166): everything that does not depend on
// pixel data is evaluated on the host in f64.
struct SmudgeStep {
  double cx, cy;    // spine position
  double c, s;      // cos / sin of the window rotation of this step (dtheta)
  int32_t roi_x, roi_y;
};

// `rotation` is Smudge::_currentRotation (carried from stroke to stroke); size = map side, as in Smudge.hxx.
inline void build_smudge_steps(const TextureFrame& f, int size, double& rotation, std::vector<SmudgeStep>& out) {
  const double pi = 3.141592653589793238462643383279502884197169399375105820974L;
  const SplineEval spine{f.spine.data(), static_cast<int>(f.spine.size())};
  for (double u = 0.0; u <= 1.0; u += 1. / f.length) {
    const V2 center = spine.catmullRom(u);
    V2 t            = spine.catmullRomDerivativeFirst(u);
    const double tn = norm(t);
    t               = {t.x / tn, t.y / tn};
    const double theta = std::atan2(t.y, t.x);  // updateOrientation :161-166
    double dtheta      = theta - rotation;
    while (dtheta <= -0.5 * pi) dtheta += pi;  // normalizeAngle :201-211
    while (dtheta > 0.5 * pi) dtheta -= pi;
    rotation = theta;
    SmudgeStep st;
    st.cx    = center.x;
    st.cy    = center.y;
    st.c     = std::cos(dtheta);
    st.s     = std::sin(dtheta);
    st.roi_x = static_cast<int32_t>(center.x - size / 2.0);
    st.roi_y = static_cast<int32_t>(center.y - size / 2.0);
    out.push_back(st);
  }
}

inline TextureFrame build_texture_frame(const V2* in, int n_in, double radius, int canvas_rows, int canvas_cols) {
  TextureFrame f;
  if (n_in < 2) return f;  // :53-55
  std::vector<V2> v;
  v.reserve(static_cast<size_t>(n_in) + 2);
  {  // :57-66
    const V2 d0     = {in[1].x - in[0].x, in[1].y - in[0].y};
    const double n0 = norm(d0);
    v.push_back({in[0].x - (d0.x / n0) * radius, in[0].y - (d0.y / n0) * radius});
    for (int i = 0; i < n_in; ++i) v.push_back(in[i]);
    const V2 d1     = {in[n_in - 1].x - in[n_in - 2].x, in[n_in - 1].y - in[n_in - 2].y};
    const double n1 = norm(d1);
    v.push_back({in[n_in - 1].x + (d1.x / n1) * radius, in[n_in - 1].y + (d1.y / n1) * radius});
  }
  const int n = static_cast<int>(v.size());
  V2 lo = v[0], hi = v[0];  // :69-84
  for (const V2& q : v) {
    lo.x = std::min(lo.x, q.x);
    lo.y = std::min(lo.y, q.y);
    hi.x = std::max(hi.x, q.x);
    hi.y = std::max(hi.y, q.y);
  }
  lo.x = std::max(lo.x - radius, 0.0);
  hi.x = std::min(hi.x + radius, static_cast<double>(canvas_cols - 1));
  lo.y = std::max(lo.y - radius, 0.0);
  hi.y = std::min(hi.y + radius, static_cast<double>(canvas_rows - 1));

  const SplineEval spine{v.data(), n};
  f.poly.resize(static_cast<size_t>(2 * n));
  f.uv.resize(static_cast<size_t>(2 * n));
  for (int i = 0; i < n; ++i) {  // :107-131
    const double u = static_cast<double>(i) / static_cast<double>(n - 1);
    const V2 c     = spine.catmullRom(u);
    V2 t           = spine.catmullRomDerivativeFirst(u);
    const double tn = norm(t);
    t               = {t.x / tn, t.y / tn};
    const V2 d      = {-t.y, t.x};
    f.poly[static_cast<size_t>(n - 1 - i)] = {c.x + radius * d.x, c.y + radius * d.y};
    f.uv[static_cast<size_t>(n - 1 - i)]   = {u, 1.0};
    f.poly[static_cast<size_t>(n + i)]     = {c.x - radius * d.x, c.y - radius * d.y};
    f.uv[static_cast<size_t>(n + i)]       = {u, 0.0};
  }
  f.x0 = static_cast<int32_t>(lo.x);
  f.x1 = static_cast<int32_t>(hi.x);
  f.y0 = static_cast<int32_t>(lo.y);
  f.y1 = static_cast<int32_t>(hi.y);
  f.local_rows = static_cast<int32_t>(hi.y - lo.y + 1);  // :135-136
  f.local_cols = static_cast<int32_t>(hi.x - lo.x + 1);
  f.bound_min  = lo;
  for (int i = 1; i < n; ++i) f.length += norm({v[i].x - v[i - 1].x, v[i].y - v[i - 1].y});
  f.spine      = v;
  f.valid      = true;
  return f;
}

}  // namespace host
}  // namespace pb
