// Host-side f64 scalar maths that stays on the CPU in the product (per-stroke / per-colour-pick
// work): bit-exact restatements, each citing the reference.
// Compiled with -ffp-contract=off semantics (nvcc host pass: -Xcompiler -ffp-contract=off).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

namespace pb {
namespace host {

struct V2 {
  double x, y;
};

// painty/core/Math.hxx:159-169
inline double coth(double x) {
  if (x > 20.0) return 1.0;
  if (std::fabs(x) > 0.0) {
    const double r = std::cosh(x) / std::sinh(x);
    return std::isnan(r) ? 1.0 : r;
  }
  return std::numeric_limits<double>::infinity();
}

// painty/core/Math.hxx:183-192
inline double acoth(double x) {
  if (std::fabs(x - 1.0) < 100.0 * DBL_EPSILON) return std::numeric_limits<double>::infinity();
  return std::log((x + 1.0) / (x - 1.0)) / 2.0;
}

// painty/core/KubelkaMunk.hxx:28-83
inline void compute_reflectance(const double K[3], const double S_in[3], const double R0[3], double d, double out[3]) {
  const double thr = DBL_EPSILON * 10000.0;
  if (std::fabs(d) < thr) {
    for (int i = 0; i < 3; ++i) out[i] = R0[i];
    return;
  }
  for (int i = 0; i < 3; ++i) {
    const double S  = (std::fabs(S_in[i]) > thr) ? S_in[i] : 0.00000000001;
    const double a  = 1.0 + K[i] / S;
    const double a2 = a * a - 1.0;
    const double b  = (a2 < 0.0) ? 0.0 : std::sqrt(a2);
    const double c  = b * coth(b * S * d);
    out[i]          = (1.0 - R0[i] * (a - c)) / (a - R0[i] + c);
  }
}

// painty/core/KubelkaMunk.hxx:92-124. false = the reference throws std::invalid_argument
inline bool compute_scattering_absorption(const double Rb[3], const double Rw[3], double K[3], double S[3]) {
  for (int i = 0; i < 3; ++i)
    if (!(Rb[i] < Rw[i] && Rb[i] > 0 && Rb[i] < 1.0 && Rw[i] > 0 && Rw[i] < 1.0)) return false;
  for (int i = 0; i < 3; ++i) {
    const double a   = 0.5 * (Rw[i] + (Rb[i] - Rw[i] + 1.) / Rb[i]);
    const double b   = std::sqrt(a * a - 1.);
    const double arg = (b * b - (a - Rw[i]) * (a - 1.)) / (b * (1. - Rw[i]));
    S[i]             = (1. / b) * acoth(arg);
    K[i]             = S[i] * (a - 1.);
  }
  return true;
}

// painty/core/Spline.hxx:28-47
inline double catmull_rom(double pm, double p0, double p1, double p2, double t) {
  const double tau = 0.5, t2 = t * t, t3 = t2 * t;
  const double w0 = -tau * t + 2.0 * tau * t2 - tau * t3;
  const double w1 = 1.0 + (tau - 3.0) * t2 + (2.0 - tau) * t3;
  const double w2 = tau * t + (3.0 - 2.0 * tau) * t2 + (tau - 2.0) * t3;
  const double w3 = -tau * t2 + tau * t3;
  return pm * w0 + p0 * w1 + p1 * w2 + p2 * w3;
}
// painty/core/Spline.hxx:50-73
inline double catmull_rom_d1(double pm, double p0, double p1, double p2, double t) {
  const double tau = 0.5, t2 = t * t;
  const double w0 = tau * (-3.0 * t2 + 4.0 * t - 1.0);
  const double w1 = -t * (-2.0 * tau + 3.0 * (tau - 2.0) * t + 6.0);
  const double w2 = (t - 1.0) * (3.0 * (tau - 2.0) * t - tau);
  const double w3 = tau * t * (3.0 * t - 2.0);
  return pm * w0 + p0 * w1 + p1 * w2 + p2 * w3;
}
inline V2 catmull_rom(V2 a, V2 b, V2 c, V2 d, double t) {
  return {catmull_rom(a.x, b.x, c.x, d.x, t), catmull_rom(a.y, b.y, c.y, d.y, t)};
}
inline V2 catmull_rom_d1(V2 a, V2 b, V2 c, V2 d, double t) {
  return {catmull_rom_d1(a.x, b.x, c.x, d.x, t), catmull_rom_d1(a.y, b.y, c.y, d.y, t)};
}
inline double norm(V2 a) { return std::sqrt(a.x * a.x + a.y * a.y); }

// SplineEval over a point list (painty/core/Spline.hxx:128-219): clamped control points, u in [0,1]
struct SplineEval {
  const V2* p;
  int n;
  const V2& clamped(int i) const { return i < 0 ? p[0] : (i >= n ? p[n - 1] : p[i]); }
  void control(double u, int& index, double& t) const {
    const double x = static_cast<double>(n - 1) * u;
    index          = static_cast<int32_t>(x);
    t              = x - std::floor(x);
  }
  V2 catmullRom(double u) const {
    int i;
    double t;
    control(u, i, t);
    return catmull_rom(clamped(i - 1), clamped(i), clamped(i + 1), clamped(i + 2), t);
  }
  V2 catmullRomDerivativeFirst(double u) const {
    int i;
    double t;
    control(u, i, t);
    return catmull_rom_d1(clamped(i - 1), clamped(i), clamped(i + 1), clamped(i + 2), t);
  }
};

struct Imprint {
  double cx, cy, theta;
};

// Stroke -> imprints. mode 0: FootprintBrush::paintStroke (FootprintBrush.hxx:251-267) with
// p_pre = path[0] on the first segment (documented deviation from the reference's out-of-bounds
// read). mode 1: the GUI's incremental loop (DigitalCanvas.cxx:107-123).
inline void expand_stroke(int mode, const V2* path, int n, std::vector<Imprint>& out) {
  auto emit = [&](V2 a, V2 b, V2 c, V2 d) {
    const double dist = norm({c.x - b.x, c.y - b.y});
    for (int32_t pd = 1; pd <= static_cast<int32_t>(dist); ++pd) {
      const double t = static_cast<double>(pd) / dist;
      const V2 dir   = catmull_rom_d1(a, b, c, d, t);
      const V2 pos   = catmull_rom(a, b, c, d, t);
      out.push_back({pos.x, pos.y, std::atan2(dir.y, dir.x)});
    }
  };
  if (mode == 0) {
    for (int i = 0; i + 1 < n; ++i) emit(path[i > 0 ? i - 1 : 0], path[i], path[i + 1], path[i + 2 < n ? i + 2 : n - 1]);
  } else {
    for (int k = 2; k <= n; ++k)  // k = number of points received so far
      emit(path[k - 3 > 0 ? k - 3 : 0], path[k - 2], path[k - 1], path[k - 1]);
  }
}

}  // namespace host
}  // namespace pb
