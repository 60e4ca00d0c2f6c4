// plane_lambda.cpp — host-side λ(d), λ′(d) and the two 10001-entry fp32 lookup tables the boundary kernel reads.
//
// Replaces boundary_handler/sdf_boundary_handler/plane_numerics.rs:19-172 and lookup_table.rs:11-49 of the
// reference.  λ(d) is the integral of the 2-D cubic spline (support radius 1) over the half plane y >= d.
// The reference evaluates Maxima-generated closed forms; here
//   λ′(d) = -2 ∫_0^sqrt(1-d²) W(sqrt(x²+d²)) dx      closed form of the chord integral (derived below)
//   λ(d)  = ∫_d^1 -λ′(t) dt                            composite Gauss-Legendre on [d,1/2] and [1/2,1]
// Both agree with the reference's Maxima tables to 1e-8 (tests/test_oracle_golden.py).
#include <cmath>
#include <mutex>
#include <vector>

#include "sim.cuh"

namespace {
const double kPi = 3.14159265358979323846;

// ∫ r^k dx with r = sqrt(x² + d²), evaluated between 0 and x
struct ChordMoments { double m0, m1, m2, m3; };
ChordMoments chord_moments(double x, double d) {
  const double dd = d * d, r = std::sqrt(x * x + dd);
  // ln((x + r) / d): difference of the antiderivative's logarithm between x and 0
  const double lg = d > 0.0 ? std::log((x + r) / d) : 0.0;
  ChordMoments m;
  m.m0 = x;
  m.m1 = 0.5 * (x * r + dd * lg);
  m.m2 = x * x * x / 3.0 + dd * x;
  m.m3 = 0.25 * x * r * r * r + 0.375 * dd * x * r + 0.375 * dd * dd * lg;
  return m;
}
// ∫_0^x w(r) dx for the inner polynomial 6r³ - 6r² + 1 and the outer one 2(1-r)³
double inner_poly(const ChordMoments& m) { return 6.0 * m.m3 - 6.0 * m.m2 + m.m0; }
double outer_poly(const ChordMoments& m) { return 2.0 * (m.m0 - 3.0 * m.m1 + 3.0 * m.m2 - m.m3); }

double chord(double d) {  // -λ′(d) for 0 <= d
  if (d >= 1.0) return 0.0;
  if (d < 1e-10) return 1.36418522650196;  // limit d -> 0 (plane_numerics.rs:83-84)
  const double sigma = 40.0 / (7.0 * kPi);  // 10 / (7π h²), h = 1/2
  const double a = std::sqrt((1.0 - d) * (1.0 + d));
  double I;
  if (d < 0.5) {
    const double b = std::sqrt((0.5 - d) * (0.5 + d));
    I = inner_poly(chord_moments(b, d)) + outer_poly(chord_moments(a, d)) - outer_poly(chord_moments(b, d));
  } else {
    I = outer_poly(chord_moments(a, d));
  }
  return 2.0 * sigma * I;
}

struct GaussLegendre {
  std::vector<double> x, w;
  explicit GaussLegendre(int n) : x(n), w(n) {
    for (int i = 0; i < n; i++) {
      double z = std::cos(kPi * (i + 0.75) / (n + 0.5)), pp = 0;
      for (int it = 0; it < 100; it++) {
        double p1 = 1.0, p2 = 0.0;
        for (int j = 0; j < n; j++) { double p3 = p2; p2 = p1; p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0); }
        pp = n * (z * p1 - p2) / (z * z - 1.0);
        double dz = p1 / pp;
        z -= dz;
        if (std::fabs(dz) < 1e-15) break;
      }
      x[i] = z;
      w[i] = 2.0 / ((1.0 - z * z) * pp * pp);
    }
  }
  template <class F> double integrate(F f, double a, double b, int panels) const {
    double s = 0;
    for (int p = 0; p < panels; p++) {
      // panels graded towards a (the integrand has a t² ln t term at 0)
      double t0 = double(p) / panels, t1 = double(p + 1) / panels;
      double lo = a + (b - a) * t0 * t0, hi = a + (b - a) * t1 * t1;
      double c = 0.5 * (lo + hi), hw = 0.5 * (hi - lo);
      for (size_t i = 0; i < x.size(); i++) s += w[i] * hw * f(c + hw * x[i]);
    }
    return s;
  }
};

double lambda_nonneg(double d) {
  static const GaussLegendre gl(24);
  if (d >= 1.0) return 0.0;
  if (d < 1e-9) return 0.5;
  if (d < 0.5) return gl.integrate(chord, d, 0.5, 6) + gl.integrate(chord, 0.5, 1.0, 6);
  return gl.integrate(chord, d, 1.0, 6);
}
}  // namespace

double asph_host_dlambda(double d) { return -chord(std::fabs(d)); }
double asph_host_lambda(double d) { return d >= 0.0 ? lambda_nonneg(d) : 1.0 - lambda_nonneg(-d); }

// LookupTable1D::new over [-1, 1] with 10000 steps (boundary_winchenbach2020.rs:33-37): sample positions are
// computed in fp32, the function in double, the stored value rounded to fp32.
void asph_host_build_luts(std::vector<float>& lam, std::vector<float>& dlam) {
  const int steps = 10000;
  lam.resize(steps + 1);
  dlam.resize(steps + 1);
  for (int i = 0; i <= steps; i++) {
    float x = (float(i) / float(steps)) * (1.f - (-1.f)) + (-1.f);
    lam[i] = float(asph_host_lambda(double(x)));
    dlam[i] = float(asph_host_dlambda(double(x)));
  }
}
// LookupTable1D::get (lookup_table.rs:32-48)
float asph_host_lut_get(const std::vector<float>& data, float x) {
  const float len_inv = 1.f / (1.f - (-1.f));
  float fidx = (x - (-1.f)) * len_inv * float(data.size() - 1);
  float fl = std::floor(fidx);
  float t = fidx - fl;
  size_t idx = size_t(fl);
  if (idx + 1 >= data.size()) return data[idx];
  return data[idx] * (1.f - t) + data[idx + 1] * t;
}
