// sph_oracle.hpp — CPU restatement of the per-particle step loop of kaegi/adaptive-sph.
//
// TEST INFRASTRUCTURE ONLY.  This file is the parity oracle and the timed CPU baseline; nothing under
// adaptive-sph_b200/ may include, link or call it.  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py use it.
//
// PARITY PIN STATUS: the reference cannot be compiled here (no cargo/rustc), and it holds
// no step-level golden vectors.  What IS pinned against the reference's own known-answer tests
// (tests/test_oracle_golden.py): cubic-spline W integral / ∇W vs finite differences
// (sph_kernels.rs:88,116), volume<->radius round trip (sph_kernels.rs:214), the 23 Maxima values of
// λ/λ′ (plane_numerics.rs:182-195,229-241), λ′ vs finite differences and λ vs numeric half-plane
// integral (plane_numerics.rs:205,251), split-pattern invariants (splitting.rs:102-108).  The step itself
// (neighbour order, reduction order) is "parity unpinned": the reference is not bit-reproducible even
// against itself (R*-tree traversal order, rayon reduce shape).  This restatement fixes
// the order: neighbours ascending j, reductions in fixed chunks of 4096 in index order.
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference/src/simulation/; "sim.rs" = simulation.rs).  FT = float reproduces the default build,
// FT = double the `double-precision` cargo feature (mod.rs:17-27).  Arithmetic is written operation by
// operation in the reference's evaluation order; compile with -ffp-contract=off (Rust never fuses).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>
#include <chrono>

#include "../include/asph.h"

namespace oracle {

template <class FT> struct V2 {
  FT x, y;
  V2() : x(0), y(0) {}
  V2(FT x_, FT y_) : x(x_), y(y_) {}
  V2 operator+(const V2& o) const { return V2(x + o.x, y + o.y); }
  V2 operator-(const V2& o) const { return V2(x - o.x, y - o.y); }
  V2 operator*(FT s) const { return V2(x * s, y * s); }
  V2 operator/(FT s) const { return V2(x / s, y / s); }
  V2& operator+=(const V2& o) { x += o.x; y += o.y; return *this; }
  V2& operator-=(const V2& o) { x -= o.x; y -= o.y; return *this; }
  // nalgebra: norm_squared = x*x + y*y, dot = a.x*b.x + a.y*b.y (no FMA)
  FT norm_squared() const { return x * x + y * y; }
  FT norm() const { return std::sqrt(norm_squared()); }
  FT dot(const V2& o) const { return x * o.x + y * o.y; }
  bool finite() const { return std::isfinite(x) && std::isfinite(y); }
};
template <class FT> inline V2<FT> operator*(FT s, const V2<FT>& v) { return V2<FT>(s * v.x, s * v.y); }

template <class FT> struct Consts;
template <> struct Consts<float> {
  static constexpr float PI = 3.14159265358979323846f;       // std::f32::consts::PI
  static constexpr float FRAC_1_PI = 0.318309886183790671538f;
};
template <> struct Consts<double> {
  static constexpr double PI = 3.14159265358979323846;
  static constexpr double FRAC_1_PI = 0.318309886183790671538;
};

struct StepError {
  int code;
  std::string msg;
};

// ---------------------------------------------------------------------------------------------
// λ(d), λ′(d): half-plane integral of the 2-D cubic spline with support radius 1, in double.
// Restates boundary_handler/sdf_boundary_handler/plane_numerics.rs:19-152 (values from Maxima,
// semi-analytic-boundary-handling-terms.maxima).  λ keeps the reference's closed form with the shared
// sub-expressions named; λ′ is re-derived here as minus the chord integral of W along the line y = d,
// which is what the Maxima expression simplifies to (checked to 1e-8 against the reference's table).
// ---------------------------------------------------------------------------------------------
inline double lambda2_nonneg(double d) {
  const double PI = 3.14159265358979323846;
  if (d < 1e-9) return 0.5;
  if (d >= 1.0) return 0.0;
  const double d3 = d * d * d, d5 = d3 * d * d;
  const double s1 = std::sqrt(1.0 - d) * std::sqrt(d + 1.0);  // sqrt(1-d^2)
  const double L1 = std::log(s1 + 1.0);
  const double ld = std::log(d), l2 = std::log(2.0);
  if (d < 0.5) {
    const double s2 = std::sqrt(1.0 - 2.0 * d) * std::sqrt(2.0 * d + 1.0);  // sqrt(1-4d^2)
    const double L2 = std::log(s2 + 1.0);
    double t = (-48.0 * d5 - 80.0 * d3) * L2 + (12.0 * d5 + 80.0 * d3) * L1 - std::acos(2.0 * d) + 36.0 * ld * d5 +
               48.0 * l2 * d5 + s2 * (68.0 * d3 + 8.0 * d) + 80.0 * l2 * d3 + s1 * (-68.0 * d3 - 32.0 * d) +
               8.0 * std::acos(d);
    return t / (7.0 * PI);
  }
  double t = (-12.0 * d5 - 80.0 * d3) * L1 + ld * (12.0 * d5 + 80.0 * d3) + s1 * (68.0 * d3 + 32.0 * d) -
             8.0 * std::acos(d);
  return -t / (7.0 * PI);
}
inline double lambda2(double d) { return d >= 0.0 ? lambda2_nonneg(d) : 1.0 - lambda2_nonneg(-d); }

// antiderivatives in x of r^k, r = sqrt(x^2 + d^2)
inline double chord_F(double x, double d, bool inner) {
  const double r = std::sqrt(x * x + d * d);
  const double d2 = d * d;
  const double lg = (x + r > 0.0) ? std::log(x + r) : 0.0;
  const double I1 = 0.5 * (x * r + d2 * lg);                                     // ∫ r dx
  const double I2 = x * x * x / 3.0 + d2 * x;                                    // ∫ r^2 dx
  const double I3 = 0.25 * x * r * r * r + 0.375 * d2 * x * r + 0.375 * d2 * d2 * lg;  // ∫ r^3 dx
  if (inner) return 6.0 * I3 - 6.0 * I2 + x;                                     // w = 6(q^3-q^2)+1
  return 2.0 * (x - 3.0 * I1 + 3.0 * I2 - I3);                                   // w = 2(1-q)^3
}
inline double dlambda2_nonneg(double d) {
  const double PI = 3.14159265358979323846;
  if (d >= 1.0) return 0.0;
  if (d < 1e-10) return -1.36418522650196;  // plane_numerics.rs:83-84
  const double c = 40.0 / (7.0 * PI);       // 10/(7π h²) with h = 1/2
  const double a = std::sqrt((1.0 - d) * (1.0 + d));
  double I;
  if (d < 0.5) {
    const double b = std::sqrt((0.5 - d) * (0.5 + d));
    I = (chord_F(b, d, true) - chord_F(0.0, d, true)) + (chord_F(a, d, false) - chord_F(b, d, false));
  } else {
    I = chord_F(a, d, false) - chord_F(0.0, d, false);
  }
  return -2.0 * c * I;
}
inline double dlambda2(double d) { return dlambda2_nonneg(d >= 0.0 ? d : -d); }

// ---------------------------------------------------------------------------------------------
// sph_kernels.rs
// ---------------------------------------------------------------------------------------------
template <class FT> inline FT cubic_unnorm(FT q) {  // sph_kernels.rs:23-32
  if (q < FT(0.5)) return FT(6) * (q * q * q - q * q) + FT(1);
  if (q < FT(1)) {
    FT v = FT(1) - q;
    return FT(2) * (v * v * v);
  }
  return FT(0);
}
template <class FT> inline FT cubic_unnorm_deriv(FT q) {  // sph_kernels.rs:34-43
  if (q < FT(0.5)) return FT(18) * q * q - FT(12) * q;
  if (q < FT(1)) {
    FT v = FT(1) - q;
    return FT(-6) * v * v;
  }
  return FT(0);
}
template <class FT> inline FT kernel_w(FT r, FT h) {  // cubic_kernel_2d sph_kernels.rs:49-52
  FT norm_factor = FT(10) / (FT(7) * Consts<FT>::PI * (h * h));
  return norm_factor * cubic_unnorm<FT>(r / (FT(2) * h));
}
template <class FT> inline FT kernelh(V2<FT> diff, FT h) { return kernel_w<FT>(diff.norm(), h); }  // :190-192
template <class FT> inline V2<FT> kernel_derivh(V2<FT> diff, FT h) {  // cubic_kernel_2d_deriv :61-71
  FT r = diff.norm();
  FT q = r / (FT(2) * h);
  if (q <= FT(1.0e-5)) return V2<FT>();
  diff = diff / r;
  FT norm_factor = FT(10) / (FT(7) * Consts<FT>::PI * (h * h));
  return (norm_factor * cubic_unnorm_deriv<FT>(q) / (FT(2) * h)) * diff;
}
template <class FT> inline FT volume_to_radius(FT area) { return std::sqrt(area * Consts<FT>::FRAC_1_PI); }  // :203-206
template <class FT> inline FT radius_to_volume(FT r) { return Consts<FT>::PI * r * r; }                      // :209-211
constexpr double ETA = 1.9;  // sim.rs:369
template <class FT> inline FT h_from_mass(FT mass, FT rest_density) {  // sim.rs:372-380
  FT volume = mass / rest_density;
  return FT(ETA) * volume_to_radius<FT>(volume);
}

// ---------------------------------------------------------------------------------------------
// LookupTable1D (boundary_handler/sdf_boundary_handler/lookup_table.rs:11-49)
// ---------------------------------------------------------------------------------------------
template <class FT> struct Lut {
  FT min, max, len_inv;
  size_t steps;
  std::vector<FT> data;
  template <class F> void init(FT mn, FT mx, size_t st, F f) {
    min = mn; max = mx; steps = st; len_inv = FT(1) / (mx - mn);
    data.resize(st + 1);
    for (size_t i = 0; i <= st; i++) {
      FT x = (FT(i) / FT(st)) * (mx - mn) + mn;
      data[i] = FT(f(double(x)));
    }
  }
  FT get(FT x) const {
    FT fidx = (x - min) * len_inv * FT(steps);
    FT fl = std::floor(fidx);
    FT interp = fidx - fl;
    size_t idx = size_t(fl);
    if (idx + 1 >= data.size()) return data[idx];
    return data[idx] * (FT(1) - interp) + data[idx + 1] * interp;
  }
};

// ---------------------------------------------------------------------------------------------
// parameters rounded once to FT (serde parses YAML literals straight into FT)
// ---------------------------------------------------------------------------------------------
template <class FT> struct Params {
  asph_params raw;
  FT rest_density, cfl_factor, max_dt, viscosity, gravity, maximum_range, jacobi_omega, sdf_gradient_eps;
  FT particle_radius_fine, particle_radius_base, maximum_surface_distance;
  FT max_mass_transfer_sharing, max_share_distance, max_merge_distance;
  FT iisph_max_avg_density_error, hybrid_dfsph_factor, hybrid_dfsph_max_avg_density_error,
      hybrid_dfsph_max_avg_divergence_error, level_estimation_range;
  V2<FT> pull;
  explicit Params(const asph_params& p) : raw(p) {
    rest_density = FT(p.rest_density); cfl_factor = FT(p.cfl_factor); max_dt = FT(p.max_dt);
    viscosity = FT(p.viscosity); gravity = FT(p.gravity); maximum_range = FT(p.maximum_range);
    jacobi_omega = FT(p.jacobi_omega); sdf_gradient_eps = FT(p.sdf_gradient_eps);
    particle_radius_fine = FT(p.particle_radius_fine); particle_radius_base = FT(p.particle_radius_base);
    maximum_surface_distance = FT(p.maximum_surface_distance);
    max_mass_transfer_sharing = FT(p.max_mass_transfer_sharing);
    max_share_distance = FT(p.max_share_distance); max_merge_distance = FT(p.max_merge_distance);
    iisph_max_avg_density_error = FT(p.iisph_max_avg_density_error);
    hybrid_dfsph_factor = FT(p.hybrid_dfsph_factor);
    hybrid_dfsph_max_avg_density_error = FT(p.hybrid_dfsph_max_avg_density_error);
    hybrid_dfsph_max_avg_divergence_error = FT(p.hybrid_dfsph_max_avg_divergence_error);
    level_estimation_range = FT(p.level_estimation_range);
    pull = V2<FT>(FT(p.pull_fluid_to[0]), FT(p.pull_fluid_to[1]));
  }
  FT mass_fine() const { return radius_to_volume<FT>(particle_radius_fine) * rest_density; }  // simulation_parameters.rs:125
  FT mass_base() const { return radius_to_volume<FT>(particle_radius_base) * rest_density; }  // :129
};

// LevelEstimationState (sim.rs:197-238): interior flag + value
template <class FT> struct Level {
  bool surface = false;  // FluidSurface(v) vs FluidInterior
  FT v = 0;
};

template <class FT> FT target_mass(const Level<FT>& l, const Params<FT>& P) {  // sim.rs:213-237
  FT level = std::max(l.v, -P.maximum_surface_distance);
  FT interpolation = level / -P.maximum_surface_distance;
  switch (P.raw.sizing_function) {
    case ASPH_SIZING_MASS: return P.mass_fine() * (FT(1) - interpolation) + P.mass_base() * interpolation;
    case ASPH_SIZING_RADIUS: {
      FT tr = P.particle_radius_fine * (FT(1) - interpolation) + P.particle_radius_base * interpolation;
      return radius_to_volume<FT>(tr) * P.rest_density;
    }
    default: {
      FT e = FT(1) / FT(2);
      FT tr = P.particle_radius_fine * (FT(1) - std::pow(interpolation, e)) +
              P.particle_radius_base * std::pow(interpolation, e);
      return radius_to_volume<FT>(tr) * P.rest_density;
    }
  }
}

template <class FT> uint8_t classify(const Level<FT>& l, FT mass, const Params<FT>& P) {  // adaptivity/mod.rs:32-48
  FT mrel = mass / target_mass<FT>(l, P);
  if (mrel <= FT(0.5)) return ASPH_CLASS_TOO_SMALL;
  if (mrel <= FT(1) / FT(1.1)) return ASPH_CLASS_SMALL;
  if (mrel < FT(1.1)) return ASPH_CLASS_OPTIMAL;
  if (mrel < FT(2.0)) return ASPH_CLASS_LARGE;
  return ASPH_CLASS_TOO_LARGE;
}

struct PcTimer {
  double ms[ASPH_PC_COUNT] = {0};
  uint64_t calls[ASPH_PC_COUNT] = {0};
  std::chrono::steady_clock::time_point t0[ASPH_PC_COUNT];
  bool enabled = false;
  void begin(int id) { if (enabled) t0[id] = std::chrono::steady_clock::now(); }
  void end(int id) {  // Counter::end sim.rs:119
    if (!enabled) return;
    ms[id] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0[id]).count();
    calls[id]++;
  }
  void end_add_to_last(int id) {  // sim.rs:123
    if (!enabled) return;
    ms[id] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0[id]).count();
  }
};

// ---------------------------------------------------------------------------------------------
// FluidSimulation (sim.rs:471-484) with ParticleVec SoA (sim.rs:284-334)
// ---------------------------------------------------------------------------------------------
template <class FT> struct Sim {
  typedef V2<FT> V;
  // persistent
  std::vector<FT> mass;
  std::vector<V> position, velocity, velocity_temp, pressure_accel;
  std::vector<FT> density, ppe_source_term, pressure, pressure_next_iter, aii, density_error, h2, constant_field;
  std::vector<FT> h2_next, omega;  // FromDistribution* / constrain_neighborhood_count state (sim.rs:309-311); IISPH2 (sim.rs:307)
  std::vector<uint8_t> flag_neighborhood_reduced;
  std::vector<Level<FT>> level_estimation, level_estimation_temp;
  std::vector<uint8_t> size_class, flag_is_fluid_surface, flag_insufficient_neighs;
  std::vector<uint32_t> merge_partner;
  std::vector<uint16_t> merge_counter;
  std::vector<std::vector<uint32_t>> neighs;  // NeighborhoodCache neighborhood_search.rs:13
  // boundary handler BoundaryWinchenbach2020 (boundary_winchenbach2020.rs:21-31)
  asph_boundary boundary;
  Lut<FT> lambda_lut, dlambda_lut;
  struct LamEntry { FT lambda; V grad; };
  std::vector<std::vector<LamEntry>> lambda;
  // split patterns
  int max_children = 0;
  std::vector<int32_t> split_offset;
  std::vector<float> split_pos;

  double time = 0;  // FT in the reference; kept in FT arithmetic below
  FT time_ft = 0;
  uint64_t step_number = 0;
  asph_step_info info;
  PcTimer pc;
  std::string last_error;
  int num_threads_hint = 0;

  size_t n() const { return position.size(); }

  void resize_all(size_t N) {
    mass.resize(N); position.resize(N); velocity.resize(N); velocity_temp.resize(N); pressure_accel.resize(N);
    density.resize(N); ppe_source_term.resize(N); pressure.resize(N); pressure_next_iter.resize(N); aii.resize(N);
    density_error.resize(N); h2.resize(N); constant_field.resize(N); h2_next.resize(N); omega.resize(N);
    flag_neighborhood_reduced.resize(N);
    level_estimation.resize(N); level_estimation_temp.resize(N);
    size_class.resize(N, ASPH_CLASS_OPTIMAL); flag_is_fluid_surface.resize(N); flag_insufficient_neighs.resize(N);
    merge_partner.resize(N); merge_counter.resize(N);
    neighs.resize(N); lambda.resize(N);
  }

  void init_luts() {  // BoundaryWinchenbach2020::new boundary_winchenbach2020.rs:33-45
    lambda_lut.init(FT(-1), FT(1), 10000, [](double x) { return lambda2(x); });
    dlambda_lut.init(FT(-1), FT(1), 10000, [](double x) { return dlambda2(x); });
  }

  // ---------------------------------------------------------------- sdf (sdf/sdf.rs, sdf_plane.rs)
  FT probe_plane(int s, V x) const {  // SdfPlane::probe sdf_plane.rs:36-38: dir.dot(x) + delta
    FT nx = FT(boundary.planes[s][0]), ny = FT(boundary.planes[s][1]), dl = FT(boundary.planes[s][2]);
    return (nx * x.x + ny * x.y) + dl;
  }
  // Sdf2D polygon (sdf/sdf2d.rs:36-143, one connected component): positive on the left-hand (air) side
  std::vector<V> poly_pt, poly_dir, poly_pn;
  void init_polygon() {  // Sdf2DConnectedComponents::from_points sdf2d.rs:37-71
    const int np = boundary.n_poly;
    poly_pt.clear(); poly_dir.clear(); poly_pn.clear();
    for (int i = 0; i < np; i++) poly_pt.push_back(V(FT(boundary.poly[i][0]), FT(boundary.poly[i][1])));
    for (int i = 0; i < np; i++) {
      V d = poly_pt[(i + 1) % np] - poly_pt[i];
      poly_dir.push_back(d / d.norm());
    }
    for (int i = 0; i < np; i++) {
      V a = poly_dir[i == 0 ? np - 1 : i - 1], b = poly_dir[i];
      poly_pn.push_back(V(-a.y, a.x) + V(-b.y, b.x));
    }
  }
  FT probe_polygon(V x) const {  // find_min_dist_object + to_dist_and_dir sdf2d.rs:73-141
    const int np = int(poly_pt.size());
    FT min_dist_sq = std::numeric_limits<FT>::infinity();
    bool is_line = false; FT line_dist = 0; int pidx = 0; V pdir; FT pdist_sq = 0;
    for (int s = 0; s < np; s++) {
      V ls = poly_pt[s], le = poly_pt[(s + 1) % np];
      FT len_sq = (le - ls).norm_squared();
      V ld = poly_dir[s];
      V pd = x - ls;
      V left(-ld.y, ld.x);
      FT proj = pd.dot(ld);
      if (proj > FT(0) && proj * proj < len_sq) {
        FT dl = pd.dot(left);
        if (dl * dl < min_dist_sq) { is_line = true; line_dist = dl; min_dist_sq = dl * dl; }
      }
      FT c = pd.norm_squared();
      if (c < min_dist_sq) { is_line = false; pidx = s; pdir = pd; pdist_sq = c; min_dist_sq = c; }
    }
    if (is_line) return line_dist;
    FT sign = poly_pn[pidx].dot(pdir) >= FT(0) ? FT(1) : FT(-1);
    return std::sqrt(pdist_sq) * sign;
  }
  int n_sdf() const {
    if (boundary.kind == ASPH_BND_PLANES) return boundary.n_planes;
    if (boundary.kind == ASPH_BND_POLYGON) return 1;
    return 0;
  }
  FT probe(int s, V x) const { return boundary.kind == ASPH_BND_PLANES ? probe_plane(s, x) : probe_polygon(x); }
  V finite_diff_gradient(int s, V x, FT eps) const {  // sdf.rs:50-62
    FT inv_2eps = FT(1) / (FT(2) * eps);
    V xp = x, xn = x;
    xp.x += eps; xn.x -= eps;
    FT gx = (probe(s, xp) - probe(s, xn)) * inv_2eps;
    xp = x; xn = x;
    xp.y += eps; xn.y -= eps;
    FT gy = (probe(s, xp) - probe(s, xn)) * inv_2eps;
    return V(gx, gy);
  }

  // ---------------------------------------------------------------- smoothing length helpers (adaptive build)
  FT hij(size_t i, size_t j) const { return (h2[i] + h2[j]) * FT(0.5); }  // sph_kernels.rs:273-278

  // ---------------------------------------------------------------- neighbour search
  // Final set of build_neighborhood_list_rstar (neighborhood_search.rs:73-185) after the symmetrize pass:
  // N_f(i) = { j : |x_ij|^2 < ((h_i+h_j)*0.5*f)^2 }, self included; rows ascending j (canonical order).
  // A uniform grid with cell = f*h_max replaces the R*-tree (stated deviation; the set is identical).
  bool build_neighbors(FT f, StepError& err) {
    const size_t N = n();
    if (N == 0) return true;
    FT hmax = 0;
    FT minx = position[0].x, maxx = minx, miny = position[0].y, maxy = miny;
    for (size_t i = 0; i < N; i++) {
      hmax = std::max(hmax, h2[i]);
      minx = std::min(minx, position[i].x); maxx = std::max(maxx, position[i].x);
      miny = std::min(miny, position[i].y); maxy = std::max(maxy, position[i].y);
    }
    double cell = double(hmax) * double(f) * 1.001;
    if (!(cell > 0)) cell = 1;
    // limit the number of cells
    double ex = double(maxx) - double(minx), ey = double(maxy) - double(miny);
    while ((ex / cell + 3) * (ey / cell + 3) > 4.0 * double(N) + 1024.0) cell *= 1.5;
    const int64_t nx = int64_t(ex / cell) + 3, ny = int64_t(ey / cell) + 3;
    auto cellof = [&](V p, int64_t& cx, int64_t& cy) {
      cx = int64_t(std::floor((double(p.x) - double(minx)) / cell)) + 1;
      cy = int64_t(std::floor((double(p.y) - double(miny)) / cell)) + 1;
    };
    std::vector<uint32_t> start(size_t(nx * ny) + 1, 0), order(N);
    std::vector<uint32_t> cidx(N);
    for (size_t i = 0; i < N; i++) {
      int64_t cx, cy; cellof(position[i], cx, cy);
      cidx[i] = uint32_t(cy * nx + cx);
      start[cidx[i] + 1]++;
    }
    for (size_t c = 0; c < size_t(nx * ny); c++) start[c + 1] += start[c];
    {
      std::vector<uint32_t> cur(start.begin(), start.end() - 1);
      for (size_t i = 0; i < N; i++) order[cur[cidx[i]]++] = uint32_t(i);  // ascending i within a cell
    }
    bool overflow = false;
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t ii = 0; ii < int64_t(N); ii++) {
      const size_t i = size_t(ii);
      std::vector<uint32_t>& out = neighs[i];
      out.clear();
      const V xi = position[i];
      // search radius: pairs with larger partners can reach up to f*(h_i+h_max)/2 <= cell
      int64_t cx, cy; cellof(xi, cx, cy);
      for (int64_t dy = -1; dy <= 1; dy++)
        for (int64_t dx = -1; dx <= 1; dx++) {
          int64_t c = (cy + dy) * nx + (cx + dx);
          for (uint32_t k = start[c]; k < start[c + 1]; k++) {
            uint32_t j = order[k];
            FT x_ij_sq = (xi - position[j]).norm_squared();
            FT s_ij = hij(i, j) * f;  // smoothing_length * support_length_by_smoothing_length :144
            if (x_ij_sq < s_ij * s_ij) out.push_back(j);
          }
        }
      std::sort(out.begin(), out.end());
      if (out.size() > 20000) overflow = true;  // MAX_NEIGHBOR_COUNT neighborhood_search.rs:3,148-150
    }
    if (overflow) { err = {ASPH_ERR_NEIGHBOR_OVERFLOW, "exceeded maximum allowed number of 20000 neighbors"}; return false; }
    return true;
  }
  // O(N^2) statement of the same set: check_correct_neighborhood sim.rs:1810-1863 /
  // neighborhood_search.rs:216-237.  Used by tests to pin build_neighbors.
  void build_neighbors_bruteforce(FT f) {
    const size_t N = n();
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t ii = 0; ii < int64_t(N); ii++) {
      size_t i = size_t(ii);
      neighs[i].clear();
      for (size_t j = 0; j < N; j++) {
        FT x_ij_sq = (position[i] - position[j]).norm_squared();
        FT s_ij = hij(i, j) * f;
        if (x_ij_sq < s_ij * s_ij) neighs[i].push_back(uint32_t(j));
      }
    }
  }
  void filter_down(FT f) {  // NeighborhoodCache::filter_down neighborhood_search.rs:56-70
    const size_t N = n();
#pragma omp parallel for schedule(static)
    for (int64_t ii = 0; ii < int64_t(N); ii++) {
      size_t i = size_t(ii);
      auto& l = neighs[i];
      size_t w = 0;
      for (size_t k = 0; k < l.size(); k++) {
        uint32_t j = l[k];
        FT x_ij_sq = (position[i] - position[j]).norm_squared();
        FT s_ij = hij(i, j) * f;
        if (x_ij_sq < s_ij * s_ij) l[w++] = j;
      }
      l.resize(w);
    }
  }

  // ---------------------------------------------------------------- boundary handler
  void boundary_update_after_advect(const Params<FT>& P) {  // boundary_winchenbach2020.rs:58-152
    const size_t N = n();
    const int ns = n_sdf();
#pragma omp parallel for schedule(static)
    for (int64_t ii = 0; ii < int64_t(N); ii++) {
      size_t i = size_t(ii);
      auto& pl = lambda[i];
      pl.clear();
      if (ns == 0) continue;
      V xi = position[i];
      FT sr_i = h2[i] * FT(2);  // support_radius_single sph_kernels.rs:291-297
      for (int s = 0; s < ns; s++) {
        FT d = probe(s, xi) / sr_i;
        if (d < FT(1)) {
          V g = finite_diff_gradient(s, xi, P.sdf_gradient_eps);
          FT gn = g.norm();
          if (gn >= FT(0.00001)) {
            g = g / gn;
            FT penalty, pder;
            switch (P.raw.boundary_penalty_term) {  // :83-124
              case ASPH_PENALTY_NONE: penalty = 1; pder = 0; break;
              case ASPH_PENALTY_LINEAR: penalty = FT(1) - d; pder = -1; break;
              case ASPH_PENALTY_QUADRATIC1:
                if (d > FT(0)) { penalty = 1; pder = 0; }
                else if (d > FT(-1)) { penalty = FT(0.5) * d * d + FT(1); pder = d; }
                else { penalty = FT(0.5) - d; pder = -1; }
                break;
              default:
                if (d > FT(0)) { penalty = 1; pder = 0; }
                else if (d > FT(-0.5)) { penalty = d * d + FT(1); pder = FT(2) * d; }
                else { penalty = FT(0.75) - d; pder = -1; }
                break;
            }
            FT lam, lamd;
            if (d <= FT(-1)) { lam = 1; lamd = 0; }
            else { lam = lambda_lut.get(d); lamd = dlambda_lut.get(d); }
            FT lambda_penalty = lam * penalty;
            V grad = (g / sr_i) * (pder * lam + penalty * lamd);
            pl.push_back({lambda_penalty, grad});
          }
        }
      }
    }
  }
  FT lambda_sum(size_t i) const {  // density_boundary_term :154-162 (iterator .sum() from 0)
    FT s = 0;
    for (auto& e : lambda[i]) s += e.lambda;
    return s;
  }
  FT distance_to_boundary(size_t i) const {  // :319-324
    FT m = std::numeric_limits<FT>::infinity();
    for (int s = 0; s < n_sdf(); s++) m = std::min(m, probe(s, position[i]));
    return m;
  }
  V boundary_pressure_accel(size_t i, const std::vector<FT>& p, const Params<FT>& P) const {  // :164-193
    V result;
    for (auto& e : lambda[i]) {
      FT p_i = p[i];
      FT p_ib = (P.raw.operator_discretization == ASPH_OP_CONSISTENT_SYMMETRIC_GRADIENT) ? p_i : FT(0);
      FT rho_i = density[i];
      FT rho_b = P.rest_density;
      result += (-rho_b * (p_i / (rho_i * rho_i) + p_ib / (rho_b * rho_b))) * e.grad;
    }
    return result;
  }
  template <class QF> FT boundary_divergence(size_t i, QF qf, V qb, const Params<FT>& P) const {  // :195-223
    FT result = 0;
    for (auto& e : lambda[i]) {
      FT rho_i = density[i];
      FT rho_b = P.rest_density;
      if (P.raw.operator_discretization == ASPH_OP_WINCHENBACH2020)
        result += (qb - qf(i)).dot(e.grad);
      else
        result += rho_b / rho_i * (qb - qf(i)).dot(e.grad);
    }
    return result;
  }
  FT boundary_aii(size_t i, const Params<FT>& P) const {  // iisph_aii :225-306
    FT mi = mass[i], rho_i = density[i], rho_0 = P.rest_density, rho_i_sq = rho_i * rho_i;
    FT rho_b = rho_0;
    if (P.raw.operator_discretization == ASPH_OP_WINCHENBACH2020) {
      V mj_wij, mj_by_rhoj_wij; FT mj_by_rhoj_wij_sq = 0;
      for (uint32_t j : neighs[i]) {
        V gw = kernel_derivh<FT>(position[i] - position[j], hij(i, j));
        mj_wij += mass[j] * gw;
        mj_by_rhoj_wij += (mass[j] / density[j]) * gw;
        mj_by_rhoj_wij_sq += mass[j] / density[j] * gw.norm_squared();
      }
      V sum_glambda, sum_boundary;
      for (auto& e : lambda[i]) {
        FT p_ib_coeff = 0;
        sum_glambda += e.grad;
        sum_boundary += (rho_b * (FT(1) / (rho_i * rho_i) + p_ib_coeff / (rho_b * rho_b))) * e.grad;
      }
      return (mj_wij / rho_i_sq + sum_boundary).dot(mj_by_rhoj_wij + sum_glambda) + (mi * mj_by_rhoj_wij_sq / rho_i_sq);
    }
    FT rho_i_cu = rho_i * rho_i * rho_i;
    V mj_wij; FT mj_wij_sq = 0;
    for (uint32_t j : neighs[i]) {
      V gw = kernel_derivh<FT>(position[i] - position[j], hij(i, j));
      mj_wij += mass[j] * gw;
      mj_wij_sq += mass[j] * gw.norm_squared();
    }
    V rhob_glambda, sum_boundary;
    for (auto& e : lambda[i]) {
      FT p_ib_coeff = (P.raw.operator_discretization == ASPH_OP_CONSISTENT_SIMPLE_GRADIENT) ? FT(0) : FT(1);
      rhob_glambda += rho_b * e.grad;
      sum_boundary += (rho_b * (FT(1) / (rho_i * rho_i) + p_ib_coeff / (rho_b * rho_b))) * e.grad;
    }
    return (mj_wij / rho_i_sq + sum_boundary).dot(mj_wij / rho_i + rhob_glambda / rho_i) + (mi * mj_wij_sq) / rho_i_cu;
  }

  // ---------------------------------------------------------------- level estimation
  void surface_detection_by_empty_angle(const Params<FT>& P) {  // sim.rs:539-625
    const size_t N = n();
    const FT threshold = std::cos(FT(50) * (Consts<FT>::PI / FT(180)));
#pragma omp parallel for schedule(static)
    for (int64_t ii = 0; ii < int64_t(N); ii++) {
      size_t i = size_t(ii);
      V normal;
      for (uint32_t j : neighs[i]) {
        V x_ij = position[i] - position[j];
        V dg = kernel_derivh<FT>(x_ij, hij(i, j));
        normal -= (mass[i] / P.rest_density) * dg;
      }
      bool interior;
      flag_insufficient_neighs[i] = 0;
      if (neighs[i].size() < 3) {  // D*2-1
        interior = false;
        flag_insufficient_neighs[i] = 1;
      } else if (normal.norm_squared() < FT(0.00001)) {
        interior = true;
      } else if (!P.raw.boundary_is_fluid_surface && distance_to_boundary(i) < h2[i] * FT(1.5)) {
        interior = true;
      } else {
        interior = false;
        normal = normal / normal.norm();  // normalize_mut
        // is_neighbor_in_level_estimation_range (sim.rs:698-723): only FromDistribution / FromDistribution2 cut the range
        const bool cut = P.raw.support_length_estimation == ASPH_H_FROM_DISTRIBUTION ||
                         P.raw.support_length_estimation == ASPH_H_FROM_DISTRIBUTION2;
        const FT particle_radius = volume_to_radius<FT>(mass[i] / P.rest_density);
        const FT cut_r = particle_radius * FT(P.raw.maximum_range);
        for (uint32_t j : neighs[i]) {
          V xji = position[j] - position[i];
          if (cut && xji.norm_squared() > cut_r * cut_r) continue;
          xji = xji / (xji.norm() + FT(0.000001));
          if (xji.dot(normal) > threshold) { interior = true; break; }
        }
      }
      level_estimation[i].surface = !interior;
      level_estimation[i].v = 0;
      flag_is_fluid_surface[i] = interior ? 0 : 1;
    }
  }
  void surface_detection_by_center_diff(const Params<FT>& P) {  // sim.rs:631-695
    const size_t N = n();
#pragma omp parallel for schedule(static)
    for (int64_t ii = 0; ii < int64_t(N); ii++) {
      size_t i = size_t(ii);
      FT weight_sum = 0, avg_radius = 0;
      V avg_center;
      int num_neighbors = 0;
      for (uint32_t j : neighs[i]) {
        FT vol = mass[j] / P.rest_density;
        FT rad = volume_to_radius<FT>(vol);
        FT weight = kernelh<FT>(position[i] - position[j], hij(i, j)) * vol;
        avg_center += position[j] * weight;
        avg_radius += rad * weight;
        weight_sum += weight;
        num_neighbors++;
      }
      avg_radius /= weight_sum;
      FT surface_level = FT(-0.85) * avg_radius;
      FT phi;
      if (num_neighbors < 5) phi = surface_level;
      else { avg_center = avg_center / weight_sum; phi = (position[i] - avg_center).norm() - avg_radius; }
      if (phi >= surface_level) { level_estimation[i].surface = true; level_estimation[i].v = phi; flag_is_fluid_surface[i] = 1; }
      else { level_estimation[i].surface = false; level_estimation[i].v = 0; flag_is_fluid_surface[i] = 0; }
    }
  }
  int propagate_level_set(const Params<FT>&) {  // sim.rs:729-801
    const size_t N = n();
    level_estimation_temp = level_estimation;
    int num_iter = 0;
    bool changed = true;
    while (changed) {
      changed = false;
#pragma omp parallel for schedule(static) reduction(|| : changed)
      for (int64_t ii = 0; ii < int64_t(N); ii++) {
        size_t i = size_t(ii);
        if (level_estimation[i].surface) { level_estimation_temp[i] = level_estimation[i]; continue; }
        bool have = false; FT best = 0;
        for (uint32_t j : neighs[i]) {
          if (level_estimation[j].surface) {
            FT est = level_estimation[j].v - (position[j] - position[i]).norm();
            if (have) best = std::max(best, est); else { best = est; have = true; }
          }
        }
        if (have) { level_estimation_temp[i].surface = true; level_estimation_temp[i].v = best; changed = true; }
        else { level_estimation_temp[i].surface = false; level_estimation_temp[i].v = 0; }
      }
      std::swap(level_estimation, level_estimation_temp);
      num_iter++;
    }
    return num_iter;
  }
  bool smooth_level_field(const Params<FT>& P, StepError& err) {  // sim.rs:804-857
    if (P.raw.level_estimation_method == ASPH_LEVEL_NONE) return true;
    const size_t N = n();
    bool bad = false;
#pragma omp parallel for schedule(static)
    for (int64_t ii = 0; ii < int64_t(N); ii++) {
      size_t i = size_t(ii);
      FT level = 0, weight = 0;
      for (uint32_t j : neighs[i]) {
        V x_ij = position[i] - position[j];
        FT w_ij = kernelh<FT>(x_ij, hij(i, j));
        FT dist = level_estimation[j].surface ? std::max(level_estimation[j].v, -P.maximum_surface_distance)
                                              : -P.maximum_surface_distance;
        level += dist * mass[j] / density[j] * w_ij;
        weight += mass[j] / density[j] * w_ij;
      }
      if (!std::isfinite(weight) || weight <= FT(0)) { bad = true; continue; }
      level /= weight;
      level_estimation_temp[i].surface = true;
      level_estimation_temp[i].v = level;
    }
    if (bad) { err = {ASPH_ERR_NONFINITE, "smooth_level_estimation_field: weight <= 0"}; return false; }
    std::swap(level_estimation, level_estimation_temp);
    return true;
  }
  bool perform_level_estimation(const Params<FT>& P, StepError& err) {  // sim.rs:863-927
    switch (P.raw.level_estimation_method) {
      case ASPH_LEVEL_NONE: return true;
      case ASPH_LEVEL_EMPTY_ANGLE: surface_detection_by_empty_angle(P); break;
      default: surface_detection_by_center_diff(P); break;
    }
    (void)err;
    info.level_sweeps = propagate_level_set(P);
    return true;
  }

  // ---------------------------------------------------------------- per-particle physics
  V non_pressure_accel(size_t i, const std::vector<V>& vel, const Params<FT>& P) const {  // sim.rs:931-1005
    const FT speed_of_sound = 88;
    V acc;
    if (P.raw.viscosity_type == ASPH_VISC_WCSPH) {
      for (uint32_t j : neighs[i]) {
        V x_ab = position[i] - position[j];
        V v_ab = vel[i] - vel[j];
        FT h_ij = hij(i, j);
        V dg = kernel_derivh<FT>(x_ab, h_ij);
        FT est = v_ab.dot(x_ab);
        if (est < FT(0)) {
          FT viscous_term = FT(2) * P.viscosity * h_ij * speed_of_sound / (density[i] + density[j]);
          FT pi_ab = -viscous_term * est / (x_ab.norm_squared() + FT(0.001) * h_ij * h_ij);
          acc += (-mass[j] * pi_ab) * dg;
        }
      }
    } else if (P.raw.viscosity_type == ASPH_VISC_APPROX_LAPLACE) {
      for (uint32_t j : neighs[i]) {
        V x_ab = position[i] - position[j];
        V v_ab = vel[i] - vel[j];
        if (x_ab.dot(v_ab) >= FT(0)) continue;
        FT h_ij = hij(i, j);
        V dg = kernel_derivh<FT>(x_ab, h_ij);
        FT rho_ij = (density[i] + density[j]) * FT(0.5);
        FT coeff = FT(2) * FT(4) * (mass[j] / rho_ij) * x_ab.dot(v_ab) / (x_ab.norm_squared() + FT(0.01) * h_ij * h_ij);
        acc += (P.viscosity * coeff) * dg;
      }
    }
    V pull;
    if (P.raw.has_pull_fluid_to) {
      V d = P.pull - position[i];
      pull = (d / d.norm()) * FT(13);
    }
    return acc + V(0, P.gravity) + pull;
  }
  void update_velocity_with_non_pressure_accel(const Params<FT>& P, FT dt) {  // sim.rs:1051-1077
    const size_t N = n();
#pragma omp parallel for schedule(static)
    for (int64_t ii = 0; ii < int64_t(N); ii++) {
      size_t i = size_t(ii);
      velocity_temp[i] = velocity[i] + dt * non_pressure_accel(i, velocity, P);
    }
    std::swap(velocity, velocity_temp);
  }
  bool calculate_all_densities(const Params<FT>&, StepError& err) {  // sim.rs:1007-1049
    const size_t N = n();
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(max : bad)
    for (int64_t ii = 0; ii < int64_t(N); ii++) {
      size_t i = size_t(ii);
      FT acc = 0;
      for (uint32_t j : neighs[i]) {
        V x_ab = position[i] - position[j];
        acc += mass[j] * kernelh<FT>(x_ab, hij(i, j));
      }
      acc += lambda_sum(i);
      density[i] = acc;
      if (!std::isfinite(acc)) bad = std::max(bad, 2);
      else if (!(acc > FT(0.0001))) bad = std::max(bad, 1);
    }
    if (bad == 2) { err = {ASPH_ERR_NONFINITE, "density not finite"}; return false; }
    if (bad == 1) { err = {ASPH_ERR_DENSITY, "density <= 0.0001"}; return false; }
    return true;
  }
  void calculate_constant_field(const Params<FT>& P) {  // sim.rs:2235-2248
    const size_t N = n();
#pragma omp parallel for schedule(static)
    for (int64_t ii = 0; ii < int64_t(N); ii++) {
      size_t i = size_t(ii);
      FT c = 0;
      for (uint32_t j : neighs[i]) c += mass[j] / density[j] * kernelh<FT>(position[i] - position[j], hij(i, j));
      c += lambda_sum(i) / P.rest_density;
      constant_field[i] = c;
    }
  }
  template <class QF> FT divergence_iisph(size_t i, QF qf, V qb, const Params<FT>& P) const {  // sim.rs:1552-1592
    FT sum = 0;
    V qi = qf(i);
    for (uint32_t j : neighs[i]) {
      V dg = kernel_derivh<FT>(position[i] - position[j], hij(i, j));
      if (P.raw.operator_discretization == ASPH_OP_WINCHENBACH2020)
        sum += mass[j] / density[j] * (qf(j) - qi).dot(dg);
      else
        sum += mass[j] / density[i] * (qf(j) - qi).dot(dg);
    }
    return sum + boundary_divergence(i, qf, qb, P);
  }
  V pressure_accel_of(size_t i, const std::vector<FT>& p, const Params<FT>& P) const {  // sim.rs:1751-1808
    FT p1 = p[i] / (density[i] * density[i]);
    V acc;
    for (uint32_t j : neighs[i]) {
      V dg = kernel_derivh<FT>(position[i] - position[j], hij(i, j));
      FT p2 = p[j] / (density[j] * density[j]);
      acc += (-mass[j] * (p1 + p2)) * dg;
    }
    return acc + boundary_pressure_accel(i, p, P);
  }
  void calculate_pressure_accels(const std::vector<FT>& p, const Params<FT>& P) {  // sim.rs:1518-1543
    const size_t N = n();
#pragma omp parallel for schedule(static)
    for (int64_t ii = 0; ii < int64_t(N); ii++) pressure_accel[size_t(ii)] = pressure_accel_of(size_t(ii), p, P);
  }
  FT next_density_estimate(size_t i, const Params<FT>& P) const {
    return P.raw.operator_discretization == ASPH_OP_WINCHENBACH2020 ? P.rest_density : density[i];
  }
  enum SourceKind { SRC_DIVERGENCE, SRC_ONLY_DENSITY, SRC_FULL, SRC_FULL_WITH_OMEGA };
  void prepare_ppe(SourceKind kind, const Params<FT>& P, FT dt) {  // sim.rs:1127-1204, 1633-1748
    const size_t N = n();
    auto vf = [&](size_t j) { return velocity[j]; };
#pragma omp parallel for schedule(static)
    for (int64_t ii = 0; ii < int64_t(N); ii++) {
      size_t i = size_t(ii);
      pressure[i] = 0;
      FT s;
      if (kind == SRC_DIVERGENCE) {
        FT div = divergence_iisph(i, vf, V(), P);
        s = -div / dt;
      } else if (kind == SRC_ONLY_DENSITY) {
        s = -(P.rest_density - density[i]) / (next_density_estimate(i, P) * dt * dt);
      } else if (kind == SRC_FULL) {
        FT div = divergence_iisph(i, vf, V(), P);
        s = -(P.rest_density - density[i]) / (next_density_estimate(i, P) * dt * dt) - div / dt;
      } else {  // calculate_source_term_full_with_omega sim.rs:1678-1710: next_density_estimate is the rest density
        FT div = divergence_iisph(i, vf, V(), P);
        s = -(P.rest_density - density[i]) / (P.rest_density * dt * dt) - div / (dt * omega[i]);
      }
      ppe_source_term[i] = s;
    }
  }

  struct Stats {  // PressureSolverStatistics sim.rs:397-469
    uint64_t normal = 0, singular = 0, negative = 0;
    FT err_sum = 0, max_error = 0;
  };
  // one relaxed-Jacobi sweep, sim.rs:1207-1322.  The reduce is done in fixed chunks of 4096 particles,
  // chunk partials added in index order (the reference's rayon tree shape is scheduling dependent).
  bool single_pressure_iteration(bool density_residual, const Params<FT>& P, FT dt, Stats& out, StepError& err) {
    const size_t N = n();
    const FT w = P.jacobi_omega;
    calculate_pressure_accels(pressure, P);
    const size_t CH = 4096, nch = (N + CH - 1) / CH;
    std::vector<Stats> part(nch);
    int bad = 0;
    auto af = [&](size_t j) { return pressure_accel[j]; };
#pragma omp parallel for schedule(static) reduction(max : bad)
    for (int64_t c = 0; c < int64_t(nch); c++) {
      Stats st;
      for (size_t i = size_t(c) * CH; i < std::min(N, size_t(c + 1) * CH); i++) {
        if (std::fabs(aii[i]) < FT(10e-4)) { pressure_next_iter[i] = 0; st.singular++; continue; }
        FT a_p = divergence_iisph(i, af, V(), P);
        FT source = ppe_source_term[i];
        if (!std::isfinite(a_p)) { bad = 1; continue; }
        FT pn = pressure[i] + w * (source - a_p) / aii[i];
        if (!std::isfinite(pn)) { bad = 1; continue; }
        FT perr;
        if (density_residual) { perr = density[i] * dt * dt * (source - a_p); density_error[i] = perr; }
        else perr = dt * (source - a_p);
        if (pn <= FT(0)) { pressure_next_iter[i] = 0; st.negative++; }
        else { pressure_next_iter[i] = pn; st.normal++; st.err_sum += perr; st.max_error = std::max(st.max_error, std::fabs(perr)); }
      }
      part[size_t(c)] = st;
    }
    if (bad) { err = {ASPH_ERR_NONFINITE, "'!a_p.is_finite()' failed. Pressure values probably have exploded!"}; return false; }
    Stats tot;
    for (auto& s : part) {
      tot.normal += s.normal; tot.singular += s.singular; tot.negative += s.negative;
      tot.err_sum += s.err_sum; tot.max_error = std::max(tot.max_error, s.max_error);
    }
    out = tot;
    return true;
  }
  // iisph_pressure_iterations sim.rs:1378-1516; returns num_pressure_iters
  bool pressure_iterations(FT max_avg_error, bool density_residual, const Params<FT>& P, FT dt, int& iters_out,
                           double& last_avg, StepError& err) {
    const size_t N = n();
    for (size_t i = 0; i < N; i++)
      if (aii[i] < FT(0)) { err = {ASPH_ERR_NEG_AII, "AII should not be negative! i=" + std::to_string(i)}; return false; }
    int64_t k = 0;
    for (;;) {
      Stats st;
      if (!single_pressure_iteration(density_residual, P, dt, st, err)) return false;
      std::swap(pressure, pressure_next_iter);
      FT avg = st.normal > 0 ? st.err_sum / FT(st.normal) : std::numeric_limits<FT>::quiet_NaN();
      last_avg = double(avg);
      if (density_residual) {
        if (st.normal == 0 || (std::fabs(avg / P.rest_density) < max_avg_error && k > 1)) break;
      } else {
        if (st.normal == 0 || (std::fabs(avg) < max_avg_error / dt && k > 1)) break;
      }
      if (k == P.raw.max_iters) break;  // "Pressure sover not converged" — not an error
      k++;
    }
    calculate_pressure_accels(pressure, P);
    iters_out = int(k);
    return true;
  }
  bool compute_aii(const Params<FT>& P, StepError& err) {  // sim.rs:1080-1125
    const size_t N = n();
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(max : bad)
    for (int64_t ii = 0; ii < int64_t(N); ii++) {
      aii[size_t(ii)] = boundary_aii(size_t(ii), P);
      if (!std::isfinite(aii[size_t(ii)])) bad = 1;
    }
    if (bad) { err = {ASPH_ERR_NONFINITE, "aii not finite"}; return false; }
    return true;
  }
  // operator diagonal by applying the operator to a unit pressure vector: check_aii sim.rs:1324-1375
  FT aii_inefficient(size_t check_i, const Params<FT>& P) const {
    std::vector<FT> unit(n(), FT(0));
    unit[check_i] = 1;
    auto af = [&](size_t j) { return pressure_accel_of(j, unit, P); };
    return divergence_iisph(check_i, af, V(), P);
  }

  // support length from the particle distribution, sim.rs:1873-1971 ("Constrained Neighbor Lists for SPH-based Fluid
  // Simulations" eq. 3-4), and the neighbour-count constraint, sim.rs:2145-2177.  lambda_sum is the boundary handler's
  // state of the PREVIOUS step (update_after_advect runs afterwards, sim.rs:2179).
  bool estimate_support_lengths(const Params<FT>& P, StepError& err) {
    const size_t N = n();
    const int mode = P.raw.support_length_estimation;
    if (mode != ASPH_H_FROM_MASS) {
      int bad = 0;
#pragma omp parallel for schedule(static) reduction(max : bad)
      for (int64_t ii = 0; ii < int64_t(N); ii++) {
        size_t i = size_t(ii);
        const FT w = FT(0.5);
        FT volume_estimate;
        if (mode == ASPH_H_FROM_DISTRIBUTION2) {
          FT v_w_sum = 0;
          for (uint32_t j : neighs[i]) v_w_sum += (mass[j] / P.rest_density) * kernelh<FT>(position[i] - position[j], hij(i, j));
          FT vi = mass[i] / P.rest_density;
          volume_estimate = vi / (v_w_sum + lambda_sum(i));
        } else {
          FT w_sum = 0;
          for (uint32_t j : neighs[i]) w_sum += kernelh<FT>(position[i] - position[j], hij(i, j));
          volume_estimate = (FT(1) - std::min(lambda_sum(i), FT(0.5))) / w_sum;
        }
        if (!(volume_estimate >= FT(0))) { bad = 1; continue; }
        FT h_new = FT(ETA) * volume_to_radius<FT>(volume_estimate);
        FT hn = w * h_new + (FT(1) - w) * h2[i];
        if (mode == ASPH_H_FROM_DISTRIBUTION_CLAMPED1) hn = std::min(hn, FT(1) * h_from_mass<FT>(mass[i], P.rest_density));
        if (mode == ASPH_H_FROM_DISTRIBUTION_CLAMPED2) hn = std::min(hn, FT(2) * h_from_mass<FT>(mass[i], P.rest_density));
        h2_next[i] = hn;
      }
      if (bad) { err = {ASPH_ERR_INVALID, "assert!(volume_estimate >= 0.)"}; return false; }
    }
    if (P.raw.constrain_neighborhood_count) {
      const size_t target = size_t(FT(ETA * 2) * FT(ETA * 2)) + 5;  // optimal_neighbor_number (sim.rs:386) as usize + 5
      int bad = 0;
#pragma omp parallel for schedule(static) reduction(max : bad)
      for (int64_t ii = 0; ii < int64_t(N); ii++) {
        size_t i = size_t(ii);
        const size_t cnt = neighs[i].size();
        if (cnt > target) {
          std::vector<FT> fringe;
          fringe.reserve(cnt);
          for (uint32_t j : neighs[i]) fringe.push_back(FT(2) * (position[i] - position[j]).norm() - h2[j] * FT(2));
          std::sort(fringe.begin(), fringe.end(), [](FT a, FT b) { return a > b; });
          FT hn = fringe[cnt - target];
          if (!(hn < h2[i]) || !(hn >= FT(0))) { bad = 1; continue; }
          h2_next[i] = hn;
          flag_neighborhood_reduced[i] = 1;
        } else {
          h2_next[i] = h2[i];
          flag_neighborhood_reduced[i] = 0;
        }
      }
      if (bad) { err = {ASPH_ERR_INVALID, "constrain_neighborhood_count: assert!(*p_h_next < h) / assert!(*p_h_next >= 0.)"}; return false; }
      std::swap(h2, h2_next);
    }
    return true;
  }
  // IISPH2 correction factor, sim.rs:2263-2311
  void compute_omega(const Params<FT>& P) {
    const size_t N = n();
    auto dwdh = [](FT d, FT H) {
      FT q = d / H;
      FT cd = FT(40) / (FT(7) * Consts<FT>::PI);
      FT w = cubic_unnorm<FT>(q), wd = cubic_unnorm_deriv<FT>(q);
      return cd * -FT(2) / (H * H * H) * w + cd / (H * H) * wd * (-d / (H * H));
    };
    (void)P;
    for (size_t i = 0; i < N; i++) {
      FT om = 1;
      const FT H_i = h2[i] * FT(2);
      if (size_class[i] == ASPH_CLASS_LARGE) {
        om += H_i / (FT(3) * density[i]) * mass[i] * dwdh(FT(0), h2[i] * FT(2));
      } else {
        for (uint32_t j : neighs[i]) {
          FT d = (position[i] - position[j]).norm();
          om += H_i / (FT(3) * density[i]) * mass[j] * dwdh(d, hij(i, j) * FT(2));
        }
      }
      omega[i] = std::min(FT(2.5), std::max(om, FT(0.125)));
    }
  }

  // ---------------------------------------------------------------- the step, sim.rs:1980-2730
  bool step_physics(const Params<FT>& P, FT& dt_out, StepError& err) {
    const size_t N = n();
    info = asph_step_info();
    info.n_particles_begin = N;
    pc.begin(ASPH_PC_SIMULATION_STEP);
    // 1: kernel support, sim.rs:1998-2016: from mass (sim.rs:1865-1871), or the length estimated in the last step
    if (P.raw.support_length_estimation == ASPH_H_FROM_MASS) {
#pragma omp parallel for schedule(static)
      for (int64_t ii = 0; ii < int64_t(N); ii++) h2[size_t(ii)] = h_from_mass<FT>(mass[size_t(ii)], P.rest_density);
    } else {
      std::swap(h2, h2_next);
    }

    if (!P.raw.level_estimation_after_advection) {  // sim.rs:2018-2058
      if (!P.raw.use_extended_range_for_level_estimation || P.raw.level_estimation_method == ASPH_LEVEL_CENTER_DIFF) {
        err = {ASPH_ERR_INVALID, "level estimation before advection needs the extended range and not CenterDiff"}; return false;
      }
      pc.begin(ASPH_PC_NEIGHBORHOOD);
      if (!build_neighbors(P.level_estimation_range / FT(ETA), err)) return false;
      pc.end(ASPH_PC_NEIGHBORHOOD);
      pc.begin(ASPH_PC_LEVEL_ESTIMATION);
      if (!perform_level_estimation(P, err)) return false;
      pc.end(ASPH_PC_LEVEL_ESTIMATION);
      pc.begin(ASPH_PC_NEIGHBORHOOD);
      filter_down(FT(2));
      pc.end_add_to_last(ASPH_PC_NEIGHBORHOOD);
    } else {
      pc.begin(ASPH_PC_NEIGHBORHOOD);
      if (!build_neighbors(FT(2), err)) return false;
      pc.end(ASPH_PC_NEIGHBORHOOD);
    }
    if (!estimate_support_lengths(P, err)) return false;  // sim.rs:2090-2177
    boundary_update_after_advect(P);                      // sim.rs:2179
    // CFL, sim.rs:2182-2191
    FT min_cfl = std::numeric_limits<FT>::infinity();
    for (size_t i = 0; i < N; i++) {
      FT sr = h2[i] * FT(2);
      FT c = sr * sr / (velocity[i].norm_squared() + FT(0.01));
      min_cfl = std::min(min_cfl, c);
    }
    FT cfl_dt = P.cfl_factor * std::sqrt(min_cfl);
    FT dt = std::min(P.max_dt, cfl_dt);
    info.dt = float(dt);

    if (!calculate_all_densities(P, err)) return false;  // sim.rs:2204
    calculate_constant_field(P);                         // sim.rs:2235
    if (!compute_aii(P, err)) return false;              // sim.rs:2250

    int iters = 0;
    switch (P.raw.pressure_solver_method) {
      case ASPH_SOLVER_IISPH: {  // sim.rs:2389-2446
        update_velocity_with_non_pressure_accel(P, dt);
        prepare_ppe(SRC_FULL, P, dt);
        pc.begin(ASPH_PC_DENSITY_SOLVER);
        if (!pressure_iterations(P.iisph_max_avg_density_error, true, P, dt, iters, info.last_avg_error_density, err)) return false;
        pc.end(ASPH_PC_DENSITY_SOLVER);
        info.density_iterations = iters; info.density_sweeps = iters + 1;
#pragma omp parallel for schedule(static)
        for (int64_t ii = 0; ii < int64_t(N); ii++) {
          size_t i = size_t(ii);
          velocity[i] += dt * pressure_accel[i];
          position[i] += dt * velocity[i];
        }
        break;
      }
      case ASPH_SOLVER_ONLY_DIVERGENCE: {  // sim.rs:2448-2500
        update_velocity_with_non_pressure_accel(P, dt);
        prepare_ppe(SRC_DIVERGENCE, P, dt);
        pc.begin(ASPH_PC_DIV_SOLVER);
        if (!pressure_iterations(P.hybrid_dfsph_max_avg_divergence_error, false, P, dt, iters, info.last_avg_error_div, err)) return false;
        pc.end(ASPH_PC_DIV_SOLVER);
        info.div_iterations = iters; info.div_sweeps = iters + 1;
#pragma omp parallel for schedule(static)
        for (int64_t ii = 0; ii < int64_t(N); ii++) {
          size_t i = size_t(ii);
          velocity[i] += dt * pressure_accel[i];
          position[i] += dt * velocity[i];
        }
        break;
      }
      case ASPH_SOLVER_HYBRID_DFSPH: {  // sim.rs:2502-2670
        if (P.raw.hybrid_dfsph_non_pressure_accel_before_divergence_free) update_velocity_with_non_pressure_accel(P, dt);
        pc.begin(ASPH_PC_DIV_SOLVER);
        prepare_ppe(SRC_DIVERGENCE, P, dt);
        if (!pressure_iterations(P.hybrid_dfsph_max_avg_divergence_error, false, P, dt, iters, info.last_avg_error_div, err)) return false;
        pc.end(ASPH_PC_DIV_SOLVER);
        info.div_iterations = iters; info.div_sweeps = iters + 1;
#pragma omp parallel for schedule(static)
        for (int64_t ii = 0; ii < int64_t(N); ii++) velocity[size_t(ii)] += dt * pressure_accel[size_t(ii)];
        if (!P.raw.hybrid_dfsph_non_pressure_accel_before_divergence_free) update_velocity_with_non_pressure_accel(P, dt);
        pc.begin(ASPH_PC_DENSITY_SOLVER);
        prepare_ppe(P.raw.hybrid_dfsph_density_source_term == ASPH_SRC_ONLY_DENSITY ? SRC_ONLY_DENSITY : SRC_FULL, P, dt);
        if (!pressure_iterations(P.hybrid_dfsph_max_avg_density_error, true, P, dt, iters, info.last_avg_error_density, err)) return false;
        pc.end(ASPH_PC_DENSITY_SOLVER);
        info.density_iterations = iters; info.density_sweeps = iters + 1;
        const FT fac = std::min(dt * P.hybrid_dfsph_factor, FT(1));
#pragma omp parallel for schedule(static)
        for (int64_t ii = 0; ii < int64_t(N); ii++) {
          size_t i = size_t(ii);
          position[i] += dt * velocity[i] + (dt * dt) * pressure_accel[i];
          velocity[i] += (dt * pressure_accel[i]) * fac;
        }
        break;
      }
      default: {  // IISPH2, sim.rs:2262-2387
        compute_omega(P);
        update_velocity_with_non_pressure_accel(P, dt);
        prepare_ppe(SRC_FULL_WITH_OMEGA, P, dt);
        pc.begin(ASPH_PC_DENSITY_SOLVER);
        if (!pressure_iterations(P.iisph_max_avg_density_error, true, P, dt, iters, info.last_avg_error_density, err)) return false;
        pc.end(ASPH_PC_DENSITY_SOLVER);
        info.density_iterations = iters; info.density_sweeps = iters + 1;
        for (size_t i = 0; i < N; i++) pressure[i] /= std::sqrt(omega[i]);
        calculate_pressure_accels(pressure, P);
#pragma omp parallel for schedule(static)
        for (int64_t ii = 0; ii < int64_t(N); ii++) {
          size_t i = size_t(ii);
          velocity[i] += dt * pressure_accel[i];
          position[i] += dt * velocity[i];
        }
        break;
      }
    }
    for (size_t i = 0; i < N; i++)
      if (!position[i].finite() || !velocity[i].finite()) { err = {ASPH_ERR_NONFINITE, "position/velocity not finite"}; return false; }

    if (P.raw.level_estimation_after_advection) {  // sim.rs:2678-2707
      if (P.raw.use_extended_range_for_level_estimation)
        if (!build_neighbors(P.level_estimation_range / FT(ETA), err)) return false;
      pc.begin(ASPH_PC_LEVEL_ESTIMATION);
      if (!perform_level_estimation(P, err)) return false;
      pc.end(ASPH_PC_LEVEL_ESTIMATION);
    }
    pc.begin(ASPH_PC_LEVEL_ESTIMATION);
    if (!smooth_level_field(P, err)) return false;  // sim.rs:2710
    pc.end_add_to_last(ASPH_PC_LEVEL_ESTIMATION);

    time_ft += dt;
    time = double(time_ft);
    step_number += 1;
    pc.end(ASPH_PC_SIMULATION_STEP);
    dt_out = dt;
    info.n_particles_end = n();
    return true;
  }

  // ---------------------------------------------------------------- adaptivity
  void classify_particles(const Params<FT>& P) {  // adaptivity/mod.rs:50-59
    const size_t N = n();
#pragma omp parallel for schedule(static)
    for (int64_t ii = 0; ii < int64_t(N); ii++)
      size_class[size_t(ii)] = classify<FT>(level_estimation[size_t(ii)], mass[size_t(ii)], P);
  }
  FT dropped_mass_sharing(size_t i, FT m, FT dt, const Params<FT>& P) const {  // particle_sharing.rs:242-253
    FT tm = target_mass<FT>(level_estimation[i], P);
    return std::min(m - tm, tm * P.max_mass_transfer_sharing * dt);
  }
  // find_share_partner_sequential (particle_sharing.rs:14-111) and find_merge_partner_sequential
  // (particle_merging.rs:16-122) are the same greedy with different donor class / eligibility / drop.
  int find_partner_sequential(bool merging, const Params<FT>& P, FT dt) {
    const size_t N = n();
    for (size_t i = 0; i < N; i++) merge_partner[i] = ASPH_MERGE_PARTNER_AVAILABLE;
    int count = 0;
    const uint8_t donor_class = merging ? ASPH_CLASS_TOO_SMALL : ASPH_CLASS_LARGE;
    const FT dist_factor = merging ? P.max_merge_distance : P.max_share_distance;
    for (size_t i = 0; i < N; i++) {
      merge_counter[i] = 0;
      if (size_class[i] != donor_class) continue;
      for (uint32_t j : neighs[i]) {
        if (i == j) continue;
        bool can;
        if (merging) {
          switch (size_class[j]) {
            case ASPH_CLASS_LARGE: case ASPH_CLASS_TOO_LARGE: can = false; break;
            case ASPH_CLASS_OPTIMAL: can = P.raw.allow_merge_with_optimal_particle != 0; break;
            default: can = true;
          }
          if (P.raw.allow_merge_on_size_difference && mass[j] > FT(5) * mass[i]) can = true;
        } else {
          switch (size_class[j]) {
            case ASPH_CLASS_SMALL: can = true; break;
            case ASPH_CLASS_TOO_SMALL: can = P.raw.allow_share_with_too_small_particle != 0; break;
            case ASPH_CLASS_OPTIMAL: can = P.raw.allow_share_with_optimal_particle != 0; break;
            default: can = false;
          }
        }
        if (!can) continue;
        V xij = position[i] - position[j];
        FT max_dist = hij(i, j) * dist_factor;
        if (xij.norm_squared() > max_dist * max_dist) continue;
        FT dropped = merging ? mass[i] : dropped_mass_sharing(i, mass[i], dt, P);
        FT new_mass_j = mass[j] + dropped / FT(merge_counter[i] + 1);
        FT target_j = target_mass<FT>(level_estimation[j], P);
        if (new_mass_j >= target_j * FT(1.1)) continue;
        if (new_mass_j > P.mass_base()) continue;
        if (merge_partner[j] != ASPH_MERGE_PARTNER_AVAILABLE) continue;
        if (merge_counter[i] == 0) {
          if (merge_partner[i] != ASPH_MERGE_PARTNER_AVAILABLE) continue;
          merge_partner[i] = ASPH_MERGE_PARTNER_DELETE;
        }
        merge_partner[j] = uint32_t(i);
        merge_counter[i] += 1;
        count++;
      }
    }
    return count;
  }
  // validate_share_partners particle_sharing.rs:113-150 / validate_merge_partners particle_merging.rs:230-268
  bool validate_partners(uint8_t donor_class) const {
    for (size_t i = 0; i < n(); i++) {
      if (merge_counter[i] > 0) {
        if (size_class[i] != donor_class) return false;
        if (merge_partner[i] != ASPH_MERGE_PARTNER_DELETE) return false;
        uint32_t c2 = 0;
        for (uint32_t j : neighs[i]) if (merge_partner[j] == uint32_t(i)) c2++;
        if (c2 != merge_counter[i]) return false;
      } else {
        if (merge_partner[i] == ASPH_MERGE_PARTNER_DELETE) return false;
        if (merge_partner[i] != ASPH_MERGE_PARTNER_AVAILABLE)
          if (merge_partner[merge_partner[i]] != ASPH_MERGE_PARTNER_DELETE) return false;
      }
    }
    return true;
  }
  void apply_receivers(bool merging, const Params<FT>& P, FT dt) {  // particle_sharing.rs:164-211, particle_merging.rs:282-324
    const size_t N = n();
    const int minp = merging ? P.raw.minimum_merge_partners : P.raw.minimum_share_partners;
    // receivers only read donor attributes, and donors are never receivers: order independent
    for (size_t i = 0; i < N; i++) {
      uint32_t j = merge_partner[i];
      if (j == ASPH_MERGE_PARTNER_AVAILABLE || j == ASPH_MERGE_PARTNER_DELETE) continue;
      if (int(merge_counter[j]) < minp) continue;
      FT mass_i = mass[i], mass_j = mass[j];
      FT dropped = merging ? mass_j : dropped_mass_sharing(j, mass_j, dt, P);
      FT mass_n = dropped / FT(merge_counter[j]);
      FT m = mass_i + mass_n;
      velocity[i] = (mass_i * velocity[i] + mass_n * velocity[j]) / m;
      position[i] = (mass_i * position[i] + mass_n * position[j]) / m;
      mass[i] = m;
      h2_next[i] = h_from_mass<FT>(m, P.rest_density);  // particle_sharing.rs:206, particle_merging.rs:323
    }
  }
  void share_particles(const Params<FT>& P, FT dt) {  // particle_sharing.rs:152-240
    apply_receivers(false, P, dt);
    for (size_t i = 0; i < n(); i++) {
      if (merge_partner[i] != ASPH_MERGE_PARTNER_DELETE) continue;
      if (int(merge_counter[i]) < P.raw.minimum_share_partners) continue;
      mass[i] -= dropped_mass_sharing(i, mass[i], dt, P);
      h2_next[i] = h_from_mass<FT>(mass[i], P.rest_density);  // particle_sharing.rs:238
    }
  }
  void swap_particles(size_t a, size_t b) {  // ParticleVec::swap sim.rs:249-253 (+ neighs, boundary)
    std::swap(mass[a], mass[b]); std::swap(position[a], position[b]); std::swap(velocity[a], velocity[b]);
    std::swap(velocity_temp[a], velocity_temp[b]); std::swap(pressure_accel[a], pressure_accel[b]);
    std::swap(density[a], density[b]); std::swap(ppe_source_term[a], ppe_source_term[b]);
    std::swap(pressure[a], pressure[b]); std::swap(pressure_next_iter[a], pressure_next_iter[b]);
    std::swap(aii[a], aii[b]); std::swap(density_error[a], density_error[b]); std::swap(h2[a], h2[b]);
    std::swap(constant_field[a], constant_field[b]); std::swap(h2_next[a], h2_next[b]); std::swap(omega[a], omega[b]);
    std::swap(flag_neighborhood_reduced[a], flag_neighborhood_reduced[b]);
    std::swap(level_estimation[a], level_estimation[b]); std::swap(level_estimation_temp[a], level_estimation_temp[b]);
    std::swap(size_class[a], size_class[b]);
    std::swap(flag_is_fluid_surface[a], flag_is_fluid_surface[b]);
    std::swap(flag_insufficient_neighs[a], flag_insufficient_neighs[b]);
    std::swap(merge_partner[a], merge_partner[b]); std::swap(merge_counter[a], merge_counter[b]);
    neighs[a].swap(neighs[b]); lambda[a].swap(lambda[b]);
  }
  void merge_particles(const Params<FT>& P, FT dt) {  // particle_merging.rs:270-371
    apply_receivers(true, P, dt);
    const size_t N = n();
    if (N == 0) return;
    int64_t last = int64_t(N) - 1, i = 0;
    while (i <= last) {
      if (merge_partner[size_t(i)] == ASPH_MERGE_PARTNER_DELETE &&
          int(merge_counter[size_t(i)]) >= P.raw.minimum_merge_partners) {
        mass[size_t(i)] -= mass[size_t(i)];  // dropped_mass_merging == mass particle_merging.rs:373-385
        if (mass[size_t(i)] < FT(0.000001)) {
          swap_particles(size_t(i), size_t(last));
          last--;
          continue;
        }
      }
      i++;
    }
    resize_all(size_t(last + 1));
  }
  bool split_particles(const Params<FT>& P, int& parents, StepError& err) {  // splitting.rs:19-81
    const size_t N = n();
    size_t new_id = N;
    parents = 0;
    for (size_t i = 0; i < N; i++) {
      if (size_class[i] != ASPH_CLASS_TOO_LARGE) continue;
      FT tm = target_mass<FT>(level_estimation[i], P);
      size_t nc = size_t(std::round(mass[i] / tm));
      if (nc > size_t(max_children)) {
        if (P.raw.fail_on_missing_split_pattern) { err = {ASPH_ERR_INVALID, "no split pattern for a 1-to-n split"}; return false; }
        nc = size_t(max_children);
      }
      if (nc < 2) { err = {ASPH_ERR_INVALID, "assert!(num_children > 1)"}; return false; }
      const float* pat = &split_pos[2 * size_t(split_offset[nc - 2])];
      FT radius = volume_to_radius<FT>(mass[i] / FT(1));  // INIT_REST_DENSITY
      FT child_mass = mass[i] / FT(nc);
      FT child_h = h_from_mass<FT>(child_mass, P.rest_density);
      V ov = velocity[i], op = position[i];
      Level<FT> ol = level_estimation[i];
      resize_all(n() + nc - 1);
      for (size_t c = 0; c < nc; c++) {
        V off = V(FT(pat[2 * c]), FT(pat[2 * c + 1])) * radius;
        size_t t = (c == 0) ? i : new_id++;
        mass[t] = child_mass; velocity[t] = ov; position[t] = op + off; level_estimation[t] = ol;
        h2[i] = child_h; h2_next[t] = child_h;  // splitting.rs:65-74: h2 of an appended child stays 0 until the next step
      }
      parents++;
    }
    return true;
  }
  // The reference sums with iter().sum() in FT (sim.rs:2745, 2791).  Its f32 running sum drifts by a fraction of an ulp
  // per addend and trips the 0.005 tolerance by itself near a million equal particles (before vs after a merge step the
  // addends differ), which says nothing about conservation; the check is kept, the sum is accumulated in double.
  FT total_mass() const {
    double s = 0;
    for (FT m : mass) s += double(m);
    return FT(s);
  }
  bool step_adaptivity(const Params<FT>& P, FT dt, StepError& err) {  // sim.rs:2732-2796
    pc.begin(ASPH_PC_SIMULATION_STEP);
    pc.begin(ASPH_PC_ADAPTIVITY);
    if ((P.raw.sharing || P.raw.merging || P.raw.splitting) && P.raw.level_estimation_method == ASPH_LEVEL_NONE) {
      err = {ASPH_ERR_INVALID, "resampling needs a level estimation (level() on FluidInterior is unreachable!, sim.rs:204-211)"};
      return false;
    }
    FT m1 = total_mass();
    if (P.raw.sharing) {
      classify_particles(P);
      info.n_shared = find_partner_sequential(false, P, dt);
      if (!validate_partners(ASPH_CLASS_LARGE)) { err = {ASPH_ERR_INVALID, "validate_share_partners"}; return false; }
      share_particles(P, dt);
    }
    if (step_number % 2 == 0) {
      if (P.raw.merging) {
        classify_particles(P);
        info.n_merged = find_partner_sequential(true, P, dt);
        if (!validate_partners(ASPH_CLASS_TOO_SMALL)) { err = {ASPH_ERR_INVALID, "validate_merge_partners"}; return false; }
        merge_particles(P, dt);
      }
    } else {
      if (P.raw.splitting) {
        classify_particles(P);
        if (!split_particles(P, info.n_split_parents, err)) return false;
      }
    }
    FT m2 = total_mass();
    if (!(m2 <= m1 + FT(0.005) && m2 >= m1 - FT(0.005))) { err = {ASPH_ERR_MASS_CONSERVATION, "mass sum"}; return false; }
    pc.end(ASPH_PC_ADAPTIVITY);
    pc.end_add_to_last(ASPH_PC_SIMULATION_STEP);
    info.n_particles_end = n();
    return true;
  }
};

}  // namespace oracle
