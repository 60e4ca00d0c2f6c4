// sim.cuh — internal declarations of libasph_b200.so (hand-written sm_100a CUDA; no tensor cores: the
// path has no dense contraction).  Device state layout, device math, launcher prototypes.
//
// Data layout in HBM (all SoA, fp32, particles kept SORTED by (size level, cell) every step):
//   persistent  pos float2 | vel float2 | mass f32 | refid u32 (index in the reference's ParticleVec) | level f32
//   per step    xyhm float4 {x, y, h, m}  (one 16 B gather per neighbour candidate)
//               neighbour lists in sliced-ELL: slice = 32 consecutive particles = one warp; entry k of lane l at
//               slice_base[s] + 32*k + l  -> every warp load of idx/coef is one fully coalesced 128 B line.
//               Row k < cnt_near holds the 2h neighbours, cnt_near <= k < cnt_ext the extended-range
//               (level-set) ones, so NeighborhoodCache::filter_down (neighborhood_search.rs:56) is free.
//               coef[k] = m_j * dW/dr / r  so that  m_j * gradW_ij = coef * (x_i - x_j)
//   solver      packP float4 {x, y, p/rho^2, p} | packA float4 {x, y, a^p_x, a^p_y} | jc float4 {rho0*G, s, a_ii}
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/asph.h"

#define ASPH_MAX_LEVELS 10
#define ASPH_SLACK 1.00390625f  // 1 + 1/256: cell / search-radius safety factor against fp32 binning error

#define CUDA_TRY(x)                                                                              \
  do {                                                                                           \
    cudaError_t e__ = (x);                                                                       \
    if (e__ != cudaSuccess) {                                                                    \
      sim->last_error = std::string(#x) + ": " + cudaGetErrorString(e__);                        \
      return ASPH_ERR_CUDA;                                                                      \
    }                                                                                            \
  } while (0)

// ------------------------------------------------------------------------------------------------
// device-visible control block: everything data-dependent that decides control flow lives here so that the
// step needs ONE host synchronisation (at its end) plus one per solver batch.
// ------------------------------------------------------------------------------------------------
enum {
  ERRF_NONFINITE = 1, ERRF_NEG_AII = 2, ERRF_DENSITY = 4, ERRF_NEIGHBOR_OVERFLOW = 8, ERRF_LIST_CAPACITY = 16,
  ERRF_PARTICLE_CAPACITY = 32, ERRF_SPLIT_PATTERN = 64, ERRF_LEVEL_WEIGHT = 128
};

struct GridLevel {
  float cell, inv_cell, hmax;
  int nx, ny;
  uint32_t base;   // first cell of this level in the concatenated cell array
  uint32_t count;  // particles in this level
};

struct SolverCtl {
  int k;        // index of the sweep being executed (num_pressure_iters, simulation.rs:1388)
  int done;     // set by the sweep that satisfies the stop rule of simulation.rs:1453-1477
  int sweeps;   // sweeps executed
  unsigned long long normal, singular, negative;
  float err_sum, max_err, avg;
  unsigned int ticket;
};

struct StepCtl {
  // order-preserving encodings (see enc_f / dec_f) for atomicMin / atomicMax on floats
  unsigned int hmin_enc, hmax_enc, minx_enc, miny_enc, maxx_enc, maxy_enc, cfl_enc;
  unsigned int lvl_hmax_enc[ASPH_MAX_LEVELS];
  unsigned int lvl_count[ASPH_MAX_LEVELS];
  int nlevels;
  float hmin, origin_x, origin_y;
  GridLevel lv[ASPH_MAX_LEVELS];
  uint32_t total_cells;
  float dt;
  unsigned int error_flags;
  unsigned long long list_entries;  // total sliced-ELL entries needed this step
  uint32_t max_count;               // largest neighbour count
  SolverCtl solver;
  // level set
  uint32_t front_size[2];
  int level_sweeps;
  // adaptivity
  uint32_t n_new;                 // particle count after merge / split
  uint32_t n_shared, n_merged, n_split_parents;
  uint32_t undecided;
  double mass_before, mass_after;
};

struct PackedParams {  // SimulationParams rounded once to fp32 (what serde does for the f32 build)
  float rest_density, cfl_factor, max_dt, viscosity, gravity, jacobi_omega, sdf_gradient_eps;
  float particle_radius_fine, particle_radius_base, maximum_surface_distance, mass_fine, mass_base;
  float max_mass_transfer_sharing, max_share_distance, max_merge_distance;
  float max_avg_density_error_iisph, hybrid_factor, max_avg_density_error, max_avg_divergence_error;
  float f_ext, f_near;  // range factors: level_estimation_range / ETA and 2
  float pull_x, pull_y;
  int has_pull, viscosity_type, level_method, solver, density_source, np_before_div, penalty, sizing, opdisc;
  int boundary_is_fluid_surface, max_iters;
  int min_share_partners, min_merge_partners, allow_merge_optimal, allow_share_optimal, allow_share_too_small,
      allow_merge_size_diff, fail_on_missing_split_pattern;
  int n_planes;
  float planes[ASPH_MAX_PLANES][3];
};

template <class T> struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = n + n / 8 + 64;
    cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct asph_sim {
  int device = 0;
  cudaStream_t stream = nullptr;
  uint32_t n = 0, cap = 0;
  uint32_t cells_budget = 0;
  // persistent, double buffered for the per-step reorder
  DevBuf<float2> pos[2], vel[2];
  DevBuf<float> mass[2], level[2];
  DevBuf<uint32_t> refid[2];
  int cur = 0;
  // per step
  DevBuf<float4> xyhm, xyv, packP, packA, jc;
  DevBuf<float> h_unsorted, rho, aii, src, lam_sum, dens_err;
  DevBuf<float2> lam_grad, sumgrad, paccel;
  DevBuf<uint32_t> cellkey, cellcount, cellstart, order_tmp, order, scan_tmp;
  DevBuf<uint32_t> cnt_near, cnt_ext, slice_width, slice_base;
  DevBuf<uint32_t> nidx;
  DevBuf<float> ncoef;
  DevBuf<uint8_t> size_class, flag_surface, flag_insufficient;
  DevBuf<uint32_t> merge_partner, front[2], work_a, work_b, work_c;
  DevBuf<uint16_t> merge_counter;
  DevBuf<int> assigned;
  DevBuf<float> lut;        // 2 * 10001 floats: λ then λ′
  DevBuf<float> split_pos;  // flattened patterns
  DevBuf<int> split_off;
  DevBuf<float> blockstats; // per-block partials of the Jacobi reduction
  StepCtl* ctl = nullptr;   // device
  StepCtl* ctl_host = nullptr;  // pinned mirror
  PackedParams pp;
  int max_children = 0;
  asph_boundary boundary;
  bool lists_valid = false;
  float lists_factor = 0;
  bool have_level = false;
  // bookkeeping
  double time = 0;
  float time_f = 0;
  uint64_t step_number = 0;
  asph_step_info info;
  bool counters = false;
  double pc_ms[ASPH_PC_COUNT] = {0};
  uint64_t pc_calls[ASPH_PC_COUNT] = {0};
  cudaEvent_t ev[16];
  int sm_count = 148;
  std::string last_error;
  uint64_t kernel_launches = 0;
  // multi-GPU
  int rank = 0, n_ranks = 1;
  void* comm = nullptr;
};

// ------------------------------------------------------------------------------------------------
// host-side λ tables (plane_lambda.cpp)
// ------------------------------------------------------------------------------------------------
double asph_host_lambda(double d);
double asph_host_dlambda(double d);
void asph_host_build_luts(std::vector<float>& lam, std::vector<float>& dlam);
float asph_host_lut_get(const std::vector<float>& data, float x);

// ------------------------------------------------------------------------------------------------
// launchers (sim_core.cu / sim_adapt.cu)
// ------------------------------------------------------------------------------------------------
int launch_sort_and_grid(asph_sim* sim, float f_search);
int launch_neighbors(asph_sim* sim, float f_ext, float f_near, bool physics);
int launch_solver(asph_sim* sim, bool density_mode, float max_avg_error);
int launch_viscosity(asph_sim* sim);
int launch_source(asph_sim* sim, int kind);  // 0 divergence, 1 only density, 2 full
int launch_final_accel(asph_sim* sim, int mode);  // see sim_core.cu
int launch_level_estimation(asph_sim* sim);
int launch_level_smoothing(asph_sim* sim);
int launch_adaptivity(asph_sim* sim, float dt);
int launch_exclusive_scan(asph_sim* sim, uint32_t* data, uint32_t n, DevBuf<uint32_t>& tmp);

// ------------------------------------------------------------------------------------------------
// device math
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
#define ASPH_PI_F 3.14159265358979323846f
#define ASPH_FRAC_1_PI_F 0.318309886183790671538f

__device__ __forceinline__ unsigned int enc_f(float f) {
  unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_f(unsigned int e) {
  return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e);
}

// h = ETA * sqrt((m / rho0) / pi), simulation.rs:372-380 — IEEE ops in the reference's order (bit exact: the
// neighbour predicate depends on it)
__device__ __forceinline__ float h_from_mass(float m, float rho0) {
  return __fmul_rn(1.9f, __fsqrt_rn(__fmul_rn(__fdiv_rn(m, rho0), ASPH_FRAC_1_PI_F)));
}
// strict predicate |x_ij|^2 < ((h_i + h_j) * 0.5 * f)^2 without FMA contraction
// (neighborhood_search.rs:141-146, 63-67)
__device__ __forceinline__ float dist_sq_exact(float dx, float dy) {
  return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}
__device__ __forceinline__ float support_sq_exact(float hi, float hj, float f) {
  float s = __fmul_rn(__fmul_rn(__fadd_rn(hi, hj), 0.5f), f);
  return __fmul_rn(s, s);
}

// cubic spline, sph_kernels.rs:23-71; h = smoothing length (support 2h)
__device__ __forceinline__ float cubic_w(float q) {
  if (q < 0.5f) return 6.f * (q * q * q - q * q) + 1.f;
  if (q < 1.f) { float v = 1.f - q; return 2.f * (v * v * v); }
  return 0.f;
}
__device__ __forceinline__ float cubic_dw(float q) {
  if (q < 0.5f) return 18.f * q * q - 12.f * q;
  if (q < 1.f) { float v = 1.f - q; return -6.f * v * v; }
  return 0.f;
}
__device__ __forceinline__ float kernel_norm(float h) { return 10.f / (7.f * ASPH_PI_F * (h * h)); }
__device__ __forceinline__ float kernel_w(float r, float h) { return kernel_norm(h) * cubic_w(r / (2.f * h)); }
// dW/dr / r, i.e. gradW = kernel_dcoef * x_ij (zero when q <= 1e-5, sph_kernels.rs:64-66)
__device__ __forceinline__ float kernel_dcoef(float r, float h) {
  float q = r / (2.f * h);
  if (q <= 1.0e-5f) return 0.f;
  return kernel_norm(h) * cubic_dw(q) / (2.f * h) / r;
}
__device__ __forceinline__ float radius_to_volume(float r) { return ASPH_PI_F * r * r; }
__device__ __forceinline__ float volume_to_radius(float a) { return sqrtf(a * ASPH_FRAC_1_PI_F); }

// LevelEstimationState::target_mass simulation.rs:213-237
__device__ __forceinline__ float target_mass(float level, const PackedParams& P) {
  float lv = fmaxf(level, -P.maximum_surface_distance);
  float t = lv / -P.maximum_surface_distance;
  if (P.sizing == ASPH_SIZING_MASS) return P.mass_fine * (1.f - t) + P.mass_base * t;
  if (P.sizing == ASPH_SIZING_RADIUS) {
    float tr = P.particle_radius_fine * (1.f - t) + P.particle_radius_base * t;
    return radius_to_volume(tr) * P.rest_density;
  }
  float st = powf(t, 0.5f);
  float tr = P.particle_radius_fine * (1.f - st) + P.particle_radius_base * st;
  return radius_to_volume(tr) * P.rest_density;
}
// classify_particle adaptivity/mod.rs:32-48
__device__ __forceinline__ uint8_t classify_particle(float level, float mass, const PackedParams& P) {
  float mrel = mass / target_mass(level, P);
  if (mrel <= 0.5f) return ASPH_CLASS_TOO_SMALL;
  if (mrel <= 1.f / 1.1f) return ASPH_CLASS_SMALL;
  if (mrel < 1.1f) return ASPH_CLASS_OPTIMAL;
  if (mrel < 2.0f) return ASPH_CLASS_LARGE;
  return ASPH_CLASS_TOO_LARGE;
}
#endif
