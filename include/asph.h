/*
 * asph.h — C ABI of the B200-native SPH step loop (drop-in for the step path of kaegi/adaptive-sph).
 *
 * The reference has no FFI layer; its seam is the Rust API the three front ends call
 * (SURVEY.md §8b).  Every entry point below names the reference interface it replaces
 * (paths relative to /root/reference/src/simulation/).  Plain pointers and sizes only; no
 * torch/C++ types.  All host arrays are in REFERENCE PARTICLE ORDER (index i here == index i of
 * the reference's `ParticleVec`), independent of how the device sorts particles internally.
 *
 * Two libraries export exactly this ABI:
 *   adaptive-sph_b200/csrc/libasph_b200.so   the product (sm_100a CUDA kernels; needs a GPU to compute)
 *   oracle/liboracle_f32.so, liboracle_f64.so  the CPU restatement used ONLY as test oracle / CPU baseline
 */
#ifndef ASPH_H
#define ASPH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ status codes */
/* The reference panics (assert!/panic!) and the front end wraps the step in catch_unwind
 * (platform/desktop/main_loop.rs:295-311).  Here each panic site becomes a status code. */
enum {
  ASPH_OK = 0,
  ASPH_ERR_INVALID = 1,          /* bad argument / inconsistent params (simulation.rs:2020-2021 asserts)   */
  ASPH_ERR_UNSUPPORTED = 2,      /* a mode SURVEY.md §8 marks "next"/out of scope                          */
  ASPH_ERR_NONFINITE = 3,        /* assert!(is_finite) family (simulation.rs:391-395,1046,1106,1269-1281)  */
  ASPH_ERR_NEG_AII = 4,          /* "AII should not be negative" simulation.rs:1390-1403                   */
  ASPH_ERR_DENSITY = 5,          /* assert!(*p_density > 0.0001) simulation.rs:1047                        */
  ASPH_ERR_MASS_CONSERVATION = 6,/* |Σm before − Σm after| > 0.005 simulation.rs:2791-2792                 */
  ASPH_ERR_NEIGHBOR_OVERFLOW = 7,/* MAX_NEIGHBOR_COUNT = 20000 neighborhood_search.rs:3,149                */
  ASPH_ERR_CUDA = 8,
  ASPH_ERR_NCCL = 9,
  ASPH_ERR_CAPACITY = 10,        /* particle capacity exhausted by splitting                                */
  ASPH_ERR_NO_DEVICE = 11        /* CUDA library built but no usable GPU: the product never falls back to CPU */
};

/* ------------------------------------------------------------------ enums (simulation_parameters.rs:4-213) */
enum { ASPH_VISC_WCSPH = 0, ASPH_VISC_APPROX_LAPLACE = 1, ASPH_VISC_XSPH = 2 };
enum { ASPH_LEVEL_NONE = 0, ASPH_LEVEL_CENTER_DIFF = 1, ASPH_LEVEL_EMPTY_ANGLE = 2 };
enum { ASPH_NS_GRID = 0, ASPH_NS_RSTAR = 1 };
enum { ASPH_BOUNDARY_PARTICLES = 0, ASPH_BOUNDARY_ANALYTIC_UNDERESTIMATE = 1,
       ASPH_BOUNDARY_ANALYTIC_OVERESTIMATE = 2, ASPH_BOUNDARY_NONE = 3 };
enum { ASPH_H_FROM_DISTRIBUTION = 0, ASPH_H_FROM_DISTRIBUTION_CLAMPED1 = 1, ASPH_H_FROM_DISTRIBUTION_CLAMPED2 = 2,
       ASPH_H_FROM_DISTRIBUTION2 = 3, ASPH_H_FROM_MASS = 4 };
enum { ASPH_SOLVER_IISPH = 0, ASPH_SOLVER_IISPH2 = 1, ASPH_SOLVER_HYBRID_DFSPH = 2, ASPH_SOLVER_ONLY_DIVERGENCE = 3 };
enum { ASPH_SRC_DENSITY_AND_DIVERGENCE = 0, ASPH_SRC_ONLY_DENSITY = 1 };
enum { ASPH_PENALTY_NONE = 0, ASPH_PENALTY_LINEAR = 1, ASPH_PENALTY_QUADRATIC1 = 2, ASPH_PENALTY_QUADRATIC2 = 3 };
enum { ASPH_SIZING_RADIUS2 = 0, ASPH_SIZING_RADIUS = 1, ASPH_SIZING_MASS = 2 };
enum { ASPH_OP_CONSISTENT_SIMPLE_GRADIENT = 0, ASPH_OP_CONSISTENT_SYMMETRIC_GRADIENT = 1, ASPH_OP_WINCHENBACH2020 = 2 };
enum { ASPH_STASH_NONE = 0, ASPH_STASH_SURFACE_DISTANCE_FIRST_ITERATION = 1, ASPH_STASH_SURFACE_DISTANCE_MIDDLE = 2 };

/* ParticleSizeClass, adaptivity/mod.rs:11-23 */
enum { ASPH_CLASS_TOO_SMALL = 0, ASPH_CLASS_SMALL = 1, ASPH_CLASS_OPTIMAL = 2, ASPH_CLASS_LARGE = 3, ASPH_CLASS_TOO_LARGE = 4 };

/* ------------------------------------------------------------------ SimulationParams
 * Field names == YAML keys == fields of `SimulationParams` (simulation_parameters.rs:25-108).
 * Reals are carried as double (the YAML literal); the fp32 implementations round them to float once,
 * which is what serde does when it parses the literal into an f32.  Passed BY VALUE to every step:
 * the reference re-copies the params every step (platform/desktop/main_loop.rs:280). */
typedef struct asph_params {
  double rest_density, cfl_factor, max_dt, h;
  int32_t use_iisph;
  double viscosity;
  int32_t viscosity_type;
  double gravity;
  int32_t check_aii;
  int32_t level_estimation_method;
  double maximum_range;
  double jacobi_omega;
  double eos_stiffness;
  int32_t eos_power;
  int32_t neighborhood_search_algorithm;
  int32_t init_boundary_handler;
  int32_t support_length_estimation;
  double sdf_gradient_eps;
  int32_t fail_on_missing_split_pattern;
  int32_t has_pull_fluid_to;
  double pull_fluid_to[3];
  int32_t constrain_neighborhood_count;
  double particle_radius_fine, particle_radius_base, maximum_surface_distance;
  int32_t minimum_share_partners, minimum_merge_partners;
  int32_t merging, sharing, splitting;
  double max_mass_transfer_sharing, max_mass_transfer_merging, max_share_distance, max_merge_distance;
  int32_t allow_merge_with_optimal_particle, allow_share_with_optimal_particle,
          allow_share_with_too_small_particle, allow_merge_on_size_difference;
  int32_t boundary_is_fluid_surface, use_extended_range_for_level_estimation;
  int32_t pressure_solver_method;
  double iisph_max_avg_density_error, hybrid_dfsph_factor, hybrid_dfsph_max_avg_density_error,
         hybrid_dfsph_max_avg_divergence_error;
  int32_t hybrid_dfsph_density_source_term;
  int32_t hybrid_dfsph_non_pressure_accel_before_divergence_free;
  int32_t check_neighborhood;
  int32_t fill_stash_with;
  int32_t boundary_penalty_term;
  int32_t sizing_function;
  int32_t level_estimation_after_advection;
  double level_estimation_range;
  int32_t operator_discretization;
  int32_t operator_discretization_for_diagonal; /* -1 = None */
  int64_t max_iters;
} asph_params;

/* ------------------------------------------------------------------ boundary (simulation.rs:3137-3213)
 * kind PLANES: each plane is (nx, ny, delta), probe(x) = n·x + delta (sdf/sdf_plane.rs:13-38);
 * the host builds the 4 planes of `SdfPlane::new_boundary_box`.  kind POLYGON: closed polygon, the
 * `Sdf2D` of sdf/sdf2d.rs (AnalyticUnderestimate). */
enum { ASPH_BND_NONE = 0, ASPH_BND_PLANES = 1, ASPH_BND_POLYGON = 2 };
#define ASPH_MAX_PLANES 8
#define ASPH_MAX_POLY_VERTS 64
typedef struct asph_boundary {
  int32_t kind;
  int32_t n_planes;
  float planes[ASPH_MAX_PLANES][3];
  int32_t n_poly;
  float poly[ASPH_MAX_POLY_VERTS][2];
} asph_boundary;

/* ------------------------------------------------------------------ split patterns (adaptivity/splitting.rs:84-120)
 * Pattern for a 1→n split (n = 2..max_children) occupies pos_xy[2*offset[n-2] .. 2*(offset[n-2]+n)).
 * Only `pos_s` of split-patterns.yaml is used at run time. */
typedef struct asph_split_patterns {
  int32_t max_children;       /* 59 for the shipped file: get_max_num_children() = len + 1 */
  const int32_t* offset;      /* max_children - 1 entries */
  const float* pos_xy;
} asph_split_patterns;

/* ------------------------------------------------------------------ per-step report
 * What the reference prints / records in `ValueCounters` (simulation.rs:137-157,1990,2202,2543,2618). */
typedef struct asph_step_info {
  float dt;
  int32_t div_iterations;       /* `num_pressure_iters` of the divergence solve (sweeps executed − 1) */
  int32_t density_iterations;   /* same for the density solve                                         */
  int32_t div_sweeps, density_sweeps;
  int32_t level_sweeps;         /* sweeps of propagate_level_set_from_surface_detection (the CUDA path stops
                                   once a sweep assigned only values below -maximum_surface_distance, which
                                   the smoothing clamps anyway: fewer sweeps than the reference, same field) */
  int32_t n_shared, n_merged, n_split_parents;
  uint64_t n_particles_begin, n_particles_end;
  double last_avg_error_div, last_avg_error_density;
} asph_step_info;

/* PerformanceCounters labels (simulation.rs:1993,2023,2033,2517,2578,2734), accumulated milliseconds */
enum { ASPH_PC_SIMULATION_STEP = 0, ASPH_PC_NEIGHBORHOOD = 1, ASPH_PC_LEVEL_ESTIMATION = 2, ASPH_PC_DIV_SOLVER = 3,
       ASPH_PC_DENSITY_SOLVER = 4, ASPH_PC_ADAPTIVITY = 5, ASPH_PC_COUNT = 6 };

/* ------------------------------------------------------------------ fields for read-back
 * Replaces the public `particles`/`neighs` fields read by SimulationVisualizer::present
 * (simulation.rs:471-476, 2903-2913; colors.rs:380-492; vtk_exporter.rs:96-115). */
enum {
  ASPH_F_POSITION = 0,        /* float[2n]  */
  ASPH_F_VELOCITY = 1,        /* float[2n]  */
  ASPH_F_MASS = 2,            /* float[n]   */
  ASPH_F_H = 3,               /* float[n]   particles.h2 */
  ASPH_F_DENSITY = 4,         /* float[n]   */
  ASPH_F_PRESSURE = 5,        /* float[n]   */
  ASPH_F_AII = 6,             /* float[n]   */
  ASPH_F_SOURCE_TERM = 7,     /* float[n]   ppe_source_term */
  ASPH_F_PRESSURE_ACCEL = 8,  /* float[2n]  */
  ASPH_F_LEVEL = 9,           /* float[n]   FluidSurface(x) -> x (<= 0); FluidInterior -> ASPH_LEVEL_INTERIOR */
  ASPH_F_SIZE_CLASS = 10,     /* uint8[n]   ASPH_CLASS_* */
  ASPH_F_NEIGHBOR_COUNT = 11, /* uint32[n]  */
  ASPH_F_FLAG_SURFACE = 12,   /* uint8[n]   flag_is_fluid_surface */
  ASPH_F_FLAG_INSUFFICIENT = 13, /* uint8[n] flag_insufficient_neighs */
  ASPH_F_LAMBDA_SUM = 14,     /* float[n]   Σλ·penalty of the boundary handler (boundary_winchenbach2020.rs:47-49) */
  ASPH_F_LAMBDA_GRAD = 15,    /* float[2n]  Σ∇(λ·penalty) */
  ASPH_F_MERGE_PARTNER = 16,  /* uint32[n]  */
  ASPH_F_MERGE_COUNTER = 17,  /* uint16[n]  */
  ASPH_F_DENSITY_ERROR = 18,  /* float[n]   */
  ASPH_F_CONSTANT_FIELD = 19  /* float[n]   visualisation only (simulation.rs:2235) */
};
#define ASPH_LEVEL_INTERIOR 1.0f
#define ASPH_MERGE_PARTNER_AVAILABLE 0xFFFFFFFFu /* adaptivity/mod.rs:29 */
#define ASPH_MERGE_PARTNER_DELETE 0xFFFFFFFEu    /* adaptivity/mod.rs:30 */

typedef struct asph_sim asph_sim;

/* ------------------------------------------------------------------ construction
 * Replaces init_fluid_sim (simulation.rs:3074) / FluidSimulation::new (simulation.rs:487).
 * Scene YAML → arrays stays on the host side (loader replicating add_fluid_block, simulation.rs:2915).
 * `split` may be NULL when params.splitting is off.  `capacity` = max particles the handle can hold
 * (0 → 2·n + 1024). */
int asph_create(const asph_params* params, const float* pos_xy, const float* vel_xy, const float* mass, uint64_t n,
                const asph_boundary* boundary, const asph_split_patterns* split, int counters_enabled,
                uint64_t capacity, asph_sim** out);
void asph_destroy(asph_sim* sim);

/* Overwrite the persistent state (x, v, m) — the only state that survives between steps.
 * Used by the end-to-end path (host buffers in, host buffers out) and by parity tests. */
int asph_set_state(asph_sim* sim, const float* pos_xy, const float* vel_xy, const float* mass, uint64_t n);

/* ------------------------------------------------------------------ stepping
 * asph_step            = FluidSimulation::single_step                     simulation.rs:1973
 * asph_step_physics    = single_step_without_adaptivity (returns dt)      simulation.rs:1980
 * asph_step_adaptivity = single_step_adaptivity(params, dt)               simulation.rs:2732
 * The exporter (platform/desktop/animation/mod.rs:138-273) needs the split form. */
int asph_step(asph_sim* sim, const asph_params* params, float* dt_out);
int asph_step_physics(asph_sim* sim, const asph_params* params, float* dt_out);
int asph_step_adaptivity(asph_sim* sim, const asph_params* params, float dt);

/* ------------------------------------------------------------------ read-back */
uint64_t asph_num_particles(const asph_sim* sim);   /* FluidSimulation::num_fluid_particles simulation.rs:535 */
double asph_time(const asph_sim* sim);              /* pub time simulation.rs:475 */
uint64_t asph_step_number(const asph_sim* sim);     /* step_number simulation.rs:481 */
int asph_get_field(asph_sim* sim, int field, void* dst, uint64_t dst_bytes);
/* NeighborhoodCache (neighborhood_search.rs:13): CSR in reference indices, each row sorted ascending
 * (the reference's order is R*-tree traversal order and is not part of its contract).
 * offsets has n+1 entries.  If idx is NULL only *nnz_out is written. */
int asph_get_neighbors_csr(asph_sim* sim, uint64_t* offsets, uint32_t* idx, uint64_t idx_capacity, uint64_t* nnz_out);
/* Build the neighbour lists of the CURRENT state for range factor f (build_neighborhood_list,
 * neighborhood_search.rs:325) without stepping: h is refreshed from mass first (simulation.rs:1999-2003). */
int asph_build_neighbors(asph_sim* sim, const asph_params* params, float range_factor);
int asph_get_step_info(const asph_sim* sim, asph_step_info* out);
/* write_statistics (simulation.rs:3279): ms_sum/calls per PerformanceCounters label */
int asph_get_counters(const asph_sim* sim, double ms_sum[ASPH_PC_COUNT], uint64_t calls[ASPH_PC_COUNT]);
const char* asph_last_error(const asph_sim* sim);
const char* asph_backend_name(void);               /* "cuda-sm100a" | "oracle-f32" | "oracle-f64" */
/* number of GPU kernels this handle has launched so far (0 on the CPU oracle); bench.py reports it */
uint64_t asph_kernel_launches(const asph_sim* sim);

/* ------------------------------------------------------------------ kernel timing (measurement only)
 * CUDA-event durations, on the handle's own stream, of sampled launches of the hot kernels, accumulated since
 * asph_set_kernel_timing was last called.  sample_every = 0 switches it off; k > 0 times every k-th Jacobi sweep
 * (its pressure-acceleration pass and its update pass separately) and every neighbour build. */
enum { ASPH_KT_ACCEL_SWEEP = 0, ASPH_KT_JACOBI_SWEEP = 1, ASPH_KT_NEIGHBORS = 2, ASPH_KT_SORT_GRID = 3,
       ASPH_KT_LEVEL_PROPAGATE = 4,  /* the persistent level-set propagation kernel (one launch per step) */
       ASPH_KT_PARTNER_SEARCH = 5,   /* the persistent greedy partner-search kernel (one launch per share / merge phase) */
       ASPH_KT_COUNT = 6 };
int asph_set_kernel_timing(asph_sim* sim, int sample_every);
int asph_get_kernel_timing(asph_sim* sim, double ms_sum[ASPH_KT_COUNT], uint64_t samples[ASPH_KT_COUNT]);

/* ------------------------------------------------------------------ diagnostics (parity tests, restart files)
 * asph_set_level: prescribe the level field (reference particle order; value <= 0 = FluidSurface(value),
 * ASPH_LEVEL_INTERIOR = FluidInterior) so that single_step_adaptivity (simulation.rs:2732) can be driven on its own;
 * asph_set_step_number: the step counter that decides merge (even) / split (odd) steps (simulation.rs:2725, 2760);
 * asph_adapt_rounds: dependency rounds the greedy partner searches of the last resampling phase took (0 on the CPU
 * oracle, whose searches are the reference's serial loops). */
int asph_set_level(asph_sim* sim, const float* level_ref_order, uint64_t n);
void asph_set_step_number(asph_sim* sim, uint64_t step_number);
uint64_t asph_adapt_rounds(const asph_sim* sim);
/* ready-list entries of the partner searches whose donor had been decided already, over the handle's lifetime (a
 * consistency counter of the CUDA implementation; expected 0, always 0 on the CPU oracle) */
uint64_t asph_debug_greedy_duplicates(const asph_sim* sim);

/* ------------------------------------------------------------------ pure helpers (host side, no GPU needed)
 * sph_kernels.rs:49-71 (cubic spline, h = smoothing length, support 2h) and
 * boundary_handler/sdf_boundary_handler/plane_numerics.rs:19-152 (λ, λ′ in double). */
float asph_kernel_w(float r, float h);
void asph_kernel_grad(float dx, float dy, float h, float* gx, float* gy);
double asph_lambda(double d);
double asph_dlambda(double d);
/* LookupTable1D::get over the 10001-entry tables (lookup_table.rs:32-48) */
float asph_lambda_lut(float d);
float asph_dlambda_lut(float d);

/* ------------------------------------------------------------------ multi-GPU
 * One process per GPU.  Slab decomposition along x; `asph_create_distributed` takes this rank's share
 * of a global scene (any subset; particles are migrated to their owner slab at the first step).
 * The 128-byte NCCL unique id is produced on rank 0 by asph_comm_unique_id and broadcast by the host
 * (torch.distributed in the Python harness). */
int asph_comm_unique_id(uint8_t id_out[128]);
int asph_create_distributed(const asph_params* params, const float* pos_xy, const float* vel_xy, const float* mass,
                            const uint32_t* global_index, uint64_t n_local, uint64_t n_global,
                            const asph_boundary* boundary, const asph_split_patterns* split, int counters_enabled,
                            uint64_t capacity, const uint8_t nccl_id[128], int rank, int n_ranks, int device,
                            asph_sim** out);
/* global index of each locally owned particle, for assembling read-backs across ranks */
int asph_get_global_index(asph_sim* sim, uint32_t* dst, uint64_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* ASPH_H */
