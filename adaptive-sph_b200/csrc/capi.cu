// capi.cu — the C ABI of include/asph.h on top of the CUDA kernels: handle lifetime, the step orchestration of
// FluidSimulation::single_step (simulation.rs:1973-2796) and read-back in reference particle order.
//
// There is no CPU path here: without a usable CUDA device asph_create returns ASPH_ERR_NO_DEVICE.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <type_traits>

#include "lists.cuh"


namespace {

constexpr int kThreads = 256;

int pack_params(asph_sim* sim, const asph_params* p) {
  PackedParams& q = sim->pp;
  memset(&q, 0, sizeof(q));
  q.rest_density = float(p->rest_density); q.cfl_factor = float(p->cfl_factor); q.max_dt = float(p->max_dt);
  q.viscosity = float(p->viscosity); q.gravity = float(p->gravity); q.jacobi_omega = float(p->jacobi_omega);
  q.sdf_gradient_eps = float(p->sdf_gradient_eps);
  q.particle_radius_fine = float(p->particle_radius_fine); q.particle_radius_base = float(p->particle_radius_base);
  q.maximum_surface_distance = float(p->maximum_surface_distance);
  // mass_fine / mass_base: radius_to_volume(r) * rest_density in fp32 (simulation_parameters.rs:124-131)
  const float PI_F = 3.14159265358979323846f;
  q.mass_fine = ((PI_F * q.particle_radius_fine) * q.particle_radius_fine) * q.rest_density;
  q.mass_base = ((PI_F * q.particle_radius_base) * q.particle_radius_base) * q.rest_density;
  q.max_mass_transfer_sharing = float(p->max_mass_transfer_sharing);
  q.max_share_distance = float(p->max_share_distance); q.max_merge_distance = float(p->max_merge_distance);
  q.max_avg_density_error_iisph = float(p->iisph_max_avg_density_error); q.hybrid_factor = float(p->hybrid_dfsph_factor);
  q.max_avg_density_error = float(p->hybrid_dfsph_max_avg_density_error);
  q.max_avg_divergence_error = float(p->hybrid_dfsph_max_avg_divergence_error);
  q.f_ext = float(p->level_estimation_range) / 1.9f;  // level_estimation_range / ETA in FT, simulation.rs:2028
  q.f_near = 2.f;
  q.pull_x = float(p->pull_fluid_to[0]); q.pull_y = float(p->pull_fluid_to[1]);
  q.has_pull = p->has_pull_fluid_to; q.viscosity_type = p->viscosity_type; q.level_method = p->level_estimation_method;
  q.solver = p->pressure_solver_method; q.density_source = p->hybrid_dfsph_density_source_term;
  q.np_before_div = p->hybrid_dfsph_non_pressure_accel_before_divergence_free; q.penalty = p->boundary_penalty_term;
  q.sizing = p->sizing_function; q.opdisc = p->operator_discretization;
  q.self_last = 1;
  q.constrain = p->constrain_neighborhood_count ? 1 : 0;
  q.h_mode = p->support_length_estimation;
  q.level_cut = (q.h_mode == ASPH_H_FROM_DISTRIBUTION || q.h_mode == ASPH_H_FROM_DISTRIBUTION2) ? float(p->maximum_range) : 0.f;
  q.boundary_is_fluid_surface = p->boundary_is_fluid_surface;
  q.max_iters = int(std::min<int64_t>(p->max_iters, 1 << 28));
  q.min_share_partners = p->minimum_share_partners; q.min_merge_partners = p->minimum_merge_partners;
  q.allow_merge_optimal = p->allow_merge_with_optimal_particle; q.allow_share_optimal = p->allow_share_with_optimal_particle;
  q.allow_share_too_small = p->allow_share_with_too_small_particle; q.allow_merge_size_diff = p->allow_merge_on_size_difference;
  q.fail_on_missing_split_pattern = p->fail_on_missing_split_pattern;
  q.n_planes = sim->boundary.kind == ASPH_BND_PLANES ? sim->boundary.n_planes : 0;
  for (int s = 0; s < q.n_planes; s++) for (int k = 0; k < 3; k++) q.planes[s][k] = sim->boundary.planes[s][k];
  // Sdf2DConnectedComponents::from_points (sdf/sdf2d.rs:37-71): normalised edge directions and vertex pseudo-normals, fp32
  q.n_poly = 0;
  if (sim->boundary.kind == ASPH_BND_POLYGON) {
    const int np = sim->boundary.n_poly;
    if (np < 3 || np > ASPH_POLY_DEV) { sim->last_error = "polygon boundary: 3..64 vertices"; return ASPH_ERR_INVALID; }
    q.n_poly = np;
    for (int i = 0; i < np; i++) { q.poly_pt[i][0] = sim->boundary.poly[i][0]; q.poly_pt[i][1] = sim->boundary.poly[i][1]; }
    for (int i = 0; i < np; i++) {
      const int j = (i + 1) % np;
      volatile float dx = q.poly_pt[j][0] - q.poly_pt[i][0], dy = q.poly_pt[j][1] - q.poly_pt[i][1];
      volatile float xx = dx * dx, yy = dy * dy;  // volatile: no contraction on the host either
      volatile float n2 = xx + yy;
      if (!(n2 > 0.00001f)) { sim->last_error = "polygon boundary: degenerate edge"; return ASPH_ERR_INVALID; }
      const float n = std::sqrt(n2);
      q.poly_dir[i][0] = dx / n; q.poly_dir[i][1] = dy / n;
    }
    for (int i = 0; i < np; i++) {
      const int a = i == 0 ? np - 1 : i - 1;
      q.poly_pn[i][0] = -q.poly_dir[a][1] + -q.poly_dir[i][1];
      q.poly_pn[i][1] = q.poly_dir[a][0] + q.poly_dir[i][0];
    }
  }

  // A refused mode must not leak into the calls that go on after ASPH_ERR_UNSUPPORTED (asph_create, asph_build_neighbors
  // only need the neighbour pass): the packed copy falls back to the plain variants of everything mode dependent.
  auto unsupported = [&](const char* what) {
    sim->last_error = std::string(what) + " is not implemented yet (SURVEY.md §8f)";
    q.h_mode = ASPH_H_FROM_MASS; q.level_cut = 0.f; q.opdisc = ASPH_OP_CONSISTENT_SIMPLE_GRADIENT; q.solver = ASPH_SOLVER_HYBRID_DFSPH;
    return ASPH_ERR_UNSUPPORTED;
  };
  // constrain_neighborhood_count rewrites h in the middle of the step: single GPU, h from the masses
  if (p->constrain_neighborhood_count && (sim->dist || p->support_length_estimation != ASPH_H_FROM_MASS)) {
    q.constrain = 0;
    return unsupported("constrain_neighborhood_count across GPU slabs or with support_length_estimation != FromMass");
  }
  // single-GPU only so far: the per-particle state these modes carry from step to step does not migrate between slabs
  if (p->support_length_estimation != ASPH_H_FROM_MASS && sim->dist) return unsupported("support_length_estimation != FromMass across GPU slabs");
  if (p->pressure_solver_method == ASPH_SOLVER_IISPH2 && sim->dist) return unsupported("pressure_solver_method IISPH2 across GPU slabs");
  if (p->viscosity_type == ASPH_VISC_XSPH) return unsupported("viscosity_type XSPH (todo!() in the reference)");
  // level_estimation_after_advection (simulation.rs:2678-2707) sorts and searches a second time in the middle of the step
  if (p->level_estimation_after_advection && p->level_estimation_method != ASPH_LEVEL_NONE) {
    if (sim->dist) return unsupported("level_estimation_after_advection across GPU slabs");
    if (!p->use_extended_range_for_level_estimation) return unsupported("level_estimation_after_advection without use_extended_range_for_level_estimation");
    if (p->support_length_estimation != ASPH_H_FROM_MASS) return unsupported("level_estimation_after_advection with support_length_estimation != FromMass");
    if (p->constrain_neighborhood_count) return unsupported("level_estimation_after_advection with constrain_neighborhood_count");
    // the reference's resampling then walks the extended-range lists; here it walks N_2, the same thing as long as no
    // partner can be further away than the support
    if ((p->sharing && p->max_share_distance > 2.0) || (p->merging && p->max_merge_distance > 2.0))
      return unsupported("level_estimation_after_advection with a share / merge distance beyond the kernel support");
  }
  return ASPH_OK;
}

// ---- PerformanceCounters (simulation.rs:159-189): CUDA-event intervals per label ----------------------------
struct PcInterval { int label; cudaEvent_t b, e; bool counts_call; bool ended; };
struct PcState { std::vector<PcInterval> open; std::vector<cudaEvent_t> pool; };
PcState& pc_of(asph_sim* sim) {  // owned by the handle (asph_sim::pc), freed by asph_destroy
  if (!sim->pc) sim->pc = new PcState();
  return *static_cast<PcState*>(sim->pc);
}
void pc_destroy(asph_sim* sim) {
  if (!sim->pc) return;
  PcState* s = static_cast<PcState*>(sim->pc);
  for (auto& iv : s->open) { cudaEventDestroy(iv.b); cudaEventDestroy(iv.e); }
  for (cudaEvent_t e : s->pool) cudaEventDestroy(e);
  delete s;
  sim->pc = nullptr;
}
// an error return leaves intervals open whose end was never recorded: drop them instead of timing them later
void pc_discard(asph_sim* sim) {
  if (!sim->counters || !sim->pc) return;
  PcState& s = pc_of(sim);
  cudaStreamSynchronize(sim->stream);
  for (auto& iv : s.open) { s.pool.push_back(iv.b); s.pool.push_back(iv.e); }
  s.open.clear();
  cudaGetLastError();
}
cudaEvent_t pc_event(PcState& s) {
  if (!s.pool.empty()) { cudaEvent_t e = s.pool.back(); s.pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
void pc_begin(asph_sim* sim, int label, bool counts_call = true) {
  if (!sim->counters) return;
  PcState& s = pc_of(sim);
  PcInterval iv{label, pc_event(s), pc_event(s), counts_call, false};
  cudaEventRecord(iv.b, sim->stream);
  s.open.push_back(iv);
}
void pc_end(asph_sim* sim, int label) {
  if (!sim->counters) return;
  PcState& s = pc_of(sim);
  for (int k = int(s.open.size()) - 1; k >= 0; k--)
    if (s.open[k].label == label && !s.open[k].ended) { cudaEventRecord(s.open[k].e, sim->stream); s.open[k].ended = true; return; }
}
void pc_collect(asph_sim* sim) {
  if (!sim->counters) return;
  PcState& s = pc_of(sim);
  cudaStreamSynchronize(sim->stream);
  for (auto& iv : s.open) {
    float ms = 0.f;
    if (iv.ended && cudaEventElapsedTime(&ms, iv.b, iv.e) == cudaSuccess) {
      sim->pc_ms[iv.label] += ms;
      if (iv.counts_call) sim->pc_calls[iv.label]++;
    }
    s.pool.push_back(iv.b); s.pool.push_back(iv.e);
  }
  s.open.clear();
  cudaGetLastError();
}

__global__ void k_init_ids(uint32_t n, uint32_t* refid, float* level) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { refid[i] = i; level[i] = ASPH_LEVEL_INTERIOR; }
}

// field extraction into reference order: out[refid[i] * comps + c]
__global__ void k_extract(uint32_t n, int field, const uint32_t* __restrict__ refid, const float2* pos, const float2* vel,
                          const float* mass, const float4* xyhm, const float* rho, const float4* packP, const float4* pconst,
                          const float4* packA, const float* level, const uint8_t* size_class, const uint32_t* cnt,
                          const uint8_t* flags, const float* lam_sum, const float2* lam_grad, const uint32_t* merge_partner,
                          const uint32_t* merge_counter, void* out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t r = refid[i];
  if (r == 0xFFFFFFFFu) return;  // multi-GPU read-back map: ghost particle
  float* f = (float*)out;
  switch (field) {
    case ASPH_F_POSITION: f[2 * r] = pos[i].x; f[2 * r + 1] = pos[i].y; break;
    case ASPH_F_VELOCITY: f[2 * r] = vel[i].x; f[2 * r + 1] = vel[i].y; break;
    case ASPH_F_MASS: f[r] = mass[i]; break;
    case ASPH_F_H: f[r] = xyhm[i].z; break;
    case ASPH_F_DENSITY: f[r] = rho[i]; break;
    case ASPH_F_PRESSURE: f[r] = packP[i].w; break;
    case ASPH_F_AII: f[r] = pconst[i].z; break;
    case ASPH_F_SOURCE_TERM: f[r] = pconst[i].w; break;
    case ASPH_F_PRESSURE_ACCEL: f[2 * r] = packA[i].z; f[2 * r + 1] = packA[i].w; break;
    case ASPH_F_LEVEL: f[r] = level[i]; break;
    case ASPH_F_SIZE_CLASS: ((uint8_t*)out)[r] = size_class[i]; break;
    case ASPH_F_NEIGHBOR_COUNT: ((uint32_t*)out)[r] = nb_cn(cnt[i]); break;
    case ASPH_F_FLAG_SURFACE: ((uint8_t*)out)[r] = flags[i] & 1u; break;
    case ASPH_F_FLAG_INSUFFICIENT: ((uint8_t*)out)[r] = (flags[i] >> 1) & 1u; break;
    case ASPH_F_LAMBDA_SUM: f[r] = lam_sum[i]; break;
    case ASPH_F_LAMBDA_GRAD: f[2 * r] = lam_grad[i].x; f[2 * r + 1] = lam_grad[i].y; break;
    case ASPH_F_MERGE_PARTNER: {
      // partner indices are device indices internally; report reference indices
      uint32_t p = merge_partner[i];
      ((uint32_t*)out)[r] = (p >= ASPH_MERGE_PARTNER_DELETE) ? p : refid[p];
      break;
    }
    case ASPH_F_MERGE_COUNTER: ((uint16_t*)out)[r] = uint16_t(merge_counter[i]); break;
  }
}

// multi-GPU asph_set_state: host arrays are in the handle's read-back order; map[i] = slot of sorted particle i
template <class T>
__global__ void k_scatter_state(uint32_t n, const uint32_t* __restrict__ map, const T* __restrict__ src, T* __restrict__ dst) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t r = map[i];
  if (r != 0xFFFFFFFFu) dst[i] = src[r];
}

int field_elem_bytes(int field, int* comps) {
  *comps = 1;
  switch (field) {
    case ASPH_F_POSITION: case ASPH_F_VELOCITY: case ASPH_F_PRESSURE_ACCEL: case ASPH_F_LAMBDA_GRAD: *comps = 2; return 4;
    case ASPH_F_SIZE_CLASS: case ASPH_F_FLAG_SURFACE: case ASPH_F_FLAG_INSUFFICIENT: return 1;
    case ASPH_F_MERGE_COUNTER: return 2;
    case ASPH_F_NEIGHBOR_COUNT: case ASPH_F_MERGE_PARTNER: return 4;
    case ASPH_F_MASS: case ASPH_F_H: case ASPH_F_DENSITY: case ASPH_F_PRESSURE: case ASPH_F_AII: case ASPH_F_SOURCE_TERM:
    case ASPH_F_LEVEL: case ASPH_F_LAMBDA_SUM: return 4;
  }
  return 0;
}

}  // namespace

cudaEvent_t kt_event(asph_sim* sim) {
  if (!sim->kt_pool.empty()) { cudaEvent_t e = sim->kt_pool.back(); sim->kt_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
void kt_release(asph_sim* sim, cudaEvent_t e) { sim->kt_pool.push_back(e); }

int check_error_flags(asph_sim* sim) {
  const unsigned int f = sim->ctl_host->error_flags;
  if (!f) return ASPH_OK;
  if (f & ERRF_PEER_TIMEOUT) { sim->last_error = "timed out waiting for a neighbour GPU's data (a rank of the job has failed or stopped stepping)"; return ASPH_ERR_NCCL; }
  if (f & ERRF_CELL_BUDGET) { sim->last_error = "cell grid does not fit the cell budget: particle positions are spread over an absurd extent (the simulation has exploded?)"; return ASPH_ERR_CAPACITY; }
  if (f & ERRF_NEIGHBOR_OVERFLOW) { sim->last_error = "exceeded maximum allowed number of 20000 neighbors"; return ASPH_ERR_NEIGHBOR_OVERFLOW; }
  if (f & ERRF_NONFINITE) { sim->last_error = "assert!(is_finite) failed (density / a_ii / position / velocity)"; return ASPH_ERR_NONFINITE; }
  if (f & ERRF_DENSITY) { sim->last_error = "assert!(*p_density > 0.0001) failed"; return ASPH_ERR_DENSITY; }
  if (f & ERRF_NEG_AII) { sim->last_error = "AII should not be negative!"; return ASPH_ERR_NEG_AII; }
  if (f & ERRF_SOLVER_NONFINITE) { sim->last_error = "'!a_p.is_finite()' failed. Pressure values probably have exploded!"; return ASPH_ERR_NONFINITE; }
  if (f & ERRF_LEVEL_WEIGHT) { sim->last_error = "smooth_level_estimation_field: weight <= 0"; return ASPH_ERR_NONFINITE; }
  if (f & ERRF_CONSTRAIN) { sim->last_error = "constrain_neighborhood_count: assert!(*p_h_next < h) / assert!(*p_h_next >= 0.) failed"; return ASPH_ERR_INVALID; }
  if (f & ERRF_PARTICLE_CAPACITY) { sim->last_error = "particle capacity exhausted by splitting"; return ASPH_ERR_CAPACITY; }
  if (f & ERRF_SPLIT_PATTERN) { sim->last_error = "no split pattern for a 1-to-n split"; return ASPH_ERR_INVALID; }
  if (f & ERRF_SPLIT_CHILDREN) { sim->last_error = "assert!(num_children > 1)"; return ASPH_ERR_INVALID; }
  if (f & ERRF_PARTNER_VALIDATION) {
    const StepCtl& c = *sim->ctl_host;
    sim->last_error = "validate_share_partners / validate_merge_partners failed (invariant " + std::to_string(c.validate_why) + " at particle " +
                      std::to_string(c.validate_at[0]) + ": counter " + std::to_string(c.validate_at[1]) + ", partner " + std::to_string(c.validate_at[2]) +
                      ", found " + std::to_string(c.validate_at[3]) + "; rounds " + std::to_string(c.rounds) + ", n " + std::to_string(sim->n) + ")";
    return ASPH_ERR_INVALID;
  }
  sim->last_error = "device error flag " + std::to_string(f);
  return ASPH_ERR_INVALID;
}

// One attempt at the physics part of the step from the neighbour pass on.  Returns ASPH_RETRY_LISTS when the first
// synchronisation shows that the neighbour pool was too small (nothing irreversible has happened by then).
static int physics_after_sort(asph_sim* sim, bool lvl, float f_ext, bool smooth) {
  const PackedParams& P = sim->pp;
  cudaEvent_t kt1 = nullptr, kt2 = nullptr;
  if (sim->kt_every > 0) { kt1 = kt_event(sim); kt2 = kt_event(sim); cudaEventRecord(kt1, sim->stream); }
  TRY(launch_neighbors(sim, f_ext, P.f_near));
  if (kt2) cudaEventRecord(kt2, sim->stream);
  pc_end(sim, ASPH_PC_NEIGHBORHOOD);
  auto finish_kt = [&]() {
    if (!kt2) return;
    float b = 0.f;
    if (cudaEventSynchronize(kt2) == cudaSuccess && cudaEventElapsedTime(&b, kt1, kt2) == cudaSuccess) {
      sim->kt_ms[ASPH_KT_NEIGHBORS] += b; sim->kt_samples[ASPH_KT_NEIGHBORS]++;
    }
    kt_release(sim, kt1); kt_release(sim, kt2);
    kt1 = kt2 = nullptr;
  };
  int rc = ASPH_OK;
  auto guard = [&](int r) { if (r != ASPH_OK && rc == ASPH_OK) rc = r; return r == ASPH_OK; };
  do {
    if (lvl) {
      pc_begin(sim, ASPH_PC_LEVEL_ESTIMATION);
      if (!guard(launch_level_estimation(sim))) break;
      pc_end(sim, ASPH_PC_LEVEL_ESTIMATION);
      TRY(check_error_flags(sim));  // density / a_ii asserts of the neighbour pass (simulation.rs:1046-1047, 1390)
    }
    if (P.constrain) {  // simulation.rs:2145-2177, after the level estimation (simulation.rs:2018-2057) and before everything that uses h
      if (!lvl) {  // the lists the new lengths are chosen from must be complete before h is overwritten
        if (!guard(sync_ctl(sim))) break;
        if (sim->ctl_host->error_flags & ERRF_LIST_CAPACITY) { rc = ASPH_RETRY_LISTS; break; }
        if (!guard(check_error_flags(sim))) break;
      }
      if (!guard(launch_constrain_neighborhood(sim))) break;
      pc_begin(sim, ASPH_PC_NEIGHBORHOOD, false);
      if (!guard(launch_neighbors(sim, P.f_near, P.f_near))) break;
      pc_end(sim, ASPH_PC_NEIGHBORHOOD);
    }
    int iters = 0, sweeps = 0;
    bool first = !lvl || P.constrain;  // the first synchronisation after the neighbour pass also validates its error flags
    auto solve = [&](bool density, float tol, double* avg) {
      int r = launch_solver(sim, density, tol, &iters, &sweeps, avg);
      if (r == ASPH_OK && first) { r = check_error_flags(sim); first = false; }
      return r;
    };
    switch (P.solver) {
      case ASPH_SOLVER_IISPH2:  // simulation.rs:2262-2387
        if (!guard(launch_omega(sim)) || !guard(launch_viscosity(sim)) || !guard(launch_source(sim, 3))) break;
        pc_begin(sim, ASPH_PC_DENSITY_SOLVER);
        if (!guard(solve(true, P.max_avg_density_error_iisph, &sim->info.last_avg_error_density))) break;
        sim->info.density_iterations = iters; sim->info.density_sweeps = sweeps;
        if (!guard(launch_scale_pressure(sim)) || !guard(launch_final_accel(sim, 3))) break;
        pc_end(sim, ASPH_PC_DENSITY_SOLVER);
        break;
      case ASPH_SOLVER_IISPH:  // simulation.rs:2389-2446
        if (!guard(launch_viscosity(sim)) || !guard(launch_source(sim, 2))) break;
        pc_begin(sim, ASPH_PC_DENSITY_SOLVER);
        if (!guard(solve(true, P.max_avg_density_error_iisph, &sim->info.last_avg_error_density))) break;
        sim->info.density_iterations = iters; sim->info.density_sweeps = sweeps;
        if (!guard(launch_final_accel(sim, 3))) break;
        pc_end(sim, ASPH_PC_DENSITY_SOLVER);
        break;
      case ASPH_SOLVER_ONLY_DIVERGENCE:  // simulation.rs:2448-2500
        if (!guard(launch_viscosity(sim)) || !guard(launch_source(sim, 0))) break;
        pc_begin(sim, ASPH_PC_DIV_SOLVER);
        if (!guard(solve(false, P.max_avg_divergence_error, &sim->info.last_avg_error_div))) break;
        sim->info.div_iterations = iters; sim->info.div_sweeps = sweeps;
        if (!guard(launch_final_accel(sim, 3))) break;
        pc_end(sim, ASPH_PC_DIV_SOLVER);
        break;
      default:  // HybridDFSPH, simulation.rs:2502-2670
        if (P.np_before_div && !guard(launch_viscosity(sim))) break;
        pc_begin(sim, ASPH_PC_DIV_SOLVER);
        if (!guard(launch_source(sim, 0))) break;
        if (!guard(solve(false, P.max_avg_divergence_error, &sim->info.last_avg_error_div))) break;
        sim->info.div_iterations = iters; sim->info.div_sweeps = sweeps;
        if (!guard(launch_final_accel(sim, 1))) break;
        pc_end(sim, ASPH_PC_DIV_SOLVER);
        if (!P.np_before_div && !guard(launch_viscosity(sim))) break;
        pc_begin(sim, ASPH_PC_DENSITY_SOLVER);
        if (!guard(launch_source(sim, P.density_source == ASPH_SRC_ONLY_DENSITY ? 1 : 2))) break;
        if (!guard(solve(true, P.max_avg_density_error, &sim->info.last_avg_error_density))) break;
        sim->info.density_iterations = iters; sim->info.density_sweeps = sweeps;
        if (!guard(launch_final_accel(sim, 2))) break;
        pc_end(sim, ASPH_PC_DENSITY_SOLVER);
        break;
    }
    if (rc != ASPH_OK) break;
    if (smooth) {
      pc_begin(sim, ASPH_PC_LEVEL_ESTIMATION, false);
      if (!guard(launch_level_smoothing(sim))) break;
      pc_end(sim, ASPH_PC_LEVEL_ESTIMATION);
    }
    if (!guard(sync_ctl(sim))) break;
    guard(check_error_flags(sim));
  } while (false);
  finish_kt();
  return rc;
}

template <class T>
__global__ void k_gather(uint32_t n, const uint32_t* __restrict__ order, const T* __restrict__ in, T* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[order[i]];
}
__global__ void k_restore_dt(StepCtl* ctl, float dt) { ctl->dt = dt; }

// level_estimation_after_advection (simulation.rs:2678-2720): the advected particles are sorted and searched again with
// the extended range, the level set is estimated and smoothed on those lists, and the resampling phase that follows uses
// them too.  The per-step fields of the physics part (density — K17 reads it —, pressure, a_ii, ...) follow the
// particles through the second sort; the second neighbour pass writes lists and surface normals only.
static int level_after_advection(asph_sim* sim) {
  const PackedParams& P = sim->pp;
  const uint32_t n = sim->n;
  if (n == 0) return ASPH_OK;
  const uint32_t blocks = (n + 255u) / 256u;
  cudaStream_t st = sim->stream;
  const float dt = sim->ctl_host->dt;  // the step's dt: the second sort would compute one from the new velocities
  const unsigned int sweeps_parity = sim->p_cur;
  TRY(launch_sort_and_grid(sim, std::max(P.f_ext, P.f_near)));
  void* tmp = sim->xv[1].p;  // free between the sort (which fills xv[0]) and the next step
  auto follow = [&](auto* arr) -> int {
    using T = std::remove_pointer_t<decltype(arr)>;
    k_gather<T><<<blocks, 256, 0, st>>>(n, sim->order.p, arr, static_cast<T*>(tmp));
    LAUNCH_CHECK();
    CUDA_TRY(cudaMemcpyAsync(arr, tmp, size_t(n) * sizeof(T), cudaMemcpyDeviceToDevice, st));
    return ASPH_OK;
  };
  TRY(follow(sim->rho.p));
  TRY(follow(sim->packP[sweeps_parity].p));
  TRY(follow(sim->pconst.p));
  TRY(follow(sim->packA.p));
  TRY(follow(sim->lam_sum.p));
  TRY(follow(sim->lam_grad.p));
  for (int attempt = 0;; attempt++) {
    TRY(launch_neighbors(sim, P.f_ext, P.f_near, true));
    TRY(sync_ctl(sim));
    if (!(sim->ctl_host->error_flags & ERRF_LIST_CAPACITY)) break;
    if (attempt == 5) { sim->last_error = "neighbour list pool could not be sized"; return ASPH_ERR_CAPACITY; }
    TRY(neighbors_grow(sim));
  }
  TRY(check_error_flags(sim));
  k_restore_dt<<<1, 1, 0, st>>>(sim->ctl, dt);
  LAUNCH_CHECK();
  pc_begin(sim, ASPH_PC_LEVEL_ESTIMATION);
  TRY(launch_level_estimation(sim));
  pc_end(sim, ASPH_PC_LEVEL_ESTIMATION);
  pc_begin(sim, ASPH_PC_LEVEL_ESTIMATION, false);
  TRY(launch_level_smoothing(sim));
  pc_end(sim, ASPH_PC_LEVEL_ESTIMATION);
  TRY(sync_ctl(sim));
  return check_error_flags(sim);
}

static int step_physics(asph_sim* sim, const asph_params* params, float* dt_out) {
  TRY(pack_params(sim, params));
  const PackedParams& P = sim->pp;
  memset(&sim->info, 0, sizeof(sim->info));
  sim->info.n_particles_begin = sim->n;
  const bool after = P.level_method != ASPH_LEVEL_NONE && params->level_estimation_after_advection;
  const bool lvl = P.level_method != ASPH_LEVEL_NONE && !after;  // level estimation as the first thing in the step
  if (lvl && !params->use_extended_range_for_level_estimation) {  // simulation.rs:2020
    sim->last_error = "level estimation before advection needs use_extended_range_for_level_estimation";
    return ASPH_ERR_INVALID;
  }
  if (lvl && P.level_method == ASPH_LEVEL_CENTER_DIFF) {  // simulation.rs:2021
    sim->last_error = "center diff level estimation method needs density values which are not available when performing level estimation as first step in loop";
    return ASPH_ERR_INVALID;
  }
  pc_begin(sim, ASPH_PC_SIMULATION_STEP);
  // With no level estimation the extended-range lists are never read (perform_level_estimation is a no-op,
  // simulation.rs:2034-2057), so only N_2 is built.
  const float f_ext = lvl ? P.f_ext : P.f_near;
  if (sim->dist) {
    // Ghost zone: one support of the widest lists — two pair supports (4 h_max) when the step ends with share / merge: an
    // owned donor must see every donor that can touch its touch set (adapt.cu, k_greedy)
    float f_ghost = std::max(f_ext, P.f_near);
    if (lvl && (params->sharing || params->merging)) f_ghost = std::max(f_ghost, 4.0f);
    int rc = dist_begin_step(sim, f_ghost);  // migration + ghost exchange; sim->n = owned + ghosts
    if (rc != ASPH_OK) { pc_collect(sim); return rc; }
    sim->info.n_particles_begin = sim->n_owned;
    if (sim->n_owned == 0) {  // every collective below assumes all ranks take part
      sim->last_error = "a rank owns no particles (more GPUs than the particle set can be split over)";
      pc_collect(sim);
      return ASPH_ERR_INVALID;
    }
  }
  if (sim->n == 0) {
    sim->info.dt = P.max_dt;
  } else {
    pc_begin(sim, ASPH_PC_NEIGHBORHOOD);
    cudaEvent_t kt0 = nullptr, kt1 = nullptr;
    if (sim->kt_every > 0) { kt0 = kt_event(sim); kt1 = kt_event(sim); cudaEventRecord(kt0, sim->stream); }
    { const int rc0 = launch_sort_and_grid(sim, std::max(f_ext, P.f_near)); if (rc0 != ASPH_OK) { pc_discard(sim); return rc0; } }
    if (kt1) cudaEventRecord(kt1, sim->stream);
    int rc = ASPH_OK;
    for (int attempt = 0; attempt < 6; attempt++) {
      sim->xv_cur = 0;
      rc = physics_after_sort(sim, lvl, f_ext, lvl);
      if (rc != ASPH_RETRY_LISTS) break;
      { const int rc0 = neighbors_grow(sim); if (rc0 != ASPH_OK) { pc_discard(sim); return rc0; } }
      pc_begin(sim, ASPH_PC_NEIGHBORHOOD, false);
    }
    if (kt1) {
      float a = 0.f;
      if (cudaEventSynchronize(kt1) == cudaSuccess && cudaEventElapsedTime(&a, kt0, kt1) == cudaSuccess) {
        sim->kt_ms[ASPH_KT_SORT_GRID] += a; sim->kt_samples[ASPH_KT_SORT_GRID]++;
      }
      kt_release(sim, kt0); kt_release(sim, kt1);
    }
    if (rc == ASPH_RETRY_LISTS) { sim->last_error = "neighbour list pool could not be sized"; rc = ASPH_ERR_CAPACITY; }
    if (rc == ASPH_OK && after) rc = level_after_advection(sim);
    if (rc != ASPH_OK) { pc_collect(sim); return rc; }
    sim->info.dt = sim->ctl_host->dt;
  }
  const float dt = sim->info.dt;
  sim->last_dt = dt;
  sim->time_f += dt;  // time: FT in the reference (simulation.rs:2724)
  sim->time = double(sim->time_f);
  sim->step_number += 1;
  sim->step_fields_valid = true;
  pc_end(sim, ASPH_PC_SIMULATION_STEP);
  pc_collect(sim);
  if (dt_out) *dt_out = dt;
  sim->info.n_particles_end = sim->dist ? sim->n_owned : sim->n;
  return ASPH_OK;
}

static int step_adaptivity(asph_sim* sim, const asph_params* params, float dt) {
  TRY(pack_params(sim, params));
  const bool any = params->sharing || params->merging || params->splitting;
  if (any && params->level_estimation_method == ASPH_LEVEL_NONE) {
    sim->last_error = "resampling needs a level estimation (level() on FluidInterior is unreachable!, simulation.rs:204-211)";
    return ASPH_ERR_INVALID;
  }
  if (!any) return ASPH_OK;
  if (sim->dist && dist_ranks(sim) > 1 && !dist_p2p(sim)) { sim->last_error = "resampling across GPU slabs needs the peer-memory path (ASPH_DIST_P2P)"; return ASPH_ERR_UNSUPPORTED; }
  if (!sim->lists_valid || !sim->level_valid) {
    sim->last_error = "single_step_adaptivity needs the neighbour lists and level field of the preceding physics step";
    return ASPH_ERR_INVALID;
  }
  pc_begin(sim, ASPH_PC_SIMULATION_STEP, false);
  pc_begin(sim, ASPH_PC_ADAPTIVITY);
  sim->share_enabled = params->sharing != 0; sim->merge_enabled = params->merging != 0; sim->split_enabled = params->splitting != 0;
  int rc = launch_adaptivity(sim, dt);
  pc_end(sim, ASPH_PC_ADAPTIVITY);
  pc_end(sim, ASPH_PC_SIMULATION_STEP);
  pc_collect(sim);
  sim->info.n_particles_end = sim->dist ? sim->n_owned : sim->n;
  return rc;
}

extern "C" {

const char* asph_backend_name(void) { return "cuda-sm100a"; }

int asph_set_state(asph_sim* sim, const float* pos, const float* vel, const float* mass, uint64_t n) {
  if (!sim || (n && (!pos || !vel || !mass))) return ASPH_ERR_INVALID;
  if (n > 0x7FFFFFF0ull) return ASPH_ERR_CAPACITY;  // bit 31 of the particle index marks ghosts (multi-GPU)
  if (sim->dist) {
    // this rank's owned particles, in the order asph_get_field reports them; identities (global indices) are kept
    if (n != sim->n_owned) { sim->last_error = "asph_set_state on a distributed handle: n must equal the owned particle count"; return ASPH_ERR_INVALID; }
    CUDA_TRY(cudaSetDevice(sim->device));
    if (n == 0) return ASPH_OK;
    TRY(dist_local_map(sim));
    const int c = sim->cur;
    const uint32_t na = sim->n, blocks = (na + kThreads - 1) / kThreads;
    const uint32_t* map = sim->scratch_u[3].p;
    CUDA_TRY(cudaMemcpyAsync(sim->scratch_f.p, pos, n * sizeof(float2), cudaMemcpyHostToDevice, sim->stream));
    k_scatter_state<float2><<<blocks, kThreads, 0, sim->stream>>>(na, map, (const float2*)sim->scratch_f.p, sim->pos[c].p);
    LAUNCH_CHECK();
    CUDA_TRY(cudaMemcpyAsync(sim->scratch_f.p, vel, n * sizeof(float2), cudaMemcpyHostToDevice, sim->stream));
    k_scatter_state<float2><<<blocks, kThreads, 0, sim->stream>>>(na, map, (const float2*)sim->scratch_f.p, sim->vel[c].p);
    LAUNCH_CHECK();
    CUDA_TRY(cudaMemcpyAsync(sim->scratch_f.p, mass, n * sizeof(float), cudaMemcpyHostToDevice, sim->stream));
    k_scatter_state<float><<<blocks, kThreads, 0, sim->stream>>>(na, map, (const float*)sim->scratch_f.p, sim->mass[c].p);
    LAUNCH_CHECK();
    CUDA_TRY(cudaStreamSynchronize(sim->stream));
    sim->lists_valid = false; sim->level_valid = false; sim->step_fields_valid = false;
    return ASPH_OK;
  }
  CUDA_TRY(cudaSetDevice(sim->device));
  if (n > sim->cap) { sim->n = 0; TRY(ensure_capacity(sim, uint32_t(n + n / 4 + 1024))); }
  sim->n = uint32_t(n); sim->n_owned = uint32_t(n);
  const int c = sim->cur;
  if (n) {
    CUDA_TRY(cudaMemcpyAsync(sim->pos[c].p, pos, n * sizeof(float2), cudaMemcpyHostToDevice, sim->stream));
    CUDA_TRY(cudaMemcpyAsync(sim->vel[c].p, vel, n * sizeof(float2), cudaMemcpyHostToDevice, sim->stream));
    CUDA_TRY(cudaMemcpyAsync(sim->mass[c].p, mass, n * sizeof(float), cudaMemcpyHostToDevice, sim->stream));
    k_init_ids<<<(uint32_t(n) + kThreads - 1) / kThreads, kThreads, 0, sim->stream>>>(uint32_t(n), sim->refid[c].p, sim->level[c].p);
    LAUNCH_CHECK();
  }
  CUDA_TRY(cudaStreamSynchronize(sim->stream));
  sim->lists_valid = false; sim->level_valid = false; sim->step_fields_valid = false;
  sim->hdist_valid = false;  // h2_next restarts from the masses, as in FluidSimulation::new (simulation.rs:505-520)
  sim->cls_valid = false;    // every particle Optimal (ParticleVec default)
  return ASPH_OK;
}

int asph_create(const asph_params* params, const float* pos, const float* vel, const float* mass, uint64_t n,
                const asph_boundary* boundary, const asph_split_patterns* split, int counters_enabled, uint64_t capacity,
                asph_sim** out) {
  if (!params || !out || (n && (!pos || !vel || !mass))) return ASPH_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return ASPH_ERR_NO_DEVICE; }
  int device = 0;
  if (const char* e = getenv("ASPH_DEVICE")) device = atoi(e);
  else if (const char* e2 = getenv("LOCAL_RANK")) device = atoi(e2) % ndev;
  if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return ASPH_ERR_NO_DEVICE; }
  asph_sim* sim = new asph_sim();
  sim->device = device;
  auto fail = [&](int rc) { asph_destroy(sim); return rc; };
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(ASPH_ERR_NO_DEVICE);
  sim->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&sim->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(ASPH_ERR_CUDA);
  if (cudaMalloc((void**)&sim->ctl, sizeof(StepCtl)) != cudaSuccess) return fail(ASPH_ERR_CUDA);
  if (cudaMemset(sim->ctl, 0, sizeof(StepCtl)) != cudaSuccess) return fail(ASPH_ERR_CUDA);
  if (cudaMallocHost((void**)&sim->ctl_host, sizeof(StepCtl)) != cudaSuccess) return fail(ASPH_ERR_CUDA);
  memset(sim->ctl_host, 0, sizeof(StepCtl));
  sim->counters = counters_enabled != 0;
  if (const char* e = getenv("ASPH_BULK")) sim->bulk = atoi(e) != 0;
  if (const char* e = getenv("ASPH_CELL_SCALE")) sim->cell_scale = std::min(1.f, std::max(0.25f, float(atof(e))));
  if (boundary) sim->boundary = *boundary; else memset(&sim->boundary, 0, sizeof(sim->boundary));
  {  // λ / λ′ lookup tables (BoundaryWinchenbach2020::new, boundary_winchenbach2020.rs:33-45)
    std::vector<float> lam, dlam;
    asph_host_build_luts(lam, dlam);
    if (sim->lut.ensure(2 * 10001) != cudaSuccess) return fail(ASPH_ERR_CUDA);
    cudaMemcpy(sim->lut.p, lam.data(), 10001 * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemcpy(sim->lut.p + 10001, dlam.data(), 10001 * sizeof(float), cudaMemcpyHostToDevice);
  }
  if (split && split->max_children >= 2) {
    sim->max_children = split->max_children;
    const int total = split->offset[split->max_children - 2] + split->max_children;
    if (sim->split_off.ensure(split->max_children - 1) != cudaSuccess || sim->split_pos.ensure(2 * size_t(total)) != cudaSuccess)
      return fail(ASPH_ERR_CUDA);
    cudaMemcpy(sim->split_off.p, split->offset, (split->max_children - 1) * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(sim->split_pos.p, split->pos_xy, 2 * size_t(total) * sizeof(float), cudaMemcpyHostToDevice);
  }
  uint64_t cap = capacity ? capacity : (2 * n + 1024);
  if (cap < n) cap = n;
  if (cap > 0x7FFFFFF0ull) return fail(ASPH_ERR_CAPACITY);
  int rc = ensure_capacity(sim, uint32_t(cap));
  if (rc != ASPH_OK) return fail(rc);
  rc = pack_params(sim, params);
  if (rc != ASPH_OK && rc != ASPH_ERR_UNSUPPORTED) return fail(rc);
  rc = asph_set_state(sim, pos, vel, mass, n);
  if (rc != ASPH_OK) return fail(rc);
  memset(&sim->info, 0, sizeof(sim->info));
  *out = sim;
  return ASPH_OK;
}

void asph_destroy(asph_sim* sim) {
  if (!sim) return;
  cudaSetDevice(sim->device);
  if (sim->stream) cudaStreamSynchronize(sim->stream);
  dist_destroy(sim);
  pc_destroy(sim);
  for (int b = 0; b < 2; b++) {
    sim->pos[b].release(); sim->vel[b].release(); sim->mass[b].release(); sim->level[b].release(); sim->refid[b].release();
    sim->xv[b].release(); sim->packP[b].release(); sim->front[b].release(); sim->work[b].release();
  }
  sim->xyhm.release(); sim->packA.release(); sim->pconst.release(); sim->h_tmp.release(); sim->rho.release(); sim->lam_sum.release();
  sim->nrm.release(); sim->gB.release(); sim->lam_grad.release(); sim->key.release(); sim->cellcount.release(); sim->cellstart.release();
  sim->order.release(); sim->scan_sums.release(); sim->cnt.release(); sim->cnt_ext.release(); sim->far_idx.release(); sim->far_cnt.release(); sim->slice_base.release();
  for (int b = 0; b < 2; b++) { sim->hnext[b].release(); sim->lamprev[b].release(); sim->cls[b].release(); }
  sim->omega.release();
  sim->nbpool.release(); sim->hm.release(); sim->hv.release(); sim->size_class.release(); sim->flags.release(); sim->merge_partner.release();
  sim->cand.release(); for (int k = 0; k < 4; k++) sim->scratch_u[k].release();
  sim->merge_counter.release(); sim->stamp.release(); sim->g_info.release(); sim->g_head.release(); sim->g_next.release(); sim->g_resume.release(); sim->g_drop.release(); sim->scratch_f.release(); sim->lut.release(); sim->split_pos.release();
  sim->split_off.release(); sim->blockstats.release();
  if (sim->ctl) cudaFree(sim->ctl);
  if (sim->ctl_host) cudaFreeHost(sim->ctl_host);
  for (cudaEvent_t e : sim->kt_pool) cudaEventDestroy(e);
  if (sim->stream) cudaStreamDestroy(sim->stream);
  cudaGetLastError();
  delete sim;
}

int asph_step_physics(asph_sim* sim, const asph_params* params, float* dt_out) {
  if (!sim || !params) return ASPH_ERR_INVALID;
  CUDA_TRY(cudaSetDevice(sim->device));
  return step_physics(sim, params, dt_out);
}
int asph_step_adaptivity(asph_sim* sim, const asph_params* params, float dt) {
  if (!sim || !params) return ASPH_ERR_INVALID;
  CUDA_TRY(cudaSetDevice(sim->device));
  return step_adaptivity(sim, params, dt);
}
int asph_step(asph_sim* sim, const asph_params* params, float* dt_out) {  // simulation.rs:1973-1978
  if (!sim || !params) return ASPH_ERR_INVALID;
  CUDA_TRY(cudaSetDevice(sim->device));
  float dt = 0.f;
  TRY(step_physics(sim, params, &dt));
  if (dt_out) *dt_out = dt;
  return step_adaptivity(sim, params, dt);
}

uint64_t asph_num_particles(const asph_sim* sim) { return sim->dist ? sim->n_owned : sim->n; }
double asph_time(const asph_sim* sim) { return sim->time; }
uint64_t asph_step_number(const asph_sim* sim) { return sim->step_number; }

int asph_get_field(asph_sim* sim, int field, void* dst, uint64_t bytes) {
  if (!sim || !dst) return ASPH_ERR_INVALID;
  CUDA_TRY(cudaSetDevice(sim->device));
  int comps = 1;
  const int eb = field_elem_bytes(field, &comps);
  if (eb == 0) { sim->last_error = "field not available from the CUDA backend"; return ASPH_ERR_UNSUPPORTED; }
  const uint32_t n_out = sim->dist ? sim->n_owned : sim->n;  // ghosts are not reported
  const uint32_t n = sim->n;
  const uint64_t need = uint64_t(n_out) * comps * eb;
  if (bytes < need) return ASPH_ERR_INVALID;
  const bool persistent = field == ASPH_F_POSITION || field == ASPH_F_VELOCITY || field == ASPH_F_MASS || field == ASPH_F_LEVEL;
  if (!persistent && !sim->step_fields_valid) {
    sim->last_error = "per-step field requested before a physics step (or after resampling changed the particle set)";
    return ASPH_ERR_INVALID;
  }
  if (n_out == 0) return ASPH_OK;
  if (sim->dist) TRY(dist_local_map(sim));  // multi-GPU: owned particles in local (sorted) order; asph_get_global_index names them
  // scatter into reference order in scratch (scratch_f holds 2 * cap floats)
  void* tmp = sim->scratch_f.p;
  const int c = sim->cur;
  k_extract<<<(n + kThreads - 1) / kThreads, kThreads, 0, sim->stream>>>(
      n, field, sim->dist ? sim->scratch_u[3].p : sim->refid[c].p, sim->pos[c].p, sim->vel[c].p, sim->mass[c].p, sim->xyhm.p, sim->rho.p,
      sim->packP[sim->p_cur].p, sim->pconst.p, sim->packA.p, sim->level[c].p, sim->size_class.p, sim->cnt.p, sim->flags.p,
      sim->lam_sum.p, sim->lam_grad.p, sim->merge_partner.p, sim->merge_counter.p, tmp);
  LAUNCH_CHECK();
  CUDA_TRY(cudaMemcpyAsync(dst, tmp, need, cudaMemcpyDeviceToHost, sim->stream));
  CUDA_TRY(cudaStreamSynchronize(sim->stream));
  return ASPH_OK;
}

int asph_get_neighbors_csr(asph_sim* sim, uint64_t* offsets, uint32_t* idx, uint64_t cap, uint64_t* nnz_out) {
  if (!sim) return ASPH_ERR_INVALID;
  CUDA_TRY(cudaSetDevice(sim->device));
  if (!sim->lists_valid) { sim->last_error = "no neighbour lists (call asph_build_neighbors or step first)"; return ASPH_ERR_INVALID; }
  if (sim->dist) { sim->last_error = "neighbour read-back on a distributed handle"; return ASPH_ERR_UNSUPPORTED; }
  const uint32_t n = sim->n;
  std::vector<uint32_t> cnt(n), refid(n), sbase((n + 31) / 32);
  TRY(sync_ctl(sim));
  const size_t used = size_t(sim->ctl_host->list_used) * 64;  // uint16 units
  std::vector<uint16_t> pool(used);
  std::vector<uint32_t> far(size_t((n + ASPH_PAIR_BLOCK - 1) / ASPH_PAIR_BLOCK) * ASPH_PAIR_FAR);
  if (n) {
    CUDA_TRY(cudaMemcpy(cnt.data(), sim->cnt.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(refid.data(), sim->refid[sim->cur].p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(sbase.data(), sim->slice_base.p, sbase.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (used) CUDA_TRY(cudaMemcpy(pool.data(), sim->nbpool.p, used * sizeof(uint16_t), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(far.data(), sim->far_idx.p, far.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  }
  uint64_t nnz = 0;
  for (uint32_t i = 0; i < n; i++) nnz += nb_cn(cnt[i]);
  if (nnz_out) *nnz_out = nnz;
  if (!idx) return ASPH_OK;
  if (cap < nnz || !offsets) return ASPH_ERR_INVALID;
  std::vector<uint32_t> where(n);  // reference index -> device index
  for (uint32_t i = 0; i < n; i++) where[refid[i]] = i;
  uint64_t o = 0;
  for (uint32_t r = 0; r < n; r++) {
    const uint32_t i = where[r];
    offsets[r] = o;
    const uint32_t cw = nb_cw(cnt[i]), cf = nb_cf(cnt[i]), c = cw + cf;
    const uint32_t sb = sbase[i >> 5];
    const bool wide = (sb >> 31) != 0;
    const size_t base = size_t(sb & 0x7fffffffu) * 64;
    for (uint32_t k = 0; k < c; k++) idx[o + k] = refid[nb_get(pool.data() + base, far.data(), wide, i, k, cw, cf)];
    std::sort(idx + o, idx + o + c);
    o += c;
  }
  offsets[n] = o;
  return ASPH_OK;
}

int asph_build_neighbors(asph_sim* sim, const asph_params* params, float f) {
  if (!sim || !params) return ASPH_ERR_INVALID;
  CUDA_TRY(cudaSetDevice(sim->device));
  int rc = pack_params(sim, params);
  if (rc != ASPH_OK && rc != ASPH_ERR_UNSUPPORTED) return rc;
  if (sim->dist) { sim->last_error = "asph_build_neighbors on a distributed handle"; return ASPH_ERR_UNSUPPORTED; }
  if (sim->n == 0) { sim->lists_valid = true; return ASPH_OK; }
  TRY(launch_sort_and_grid(sim, f));
  for (int attempt = 0;; attempt++) {
    TRY(launch_neighbors(sim, f, f));
    TRY(sync_ctl(sim));
    if (!(sim->ctl_host->error_flags & ERRF_LIST_CAPACITY)) break;
    if (attempt >= 6) { sim->last_error = "neighbour list pool could not be sized"; return ASPH_ERR_CAPACITY; }
    TRY(neighbors_grow(sim));
  }
  const unsigned int fl = sim->ctl_host->error_flags;
  if (fl & (ERRF_NEIGHBOR_OVERFLOW | ERRF_CELL_BUDGET)) return check_error_flags(sim);
  sim->step_fields_valid = true;
  return ASPH_OK;
}

int asph_get_step_info(const asph_sim* sim, asph_step_info* out) { *out = sim->info; return ASPH_OK; }
int asph_get_counters(const asph_sim* sim, double ms[ASPH_PC_COUNT], uint64_t calls[ASPH_PC_COUNT]) {
  for (int i = 0; i < ASPH_PC_COUNT; i++) { ms[i] = sim->pc_ms[i]; calls[i] = sim->pc_calls[i]; }
  return ASPH_OK;
}
const char* asph_last_error(const asph_sim* sim) { return sim ? sim->last_error.c_str() : ""; }

// ---- pure helpers (host code) ------------------------------------------------------------------------------
// cubic_kernel_2d / cubic_kernel_2d_deriv, sph_kernels.rs:49-71
float asph_kernel_w(float r, float h) {
  const float nf = 10.f / (7.f * 3.14159265358979323846f * (h * h));
  const float q = r / (2.f * h);
  float w;
  if (q < 0.5f) w = 6.f * (q * q * q - q * q) + 1.f;
  else if (q < 1.f) { float v = 1.f - q; w = 2.f * (v * v * v); }
  else w = 0.f;
  return nf * w;
}
void asph_kernel_grad(float dx, float dy, float h, float* gx, float* gy) {
  const float r = std::sqrt(dx * dx + dy * dy);
  const float q = r / (2.f * h);
  if (q <= 1.0e-5f) { *gx = 0.f; *gy = 0.f; return; }
  const float nf = 10.f / (7.f * 3.14159265358979323846f * (h * h));
  float dw;
  if (q < 0.5f) dw = 18.f * q * q - 12.f * q;
  else if (q < 1.f) { float v = 1.f - q; dw = -6.f * v * v; }
  else dw = 0.f;
  const float s = nf * dw / (2.f * h);
  *gx = s * (dx / r); *gy = s * (dy / r);
}
double asph_lambda(double d) { return asph_host_lambda(d); }
double asph_dlambda(double d) { return asph_host_dlambda(d); }
static std::vector<float>& host_lut(int which) {
  static std::vector<float> lam, dlam;
  if (lam.empty()) asph_host_build_luts(lam, dlam);
  return which ? dlam : lam;
}
float asph_lambda_lut(float d) { return asph_host_lut_get(host_lut(0), d); }
float asph_dlambda_lut(float d) { return asph_host_lut_get(host_lut(1), d); }

// ---- diagnostics used by the parity tests: drive single_step_adaptivity from a prescribed level field / step parity
int asph_set_level(asph_sim* sim, const float* level_ref_order, uint64_t n) {
  // several GPUs: the array covers the whole fluid (indexed by reference index); this rank picks its particles' and its ghosts' values
  if (!sim || !level_ref_order || n != (sim->dist ? dist_n_global(sim) : uint64_t(sim->n))) return ASPH_ERR_INVALID;
  CUDA_TRY(cudaSetDevice(sim->device));
  const uint64_t nl = sim->n;
  std::vector<uint32_t> refid(nl);
  std::vector<float> lv(nl);
  if (nl) {
    CUDA_TRY(cudaMemcpy(refid.data(), sim->refid[sim->cur].p, nl * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (uint64_t i = 0; i < nl; i++) {
      const uint32_t r = refid[i] & ~ASPH_GHOST_BIT;
      if (r >= n) return ASPH_ERR_INVALID;
      lv[i] = level_ref_order[r];
    }
    CUDA_TRY(cudaMemcpy(sim->level[sim->cur].p, lv.data(), nl * sizeof(float), cudaMemcpyHostToDevice));
  }
  sim->level_valid = true;
  return ASPH_OK;
}
void asph_set_step_number(asph_sim* sim, uint64_t k) { if (sim) sim->step_number = k; }
uint64_t asph_adapt_rounds(const asph_sim* sim) { return sim ? sim->adapt_rounds : 0; }
uint64_t asph_debug_greedy_duplicates(const asph_sim* sim) { return sim ? sim->greedy_duplicates : 0; }

int asph_set_kernel_timing(asph_sim* sim, int sample_every) {
  if (!sim) return ASPH_ERR_INVALID;
  sim->kt_every = sample_every > 0 ? sample_every : 0;
  for (int k = 0; k < ASPH_KT_COUNT; k++) { sim->kt_ms[k] = 0; sim->kt_samples[k] = 0; }
  return ASPH_OK;
}
int asph_get_kernel_timing(asph_sim* sim, double ms_sum[ASPH_KT_COUNT], uint64_t samples[ASPH_KT_COUNT]) {
  if (!sim) return ASPH_ERR_INVALID;
  for (int k = 0; k < ASPH_KT_COUNT; k++) { ms_sum[k] = sim->kt_ms[k]; samples[k] = sim->kt_samples[k]; }
  return ASPH_OK;
}

// kernels launched by this handle so far (bench.py reports it as gpu_launches)
uint64_t asph_kernel_launches(const asph_sim* sim) { return sim ? sim->kernel_launches : 0; }

}  // extern "C"
