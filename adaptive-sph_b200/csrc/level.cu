// level.cu — level-set (distance to the free surface) estimation on the extended-range lists.
//
//   K3  surface_detection_by_empty_angle          simulation.rs:539-625
//   K4  propagate_level_set_from_surface_detection simulation.rs:729-801
//   K17 smooth_level_estimation_field             simulation.rs:804-857
//
// K4 in the reference re-scans all N particles every sweep until nothing changes (O(N * sweeps)).  A particle's
// value is final in the first sweep in which any neighbour already has one, i.e. in sweep t = its BFS distance from
// the detected surface, and it only reads values of the previous sweep — which, the lists being symmetric, are exactly
// the values of the particles assigned in sweep t - 1.  So here each sweep touches only that front, and the whole
// propagation is ONE persistent cooperative kernel (k_propagate): in sweep t every particle j of front(t - 1) pushes
// φ_j − |x_ij| into its still unassigned neighbours i with an integer atomic (all values are <= 0, so the float maximum is
// the unsigned minimum of the bit patterns); the first push claims i for front(t).  One grid-wide barrier per sweep, no
// host round trip, no launch.  max is order independent and the distance is computed without contraction, so the
// field is bit-identical to the reference's Jacobi sweeps.
// `level` encoding: value <= 0 = FluidSurface(value); ASPH_LEVEL_INTERIOR (1.0) = FluidInterior.
#include <cooperative_groups.h>

#include "lists.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kThreads = 256;
constexpr int kPropThreads = 512;                 // block of the persistent propagation kernel
constexpr unsigned int kUnassigned = 0xFFFFFFFFu;  // bit pattern of a level value no push has reached yet

typedef NbLists Lists;

// front_n[0] = tail of the one front array (every particle enters it once; the fronts of successive sweeps are
// consecutive segments); level_live[0] = number of the last sweep that assigned a value above the cutoff
__global__ void k_level_reset(StepCtl* ctl) {
  ctl->front_n[0] = 0; ctl->front_n[1] = 0; ctl->cand_n[0] = 0; ctl->cand_n[1] = 0;
  ctl->level_live[0] = 0; ctl->level_live[1] = 0;
  ctl->level_sweep = 0; ctl->level_done = 0;
}

// K3 for particle i; returns whether it is a surface particle
__device__ __forceinline__ bool surface_of(uint32_t i, const Lists& L, const float4* __restrict__ xyhm, const float2* __restrict__ nrm,
                                           const PackedParams& P, float cos_threshold, float* __restrict__ level, int* __restrict__ stamp,
                                           uint8_t* __restrict__ flags) {
  const float4 me = xyhm[i];
  const uint32_t ce = L.cnt_ext[i];
  const float2 nr = nrm[i];
  const float s = -(me.w / P.rest_density);
  float nx = s * nr.x, ny = s * nr.y;
  bool interior;
  uint8_t fl = 0;
  if (ce < 3u) {  // D * 2 - 1
    interior = false;
    fl |= 2u;
  } else if (nx * nx + ny * ny < 0.00001f) {
    interior = true;
  } else {
    float dmin = __int_as_float(0x7f800000);
    for (int p = 0; p < sdf_count(P); p++) dmin = fminf(dmin, sdf_probe(P, p, me.x, me.y));
    if (!P.boundary_is_fluid_surface && dmin < me.z * 1.5f) {
      interior = true;
    } else {
      interior = false;
      const float nn = sqrtf(nx * nx + ny * ny);
      nx /= nn; ny /= nn;
      const NbCol col(L, i);
      // is_neighbor_in_level_estimation_range (simulation.rs:698-723): FromDistribution / FromDistribution2 only
      const float cut = P.level_cut * sqrtf((me.w / P.rest_density) * ASPH_FRAC_1_PI_F);
      const float cut2 = P.level_cut > 0.f ? cut * cut : __int_as_float(0x7f800000);
      for (uint32_t k = 0; k < ce; k++) {
        const uint32_t j = col.get(k);
        const float4 o = __ldg(&xyhm[j]);
        const float dx = o.x - me.x, dy = o.y - me.y;
        if (dx * dx + dy * dy > cut2) continue;
        const float inv = 1.f / (sqrtf(dx * dx + dy * dy) + 0.000001f);
        if ((dx * inv) * nx + (dy * inv) * ny > cos_threshold) { interior = true; break; }
      }
    }
  }
  if (!interior) {
    fl |= 1u;
    level[i] = 0.f;
    stamp[i] = 0;
  } else {
    level[i] = __uint_as_float(kUnassigned);  // k_propagate turns what no push reaches into ASPH_LEVEL_INTERIOR
    stamp[i] = -1;
  }
  flags[i] = fl;
  return !interior;
}
__global__ void __launch_bounds__(kThreads)
k_surface(uint32_t n, Lists L, const float4* __restrict__ xyhm, const float2* __restrict__ nrm, const PackedParams P, float cos_threshold,
          float* __restrict__ level, int* __restrict__ stamp, uint8_t* __restrict__ flags, uint32_t* __restrict__ front, StepCtl* ctl) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool surf = i < n && surface_of(i, L, xyhm, nrm, P, cos_threshold, level, stamp, flags);
  // front(0): one atomic per warp
  const unsigned int mask = __ballot_sync(0xffffffffu, surf);
  if (!mask) return;
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t base = 0;
  if (lane == 0) base = atomicAdd(&ctl->front_n[0], uint32_t(__popc(mask)));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (surf) front[base + uint32_t(__popc(mask & ((1u << lane) - 1u)))] = i;
}

// K4 (simulation.rs:739-800) as one persistent cooperative kernel.  One warp per front particle j, one lane per
// neighbour i; sweep t reads front(t - 1) = front[begin, end) and appends front(t) behind it.
//   stamp[i]: -1 unassigned, otherwise the sweep that assigned i (0 = detected surface)
// Values other SMs wrote in earlier sweeps (stamp, level, the front entries) are read with ld.global.cg: an L1 line
// fetched in an earlier sweep may hold their previous contents.
__global__ void __launch_bounds__(kPropThreads)
k_propagate(uint32_t n, Lists L, const float4* __restrict__ xyhm, float* __restrict__ level, int* __restrict__ stamp,
            uint32_t* __restrict__ front, StepCtl* ctl, float neg_dmax, int use_cutoff) {
  cg::grid_group grid = cg::this_grid();
  unsigned int* level_bits = reinterpret_cast<unsigned int*>(level);
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  volatile uint32_t* tail = &ctl->front_n[0];
  volatile int* live_sweep = &ctl->level_live[0];
  uint32_t begin = 0, end = *tail;
  int sweeps = 0;
  for (int t = 1; t <= int(n) + 1; t++) {
    // the sweep runs if the previous one assigned anything (above the cutoff); the reference's last sweep changes nothing
    if (end == begin || (t > 1 && *live_sweep != t - 1)) break;
    sweeps = t;
    bool live = false;
    for (uint32_t f = begin + gwarp; f < end; f += nwarps) {
      const uint32_t j = __ldcg(front + f);
      const float4 me = __ldg(&xyhm[j]);
      const float lj = __ldcg(level + j);
      const uint32_t ce = __ldg(&L.cnt_ext[j]);
      const NbCol col(L, j);
      for (uint32_t k0 = 0; k0 < ce; k0 += 32u) {
        const uint32_t k = k0 + lane;
        bool won = false;
        uint32_t i = 0;
        if (k < ce) {
          i = col.get(k);
          const int s = __ldcg(stamp + i);
          if (s == -1 || s == t) {
            const float4 o = __ldg(&xyhm[i]);
            const float d = __fsqrt_rn(dist_sq_exact(__fsub_rn(me.x, o.x), __fsub_rn(me.y, o.y)));
            const float v = __fsub_rn(lj, d);  // <= 0: the largest float is the smallest bit pattern
            atomicMin(level_bits + i, __float_as_uint(v));
            if (s == -1) won = atomicCAS(stamp + i, -1, t) == -1;
            if (!use_cutoff || v > neg_dmax) live = true;
          }
        }
        const unsigned int mask = __ballot_sync(0xffffffffu, won);
        if (mask) {
          uint32_t base = 0;
          if (lane == 0) base = atomicAdd(&ctl->front_n[0], uint32_t(__popc(mask)));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (won) front[base + uint32_t(__popc(mask & ((1u << lane) - 1u)))] = i;
        }
      }
    }
    if (__any_sync(0xffffffffu, live) && lane == 0) *live_sweep = t;
    grid.sync();
    begin = end;
    end = *tail;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { ctl->level_sweep = sweeps; ctl->level_done = 1; }
  // particles no push has reached stay FluidInterior
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (__ldcg(level_bits + i) == kUnassigned) level[i] = ASPH_LEVEL_INTERIOR;
}

// K17: φ_i = Σ φ~_j V_j W_ij / Σ V_j W_ij with post-advection positions, pre-advection lists / densities
__global__ void __launch_bounds__(kThreads)
k_smooth(uint32_t n, Lists L, const float2* __restrict__ pos, const float4* __restrict__ xyhm, const float* __restrict__ rho, const float* __restrict__ level,
         float* __restrict__ level_out, float dmax, StepCtl* ctl) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float2 xi = pos[i];
  const float hi = xyhm[i].z;
  const uint32_t cn = nb_cn(L.cnt[i]);
  const NbCol col(L, i);
  float num = 0.f, den = 0.f;
  for (uint32_t k = 0; k < cn; k++) {
    const uint32_t j = col.get(k);
    const float2 xj = __ldg(&pos[j]);
    const float4 o = __ldg(&xyhm[j]);
    const float lj = __ldg(&level[j]);
    const float dx = xi.x - xj.x, dy = xi.y - xj.y;
    const float w = kernel_w(sqrtf(dx * dx + dy * dy), (hi + o.z) * 0.5f);
    const float dist = (lj > 0.f) ? -dmax : fmaxf(lj, -dmax);
    const float vw = o.w / __ldg(&rho[j]) * w;
    num += dist * vw;
    den += vw;
  }
  if (!isfinite(den) || !(den > 0.f)) atomicOr(&ctl->error_flags, ERRF_LEVEL_WEIGHT);
  level_out[i] = num / den;
}

}  // namespace

// perform_level_estimation (simulation.rs:863-927), EmptyAngle
int launch_level_estimation(asph_sim* sim) {
  const uint32_t n = sim->n;
  sim->level_valid = false;
  if (n == 0) return ASPH_OK;
  cudaStream_t st = sim->stream;
  const uint32_t blocks = (n + kThreads - 1) / kThreads;
  Lists L{sim->nbpool.p, sim->slice_base.p, sim->cnt.p, sim->cnt_ext.p, sim->far_idx.p, sim->far_cnt.p};
  const float cos_threshold = std::cos(50.f * (3.14159265358979323846f / 180.f));
  float* level = sim->level[sim->cur].p;
  k_level_reset<<<1, 1, 0, st>>>(sim->ctl);
  LAUNCH_CHECK();
  k_surface<<<blocks, kThreads, 0, st>>>(n, L, sim->xyhm.p, sim->nrm.p, sim->pp, cos_threshold, level, sim->stamp.p, sim->flags.p,
                                         sim->front[0].p, sim->ctl);
  LAUNCH_CHECK();
  int use_cutoff = sim->level_cutoff ? 1 : 0;
  float neg_dmax = -sim->pp.maximum_surface_distance;
  if (sim->prop_grid == 0) {  // co-resident blocks of the persistent kernel on this device
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_propagate, kPropThreads, 0));
    sim->prop_grid = std::max(1, per_sm * sim->sm_count);
  }
  // a front rarely holds more than a few ten thousand particles: one warp each
  uint32_t grid = uint32_t(std::max(1, std::min<int>(sim->prop_grid, int((n + 4u * kPropThreads - 1) / (4u * kPropThreads)))));
  uint32_t n_arg = n;
  float* level_arg = level;
  int* stamp_arg = sim->stamp.p;
  uint32_t* front_arg = sim->front[0].p;
  const float4* xyhm_arg = sim->xyhm.p;
  StepCtl* ctl_arg = sim->ctl;
  void* args[] = {&n_arg, &L, &xyhm_arg, &level_arg, &stamp_arg, &front_arg, &ctl_arg, &neg_dmax, &use_cutoff};
  cudaEvent_t kt0 = nullptr, kt1 = nullptr;
  if (sim->kt_every > 0) { kt0 = kt_event(sim); kt1 = kt_event(sim); cudaEventRecord(kt0, st); }
  CUDA_TRY(cudaLaunchCooperativeKernel((void*)k_propagate, dim3(grid), dim3(kPropThreads), args, 0, st));
  sim->kernel_launches++;
  if (kt1) cudaEventRecord(kt1, st);
  const int rc_sync = sync_ctl(sim);
  if (kt1) {
    float ms = 0.f;
    if (rc_sync == ASPH_OK && cudaEventElapsedTime(&ms, kt0, kt1) == cudaSuccess) { sim->kt_ms[ASPH_KT_LEVEL_PROPAGATE] += ms; sim->kt_samples[ASPH_KT_LEVEL_PROPAGATE]++; }
    kt_release(sim, kt0); kt_release(sim, kt1);
  }
  TRY(rc_sync);
  if (sim->ctl_host->error_flags & ERRF_LIST_CAPACITY) return ASPH_RETRY_LISTS;
  if (!sim->ctl_host->level_done) { sim->last_error = "level-set propagation did not terminate"; return ASPH_ERR_INVALID; }
  // sweeps including the final one that changes nothing, as the reference counts them (simulation.rs:739-799)
  sim->info.level_sweeps = std::max(1, sim->ctl_host->level_sweep);
  return ASPH_OK;
}

int launch_level_smoothing(asph_sim* sim) {
  const uint32_t n = sim->n;
  if (n == 0) { sim->level_valid = true; return ASPH_OK; }
  const uint32_t blocks = (n + kThreads - 1) / kThreads;
  const int c = sim->cur;
  Lists L{sim->nbpool.p, sim->slice_base.p, sim->cnt.p, sim->cnt_ext.p, sim->far_idx.p, sim->far_cnt.p};
  k_smooth<<<blocks, kThreads, 0, sim->stream>>>(n, L, sim->pos[c].p, sim->xyhm.p, sim->rho.p,
                                                 sim->level[c].p, sim->scratch_f.p, sim->pp.maximum_surface_distance, sim->ctl);
  LAUNCH_CHECK();
  CUDA_TRY(cudaMemcpyAsync(sim->level[c].p, sim->scratch_f.p, size_t(n) * sizeof(float), cudaMemcpyDeviceToDevice, sim->stream));
  sim->level_valid = true;
  return ASPH_OK;
}
