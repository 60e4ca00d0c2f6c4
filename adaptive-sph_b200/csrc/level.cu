// level.cu — level-set (distance to the free surface) estimation on the extended-range lists.
//
//   K3  surface_detection_by_empty_angle          simulation.rs:539-625
//   K4  propagate_level_set_from_surface_detection simulation.rs:729-801
//   K17 smooth_level_estimation_field             simulation.rs:804-857
//
// K4 in the reference re-scans all N particles every sweep until nothing changes (O(N * sweeps)).  A particle's
// value is final in the first sweep in which any neighbour already has one, i.e. in sweep = BFS distance from the
// detected surface, and it only reads values of the previous sweep.  So here each sweep touches only the front:
// k_expand claims the unassigned neighbours of the particles assigned in the previous sweep, k_assign computes
// max_j(φ_j − |x_ij|) for the claimed ones over neighbours stamped in EARLIER sweeps.  max is order independent and
// the distance is computed without contraction, so the field is bit-identical to the reference's Jacobi sweeps.
// `level` encoding: value <= 0 = FluidSurface(value); ASPH_LEVEL_INTERIOR (1.0) = FluidInterior.
#include "lists.cuh"

namespace {

constexpr int kThreads = 256;

typedef NbLists Lists;

__global__ void k_level_reset(StepCtl* ctl) {
  ctl->front_n[0] = 0; ctl->front_n[1] = 0; ctl->cand_n[0] = 0; ctl->cand_n[1] = 0;
  ctl->level_live[0] = 1; ctl->level_live[1] = 0;
  ctl->level_sweep = 0; ctl->level_done = 0;
}

__global__ void __launch_bounds__(kThreads)
k_surface(uint32_t n, Lists L, const float4* __restrict__ xyhm, const float2* __restrict__ nrm, const PackedParams P, float cos_threshold,
          float* __restrict__ level, int* __restrict__ stamp, uint8_t* __restrict__ flags, uint32_t* __restrict__ front, StepCtl* ctl) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
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
    front[atomicAdd(&ctl->front_n[0], 1u)] = i;
  } else {
    level[i] = ASPH_LEVEL_INTERIOR;
    stamp[i] = -1;
  }
  flags[i] = fl;
}

// sweep t: claim the unassigned neighbours of front[(t-1)&1]
__global__ void __launch_bounds__(kThreads)
k_expand(Lists L, int t, const uint32_t* __restrict__ front_in, uint32_t* __restrict__ cand, int* __restrict__ stamp, StepCtl* ctl) {
  if (ctl->level_done) return;
  const int pin = (t - 1) & 1;
  const uint32_t nf = ctl->level_live[pin] ? ctl->front_n[pin] : 0u;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ctl->front_n[t & 1] = 0;       // filled by k_assign of this sweep
    ctl->level_live[t & 1] = 0;
    if (nf == 0) ctl->level_done = 1;
  }
  for (uint32_t f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x) {
    const uint32_t j = front_in[f];
    const uint32_t ce = L.cnt_ext[j];
    const NbCol col(L, j);
    for (uint32_t k = 0; k < ce; k++) {
      const uint32_t i = col.get(k);
      if (stamp[i] == -1 && atomicCAS(&stamp[i], -1, -2) == -1) cand[atomicAdd(&ctl->cand_n[t & 1], 1u)] = i;
    }
  }
}

// sweep t: φ_i = max over neighbours assigned before sweep t of (φ_j − |x_j − x_i|)   (simulation.rs:757-785)
__global__ void __launch_bounds__(kThreads)
k_assign(Lists L, int t, const uint32_t* __restrict__ cand, const float4* __restrict__ xyhm, float* __restrict__ level,
         int* __restrict__ stamp, uint32_t* __restrict__ front_out, StepCtl* ctl, float neg_dmax, int use_cutoff) {
  if (ctl->level_done) return;
  const uint32_t nc = ctl->cand_n[t & 1];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ctl->cand_n[(t + 1) & 1] = 0;
    ctl->level_sweep = t;
  }
  bool live = false;
  for (uint32_t f = blockIdx.x * blockDim.x + threadIdx.x; f < nc; f += gridDim.x * blockDim.x) {
    const uint32_t i = cand[f];
    const float4 me = xyhm[i];
    const uint32_t ce = L.cnt_ext[i];
    const NbCol col(L, i);
    float best = -__int_as_float(0x7f800000);
    for (uint32_t k = 0; k < ce; k++) {
      const uint32_t j = col.get(k);
      const int sj = stamp[j];
      if (sj < 0 || sj >= t) continue;
      const float4 o = __ldg(&xyhm[j]);
      const float d = __fsqrt_rn(dist_sq_exact(__fsub_rn(o.x, me.x), __fsub_rn(o.y, me.y)));
      best = fmaxf(best, __fsub_rn(level[j], d));
    }
    level[i] = best;
    stamp[i] = t;
    front_out[atomicAdd(&ctl->front_n[t & 1], 1u)] = i;
    if (!use_cutoff || best > neg_dmax) live = true;
  }
  if (live) ctl->level_live[t & 1] = 1;
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
  const int grid = std::max(1, std::min<int>(int(blocks), sim->sm_count * 8));
  const int use_cutoff = sim->level_cutoff ? 1 : 0;
  int t = 1, batch = 8;
  for (;;) {
    for (int b = 0; b < batch; b++, t++) {
      k_expand<<<grid, kThreads, 0, st>>>(L, t, sim->front[(t - 1) & 1].p, sim->cand.p, sim->stamp.p, sim->ctl);
      LAUNCH_CHECK();
      k_assign<<<grid, kThreads, 0, st>>>(L, t, sim->cand.p, sim->xyhm.p, level, sim->stamp.p, sim->front[t & 1].p, sim->ctl,
                                          -sim->pp.maximum_surface_distance, use_cutoff);
      LAUNCH_CHECK();
    }
    TRY(sync_ctl(sim));
    if (sim->ctl_host->error_flags & ERRF_LIST_CAPACITY) return ASPH_RETRY_LISTS;
    if (sim->ctl_host->level_done) break;
    if (t > int(n) + 2) { sim->last_error = "level-set propagation did not terminate"; return ASPH_ERR_INVALID; }
    batch = std::min(batch * 2, 64);
  }
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
