// adapt.cu — spatially adaptive resampling: classify (K18), sharing (K19), merging (K20), splitting (K21).
// Compiled with -fmad=false: every decision (class thresholds, eligibility, mass limits) uses the reference's fp32
// arithmetic operation by operation, so given the same inputs the same particles share / merge / split.
//
//   classify_particles                 adaptivity/mod.rs:32-59
//   find_share_partner_sequential      adaptivity/particle_sharing.rs:14-111,  share_particles :152-240
//   find_merge_partner_sequential      adaptivity/particle_merging.rs:16-122,  merge_particles :270-371
//   split_particles                    adaptivity/splitting.rs:19-81
//   single_step_adaptivity             simulation.rs:2732-2796
//
// The reference's partner searches are SERIAL greedy loops over particles in index order (donor i claims the
// still-unclaimed eligible neighbours j in list order).  They are reproduced exactly by deterministic rounds: in each
// round every undecided donor stamps itself and its possible receivers with its reference index (64-bit atomicMax
// of round:~index); a donor whose stamp survived on all of them has no undecided lower-index donor that could touch
// the same particles, so it runs its inner loop (neighbours in ascending reference index = the oracle's list order)
// and is final.  Donors claimed as receivers meanwhile drop out.  Particles live on the device in grid-cell order;
// `refid` carries the reference index, and the reference's swap-with-last deletion and append-at-end splitting are
// reproduced in reference-index space with prefix sums.
#include "lists.cuh"

namespace {

constexpr int kThreads = 256;
constexpr uint32_t AVAILABLE = ASPH_MERGE_PARTNER_AVAILABLE;
constexpr uint32_t DELETE_ = ASPH_MERGE_PARTNER_DELETE;

struct AdaptArgs {
  NbLists L;
  const float4* __restrict__ xyhm;  // .z = h of this step (h2, unchanged since the step started)
  float2* pos;
  float2* vel;
  float* mass;
  const float* __restrict__ level;
  const uint32_t* __restrict__ refid;
  uint8_t* size_class;
  uint32_t* partner;
  uint32_t* counter;
  unsigned long long* stampkey;
};

__global__ void k_classify(uint32_t n, const float* __restrict__ level, const float* __restrict__ mass, const PackedParams P,
                           uint8_t* __restrict__ size_class) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) size_class[i] = classify_particle(level[i], mass[i], P);
}

__global__ void k_mass_sum(uint32_t n, const float* __restrict__ mass, double* out) {
  double s = 0.0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += double(mass[i]);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

__device__ __forceinline__ float dropped_mass_sharing(float level, float m, float dt, const PackedParams& P) {  // particle_sharing.rs:242-253
  const float tm = target_mass(level, P);
  return fminf(m - tm, tm * P.max_mass_transfer_sharing * dt);
}

// class + distance part of the eligibility test (independent of the greedy state)
__device__ __forceinline__ bool static_eligible(const AdaptArgs& A, const PackedParams& P, bool merging, uint32_t d, uint32_t j,
                                                float2 xd, float hd, float md) {
  const uint8_t cj = A.size_class[j];
  bool can;
  if (merging) {
    if (cj == ASPH_CLASS_LARGE || cj == ASPH_CLASS_TOO_LARGE) can = false;
    else if (cj == ASPH_CLASS_OPTIMAL) can = P.allow_merge_optimal != 0;
    else can = true;
    if (P.allow_merge_size_diff && A.mass[j] > 5.f * md) can = true;
  } else {
    if (cj == ASPH_CLASS_SMALL) can = true;
    else if (cj == ASPH_CLASS_TOO_SMALL) can = P.allow_share_too_small != 0;
    else if (cj == ASPH_CLASS_OPTIMAL) can = P.allow_share_optimal != 0;
    else can = false;
  }
  if (!can) return false;
  const float2 xj = A.pos[j];
  const float dx = xd.x - xj.x, dy = xd.y - xj.y;
  const float max_dist = ((hd + A.xyhm[j].z) * 0.5f) * (merging ? P.max_merge_distance : P.max_share_distance);
  return !(dx * dx + dy * dy > max_dist * max_dist);
}

__global__ void k_partner_init(uint32_t n, AdaptArgs A, uint8_t donor_class, uint32_t* __restrict__ work, StepCtl* ctl) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  A.partner[i] = AVAILABLE;
  A.counter[i] = 0;
  A.stampkey[i] = 0ull;
  if (A.size_class[i] == donor_class) work[atomicAdd(&ctl->work_n[0], 1u)] = i;
}
__global__ void k_partner_ctl_reset(StepCtl* ctl) {
  ctl->work_n[0] = 0; ctl->work_n[1] = 0; ctl->rounds = 0; ctl->n_claims = 0;
}

__device__ __forceinline__ unsigned long long stamp_of(uint32_t round, uint32_t refid) {
  return ((unsigned long long)round << 32) | (unsigned long long)(0xFFFFFFFFu - refid);
}

// round r, phase A: every live undecided donor stamps itself and its possible receivers
__global__ void __launch_bounds__(kThreads)
k_mark(AdaptArgs A, const PackedParams P, int merging, uint32_t round, const uint32_t* __restrict__ work_in, StepCtl* ctl) {
  const int pin = (round - 1) & 1;
  const uint32_t nw = ctl->work_n[pin];
  if (blockIdx.x == 0 && threadIdx.x == 0) ctl->work_n[round & 1] = 0;  // filled by k_decide of this round
  for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < nw; w += gridDim.x * blockDim.x) {
    const uint32_t d = work_in[w];
    if (A.partner[d] != AVAILABLE) continue;  // claimed as a receiver: can never donate (…rs:62-68 / :88-96)
    const unsigned long long key = stamp_of(round, A.refid[d]);
    atomicMax(&A.stampkey[d], key);
    const float2 xd = A.pos[d];
    const float hd = A.xyhm[d].z, md = A.mass[d];
    const uint32_t cn = nb_cn(A.L.cnt[d]);
    const NbCol col(A.L, d);
    for (uint32_t k = 0; k < cn; k++) {
      const uint32_t j = col.get(k);
      if (j == d || A.partner[j] != AVAILABLE) continue;
      if (static_eligible(A, P, merging != 0, d, j, xd, hd, md)) atomicMax(&A.stampkey[j], key);
    }
  }
}

// round r, phase B: donors that own all their stamps run the reference's inner loop; the rest wait
__global__ void __launch_bounds__(kThreads)
k_decide(AdaptArgs A, const PackedParams P, int merging, uint32_t round, float dt, const uint32_t* __restrict__ work_in,
         uint32_t* __restrict__ work_out, StepCtl* ctl) {
  const int pin = (round - 1) & 1;
  const uint32_t nw = ctl->work_n[pin];
  if (blockIdx.x == 0 && threadIdx.x == 0 && nw > 0) ctl->rounds = round;
  for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < nw; w += gridDim.x * blockDim.x) {
    const uint32_t d = work_in[w];
    if (A.partner[d] != AVAILABLE) continue;
    const unsigned long long key = stamp_of(round, A.refid[d]);
    const float2 xd = A.pos[d];
    const float hd = A.xyhm[d].z, md = A.mass[d];
    const uint32_t cn = nb_cn(A.L.cnt[d]);
    const NbCol col(A.L, d);
    bool ready = A.stampkey[d] == key;
    for (uint32_t k = 0; k < cn && ready; k++) {
      const uint32_t j = col.get(k);
      if (j == d || A.partner[j] != AVAILABLE) continue;
      if (static_eligible(A, P, merging != 0, d, j, xd, hd, md) && A.stampkey[j] != key) ready = false;
    }
    if (!ready) { work_out[atomicAdd(&ctl->work_n[round & 1], 1u)] = d; continue; }
    // the reference's inner loop over N(d) in ascending reference index
    const float dropped = merging ? md : dropped_mass_sharing(A.level[d], md, dt, P);
    uint32_t count = 0;
    long long last = -1;
    for (;;) {
      uint32_t best_j = 0xFFFFFFFFu;
      long long best_r = 0x7FFFFFFFFFFFFFFFll;
      for (uint32_t k = 0; k < cn; k++) {
        const uint32_t j = col.get(k);
        const long long r = (long long)A.refid[j];
        if (r > last && r < best_r) { best_r = r; best_j = j; }
      }
      if (best_j == 0xFFFFFFFFu) break;
      last = best_r;
      const uint32_t j = best_j;
      if (j == d) continue;
      if (!static_eligible(A, P, merging != 0, d, j, xd, hd, md)) continue;
      const float new_mass_j = A.mass[j] + dropped / float(count + 1u);
      const float target_j = target_mass(A.level[j], P);
      if (new_mass_j >= target_j * 1.1f) continue;
      if (new_mass_j > P.mass_base) continue;
      if (A.partner[j] != AVAILABLE) continue;
      if (count == 0) A.partner[d] = DELETE_;  // partner[d] == AVAILABLE was checked above
      A.partner[j] = d;
      count++;
    }
    A.counter[d] = count;
    if (count) atomicAdd(&ctl->n_claims, count);
  }
}

// validate_share_partners particle_sharing.rs:113-150 / validate_merge_partners particle_merging.rs:230-268
__global__ void __launch_bounds__(kThreads)
k_validate(uint32_t n, AdaptArgs A, uint8_t donor_class, StepCtl* ctl) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool ok = true;
  const uint32_t c = A.counter[i], p = A.partner[i];
  if (c > 0) {
    if (A.size_class[i] != donor_class || p != DELETE_) ok = false;
    uint32_t c2 = 0;
    const uint32_t cn = nb_cn(A.L.cnt[i]);
    const NbCol col(A.L, i);
    for (uint32_t k = 0; k < cn; k++) if (A.partner[col.get(k)] == i) c2++;
    if (c2 != c) ok = false;
  } else {
    if (p == DELETE_) ok = false;
    else if (p != AVAILABLE && A.partner[p] != DELETE_) ok = false;
  }
  if (!ok) atomicOr(&ctl->error_flags, ERRF_PARTNER_VALIDATION);
}

// receivers absorb their share (particle_sharing.rs:164-211, particle_merging.rs:282-324).  A receiver has exactly one
// donor and donors are never receivers, so this is order independent and runs in place.
__global__ void __launch_bounds__(kThreads)
k_apply_receivers(uint32_t n, AdaptArgs A, const PackedParams P, int merging, float dt) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t j = A.partner[i];
  if (j == AVAILABLE || j == DELETE_) return;
  const uint32_t cj = A.counter[j];
  if (int(cj) < (merging ? P.min_merge_partners : P.min_share_partners)) return;
  const float mass_i = A.mass[i], mass_j = A.mass[j];
  const float dropped = merging ? mass_j : dropped_mass_sharing(A.level[j], mass_j, dt, P);
  const float mass_n = dropped / float(cj);
  const float m = mass_i + mass_n;
  const float2 vi = A.vel[i], vj = A.vel[j], xi = A.pos[i], xj = A.pos[j];
  A.vel[i] = make_float2((mass_i * vi.x + mass_n * vj.x) / m, (mass_i * vi.y + mass_n * vj.y) / m);
  A.pos[i] = make_float2((mass_i * xi.x + mass_n * xj.x) / m, (mass_i * xi.y + mass_n * xj.y) / m);
  A.mass[i] = m;
}
__global__ void __launch_bounds__(kThreads)
k_share_donors(uint32_t n, AdaptArgs A, const PackedParams P, float dt) {  // particle_sharing.rs:213-240
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (A.partner[i] != DELETE_ || int(A.counter[i]) < P.min_share_partners) return;
  const float m = A.mass[i];
  A.mass[i] = m - dropped_mass_sharing(A.level[i], m, dt, P);
}

// ---- merging: deletion in reference-index space (swap-with-last loop of particle_merging.rs:341-370) --------------
// del_ref[r] = 1 if the particle with reference index r is removed; keep[i] = 1 in device order otherwise
__global__ void k_mark_deleted(uint32_t n, AdaptArgs A, int min_partners, uint32_t* __restrict__ del_ref, uint32_t* __restrict__ keep,
                               StepCtl* ctl) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool del = false;
  if (A.partner[i] == DELETE_ && int(A.counter[i]) >= min_partners) {
    const float m = A.mass[i] - A.mass[i];  // dropped_mass_merging == mass (particle_merging.rs:373-385)
    del = m < 0.000001f;
    if (!del) A.mass[i] = m;
  }
  del_ref[A.refid[i]] = del ? 1u : 0u;
  keep[i] = del ? 0u : 1u;
  if (i == 0) { del_ref[n] = 0u; keep[n] = 0u; ctl->n_new = n; }
}
// holes[k] = k-th removed reference index below n_new (ascending)
__global__ void k_holes(uint32_t n, const uint32_t* __restrict__ del_scan, uint32_t* __restrict__ holes) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const uint32_t n_new = n - del_scan[n];
  if (r < n_new && del_scan[r + 1] != del_scan[r]) holes[del_scan[r]] = r;
}
// survivors: those with reference index >= n_new fill the holes, highest index first; then compact in device order
__global__ void k_compact(uint32_t n, const uint32_t* __restrict__ del_scan, const uint32_t* __restrict__ holes,
                          const uint32_t* __restrict__ keep_scan, const float2* __restrict__ pos, const float2* __restrict__ vel,
                          const float* __restrict__ mass, const float* __restrict__ level, const uint32_t* __restrict__ refid,
                          float2* __restrict__ pos_o, float2* __restrict__ vel_o, float* __restrict__ mass_o, float* __restrict__ level_o,
                          uint32_t* __restrict__ refid_o, StepCtl* ctl) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (keep_scan[i + 1] == keep_scan[i]) return;
  const uint32_t total_del = del_scan[n];
  const uint32_t n_new = n - total_del;
  uint32_t r = refid[i];
  if (r >= n_new) {
    const uint32_t k = (n - 1u - r) - (total_del - del_scan[r + 1]);  // survivors with a higher reference index
    r = holes[k];
  }
  const uint32_t o = keep_scan[i];
  pos_o[o] = pos[i]; vel_o[o] = vel[i]; mass_o[o] = mass[i]; level_o[o] = level[i]; refid_o[o] = r;
  if (i == 0 || o == 0) ctl->n_new = n_new;
}

// ---- splitting (splitting.rs:19-81) ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t split_children(float level, float m, const PackedParams& P, int max_children, unsigned int* err) {
  const float tm = target_mass(level, P);
  const float rr = roundf(m / tm);
  uint32_t nc = rr > 4.0e9f ? 0xFFFFFFFFu : uint32_t(rr);
  if (nc > uint32_t(max_children)) {
    if (P.fail_on_missing_split_pattern) *err |= ERRF_SPLIT_PATTERN;
    nc = uint32_t(max_children);
  }
  if (nc < 2u) { *err |= ERRF_SPLIT_CHILDREN; nc = 1u; }
  return nc;
}
__global__ void k_split_count(uint32_t n, const uint8_t* __restrict__ size_class, const float* __restrict__ level,
                              const float* __restrict__ mass, const uint32_t* __restrict__ refid, const PackedParams P, int max_children,
                              uint32_t* __restrict__ extra_ref, StepCtl* ctl) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t extra = 0;
  if (size_class[i] == ASPH_CLASS_TOO_LARGE) {
    unsigned int err = 0;
    extra = split_children(level[i], mass[i], P, max_children, &err) - 1u;
    if (err) atomicOr(&ctl->error_flags, err);
    atomicAdd(&ctl->n_split_parents, 1u);
  }
  extra_ref[refid[i]] = extra;
  if (i == 0) extra_ref[n] = 0u;
}
__global__ void k_split_total(uint32_t n, const uint32_t* __restrict__ extra_scan, StepCtl* ctl) { ctl->n_new = n + extra_scan[n]; }
__global__ void k_split_apply(uint32_t n, uint32_t cap, const uint8_t* __restrict__ size_class, const uint32_t* __restrict__ extra_scan,
                              const PackedParams P, int max_children, const int* __restrict__ split_off, const float* __restrict__ split_pos,
                              float2* __restrict__ pos, float2* __restrict__ vel, float* __restrict__ mass, float* __restrict__ level,
                              uint32_t* __restrict__ refid) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (size_class[i] != ASPH_CLASS_TOO_LARGE) return;
  unsigned int err = 0;
  const float m = mass[i], lv = level[i];
  const uint32_t nc = split_children(lv, m, P, max_children, &err);
  if (nc < 2u) return;
  const float* pat = split_pos + 2 * size_t(split_off[nc - 2u]);
  const float radius = sqrtf((m / 1.0f) * ASPH_FRAC_1_PI_F);  // volume_to_radius(mass / INIT_REST_DENSITY)
  const float child_mass = m / float(nc);
  const float2 op = pos[i], ov = vel[i];
  const uint32_t first = n + extra_scan[refid[i]];
  for (uint32_t c = 0; c < nc; c++) {
    const uint32_t t = (c == 0) ? i : first + (c - 1u);
    if (t >= cap) return;
    mass[t] = child_mass; vel[t] = ov; level[t] = lv;
    pos[t] = make_float2(op.x + pat[2 * c] * radius, op.y + pat[2 * c + 1] * radius);
    if (c > 0) refid[t] = t;  // appended children: reference index == position in the appended block
  }
}

// ---- support_length_estimation != FromMass: h2_next and the boundary handler's lambda follow the particles ---------
// Receivers and sharing donors get the length of their new mass (particle_sharing.rs:206,238, particle_merging.rs:323).
// Runs after k_apply_receivers / k_share_donors (the masses are the new ones, partner / counter still say who changed).
__global__ void __launch_bounds__(kThreads)
k_hnext_after_transfer(uint32_t n, AdaptArgs A, const PackedParams P, int merging, float* __restrict__ hnext) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t j = A.partner[i];
  const int minp = merging ? P.min_merge_partners : P.min_share_partners;
  bool changed;
  if (j == AVAILABLE) changed = false;
  else if (j == DELETE_) changed = !merging && int(A.counter[i]) >= minp;  // sharing donor; a merging donor is deleted
  else changed = int(A.counter[j]) >= minp;                                // receiver
  if (changed) hnext[i] = h_from_mass(A.mass[i], P.rest_density);
}
// same destinations as k_compact
__global__ void k_compact_extra(uint32_t n, const uint32_t* __restrict__ keep_scan, const float* __restrict__ hnext, const float* __restrict__ lamprev,
                                float* __restrict__ hnext_o, float* __restrict__ lamprev_o) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (keep_scan[i + 1] == keep_scan[i]) return;
  const uint32_t o = keep_scan[i];
  hnext_o[o] = hnext[i]; lamprev_o[o] = lamprev[i];
}
// same targets as k_split_apply, launched BEFORE it (it needs the parents' unsplit masses): every child gets the length
// of the child mass (splitting.rs:50,66,74); appended children start with empty lambda lists (boundary_handler.extend)
__global__ void k_split_extra(uint32_t n, uint32_t cap, const uint8_t* __restrict__ size_class, const uint32_t* __restrict__ extra_scan,
                              const PackedParams P, int max_children, const float* __restrict__ mass, const float* __restrict__ level,
                              const uint32_t* __restrict__ refid, float* __restrict__ hnext, float* __restrict__ lamprev) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (size_class[i] != ASPH_CLASS_TOO_LARGE) return;
  unsigned int err = 0;
  const float m = mass[i];
  const uint32_t nc = split_children(level[i], m, P, max_children, &err);
  if (nc < 2u) return;
  const float child_h = h_from_mass(m / float(nc), P.rest_density);
  const uint32_t first = n + extra_scan[refid[i]];
  for (uint32_t c = 0; c < nc; c++) {
    const uint32_t t = (c == 0) ? i : first + (c - 1u);
    if (t >= cap) return;
    hnext[t] = child_h;
    if (c > 0) lamprev[t] = 0.f;
  }
}

// ---- IISPH2: particle_size_class outlives the step (the omega pass of the next step reads it) -------------------------
__global__ void k_cls_copy(uint32_t n, const uint8_t* __restrict__ size_class, uint8_t* __restrict__ cls) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cls[i] = size_class[i];
}
__global__ void k_cls_compact(uint32_t n, const uint32_t* __restrict__ keep_scan, const uint8_t* __restrict__ size_class, uint8_t* __restrict__ cls_o) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (keep_scan[i + 1] != keep_scan[i]) cls_o[keep_scan[i]] = size_class[i];
}
__global__ void k_cls_fill(uint32_t first, uint32_t n, uint8_t v, uint8_t* __restrict__ cls) {  // appended children: ParticleVec::extend default
  const uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) cls[i] = v;
}

AdaptArgs args_of(asph_sim* sim) {
  AdaptArgs A;
  const int c = sim->cur;
  A.L.pool = sim->nbpool.p; A.L.slice_base = sim->slice_base.p; A.L.cnt = sim->cnt.p; A.L.cnt_ext = sim->cnt_ext.p; A.L.far_idx = sim->far_idx.p; A.L.far_cnt = sim->far_cnt.p; A.xyhm = sim->xyhm.p;
  A.pos = sim->pos[c].p; A.vel = sim->vel[c].p; A.mass = sim->mass[c].p; A.level = sim->level[c].p; A.refid = sim->refid[c].p;
  A.size_class = sim->size_class.p; A.partner = sim->merge_partner.p; A.counter = sim->merge_counter.p;
  A.stampkey = sim->stampkey.p;
  return A;
}

int total_mass(asph_sim* sim, double* out) {
  double* acc = (double*)sim->scratch_f.p;
  CUDA_TRY(cudaMemsetAsync(acc, 0, sizeof(double), sim->stream));
  if (sim->n) {
    k_mass_sum<<<std::max(1, sim->sm_count * 4), kThreads, 0, sim->stream>>>(sim->n, sim->mass[sim->cur].p, acc);
    LAUNCH_CHECK();
  }
  CUDA_TRY(cudaMemcpyAsync(out, acc, sizeof(double), cudaMemcpyDeviceToHost, sim->stream));
  CUDA_TRY(cudaStreamSynchronize(sim->stream));
  return ASPH_OK;
}

int classify(asph_sim* sim) {
  const uint32_t n = sim->n;
  k_classify<<<(n + kThreads - 1) / kThreads, kThreads, 0, sim->stream>>>(n, sim->level[sim->cur].p, sim->mass[sim->cur].p, sim->pp,
                                                                          sim->size_class.p);
  LAUNCH_CHECK();
  return ASPH_OK;
}

// the greedy partner search in deterministic rounds; returns the number of claimed receivers
int find_partners(asph_sim* sim, bool merging, float dt, uint32_t* claims) {
  const uint32_t n = sim->n;
  cudaStream_t st = sim->stream;
  const uint32_t blocks = (n + kThreads - 1) / kThreads;
  const AdaptArgs A = args_of(sim);
  const uint8_t donor_class = merging ? ASPH_CLASS_TOO_SMALL : ASPH_CLASS_LARGE;
  k_partner_ctl_reset<<<1, 1, 0, st>>>(sim->ctl);
  LAUNCH_CHECK();
  k_partner_init<<<blocks, kThreads, 0, st>>>(n, A, donor_class, sim->work[0].p, sim->ctl);
  LAUNCH_CHECK();
  const int grid = std::max(1, std::min<int>(int(blocks), sim->sm_count * 8));
  uint32_t round = 1;
  int batch = 4;
  for (;;) {
    for (int b = 0; b < batch; b++, round++) {
      k_mark<<<grid, kThreads, 0, st>>>(A, sim->pp, merging ? 1 : 0, round, sim->work[(round - 1) & 1].p, sim->ctl);
      LAUNCH_CHECK();
      k_decide<<<grid, kThreads, 0, st>>>(A, sim->pp, merging ? 1 : 0, round, dt, sim->work[(round - 1) & 1].p, sim->work[round & 1].p,
                                          sim->ctl);
      LAUNCH_CHECK();
    }
    TRY(sync_ctl(sim));
    if (sim->ctl_host->work_n[(round - 1) & 1] == 0) break;
    if (round > 2u * n + 16u) { sim->last_error = "partner search did not terminate"; return ASPH_ERR_INVALID; }
    batch = std::min(batch * 2, 64);
  }
  sim->adapt_rounds += sim->ctl_host->rounds;
  *claims = sim->ctl_host->n_claims;
  k_validate<<<blocks, kThreads, 0, st>>>(n, A, donor_class, sim->ctl);
  LAUNCH_CHECK();
  return ASPH_OK;
}

}  // namespace

int launch_adaptivity(asph_sim* sim, float dt) {
  const uint32_t n0 = sim->n;
  if (n0 == 0) return ASPH_OK;
  cudaStream_t st = sim->stream;
  CUDA_TRY(sim->stampkey.ensure(sim->cap));
  double m1 = 0, m2 = 0;
  TRY(total_mass(sim, &m1));
  sim->adapt_rounds = 0;
  const PackedParams& P = sim->pp;
  // IISPH2 reads the classes of the last resampling phase in the next step's omega pass
  const bool track_cls = solver_iisph2(sim);
  bool cls_done = false;
  const bool classified = sim->share_enabled || (sim->step_number % 2 == 0 ? sim->merge_enabled : sim->split_enabled);
  if (track_cls) { for (int b = 0; b < 2; b++) CUDA_TRY(sim->cls[b].ensure(sim->cap)); }

  if (sim->share_enabled) {  // simulation.rs:2747-2758
    TRY(classify(sim));
    uint32_t claims = 0;
    TRY(find_partners(sim, false, dt, &claims));
    sim->info.n_shared = int(claims);
    const uint32_t blocks = (sim->n + kThreads - 1) / kThreads;
    const AdaptArgs A = args_of(sim);
    k_apply_receivers<<<blocks, kThreads, 0, st>>>(sim->n, A, P, 0, dt);
    LAUNCH_CHECK();
    k_share_donors<<<blocks, kThreads, 0, st>>>(sim->n, A, P, dt);
    LAUNCH_CHECK();
    if (sim->hdist_valid) {
      k_hnext_after_transfer<<<blocks, kThreads, 0, st>>>(sim->n, A, P, 0, sim->hnext[sim->cur].p);
      LAUNCH_CHECK();
    }
  }
  if (sim->step_number % 2 == 0) {
    if (sim->merge_enabled) {  // simulation.rs:2760-2774
      TRY(classify(sim));
      uint32_t claims = 0;
      TRY(find_partners(sim, true, dt, &claims));
      sim->info.n_merged = int(claims);
      const uint32_t n = sim->n;
      const uint32_t blocks = (n + kThreads - 1) / kThreads;
      const AdaptArgs A = args_of(sim);
      k_apply_receivers<<<blocks, kThreads, 0, st>>>(n, A, P, 1, dt);
      LAUNCH_CHECK();
      if (sim->hdist_valid) {
        k_hnext_after_transfer<<<blocks, kThreads, 0, st>>>(n, A, P, 1, sim->hnext[sim->cur].p);
        LAUNCH_CHECK();
      }
      uint32_t* del_ref = sim->scratch_u[0].p;   // n + 1
      uint32_t* keep = sim->scratch_u[1].p;      // n + 1
      uint32_t* holes = sim->scratch_u[2].p;
      k_mark_deleted<<<blocks, kThreads, 0, st>>>(n, A, P.min_merge_partners, del_ref, keep, sim->ctl);
      LAUNCH_CHECK();
      TRY(launch_exclusive_scan(sim, del_ref, del_ref, &sim->ctl->n_new, 1, n + 1));
      TRY(launch_exclusive_scan(sim, keep, keep, &sim->ctl->n_new, 1, n + 1));
      k_holes<<<blocks, kThreads, 0, st>>>(n, del_ref, holes);
      LAUNCH_CHECK();
      const int c = sim->cur;
      k_compact<<<blocks, kThreads, 0, st>>>(n, del_ref, holes, keep, sim->pos[c].p, sim->vel[c].p, sim->mass[c].p, sim->level[c].p,
                                             sim->refid[c].p, sim->pos[1 - c].p, sim->vel[1 - c].p, sim->mass[1 - c].p, sim->level[1 - c].p,
                                             sim->refid[1 - c].p, sim->ctl);
      LAUNCH_CHECK();
      if (sim->hdist_valid) {
        k_compact_extra<<<blocks, kThreads, 0, st>>>(n, keep, sim->hnext[c].p, sim->lamprev[c].p, sim->hnext[1 - c].p, sim->lamprev[1 - c].p);
        LAUNCH_CHECK();
      }
      if (track_cls) {
        k_cls_compact<<<blocks, kThreads, 0, st>>>(n, keep, sim->size_class.p, sim->cls[1 - c].p);
        LAUNCH_CHECK();
        cls_done = true;
      }
      TRY(sync_ctl(sim));
      TRY(check_error_flags(sim));
      const uint32_t n_new = sim->ctl_host->n_new;
      if (n_new != n) { sim->lists_valid = false; sim->step_fields_valid = false; }
      sim->cur = 1 - c;
      sim->n = n_new; sim->n_owned = n_new;
    }
  } else if (sim->split_enabled) {  // simulation.rs:2775-2788
    if (sim->max_children < 2) { sim->last_error = "splitting enabled but no split patterns were given"; return ASPH_ERR_INVALID; }
    const uint32_t n = sim->n;
    const uint32_t blocks = (n + kThreads - 1) / kThreads;
    uint32_t n_new = n;
    for (int attempt = 0; attempt < 2; attempt++) {
      TRY(classify(sim));
      const int c = sim->cur;
      uint32_t* extra = sim->scratch_u[0].p;
      CUDA_TRY(cudaMemsetAsync(&sim->ctl->n_split_parents, 0, sizeof(uint32_t), st));
      k_split_count<<<blocks, kThreads, 0, st>>>(n, sim->size_class.p, sim->level[c].p, sim->mass[c].p, sim->refid[c].p, P,
                                                 sim->max_children, extra, sim->ctl);
      LAUNCH_CHECK();
      CUDA_TRY(cudaMemcpyAsync(&sim->ctl->n_new, &n, sizeof(uint32_t), cudaMemcpyHostToDevice, st));  // scan length n + 1
      TRY(launch_exclusive_scan(sim, extra, extra, &sim->ctl->n_new, 1, n + 1));
      k_split_total<<<1, 1, 0, st>>>(n, extra, sim->ctl);
      LAUNCH_CHECK();
      TRY(sync_ctl(sim));
      TRY(check_error_flags(sim));
      n_new = sim->ctl_host->n_new;
      sim->info.n_split_parents = int(sim->ctl_host->n_split_parents);
      if (n_new <= sim->cap) break;
      // growing re-allocates the scratch arrays (the persistent ones are preserved): count again afterwards
      TRY(ensure_capacity(sim, n_new + n_new / 4 + 1024));
      CUDA_TRY(sim->stampkey.ensure(sim->cap));
    }
    if (n_new != n) {
      const int cc = sim->cur;
      if (sim->hdist_valid) {
        k_split_extra<<<blocks, kThreads, 0, st>>>(n, sim->cap, sim->size_class.p, sim->scratch_u[0].p, P, sim->max_children, sim->mass[cc].p,
                                                   sim->level[cc].p, sim->refid[cc].p, sim->hnext[cc].p, sim->lamprev[cc].p);
        LAUNCH_CHECK();
      }
      k_split_apply<<<blocks, kThreads, 0, st>>>(n, sim->cap, sim->size_class.p, sim->scratch_u[0].p, P, sim->max_children,
                                                 sim->split_off.p, sim->split_pos.p, sim->pos[cc].p, sim->vel[cc].p, sim->mass[cc].p,
                                                 sim->level[cc].p, sim->refid[cc].p);
      LAUNCH_CHECK();
      sim->lists_valid = false; sim->step_fields_valid = false;
      sim->n = n_new; sim->n_owned = n_new;
      if (track_cls) {  // parents keep their class, appended children are Optimal
        if (!sim->cls_valid) { for (int b = 0; b < 2; b++) CUDA_TRY(sim->cls[b].ensure(sim->cap)); }  // the capacity may just have grown
        k_cls_copy<<<blocks, kThreads, 0, st>>>(n, sim->size_class.p, sim->cls[cc].p);
        LAUNCH_CHECK();
        k_cls_fill<<<(n_new - n + kThreads - 1) / kThreads, kThreads, 0, st>>>(n, n_new, ASPH_CLASS_OPTIMAL, sim->cls[cc].p);
        LAUNCH_CHECK();
        cls_done = true;
      }
    }
  }
  if (track_cls && !cls_done && classified) {  // same particle set as the last classify
    k_cls_copy<<<(sim->n + kThreads - 1) / kThreads, kThreads, 0, st>>>(sim->n, sim->size_class.p, sim->cls[sim->cur].p);
    LAUNCH_CHECK();
  }
  if (track_cls && (cls_done || classified)) sim->cls_valid = true;
  TRY(sync_ctl(sim));
  TRY(check_error_flags(sim));
  TRY(total_mass(sim, &m2));
  if (!(m2 <= m1 + 0.005 && m2 >= m1 - 0.005)) {  // simulation.rs:2791-2792
    sim->last_error = "mass not conserved by resampling: " + std::to_string(m1) + " -> " + std::to_string(m2);
    return ASPH_ERR_MASS_CONSERVATION;
  }
  return ASPH_OK;
}
