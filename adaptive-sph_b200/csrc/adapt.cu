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
// still-unclaimed eligible neighbours j in list order).  They are reproduced exactly, without serialising, by ONE
// persistent cooperative kernel per search (k_greedy; tools/greedy_model.py is its executable specification, checked
// against the serial loop):
//   * E'(d) = what donor d would claim if every neighbour were still available (its loop run "optimistically") bounds
//     what it can ever claim: a claimed neighbour only lowers the divisor of the mass that later candidates are offered.
//     Donors with an empty E' are final at once (they are most of the donors of a settled simulation).
//   * d depends on a lower-index undecided donor y only if their touch sets {d} + E^(d), {y} + E^(y) intersect
//     (E^ = a cheap superset of E').  A donor with no such y runs the reference's inner loop (neighbours in ascending
//     reference index) and is final; a donor that finds one registers in that blocker's wait list and sleeps until the
//     blocker has decided or has been claimed itself.  So every donor is examined a handful of times, not once per round,
//     and a round costs two grid barriers instead of two launches and a host read.
// Particles live on the device in grid-cell order; `refid` carries the reference index, and the reference's
// swap-with-last deletion and append-at-end splitting are reproduced in reference-index space with prefix sums.
#include <cooperative_groups.h>

#include "lists.cuh"

namespace cg = cooperative_groups;

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
};

__global__ void k_classify(uint32_t n, const float* __restrict__ level, const float* __restrict__ mass, const PackedParams P,
                           uint8_t* __restrict__ size_class) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) size_class[i] = classify_particle(level[i], mass[i], P);
}

__global__ void k_mass_sum(uint32_t n, const float* __restrict__ mass, const uint32_t* __restrict__ refid, double* out) {
  double s = 0.0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (!(refid[i] & ASPH_GHOST_BIT)) s += double(mass[i]);  // ghosts (multi-GPU) are counted by their owners
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

__global__ void k_partner_ctl_reset(StepCtl* ctl) {
  ctl->work_n[0] = 0; ctl->work_n[1] = 0; ctl->ready_n = 0; ctl->rounds = 0; ctl->n_claims = 0; ctl->greedy_done = 0;
  ctl->mail_sent = 0; ctl->mail_n[0] = 0; ctl->mail_n[1] = 0; ctl->greedy_barriers = 0; ctl->validate_why = 0; ctl->greedy_duplicates = 0;
}

// ---- the greedy partner search as one persistent cooperative kernel ------------------------------------------------
constexpr int kGreedyThreads = 512;
constexpr uint32_t NONE_ = 0xFFFFFFFFu;
constexpr uint32_t GI_PENDING = 1u, GI_DONE = 2u;  // low byte of info[d] for a donor d; bits 8.. = |E'(d)|

struct GreedyArgs {
  AdaptArgs A;
  uint32_t* info;    // per particle: 0, or donor state | |E'| << 8
  float* drop;       // per donor: the mass it hands out (dropped_mass_sharing / its whole mass)
  uint32_t* head;    // per particle: first donor waiting for it (linked through `next`)
  uint32_t* next;
  uint32_t* resume;  // per donor: member of its touch set the blocker scan continues at
  uint32_t* work[2]; // donors to examine this round / next round
  uint32_t* ready;   // donors that decide, in the order they became ready (every donor enters once)
  StepCtl* ctl;
  uint32_t n;
  int merging;
  float dt;
  uint8_t donor_class;
};

__device__ __forceinline__ bool mass_ok(float mass_j, float target_j, float add, const PackedParams& P) {
  const float new_mass_j = mass_j + add;  // particle_sharing.rs:76-86 / particle_merging.rs:79-91
  return !(new_mass_j >= target_j * 1.1f) && !(new_mass_j > P.mass_base);
}

struct DonorRegs { uint32_t d, rid; float2 x; float h, m, drop; };

__device__ __forceinline__ DonorRegs donor_regs(const AdaptArgs& A, const PackedParams& P, int merging, float dt, uint32_t d) {
  DonorRegs D;
  D.d = d; D.rid = A.refid[d]; D.x = A.pos[d]; D.h = A.xyhm[d].z; D.m = A.mass[d];
  D.drop = merging ? D.m : dropped_mass_sharing(A.level[d], D.m, dt, P);
  return D;
}

// The reference's inner loop for donor D (one warp, every lane returns the same count): candidates in ascending
// reference index, each lane owning the list positions lane, lane + 32, ...  OPT: every receiver counts as available
// and nothing is written (E'); otherwise the claims are made, and lane c keeps the c-th claimed receiver in `mine`
// (c < 32; `walk` is called at once for the few beyond).
template <bool OPT, class Walk>
__device__ __forceinline__ uint32_t donor_loop(const AdaptArgs& A, const PackedParams& P, int merging, const DonorRegs& D, uint32_t lane,
                                               uint32_t& mine, Walk walk) {
  const uint32_t cn = nb_cn(A.L.cnt[D.d]);
  const NbCol col(A.L, D.d);
  struct Cand { uint32_t j, rid; float m, tgt; bool ok; };
  auto cand_at = [&](uint32_t k) {
    Cand c;
    c.ok = false; c.j = 0; c.rid = NONE_; c.m = 0.f; c.tgt = 0.f;
    if (k < cn) {
      c.j = col.get(k);
      if (c.j != D.d && static_eligible(A, P, merging != 0, D.d, c.j, D.x, D.h, D.m)) {
        c.ok = true; c.rid = A.refid[c.j] & ~ASPH_GHOST_BIT; c.m = A.mass[c.j]; c.tgt = target_mass(A.level[c.j], P);
      }
    }
    return c;
  };
  const Cand first = cand_at(lane);  // a column rarely has more than 32 rows: this one stays in registers
  uint32_t count = 0;
  long long last = -1;
  mine = NONE_;
  for (;;) {
    Cand best = first;
    if (!(best.ok && (long long)best.rid > last)) { best.ok = false; best.rid = NONE_; }
    for (uint32_t k = lane + 32u; k < cn; k += 32u) {
      const Cand c = cand_at(k);
      if (c.ok && (long long)c.rid > last && c.rid < best.rid) best = c;
    }
    const uint32_t r = __reduce_min_sync(0xffffffffu, best.ok ? best.rid : NONE_);
    if (r == NONE_) break;
    last = (long long)r;
    const int src = __ffs(__ballot_sync(0xffffffffu, best.ok && best.rid == r)) - 1;
    const uint32_t j = __shfl_sync(0xffffffffu, best.j, src);
    const float mj = __shfl_sync(0xffffffffu, best.m, src), tj = __shfl_sync(0xffffffffu, best.tgt, src);
    if (!mass_ok(mj, tj, D.drop / float(count + 1u), P)) continue;
    if (!OPT) {
      if (__ldcg(A.partner + j) != AVAILABLE) continue;
      if (lane == 0) {
        if (count == 0) A.partner[D.d] = DELETE_;  // the donor itself was available: checked when it was examined
        A.partner[j] = D.d;
      }
      if (count < 32u) { if (lane == count) mine = j; }
      else walk(j);
    }
    count++;
  }
  return count;
}

// ---- several GPUs (PEER): the search runs over the slabs of all ranks at once --------------------------------------
// Every rank examines and decides the donors it OWNS; its ghost zone is two pair supports wide (capi.cu), so the touch
// set of an owned donor and every donor that can touch it are present, as owned particles or as ghosts.  What a rank
// decides about a border particle is mailed to the ghost copies (CoopPeer, sim.cuh): the donor state and offered mass
// after phase I-b, and after every round's phase B the decided donors (state, partner, counter) and the claimed
// receivers; a claim on a ghost goes to the rank that owns the receiver, which also wakes the donors waiting for it.
// One cross-GPU barrier per round delivers the mail and tells every rank whether any rank has work or mail left.
// A donor blocked by a ghost cannot sleep on it (the wake-up would come from another GPU): it is examined again next round.
enum { MAIL_INFO = 0, MAIL_DROP = 1, MAIL_PARTNER = 2, MAIL_COUNTER = 3, MAIL_CLAIM = 4 };
constexpr uint32_t CLAIMED_ELSEWHERE = 0xFFFFFFFDu;  // partner of a ghost whose donor is not present on this rank
constexpr uint32_t kSlotMask = 0x0FFFFFFFu;

__device__ __forceinline__ uint32_t rid_of(const AdaptArgs& A, uint32_t i) { return A.refid[i] & ~ASPH_GHOST_BIT; }

template <bool PEER>
__global__ void __launch_bounds__(kGreedyThreads)
k_greedy(const GreedyArgs G, const PackedParams P, const CoopPeer C) {
  cg::grid_group grid = cg::this_grid();
  const AdaptArgs& A = G.A;
  StepCtl* ctl = G.ctl;
  const uint32_t n = G.n, lane = threadIdx.x & 31u;
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
  const uint32_t gwarp = gtid >> 5, nwarps = gthreads >> 5;
  const int merging = G.merging;
  volatile uint32_t* work_n = ctl->work_n;
  volatile uint32_t* ready_n = &ctl->ready_n;
  unsigned int bar = C.seq0 + 1u;  // PEER: number of the next cross-GPU barrier; mail posted before it uses its parity

  auto is_ghost = [&](uint32_t i) { return PEER && nb_ghost(__ldg(&A.L.cnt[i])); };
  // one message to the neighbour on `side` (0 = rank - 1)
  auto mail = [&](int side, uint32_t slot, uint32_t kind, uint32_t payload) { coop_mail(C, ctl, bar, side, slot | (kind << 28), payload); };
  // what the ghost copies of owned particle i have to hear
  auto mail_copies = [&](uint32_t i, uint32_t kind, uint32_t payload) {
#pragma unroll
    for (int side = 0; side < 2; side++) {
      const uint32_t sl = __ldg(&C.rslot[side][i]);
      if (sl != 0xffffffffu && C.nb_mbox[side]) mail(side, sl, kind, payload);
    }
  };

  // ---- phase I-a: reset the per-particle state, collect the donors (find_*_partner_sequential resets partner / counter)
  for (uint32_t i0 = gtid - lane; i0 < n; i0 += gthreads) {
    const uint32_t i = i0 + lane;
    bool donor = false;
    if (i < n) {
      A.partner[i] = AVAILABLE; A.counter[i] = 0; G.head[i] = NONE_; G.info[i] = 0u; G.resume[i] = 0u;
      donor = A.size_class[i] == G.donor_class && !is_ghost(i);
      if (PEER && A.size_class[i] == G.donor_class && !donor) G.info[i] = GI_PENDING;  // a ghost donor: pending until its owner says otherwise
    }
    const unsigned int mask = __ballot_sync(0xffffffffu, donor);
    if (mask) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(&ctl->work_n[0], uint32_t(__popc(mask)));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (donor) G.work[0][base + uint32_t(__popc(mask & ((1u << lane) - 1u)))] = i;
    }
  }
  grid.sync();
  // ---- phase I-b: E'(d) of every donor
  {
    const uint32_t nw = work_n[0];
    for (uint32_t w = gwarp; w < nw; w += nwarps) {
      const uint32_t d = __ldcg(G.work[0] + w);
      const DonorRegs D = donor_regs(A, P, merging, G.dt, d);
      uint32_t mine;
      const uint32_t c = donor_loop<true>(A, P, merging, D, lane, mine, [](uint32_t) {});
      if (lane == 0) {
        const uint32_t info = (c ? GI_PENDING : GI_DONE) | (min(c, 0xFFFFFFu) << 8);
        G.drop[d] = D.drop; G.info[d] = info;
        if (PEER) { mail_copies(d, MAIL_INFO, info); if (c) mail_copies(d, MAIL_DROP, __float_as_uint(D.drop)); }
      }
    }
  }

  // is x in the touch set of the lower donor y?  (x's own values are passed in: they are warp-uniform in the scan)
  auto touches = [&](uint32_t y, uint32_t info_y, uint32_t x, float mass_x, float target_x) {
    if (x == y) return true;
    const float my = A.mass[y];
    if (!static_eligible(A, P, merging != 0, y, x, A.pos[y], A.xyhm[y].z, my)) return false;
    return mass_ok(mass_x, target_x, (PEER ? __ldcg(G.drop + y) : G.drop[y]) / float(info_y >> 8), P);
  };
  // wake everything that waits for b: it goes into the next round's work list
  auto wake = [&](uint32_t b, uint32_t* __restrict__ out, uint32_t slot) {
    uint32_t z = __ldcg(G.head + b);
    if (z == NONE_) return;
    G.head[b] = NONE_;
    for (uint32_t guard = 0; z != NONE_ && guard <= n; guard++) {  // (a list is a chain of distinct donors: at most n long)
      const uint32_t nz = __ldcg(G.next + z);
      out[atomicAdd(&ctl->work_n[slot], 1u)] = z;
      z = nz;
    }
  };
  // PEER: fence, meet the other GPUs, take in their mail; returns whether any rank has work or mail outstanding
  auto exchange = [&](uint32_t pout) {
    grid.sync();
    // block 0 alone meets the other GPUs and takes in their mail (a round's mail is a handful of messages; the first
    // exchange's — the border donors' states — a few thousand), the other blocks wait at the second grid barrier
    if (blockIdx.x == 0) {
      if (threadIdx.x == 0) {
        const bool sent = coop_publish_mail(C, ctl, bar);
        const unsigned int mine = (work_n[pout] > 0u ? 1u : 0u) | (sent ? 2u : 0u);
        const unsigned int all = coop_barrier(C, bar, mine, ctl);
        *C.verdict = ((all & 3u) != 0u && !(all & 0x80000000u)) ? 1u : 0u;
      }
      __syncthreads();
      const uint32_t par = bar & 1u;
#pragma unroll
      for (int side = 0; side < 2; side++) {
        const uint32_t n_in = min(*reinterpret_cast<volatile unsigned int*>(&C.self->mbox_n[par][side]), C.mbox_cap);
        const uint2* __restrict__ box = C.mbox + size_t(par * 2u + uint32_t(side)) * C.mbox_cap;
        for (uint32_t e = threadIdx.x; e < n_in; e += blockDim.x) {
          const uint2 m = __ldcg(box + e);
          const uint32_t slot = m.x & kSlotMask;
          switch (m.x >> 28) {
            case MAIL_INFO: G.info[slot] = m.y; break;
            case MAIL_DROP: G.drop[slot] = __uint_as_float(m.y); break;
            case MAIL_PARTNER: A.partner[slot] = m.y; break;
            case MAIL_COUNTER: A.counter[slot] = m.y; break;
            case MAIL_CLAIM:  // one of my particles was claimed by a donor of the neighbour rank (m.y = that donor's ghost here)
              A.partner[slot] = m.y;
              wake(slot, G.work[pout], pout);
              break;
          }
        }
      }
    }
    grid.sync();
    const bool go = *reinterpret_cast<volatile unsigned int*>(C.verdict) != 0u;
    bar++;
    return go;
  };

  bool go = true;
  if (PEER) go = exchange(0u);  // the ghosts' donor states and offered masses; work list 0 is the first round's
  else grid.sync();

  uint32_t ready_begin = 0;
  uint32_t round = 1;
  for (;; round++) {
    const uint32_t pin = (round - 1u) & 1u, pout = round & 1u;
    const uint32_t nw = work_n[pin];
    if (PEER ? !go : nw == 0u) break;
    if (round > (PEER ? 0x3fffffffu : n + 2u)) {  // cannot happen (the lowest undecided donor is never blocked); do not hang if it does
      if (gtid == 0) atomicOr(&ctl->error_flags, ERRF_PARTNER_VALIDATION);
      break;
    }
    if (gtid == 0) ctl->rounds = round;
    // ---- phase A: examine (the state is frozen: nothing decides in this phase)
    for (uint32_t w = gwarp; w < nw; w += nwarps) {
      const uint32_t d = __ldcg(G.work[pin] + w);
      const uint32_t info_d = __ldcg(G.info + d);
      if ((info_d & 0xffu) != GI_PENDING) continue;
      if (__ldcg(A.partner + d) != AVAILABLE) {  // claimed as a receiver meanwhile: can never donate (…rs:62-68 / :88-96)
        if (lane == 0) G.info[d] = (info_d & ~0xffu) | GI_DONE;
        continue;
      }
      const DonorRegs D = donor_regs(A, P, merging, G.dt, d);
      const uint32_t rid_d = D.rid & ~ASPH_GHOST_BIT;
      const uint32_t res = __ldcg(G.resume + d);
      const uint32_t cn = nb_cn(A.L.cnt[d]);
      const NbCol col(A.L, d);
      uint32_t blocker = NONE_, at = 0;
      // scan N(x) for a lower undecided, unclaimed donor that touches x
      auto scan_x = [&](uint32_t x) {
        const float mass_x = A.mass[x], target_x = target_mass(A.level[x], P);
        const uint32_t cx = nb_cn(A.L.cnt[x]);
        const NbCol colx(A.L, x);
        for (uint32_t k0 = 0; k0 < cx; k0 += 32u) {
          const uint32_t k = k0 + lane;
          bool hit = false;
          uint32_t y = 0;
          if (k < cx) {
            y = colx.get(k);
            const uint32_t iy = __ldcg(G.info + y);
            hit = (iy & 0xffu) == GI_PENDING && rid_of(A, y) < rid_d && __ldcg(A.partner + y) == AVAILABLE && touches(y, iy, x, mass_x, target_x);
          }
          const unsigned int m = __ballot_sync(0xffffffffu, hit);
          if (m) return __shfl_sync(0xffffffffu, y, __ffs(m) - 1);
        }
        return NONE_;
      };
      if (res == 0u) blocker = scan_x(d);  // member 0: the donor itself
      for (uint32_t k0 = 0; k0 < cn && blocker == NONE_; k0 += 32u) {
        const uint32_t k = k0 + lane;
        bool member = false;
        uint32_t x = 0;
        if (k < cn && 1u + k >= res) {
          x = col.get(k);
          if (x != d && static_eligible(A, P, merging != 0, d, x, D.x, D.h, D.m))
            member = mass_ok(A.mass[x], target_mass(A.level[x], P), D.drop / float(info_d >> 8), P);
        }
        unsigned int m = __ballot_sync(0xffffffffu, member);
        while (m) {
          const int b = __ffs(m) - 1;
          m &= m - 1u;
          blocker = scan_x(__shfl_sync(0xffffffffu, x, b));
          if (blocker != NONE_) { at = 1u + k0 + uint32_t(b); break; }
        }
      }
      if (lane == 0) {
        if (blocker == NONE_) {
          G.ready[atomicAdd(&ctl->ready_n, 1u)] = d;
        } else {
          G.resume[d] = at;
          if (is_ghost(blocker)) G.work[pout][atomicAdd(&ctl->work_n[pout], 1u)] = d;  // no wake-up crosses GPUs: look again next round
          else G.next[d] = atomicExch(G.head + blocker, d);
        }
      }
    }
    grid.sync();
    // ---- phase B: the ready donors decide (their touch sets are disjoint) and wake their waiters
    const uint32_t ready_end = *ready_n;
    // a claimed receiver: its waiters wake up; on several GPUs its other copies hear of it
    auto claimed = [&](uint32_t j, uint32_t d) {
      wake(j, G.work[pout], pout);
      if (!PEER) return;
      if (is_ghost(j)) {  // to the owner: partner = this donor's ghost over there
        const uint32_t os = __ldg(&C.owner_slot[j]);
        const int side = int(os >> 31);
        mail(side, os & kSlotMask, MAIL_CLAIM, __ldg(&C.rslot[side][d]));
      } else {
#pragma unroll
        for (int side = 0; side < 2; side++) {
          const uint32_t sl = __ldg(&C.rslot[side][j]);
          if (sl == 0xffffffffu || !C.nb_mbox[side]) continue;
          const uint32_t dd = __ldg(&C.rslot[side][d]);
          mail(side, sl, MAIL_PARTNER, dd != 0xffffffffu ? dd : CLAIMED_ELSEWHERE);
        }
      }
    };
    for (uint32_t w = ready_begin + gwarp; w < ready_end; w += nwarps) {
      const uint32_t d = __ldcg(G.ready + w);
      // a donor decides ONCE: the state word changes hands atomically (a second entry for the same donor — none is known to
      // arise, the counter below would say so — must not run the loop again over receivers the first run has claimed)
      uint32_t mine_to_decide = 0;
      if (lane == 0) {
        const uint32_t iv = __ldcg(G.info + d);
        mine_to_decide = ((iv & 0xffu) == GI_PENDING && atomicCAS(G.info + d, iv, (iv & ~0xffu) | GI_DONE) == iv) ? 1u : 0u;
        if (!mine_to_decide) atomicAdd(&ctl->greedy_duplicates, 1u);
      }
      if (!__shfl_sync(0xffffffffu, mine_to_decide, 0)) continue;
      const DonorRegs D = donor_regs(A, P, merging, G.dt, d);
      uint32_t mine;
      const uint32_t c = donor_loop<false>(A, P, merging, D, lane, mine, [&](uint32_t j) { if (lane == 0) claimed(j, d); });
      if (lane == 0) {
        A.counter[d] = c;
        G.info[d] = GI_DONE;
        if (c) atomicAdd(&ctl->n_claims, c);
        if (PEER) {
          mail_copies(d, MAIL_INFO, GI_DONE);
          if (c) { mail_copies(d, MAIL_PARTNER, DELETE_); mail_copies(d, MAIL_COUNTER, c); }
        }
      }
      const uint32_t c32 = min(c, 32u);
      if (lane < c32) claimed(mine, d);
      if (lane == (c32 & 31u)) wake(d, G.work[pout], pout);
    }
    ready_begin = ready_end;
    if (gtid == 0) ctl->work_n[pin] = 0u;  // consumed in phase A; it is the work list the round after next fills
    if (PEER) go = exchange(pout);
    else grid.sync();
  }
  if (gtid == 0) { ctl->greedy_done = 1; ctl->greedy_barriers = bar - (C.seq0 + 1u); }
}

// validate_share_partners particle_sharing.rs:113-150 / validate_merge_partners particle_merging.rs:230-268
__global__ void __launch_bounds__(kThreads)
k_validate(uint32_t n, AdaptArgs A, uint8_t donor_class, StepCtl* ctl) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || nb_ghost(A.L.cnt[i])) return;  // a ghost's bookkeeping is its owner's business
  int why = 0;  // which invariant broke (reported with the first offender: diagnostics only)
  uint32_t seen = 0;
  const uint32_t c = A.counter[i], p = A.partner[i];
  if (c > 0) {
    if (A.size_class[i] != donor_class) why = 1;
    else if (p != DELETE_) why = 2;
    uint32_t c2 = 0;
    const uint32_t cn = nb_cn(A.L.cnt[i]);
    const NbCol col(A.L, i);
    for (uint32_t k = 0; k < cn; k++) if (A.partner[col.get(k)] == i) c2++;
    if (c2 != c && !why) { why = 3; seen = c2; }
  } else {
    if (p == DELETE_) why = 4;
    else if (p != AVAILABLE && p != 0xFFFFFFFDu && A.partner[p] != DELETE_) { why = 5; seen = A.partner[p]; }
  }
  if (why) {
    if (atomicOr(&ctl->error_flags, ERRF_PARTNER_VALIDATION) == 0u || true) {
      if (atomicCAS(&ctl->validate_why, 0u, uint32_t(why)) == 0u) { ctl->validate_at[0] = i; ctl->validate_at[1] = c; ctl->validate_at[2] = p; ctl->validate_at[3] = seen; }
    }
  }
}

// receivers absorb their share (particle_sharing.rs:164-211, particle_merging.rs:282-324).  A receiver has exactly one
// donor and donors are never receivers, so this is order independent and runs in place.
__global__ void __launch_bounds__(kThreads)
k_apply_receivers(uint32_t n, AdaptArgs A, const PackedParams P, int merging, float dt) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || nb_ghost(A.L.cnt[i])) return;  // (the donor may be a ghost: its state arrived with the mail of the search)
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
  if (i >= n || nb_ghost(A.L.cnt[i])) return;
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
// (n particles on this device, n_ref reference indices in all: the same number unless the fluid is spread over several GPUs)
__global__ void k_compact(uint32_t n, uint32_t n_ref, const uint32_t* __restrict__ del_scan, const uint32_t* __restrict__ holes,
                          const uint32_t* __restrict__ keep_scan, const float2* __restrict__ pos, const float2* __restrict__ vel,
                          const float* __restrict__ mass, const float* __restrict__ level, const uint32_t* __restrict__ refid,
                          float2* __restrict__ pos_o, float2* __restrict__ vel_o, float* __restrict__ mass_o, float* __restrict__ level_o,
                          uint32_t* __restrict__ refid_o, StepCtl* ctl) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i == 0) ctl->n_new = keep_scan[n];
  if (keep_scan[i + 1] == keep_scan[i]) return;
  const uint32_t total_del = del_scan[n_ref];
  const uint32_t n_new = n_ref - total_del;
  uint32_t r = refid[i] & ~ASPH_GHOST_BIT;
  if (r >= n_new) {
    const uint32_t k = (n_ref - 1u - r) - (total_del - del_scan[r + 1]);  // survivors with a higher reference index
    r = holes[k];
  }
  const uint32_t o = keep_scan[i];
  pos_o[o] = pos[i]; vel_o[o] = vel[i]; mass_o[o] = mass[i]; level_o[o] = level[i]; refid_o[o] = r;
}

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
// ---- several GPUs: the reference index space is shared by all ranks.  Every rank lists what it deletes / splits, the
// lists are gathered (dist_allgather_list), and the dense per-index arrays above are built identically on every rank.
__global__ void k_mark_deleted_dist(uint32_t n, AdaptArgs A, int min_partners, uint32_t* __restrict__ list, uint32_t* __restrict__ keep, StepCtl* ctl) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i == 0) keep[n] = 0u;
  if (nb_ghost(A.L.cnt[i])) { keep[i] = 0u; return; }  // the compaction drops this step's ghosts as well
  const bool del = A.partner[i] == DELETE_ && int(A.counter[i]) >= min_partners;  // dropped_mass_merging == mass: nothing is left
  keep[i] = del ? 0u : 1u;
  if (del) list[atomicAdd(&ctl->list_n, 1u)] = A.refid[i];
}
// gathered[q * stride + k], k < counts[q] * words: `words` 32-bit words per entry, the first is a reference index
__global__ void k_scatter_gathered(int ranks, uint32_t stride, int words, const uint32_t* __restrict__ counts, const uint32_t* __restrict__ gathered,
                                   uint32_t* __restrict__ dense) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t per = stride / uint32_t(words);
  const uint32_t q = e / per, k = e - q * per;
  if (q >= uint32_t(ranks) || k >= counts[q]) return;
  const uint32_t* ent = gathered + size_t(q) * stride + size_t(k) * words;
  dense[ent[0]] = words == 1 ? 1u : ent[1];
}
__global__ void k_split_count_dist(uint32_t n, AdaptArgs A, const PackedParams P, int max_children, uint32_t* __restrict__ list, StepCtl* ctl) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || nb_ghost(A.L.cnt[i]) || A.size_class[i] != ASPH_CLASS_TOO_LARGE) return;
  unsigned int err = 0;
  const uint32_t extra = split_children(A.level[i], A.mass[i], P, max_children, &err) - 1u;
  if (err) atomicOr(&ctl->error_flags, err);
  atomicAdd(&ctl->n_split_parents, 1u);
  atomicAdd(&ctl->local_extra, extra);
  const uint32_t k = atomicAdd(&ctl->list_n, 1u);
  list[2 * k] = A.refid[i]; list[2 * k + 1] = extra;
}
// children of an owned parent: stored behind this rank's particles (any order: the next step sorts by cell and reference
// index), numbered n_ref + (children of all parents with a lower reference index) + c - 1 as in the reference
__global__ void k_split_apply_dist(uint32_t n, uint32_t n_ref, uint32_t cap, AdaptArgs A, const uint32_t* __restrict__ extra_scan, const PackedParams P,
                                   int max_children, const int* __restrict__ split_off, const float* __restrict__ split_pos, float* __restrict__ level,
                                   uint32_t* __restrict__ refid, StepCtl* ctl) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || nb_ghost(A.L.cnt[i]) || A.size_class[i] != ASPH_CLASS_TOO_LARGE) return;
  unsigned int err = 0;
  const float m = A.mass[i], lv = level[i];
  const uint32_t nc = split_children(lv, m, P, max_children, &err);
  if (nc < 2u) return;
  const float* pat = split_pos + 2 * size_t(split_off[nc - 2u]);
  const float radius = sqrtf((m / 1.0f) * ASPH_FRAC_1_PI_F);
  const float child_mass = m / float(nc);
  const float2 op = A.pos[i], ov = A.vel[i];
  const uint32_t first_ref = n_ref + extra_scan[refid[i]];
  const uint32_t base = atomicAdd(&ctl->n_new, nc - 1u);
  for (uint32_t c = 0; c < nc; c++) {
    const uint32_t t = (c == 0) ? i : base + (c - 1u);
    if (t >= cap) return;
    A.mass[t] = child_mass; A.vel[t] = ov; level[t] = lv;
    A.pos[t] = make_float2(op.x + pat[2 * c] * radius, op.y + pat[2 * c + 1] * radius);
    if (c > 0) refid[t] = first_ref + (c - 1u);
  }
}

// ---- splitting (splitting.rs:19-81) ---------------------------------------------------------------------------------
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
  return A;
}

int total_mass(asph_sim* sim, double* out) {
  double* acc = (double*)sim->scratch_f.p;
  CUDA_TRY(cudaMemsetAsync(acc, 0, sizeof(double), sim->stream));
  if (sim->n) {
    k_mass_sum<<<std::max(1, sim->sm_count * 4), kThreads, 0, sim->stream>>>(sim->n, sim->mass[sim->cur].p, sim->refid[sim->cur].p, acc);
    LAUNCH_CHECK();
  }
  CUDA_TRY(cudaMemcpyAsync(out, acc, sizeof(double), cudaMemcpyDeviceToHost, sim->stream));
  CUDA_TRY(cudaStreamSynchronize(sim->stream));
  if (sim->dist) TRY(dist_allreduce_host(sim, out, 1, 1));  // the whole fluid, not this slab's share
  return ASPH_OK;
}

int ensure_greedy_buffers(asph_sim* sim) {
  CUDA_TRY(sim->g_info.ensure(sim->cap)); CUDA_TRY(sim->g_drop.ensure(sim->cap)); CUDA_TRY(sim->g_head.ensure(sim->cap));
  CUDA_TRY(sim->g_next.ensure(sim->cap)); CUDA_TRY(sim->g_resume.ensure(sim->cap));
  return ASPH_OK;
}

int classify(asph_sim* sim) {
  const uint32_t n = sim->n;
  k_classify<<<(n + kThreads - 1) / kThreads, kThreads, 0, sim->stream>>>(n, sim->level[sim->cur].p, sim->mass[sim->cur].p, sim->pp,
                                                                          sim->size_class.p);
  LAUNCH_CHECK();
  return ASPH_OK;
}

// the greedy partner search (k_greedy); returns the number of claimed receivers
int find_partners(asph_sim* sim, bool merging, float dt, uint32_t* claims) {
  const uint32_t n = sim->n;
  cudaStream_t st = sim->stream;
  const uint32_t blocks = (n + kThreads - 1) / kThreads;
  const uint8_t donor_class = merging ? ASPH_CLASS_TOO_SMALL : ASPH_CLASS_LARGE;
  k_partner_ctl_reset<<<1, 1, 0, st>>>(sim->ctl);
  LAUNCH_CHECK();
  const bool peer = dist_p2p(sim);
  int& co_resident = peer ? sim->greedy_grid_peer : sim->greedy_grid;
  if (co_resident == 0) {  // co-resident blocks of the persistent kernel on this device
    int per_sm = 0;
    if (peer) CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_greedy<true>, kGreedyThreads, 0));
    else CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_greedy<false>, kGreedyThreads, 0));
    co_resident = std::max(1, std::min(per_sm, 2) * sim->sm_count);
  }
  GreedyArgs G;
  G.A = args_of(sim);
  G.info = sim->g_info.p; G.drop = sim->g_drop.p; G.head = sim->g_head.p; G.next = sim->g_next.p; G.resume = sim->g_resume.p;
  G.work[0] = sim->work[0].p; G.work[1] = sim->work[1].p; G.ready = sim->cand.p;
  G.ctl = sim->ctl; G.n = n; G.merging = merging ? 1 : 0; G.dt = dt; G.donor_class = donor_class;
  PackedParams pp = sim->pp;
  CoopPeer coop = dist_coop_peer(sim);
  const uint32_t grid = uint32_t(std::max(1, std::min<int>(co_resident, int((n + kGreedyThreads - 1) / kGreedyThreads))));
  void* args[] = {&G, &pp, &coop};
  cudaEvent_t kt0 = nullptr, kt1 = nullptr;
  if (sim->kt_every > 0) { kt0 = kt_event(sim); kt1 = kt_event(sim); cudaEventRecord(kt0, st); }
  CUDA_TRY(cudaLaunchCooperativeKernel(peer ? (void*)k_greedy<true> : (void*)k_greedy<false>, dim3(grid), dim3(kGreedyThreads), args, 0, st));
  sim->kernel_launches++;
  if (kt1) cudaEventRecord(kt1, st);
  const AdaptArgs A = G.A;
  k_validate<<<blocks, kThreads, 0, st>>>(n, A, donor_class, sim->ctl);
  LAUNCH_CHECK();
  const int rc_sync = sync_ctl(sim);
  if (kt1) {
    float ms = 0.f;
    if (rc_sync == ASPH_OK && cudaEventElapsedTime(&ms, kt0, kt1) == cudaSuccess) { sim->kt_ms[ASPH_KT_PARTNER_SEARCH] += ms; sim->kt_samples[ASPH_KT_PARTNER_SEARCH]++; }
    kt_release(sim, kt0); kt_release(sim, kt1);
  }
  TRY(rc_sync);
  if (peer) dist_coop_advance(sim, sim->ctl_host->greedy_barriers);  // the same number on every rank
  if (sim->ctl_host->error_flags & ERRF_PEER_TIMEOUT) return check_error_flags(sim);
  if (!sim->ctl_host->greedy_done) { sim->last_error = "partner search did not terminate"; return ASPH_ERR_INVALID; }
  sim->adapt_rounds += sim->ctl_host->rounds;
  sim->greedy_duplicates += sim->ctl_host->greedy_duplicates;
  unsigned long long c = sim->ctl_host->n_claims;
  if (sim->dist) TRY(dist_allreduce_host(sim, &c, 1, 0));  // every rank reports (and branches on) the claims of the whole fluid
  *claims = uint32_t(c);
  return ASPH_OK;
}

// ---- several GPUs: deletion and splitting in the reference-index space all ranks share ------------------------------
int merge_compact_dist(asph_sim* sim, const AdaptArgs& A, int min_partners) {
  cudaStream_t st = sim->stream;
  const uint32_t n = sim->n, blocks = (n + kThreads - 1) / kThreads;
  const uint32_t n_ref = uint32_t(dist_n_global(sim));
  uint32_t* list = sim->scratch_u[0].p;   // reference indices this rank deletes
  uint32_t* keep = sim->scratch_u[1].p;   // n + 1
  CUDA_TRY(cudaMemsetAsync(&sim->ctl->list_n, 0, sizeof(uint32_t), st));
  k_mark_deleted_dist<<<blocks, kThreads, 0, st>>>(n, A, min_partners, list, keep, sim->ctl);
  LAUNCH_CHECK();
  TRY(sync_ctl(sim));
  const uint32_t* gathered = nullptr; const uint32_t* counts = nullptr;
  uint32_t stride = 0; unsigned long long total = 0;
  TRY(dist_allgather_list(sim, list, sim->ctl_host->list_n, 1, &gathered, &counts, &stride, &total));
  if (total == 0) return ASPH_OK;  // every donor found fewer partners than minimum_merge_partners: nothing is removed anywhere
  uint32_t *del_ref = nullptr, *holes = nullptr;
  TRY(dist_ref_buffers(sim, size_t(n_ref) + 1, &del_ref, &holes));
  CUDA_TRY(cudaMemsetAsync(del_ref, 0, (size_t(n_ref) + 1) * sizeof(uint32_t), st));
  const int R = dist_ranks(sim);
  k_scatter_gathered<<<(uint32_t(R) * stride + kThreads - 1) / kThreads, kThreads, 0, st>>>(R, stride, 1, counts, gathered, del_ref);
  LAUNCH_CHECK();
  CUDA_TRY(cudaMemcpyAsync(&sim->ctl->n_new, &n_ref, sizeof(uint32_t), cudaMemcpyHostToDevice, st));  // scan lengths
  TRY(launch_exclusive_scan(sim, del_ref, del_ref, &sim->ctl->n_new, 1, n_ref + 1));
  CUDA_TRY(cudaMemcpyAsync(&sim->ctl->n_new, &n, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  TRY(launch_exclusive_scan(sim, keep, keep, &sim->ctl->n_new, 1, n + 1));
  k_holes<<<(n_ref + kThreads - 1) / kThreads, kThreads, 0, st>>>(n_ref, del_ref, holes);
  LAUNCH_CHECK();
  const int c = sim->cur;
  k_compact<<<blocks, kThreads, 0, st>>>(n, n_ref, del_ref, holes, keep, sim->pos[c].p, sim->vel[c].p, sim->mass[c].p, sim->level[c].p,
                                         sim->refid[c].p, sim->pos[1 - c].p, sim->vel[1 - c].p, sim->mass[1 - c].p, sim->level[1 - c].p,
                                         sim->refid[1 - c].p, sim->ctl);
  LAUNCH_CHECK();
  TRY(sync_ctl(sim));
  TRY(check_error_flags(sim));
  sim->lists_valid = false; sim->step_fields_valid = false;  // this step's ghosts are gone as well
  sim->cur = 1 - c;
  sim->n = sim->ctl_host->n_new; sim->n_owned = sim->n;
  dist_set_n_global(sim, uint64_t(n_ref) - total);
  return ASPH_OK;
}

int split_dist(asph_sim* sim) {
  cudaStream_t st = sim->stream;
  const PackedParams& P = sim->pp;
  const uint32_t n = sim->n, blocks = (n + kThreads - 1) / kThreads;
  const uint32_t n_ref = uint32_t(dist_n_global(sim));
  uint32_t* list = sim->scratch_u[0].p;  // (reference index, extra children) per parent: 2 words, at most n / 2 parents fit... a parent is one of n particles
  TRY(classify(sim));
  CUDA_TRY(cudaMemsetAsync(&sim->ctl->n_split_parents, 0, sizeof(uint32_t), st));
  CUDA_TRY(cudaMemsetAsync(&sim->ctl->list_n, 0, sizeof(uint32_t), st));
  CUDA_TRY(cudaMemsetAsync(&sim->ctl->local_extra, 0, sizeof(uint32_t), st));
  {
    const AdaptArgs A = args_of(sim);
    k_split_count_dist<<<blocks, kThreads, 0, st>>>(n, A, P, sim->max_children, reinterpret_cast<uint32_t*>(sim->scratch_f.p), sim->ctl);
    LAUNCH_CHECK();
  }
  TRY(sync_ctl(sim));
  TRY(check_error_flags(sim));
  unsigned long long parents = sim->ctl_host->n_split_parents;
  TRY(dist_allreduce_host(sim, &parents, 1, 0));
  sim->info.n_split_parents = int(parents);
  const uint32_t local_extra = sim->ctl_host->local_extra, n_parents_local = sim->ctl_host->list_n;
  (void)list;
  const uint32_t* gathered = nullptr; const uint32_t* counts = nullptr;
  uint32_t stride = 0; unsigned long long total = 0;
  TRY(dist_allgather_list(sim, reinterpret_cast<uint32_t*>(sim->scratch_f.p), n_parents_local, 2, &gathered, &counts, &stride, &total));
  if (total == 0) return ASPH_OK;
  // the gathered buffers and the dense arrays live in the distributed state: growing the particle arrays keeps them
  if (uint64_t(n) + local_extra > sim->cap) {
    TRY(ensure_capacity(sim, uint32_t(std::min<uint64_t>(uint64_t(n) + local_extra + (uint64_t(n) + local_extra) / 4 + 1024, 0x7FFFFFF0ull))));
    TRY(ensure_greedy_buffers(sim));
    TRY(classify(sim));  // size_class was reallocated
  }
  uint32_t *extra_ref = nullptr, *unused = nullptr;
  TRY(dist_ref_buffers(sim, size_t(n_ref) + 1, &extra_ref, &unused));
  CUDA_TRY(cudaMemsetAsync(extra_ref, 0, (size_t(n_ref) + 1) * sizeof(uint32_t), st));
  const int R = dist_ranks(sim);
  k_scatter_gathered<<<(uint32_t(R) * (stride / 2u) + kThreads - 1) / kThreads, kThreads, 0, st>>>(R, stride, 2, counts, gathered, extra_ref);
  LAUNCH_CHECK();
  CUDA_TRY(cudaMemcpyAsync(&sim->ctl->n_new, &n_ref, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  TRY(launch_exclusive_scan(sim, extra_ref, extra_ref, &sim->ctl->n_new, 1, n_ref + 1));
  CUDA_TRY(cudaMemcpyAsync(&sim->ctl->n_new, &n, sizeof(uint32_t), cudaMemcpyHostToDevice, st));  // children are stored from here on
  const int c = sim->cur;
  const AdaptArgs A = args_of(sim);
  k_split_apply_dist<<<blocks, kThreads, 0, st>>>(n, n_ref, sim->cap, A, extra_ref, P, sim->max_children, sim->split_off.p, sim->split_pos.p,
                                                  sim->level[c].p, sim->refid[c].p, sim->ctl);
  LAUNCH_CHECK();
  uint32_t total_extra = 0;
  CUDA_TRY(cudaMemcpyAsync(&total_extra, extra_ref + n_ref, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  TRY(sync_ctl(sim));
  TRY(check_error_flags(sim));
  sim->lists_valid = false; sim->step_fields_valid = false;
  sim->n = sim->ctl_host->n_new; sim->n_owned += local_extra;
  dist_set_n_global(sim, uint64_t(n_ref) + total_extra);
  return ASPH_OK;
}

}  // namespace

int launch_adaptivity(asph_sim* sim, float dt) {
  const uint32_t n0 = sim->n;
  if (n0 == 0) return ASPH_OK;
  cudaStream_t st = sim->stream;
  TRY(ensure_greedy_buffers(sim));
  // several GPUs: the ghosts' advected velocities (their positions came with the level smoothing's halo, level.cu)
  if (sim->dist) TRY(dist_halo(sim, sim->vel[sim->cur].p, 8));
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
    if (claims) {  // nothing was claimed: no receiver, no donor changes
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
      if (sim->dist) {  // the merge search of this step reads the shared particles' new x, v, m: their ghost copies follow
        const int c = sim->cur;
        TRY(dist_halo(sim, sim->pos[c].p, 8)); TRY(dist_halo(sim, sim->vel[c].p, 8)); TRY(dist_halo_words(sim, sim->mass[c].p));
      }
    }
  }
  if (sim->step_number % 2 == 0) {
    if (sim->merge_enabled) {  // simulation.rs:2760-2774
      TRY(classify(sim));
      uint32_t claims = 0;
      TRY(find_partners(sim, true, dt, &claims));
      sim->info.n_merged = int(claims);
      if (claims == 0) {  // no receiver was claimed: no donor is removed, the particle set stays as it is
        if (track_cls) {
          k_cls_copy<<<(sim->n + kThreads - 1) / kThreads, kThreads, 0, st>>>(sim->n, sim->size_class.p, sim->cls[sim->cur].p);
          LAUNCH_CHECK();
          cls_done = true;
        }
      } else if (sim->dist) {
        const AdaptArgs A = args_of(sim);
        k_apply_receivers<<<(sim->n + kThreads - 1) / kThreads, kThreads, 0, st>>>(sim->n, A, P, 1, dt);
        LAUNCH_CHECK();
        TRY(merge_compact_dist(sim, A, P.min_merge_partners));
      } else {
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
      k_compact<<<blocks, kThreads, 0, st>>>(n, n, del_ref, holes, keep, sim->pos[c].p, sim->vel[c].p, sim->mass[c].p, sim->level[c].p,
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
    }
  } else if (sim->split_enabled) {  // simulation.rs:2775-2788
    if (sim->max_children < 2) { sim->last_error = "splitting enabled but no split patterns were given"; return ASPH_ERR_INVALID; }
    if (sim->dist) {
      TRY(split_dist(sim));
    } else {
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
      TRY(ensure_greedy_buffers(sim));
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
