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

#include <type_traits>

#include "lists.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kThreads = 256;
#ifndef ASPH_PROP_BATCH
#define ASPH_PROP_BATCH 4  // neighbours a lane of k_propagate has in flight
#endif
constexpr int kPropThreads = 512;                 // block of the persistent propagation kernel
constexpr unsigned int kUnassigned = 0xFFFFFFFFu;  // bit pattern of a level value no push has reached yet
constexpr int kGhostPending = -3;                  // stamp of a ghost copy its owner has not assigned yet (multi-GPU)
constexpr int kBorderUnassigned = -2;              // stamp of an unassigned particle that has ghost copies on a neighbour GPU

typedef NbLists Lists;

#ifdef ASPH_PROP_TRACE  // development only: per-sweep cycle counts of k_propagate (tools/prop_trace.py)
__device__ unsigned long long g_prop_trace[12][512];
#define PROP_TRACE(stmt) stmt
#else
#define PROP_TRACE(stmt)
#endif

// front_n[p] = tail of the front array of the sweeps of parity p; level_live[p] = the last sweep of parity p that assigned
// a value above the cutoff (k_propagate)
__global__ void k_level_reset(StepCtl* ctl) {
  ctl->front_n[0] = 0; ctl->front_n[1] = 0; ctl->cand_n[0] = 0; ctl->cand_n[1] = 0;
  ctl->level_live[0] = 0; ctl->level_live[1] = 0;
  ctl->level_sweep = 0; ctl->level_done = 0;
  ctl->mail_n[0] = 0; ctl->mail_n[1] = 0;
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
// CenterDiff detects surface particles with values of either sign, and the propagation keeps the LARGEST candidate with an
// unsigned atomicMin on the stored word.  For values <= 0 the float's own bit pattern orders that way (EmptyAngle: the
// array holds plain floats throughout); in signed mode (SIGNED) the array holds keys during the propagation — a negative
// float's pattern, 0x7FFFFFFF - pattern for a positive one — and k_propagate turns them back into floats at its end.
__device__ __forceinline__ unsigned int level_key(float v) {
  const unsigned int b = __float_as_uint(v);
  return (b & 0x80000000u) ? b : 0x7FFFFFFFu - b;
}
__device__ __forceinline__ float level_of_key(unsigned int k) { return __uint_as_float((k & 0x80000000u) ? k : 0x7FFFFFFFu - k); }

// K3': surface_detection_by_center_diff (simulation.rs:631-695) over the current lists (all ce entries count as
// neighbours; the kernel weight of those beyond the 2h support is zero).  Writes keys (see above).
__global__ void __launch_bounds__(kThreads)
k_center_diff(uint32_t n, Lists L, const float4* __restrict__ xyhm, const PackedParams P, float* __restrict__ level, int* __restrict__ stamp,
              uint8_t* __restrict__ flags, uint32_t* __restrict__ front, StepCtl* ctl) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool surf = false;
  if (i < n) {
    const float4 me = xyhm[i];
    const uint32_t ce = L.cnt_ext[i];
    const NbCol col(L, i);
    float wsum = 0.f, avg_r = 0.f, cx = 0.f, cy = 0.f;
    for (uint32_t k = 0; k < ce; k++) {
      const float4 o = __ldg(&xyhm[col.get(k)]);
      const float vol = o.w / P.rest_density;
      const float rad = volume_to_radius(vol);
      const float dx = me.x - o.x, dy = me.y - o.y;
      const float w = kernel_w(sqrtf(dx * dx + dy * dy), (me.z + o.z) * 0.5f) * vol;
      cx += o.x * w; cy += o.y * w;
      avg_r += rad * w;
      wsum += w;
    }
    avg_r /= wsum;
    const float surface_level = -0.85f * avg_r;
    float phi;
    if (ce < 5u) phi = surface_level;
    else {
      cx /= wsum; cy /= wsum;
      const float dx = me.x - cx, dy = me.y - cy;
      phi = sqrtf(dx * dx + dy * dy) - avg_r;
    }
    surf = phi >= surface_level;
    reinterpret_cast<unsigned int*>(level)[i] = surf ? level_key(phi) : kUnassigned;
    stamp[i] = surf ? 0 : -1;
    flags[i] = surf ? 1u : 0u;
  }
  const unsigned int mask = __ballot_sync(0xffffffffu, surf);
  if (!mask) return;
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t base = 0;
  if (lane == 0) base = atomicAdd(&ctl->front_n[0], uint32_t(__popc(mask)));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (surf) front[base + uint32_t(__popc(mask & ((1u << lane) - 1u)))] = i;
}

__global__ void __launch_bounds__(kThreads)
k_surface(uint32_t n, Lists L, const float4* __restrict__ xyhm, const float2* __restrict__ nrm, const PackedParams P, float cos_threshold,
          float* __restrict__ level, int* __restrict__ stamp, uint8_t* __restrict__ flags, uint32_t* __restrict__ front, StepCtl* ctl) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool surf = i < n && surface_of(i, L, xyhm, nrm, P, cos_threshold, level, stamp, flags);
  // multi-GPU: a ghost's neighbourhood is incomplete here; its state arrives from the rank that owns it (k_ghost_front)
  if (surf && nb_ghost(__ldg(&L.cnt[i]))) surf = false;
  // front(0): one atomic per warp
  const unsigned int mask = __ballot_sync(0xffffffffu, surf);
  if (!mask) return;
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t base = 0;
  if (lane == 0) base = atomicAdd(&ctl->front_n[0], uint32_t(__popc(mask)));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (surf) front[base + uint32_t(__popc(mask & ((1u << lane) - 1u)))] = i;
}

// multi-GPU, after the owners' level / stamp values of the ghosts have arrived: ghosts on the detected surface join front(0)
// The other ghosts get the stamp kGhostPending: no local push may assign them (k_propagate), their owners' values come by mail.
__global__ void k_ghost_front(uint32_t count, const uint32_t* __restrict__ ghost_idx, int* __restrict__ stamp, uint32_t* __restrict__ front,
                              StepCtl* ctl) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const uint32_t i = ghost_idx[k];
  if (stamp[i] == 0) front[atomicAdd(&ctl->front_n[0], 1u)] = i;
  else stamp[i] = kGhostPending;
}

// multi-GPU: unassigned particles with ghost copies on a neighbour GPU get their own stamp, so that the push that claims
// one knows without another load that the value has to be mailed
__global__ void k_mark_border(uint32_t n, const uint32_t* __restrict__ rslot0, const uint32_t* __restrict__ rslot1, int* __restrict__ stamp) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if ((rslot0[i] != 0xffffffffu || rslot1[i] != 0xffffffffu) && stamp[i] == -1) stamp[i] = kBorderUnassigned;
}

// K4 (simulation.rs:739-800) as one persistent cooperative kernel.  One warp per front particle j, one lane per
// neighbour i; sweep t reads front(t - 1) and appends front(t).
//   stamp[i]: -1 unassigned, otherwise the sweep that assigned i (0 = detected surface)
// The fronts of even sweeps live in front0, those of odd sweeps in front1, each array filled once from the start
// (front_n[p] = its tail), and level_live[p] = the last sweep of parity p that assigned a value above the cutoff: what
// sweep t reads is complete when the barrier before it falls, and nothing a block may already write in sweep t
// (parity t) touches what a slower block still has to read at the top of sweep t (parity t - 1).
// Values other SMs wrote in earlier sweeps (stamp, level, the front entries) are read with ld.global.cg: an L1 line
// fetched in an earlier sweep may hold their previous contents.
// PEER (several GPUs, x-slabs with ghost copies of the neighbours' border particles): ghosts are never assigned here —
// their neighbourhoods are incomplete; the owner's value comes in.  After the pushes of sweep t every newly assigned
// border particle is mailed to its ghost copies (CoopPeer, sim.cuh), the GPUs meet in a barrier that also tells every
// rank whether any of them assigned anything (above the cutoff), and the mail — slot, value — joins the local front(t).
template <bool PEER, bool SIGNED>
__global__ void __launch_bounds__(kPropThreads, 2)
k_propagate(uint32_t n, Lists L, const float4* __restrict__ xyhm, float* __restrict__ level, int* __restrict__ stamp,
            uint32_t* __restrict__ front0, uint32_t* __restrict__ front1, uint32_t* __restrict__ border, StepCtl* ctl, float neg_dmax,
            int use_cutoff, const CoopPeer P) {
  cg::grid_group grid = cg::this_grid();
  constexpr uint32_t kStage = 192;
  constexpr uint32_t kBatch = ASPH_PROP_BATCH;
  constexpr uint32_t kGroupRounds = 4;  // columns of up to kGroupRounds * kBatch * lanes neighbours are walked by the particle's own lanes
  __shared__ uint32_t s_stage[kPropThreads / 32][kStage];
  __shared__ uint32_t s_count[kPropThreads / 32], s_base;
  unsigned int* level_bits = reinterpret_cast<unsigned int*>(level);
  const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
  const uint32_t gwarp = gtid >> 5, nwarps = gthreads >> 5;
  volatile uint32_t* tail = ctl->front_n;
  volatile int* live_sweep = ctl->level_live;
  uint32_t consumed[2] = {0u, 0u};  // how much of each parity's array earlier sweeps have read
  uint32_t seen[2] = {0u, 0u};      // PEER: the tail at the end of the parity's previous sweep (front(0): the halo exchange delivered it)
  if (PEER) seen[0] = tail[0];
  bool go = true;                   // PEER: the verdict of the last barrier
  int sweeps = 0;
  for (int t = 1; t < (1 << 30); t++) {
    const int pin = (t - 1) & 1, pout = t & 1;
    const uint32_t begin = consumed[pin], end = tail[pin];
    // the sweep runs if the previous one assigned anything (above the cutoff); the reference's last sweep changes nothing
    if (PEER ? !go : (end == begin || (t > 1 && live_sweep[pin] != t - 1) || t > int(n) + 1)) break;
    consumed[pin] = end;
    sweeps = t;
    const uint32_t* __restrict__ fin = pin ? front1 : front0;
    uint32_t* __restrict__ fout = pout ? front1 : front0;
    bool live = false;
    uint32_t wcount = 0;  // claims staged by this warp in this sweep (warp-uniform)
    PROP_TRACE(const long long tr0 = clock64();)
    // a staged claim goes to front(t); a border particle (PEER) also to the list block 0 mails from after the barrier
    auto flush = [&](uint32_t* __restrict__ dst, uint32_t at, uint32_t v) {
      dst[at] = v & 0x7fffffffu;
      if (PEER && (v & 0x80000000u)) {  // with its ghost slots: block 0 then needs the value only
        const uint32_t i = v & 0x7fffffffu, k = atomicAdd(&ctl->cand_n[0], 1u);
        border[3u * k] = i; border[3u * k + 1u] = __ldg(&P.rslot[0][i]); border[3u * k + 2u] = __ldg(&P.rslot[1][i]);
      }
    };
    // Up to 4 * stride neighbours of front particle j, four per lane requested together: lane handles k0 + first + stride * u.
    // Called by the whole warp with warp-uniform trip counts (ballots inside); a lane without a particle passes ce = 0.
    auto push = [&](float mex, float mey, float lj, uint32_t ce, const NbCol& col, uint32_t k0, uint32_t first, uint32_t stride) {
      uint32_t iu[kBatch];
      bool cand[kBatch];
#pragma unroll
      for (int u = 0; u < int(kBatch); u++) {
        const uint32_t k = k0 + first + stride * uint32_t(u);
        cand[u] = k < ce;
        iu[u] = cand[u] ? col.get(k) : 0u;
      }
      // the stamp (L2: other SMs write it) and the position (never changes: through L1, neighbouring front particles share
      // most of their neighbours) of a neighbour are requested together: one round trip less in the chain
      int su[kBatch];
      float2 ou[kBatch];
#pragma unroll
      for (int u = 0; u < int(kBatch); u++) {
        su[u] = cand[u] ? __ldcg(stamp + iu[u]) : 0;
        ou[u] = cand[u] ? __ldg(reinterpret_cast<const float2*>(xyhm + iu[u])) : make_float2(0.f, 0.f);
      }
      bool won[kBatch];
#pragma unroll
      for (int u = 0; u < int(kBatch); u++) {
        won[u] = false;
        // a ghost is never a candidate: its stamp is kGhostPending, or the sweep its mail came in
        if (cand[u] && (su[u] == -1 || su[u] == t || (PEER && su[u] == kBorderUnassigned))) {
          const float d = __fsqrt_rn(dist_sq_exact(__fsub_rn(mex, ou[u].x), __fsub_rn(mey, ou[u].y)));
          const float v = __fsub_rn(lj, d);  // <= 0 (EmptyAngle): the largest float is the smallest bit pattern
          atomicMin(level_bits + iu[u], SIGNED ? level_key(v) : __float_as_uint(v));
          if (su[u] != t) won[u] = atomicCAS(stamp + iu[u], su[u], t) == su[u];
          if (!use_cutoff || v > neg_dmax) live = true;
        }
      }
      // the newly claimed particles go into this warp's staging buffer: the tail of front(t) is ONE word, and an atomic per
      // warp and batch on it (some ten thousand per sweep, all to the same address) was most of a sweep's time
#pragma unroll
      for (int u = 0; u < int(kBatch); u++) {
        const unsigned int mask = __ballot_sync(0xffffffffu, won[u]);
        if (mask) {
          const uint32_t cnt = uint32_t(__popc(mask));
          if (wcount + cnt > kStage) {  // full (a sweep rarely claims more than a few dozen per warp): to the front right away
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&ctl->front_n[pout], wcount);
            base = __shfl_sync(0xffffffffu, base, 0);
            for (uint32_t e = lane; e < wcount; e += 32u) flush(fout, base + e, s_stage[wid][e]);
            __syncwarp();
            wcount = 0;
          }
          if (won[u])  // bit 31: a border particle, its ghost copies on the neighbour GPUs need the value
            s_stage[wid][wcount + uint32_t(__popc(mask & ((1u << lane) - 1u)))] = iu[u] | ((PEER && su[u] == kBorderUnassigned) ? 0x80000000u : 0u);
          wcount += cnt;
        }
      }
    };
    // A sweep is a chain of dependent memory round trips (front entry -> column header -> list entry -> stamp -> atomics)
    // ended by a grid-wide barrier, so its time is the time of the SLOWEST warp.  Eight lanes per front particle, four
    // particles per warp, 32 neighbours of each per round, the four columns walked side by side; a very long column (a
    // coarse particle next to fine ones has hundreds) would take its eight lanes many rounds, so it is handed to the whole
    // warp afterwards.
    // Lanes per front particle: eight (four particles per warp) when the front has more particles than the grid has warps
    // to give them two each — the usual case on one GPU —, sixteen or all thirty-two when it is small (late sweeps, small
    // scenes, the slabs of a multi-GPU run): fewer rounds in the chain of the warp that the sweep waits for.  (A lane count
    // chosen at run time inside ONE loop cost 4 % on one GPU: three instantiations.)
    auto walk = [&](auto lanes_c) {
      constexpr uint32_t kLanes = decltype(lanes_c)::value, kPerWarp = 32u / kLanes;
      const uint32_t sub = lane & (kLanes - 1u), grp = lane / kLanes;
      for (uint32_t f0 = begin + gwarp * kPerWarp; f0 < end; f0 += nwarps * kPerWarp) {
        struct { uint32_t j, ce; float x, y, lj; } cur{0u, 0u, 0.f, 0.f, 0.f};
        NbCol col;
        if (f0 + grp < end) {
          cur.j = __ldcg(fin + f0 + grp);
          const float2 p = __ldg(reinterpret_cast<const float2*>(xyhm + cur.j));
          cur.x = p.x; cur.y = p.y;
          cur.lj = SIGNED ? level_of_key(__ldcg(level_bits + cur.j)) : __ldcg(level + cur.j);
          cur.ce = __ldg(&L.cnt_ext[cur.j]);
          col = NbCol(L, cur.j);
        }
        // an extended-range column holds about 60 neighbours (f_ext = 2.9 supports): two rounds of kBatch * kLanes for the group
        const bool big = cur.ce > kGroupRounds * kBatch * kLanes;
        const unsigned int bigmask = __ballot_sync(0xffffffffu, big && sub == 0u);
        const uint32_t ce_grp = big ? 0u : cur.ce;
        for (uint32_t k0 = 0; __any_sync(0xffffffffu, k0 < ce_grp); k0 += kBatch * kLanes) push(cur.x, cur.y, cur.lj, ce_grp, col, k0, sub, kLanes);
        for (unsigned int m = bigmask; m; m &= m - 1u) {
          const int src = __ffs(m) - 1;
          const uint32_t jb = __shfl_sync(0xffffffffu, cur.j, src);
          const float2 pb = __ldg(reinterpret_cast<const float2*>(xyhm + jb));
          const float ljb = SIGNED ? level_of_key(__ldcg(level_bits + jb)) : __ldcg(level + jb);
          const uint32_t ceb = __ldg(&L.cnt_ext[jb]);
          const NbCol colb(L, jb);
          for (uint32_t k0 = 0; k0 < ceb; k0 += 32u * kBatch) push(pb.x, pb.y, ljb, ceb, colb, k0, lane, 32u);
        }
      }
    };
    if (end - begin <= nwarps) walk(std::integral_constant<uint32_t, 32u>());
    else if (end - begin <= 2u * nwarps) walk(std::integral_constant<uint32_t, 16u>());
    else walk(std::integral_constant<uint32_t, 8u>());
    PROP_TRACE(if (t < 512 && lane == 0) { const unsigned long long d = (unsigned long long)(clock64() - tr0); atomicMax(&g_prop_trace[1][t], d); atomicAdd(&g_prop_trace[2][t], d); if (gtid == 0) g_prop_trace[0][t] = end - begin; })
    // one atomic per block: the warps' staged claims behind one another at the tail of front(t)
    __syncwarp();
    if (lane == 0) s_count[wid] = wcount;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t total = 0;
      for (int w = 0; w < kPropThreads / 32; w++) { const uint32_t c = s_count[w]; s_count[w] = total; total += c; }
      s_base = total ? atomicAdd(&ctl->front_n[pout], total) : 0u;
    }
    __syncthreads();
    {
      const uint32_t base = s_base + s_count[wid];
      for (uint32_t e = lane; e < wcount; e += 32u) flush(fout, base + e, s_stage[wid][e]);
    }
    if (__any_sync(0xffffffffu, live) && lane == 0) live_sweep[pout] = t;
    PROP_TRACE(if (t < 512 && gtid == 0) g_prop_trace[3][t] = (unsigned long long)(clock64() - tr0);)
    grid.sync();
    PROP_TRACE(if (t < 512 && gtid == 0) { g_prop_trace[4][t] = (unsigned long long)(clock64() - tr0); g_prop_trace[5][t] = nwarps; })
    if (PEER) {
      // Between two grid barriers block 0 alone talks to the other GPUs (a sweep assigns a few dozen border particles, and
      // every grid barrier less is more than a microsecond of the sweep): mail out, barrier across the GPUs, mail in.
      if (blockIdx.x == 0) {
        // (the block's own message counters and its private view of the front's tail: every round trip to ctl saved here
        // is a microsecond of every sweep)
        __shared__ uint32_t s_mail[2], s_tail;
        const unsigned int seq = P.seq0 + unsigned(t), par = seq & 1u;
        if (threadIdx.x < 2u) s_mail[threadIdx.x] = 0u;
        uint32_t tail_now = 0;
        int live_now = 0;
        if (threadIdx.x == 0) { tail_now = tail[pout]; live_now = live_sweep[pout]; }  // in flight while the mail goes out
        const uint32_t nb = *reinterpret_cast<volatile uint32_t*>(&ctl->cand_n[0]);
        __syncthreads();
        for (uint32_t e = threadIdx.x; e < nb; e += blockDim.x) {
          const uint32_t i = __ldcg(border + 3u * e), sl0 = __ldcg(border + 3u * e + 1u), sl1 = __ldcg(border + 3u * e + 2u);
          const unsigned int bits = __ldcg(level_bits + i);
#pragma unroll
          for (int side = 0; side < 2; side++) {
            const uint32_t sl = side ? sl1 : sl0;
            if (sl == 0xffffffffu || !P.nb_mbox[side]) continue;
            const uint32_t k = atomicAdd(&s_mail[side], 1u);
            if (k < P.mbox_cap) P.nb_mbox[side][size_t(par * 2u + uint32_t(1 - side)) * P.mbox_cap + k] = make_uint2(sl, bits);  // I am that neighbour's other side
            else atomicOr(&ctl->error_flags, ERRF_PEER_TIMEOUT);
          }
        }
        __syncthreads();
        PROP_TRACE(if (t < 512 && gtid == 0) { g_prop_trace[6][t] = (unsigned long long)(clock64() - tr0); g_prop_trace[10][t] = nb; })
        if (threadIdx.x == 0) {
          ctl->cand_n[0] = 0u;
          for (int side = 0; side < 2; side++)  // the counts go ahead of the barrier flag (ordered by its fence)
            if (P.nb_ctl[side]) *reinterpret_cast<volatile unsigned int*>(&P.nb_ctl[side]->mbox_n[par][1 - side]) = min(s_mail[side], P.mbox_cap);
          const unsigned int mine = (tail_now > seen[pout] ? 1u : 0u) | (live_now == t ? 2u : 0u);
          const unsigned int all = coop_barrier(P, seq, mine, ctl);
          *P.verdict = ((all & 3u) == 3u && !(all & 0x80000000u) && t <= (1 << 29)) ? 1u : 0u;
          s_tail = tail_now;
          PROP_TRACE(if (t < 512) g_prop_trace[7][t] = (unsigned long long)(clock64() - tr0);)
        }
        __syncthreads();
        // the mail of this sweep: the ghosts' values, and their place behind the local claims in front(t) (nobody else
        // appends now)
        const uint32_t n_in0 = min(*reinterpret_cast<volatile unsigned int*>(&P.self->mbox_n[par][0]), P.mbox_cap);
        const uint32_t n_in1 = min(*reinterpret_cast<volatile unsigned int*>(&P.self->mbox_n[par][1]), P.mbox_cap);
#pragma unroll
        for (int side = 0; side < 2; side++) {
          const uint32_t n_in = side ? n_in1 : n_in0, at = s_tail + (side ? n_in0 : 0u);
          const uint2* __restrict__ box = P.mbox + size_t(par * 2u + uint32_t(side)) * P.mbox_cap;
          for (uint32_t e = threadIdx.x; e < n_in; e += blockDim.x) {
            const uint2 m = __ldcg(box + e);
            level_bits[m.x] = m.y;
            stamp[m.x] = t;
            fout[at + e] = m.x;
          }
        }
        if (threadIdx.x == 0 && n_in0 + n_in1 != 0u) ctl->front_n[pout] = s_tail + n_in0 + n_in1;
      }
      PROP_TRACE(if (t < 512 && gtid == 0) g_prop_trace[8][t] = (unsigned long long)(clock64() - tr0);)
      grid.sync();
      PROP_TRACE(if (t < 512 && gtid == 0) g_prop_trace[9][t] = (unsigned long long)(clock64() - tr0);)
      go = *reinterpret_cast<volatile unsigned int*>(P.verdict) != 0u;
      seen[pout] = tail[pout];
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { ctl->level_sweep = sweeps; ctl->level_done = 1; }
  // particles no push has reached stay FluidInterior
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned int b = __ldcg(level_bits + i);
    if (b == kUnassigned) level[i] = ASPH_LEVEL_INTERIOR;
    else if (SIGNED) level[i] = level_of_key(b);
  }
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
  for (uint32_t k0 = 0; k0 < cn; k0 += 4u) {  // four neighbours' records requested together, summed in list order
    uint32_t j[4];
    float2 xj[4];
    float4 o[4];
    float lj[4], rj[4];
#pragma unroll
    for (uint32_t u = 0; u < 4u; u++) j[u] = col.get(min(k0 + u, cn - 1u));
#pragma unroll
    for (uint32_t u = 0; u < 4u; u++) { xj[u] = __ldg(&pos[j[u]]); o[u] = __ldg(&xyhm[j[u]]); lj[u] = __ldg(&level[j[u]]); rj[u] = __ldg(&rho[j[u]]); }
#pragma unroll
    for (uint32_t u = 0; u < 4u; u++) {
      if (k0 + u >= cn) break;
      const float dx = xi.x - xj[u].x, dy = xi.y - xj[u].y;
      const float w = kernel_w(sqrtf(dx * dx + dy * dy), (hi + o[u].z) * 0.5f);
      const float dist = (lj[u] > 0.f) ? -dmax : fmaxf(lj[u], -dmax);
      const float vw = o[u].w / rj[u] * w;
      num += dist * vw;
      den += vw;
    }
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
  const bool peer = dist_p2p(sim);
  const bool center_diff = sim->pp.level_method == ASPH_LEVEL_CENTER_DIFF;
  if (center_diff)
    k_center_diff<<<blocks, kThreads, 0, st>>>(n, L, sim->xyhm.p, sim->pp, level, sim->stamp.p, sim->flags.p, sim->front[0].p, sim->ctl);
  else
    k_surface<<<blocks, kThreads, 0, st>>>(n, L, sim->xyhm.p, sim->nrm.p, sim->pp, cos_threshold, level, sim->stamp.p, sim->flags.p,
                                           sim->front[0].p, sim->ctl);
  LAUNCH_CHECK();
  if (center_diff && sim->dist) { sim->last_error = "level_estimation_method CenterDiff across GPU slabs"; return ASPH_ERR_UNSUPPORTED; }
  if (sim->dist && dist_ranks(sim) > 1 && !peer) { sim->last_error = "level estimation across GPU slabs needs the peer-memory path (ASPH_DIST_P2P)"; return ASPH_ERR_UNSUPPORTED; }
  if (peer) {  // the owners' verdict on the ghosts (their own neighbourhoods are incomplete here), then front(0) with them
    TRY(dist_halo_words(sim, level));
    TRY(dist_halo_words(sim, sim->stamp.p));
    uint32_t n_ghost = 0;
    const uint32_t* ghost_idx = dist_ghost_index(sim, &n_ghost);
    if (n_ghost) {
      k_ghost_front<<<(n_ghost + kThreads - 1) / kThreads, kThreads, 0, st>>>(n_ghost, ghost_idx, sim->stamp.p, sim->front[0].p, sim->ctl);
      LAUNCH_CHECK();
    }
    const CoopPeer cp = dist_coop_peer(sim);
    k_mark_border<<<blocks, kThreads, 0, st>>>(n, cp.rslot[0], cp.rslot[1], sim->stamp.p);
    LAUNCH_CHECK();
  }
  int use_cutoff = sim->level_cutoff ? 1 : 0;
  float neg_dmax = -sim->pp.maximum_surface_distance;
  if (sim->prop_grid == 0) {  // co-resident blocks of the persistent kernel on this device
    int per_sm = 0, per_sm_peer = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_propagate<false, false>, kPropThreads, 0));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_peer, k_propagate<true, false>, kPropThreads, 0));
    sim->prop_grid = std::max(1, std::min(std::min(per_sm, per_sm_peer), 2) * sim->sm_count);  // a third block per SM (42 registers, spills) measured no faster
  }
  // a front rarely holds more than a few ten thousand particles: one warp each
  uint32_t grid = uint32_t(std::max(1, std::min<int>(sim->prop_grid, int((n + 4u * kPropThreads - 1) / (4u * kPropThreads)))));
  uint32_t n_arg = n;
  float* level_arg = level;
  int* stamp_arg = sim->stamp.p;
  const float4* xyhm_arg = sim->xyhm.p;
  uint32_t* front_arg = sim->front[0].p;
  uint32_t* front1_arg = sim->front[1].p;
  uint32_t* border_arg = sim->cand.p;
  StepCtl* ctl_arg = sim->ctl;
  CoopPeer coop = dist_coop_peer(sim);
  void* args[] = {&n_arg, &L, &xyhm_arg, &level_arg, &stamp_arg, &front_arg, &front1_arg, &border_arg, &ctl_arg, &neg_dmax, &use_cutoff, &coop};
  cudaEvent_t kt0 = nullptr, kt1 = nullptr;
  if (sim->kt_every > 0) { kt0 = kt_event(sim); kt1 = kt_event(sim); cudaEventRecord(kt0, st); }
  void* kernel = peer ? (void*)k_propagate<true, false> : center_diff ? (void*)k_propagate<false, true> : (void*)k_propagate<false, false>;
  CUDA_TRY(cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(kPropThreads), args, 0, st));
  sim->kernel_launches++;
  if (kt1) cudaEventRecord(kt1, st);
  const int rc_sync = sync_ctl(sim);
  if (kt1) {
    float ms = 0.f;
    if (rc_sync == ASPH_OK && cudaEventElapsedTime(&ms, kt0, kt1) == cudaSuccess) { sim->kt_ms[ASPH_KT_LEVEL_PROPAGATE] += ms; sim->kt_samples[ASPH_KT_LEVEL_PROPAGATE]++; }
    kt_release(sim, kt0); kt_release(sim, kt1);
  }
  TRY(rc_sync);
  if (peer) dist_coop_advance(sim, unsigned(std::max(0, sim->ctl_host->level_sweep)));  // one barrier per sweep run, on every rank alike
  if (sim->ctl_host->error_flags & ERRF_PEER_TIMEOUT) return check_error_flags(sim);
  if (sim->ctl_host->error_flags & ERRF_LIST_CAPACITY) return ASPH_RETRY_LISTS;
  if (!sim->ctl_host->level_done) { sim->last_error = "level-set propagation did not terminate"; return ASPH_ERR_INVALID; }
  // sweeps including the final one that changes nothing, as the reference counts them (simulation.rs:739-799)
  sim->info.level_sweeps = std::max(1, sim->ctl_host->level_sweep);
  return ASPH_OK;
}

#ifdef ASPH_PROP_TRACE
extern "C" int asph_debug_prop_trace(unsigned long long* out, int reset) {
  if (cudaMemcpyFromSymbol(out, g_prop_trace, sizeof(unsigned long long) * 12 * 512) != cudaSuccess) return -1;
  if (reset) { static unsigned long long z[12 * 512]; cudaMemcpyToSymbol(g_prop_trace, z, sizeof(z)); }
  return 0;
}
#endif

int launch_level_smoothing(asph_sim* sim) {
  const uint32_t n = sim->n;
  if (n == 0) { sim->level_valid = true; return ASPH_OK; }
  const uint32_t blocks = (n + kThreads - 1) / kThreads;
  const int c = sim->cur;
  Lists L{sim->nbpool.p, sim->slice_base.p, sim->cnt.p, sim->cnt_ext.p, sim->far_idx.p, sim->far_cnt.p};
  if (sim->dist) TRY(dist_halo(sim, sim->pos[c].p, 8));  // K17 reads the neighbours' advected positions: the ghosts' come from their owners
  k_smooth<<<blocks, kThreads, 0, sim->stream>>>(n, L, sim->pos[c].p, sim->xyhm.p, sim->rho.p,
                                                 sim->level[c].p, sim->scratch_f.p, sim->pp.maximum_surface_distance, sim->ctl);
  LAUNCH_CHECK();
  CUDA_TRY(cudaMemcpyAsync(sim->level[c].p, sim->scratch_f.p, size_t(n) * sizeof(float), cudaMemcpyDeviceToDevice, sim->stream));
  if (sim->dist) TRY(dist_halo_words(sim, sim->level[c].p));  // a ghost's own sum runs over an incomplete neighbourhood
  sim->level_valid = true;
  return ASPH_OK;
}
