// solver.cu — the pair-sum passes of the step that run on the stored neighbour lists: non-pressure acceleration
// (K12), PPE source terms (K13), the relaxed-Jacobi pressure sweeps (K14 + K15) with their device-side stop rule,
// and the integrators (K16).
//
// Every pass is one thread per particle.  A thread reads its column 8 rows at a time (one 16 B load of 16-bit window
// byte offsets; the warp's loads of a chunk are one 512 B run), gathers ONE float4 per neighbour from the block's
// shared-memory window (plus {h, m} when h is not uniform) and recomputes
//   m_j * gradW_ij = m_j * pair_g(|x_ij|^2, h_ij) * x_ij
// in registers (lists.cuh), so a sweep streams 2 B per neighbour from HBM instead of 8.  The two passes of a Jacobi
// sweep are persistent, software-pipelined kernels (k_sweep); on several GPUs they also carry the ghost exchange.
//
// Reference: simulation.rs:931-1005 (non-pressure accel), :1552-1592 (divergence operator), :1633-1748 (sources),
// :1751-1808 (pressure accel), :1207-1322 (Jacobi sweep + PressureSolverStatistics), :1378-1516 (loop control),
// :2389-2446 / :2502-2670 (IISPH / HybridDFSPH step orders); boundary terms boundary_winchenbach2020.rs:164-223.
#include "lists.cuh"

namespace {

__global__ void k_solver_reset(StepCtl* ctl) {
  SolverCtl& s = ctl->solver;
  if (threadIdx.x == 0) {
    s.k = 0; s.done = 0; s.sweeps = 0; s.normal = 0; s.singular = 0; s.negative = 0;
    s.err_sum = 0.f; s.max_err = 0.f; s.avg = 0.f;
    s.maxerr_enc[0] = s.maxerr_enc[1] = s.maxerr_enc[2] = 0u;
  }
  for (int t = threadIdx.x; t < 3 * ASPH_ACC_WORDS; t += blockDim.x) (&s.acc[0][0])[t] = 0ull;
}

// ---- loop control of iisph_pressure_iterations (simulation.rs:1405-1480) -------------------------------------------
struct SweepTotals {
  unsigned long long normal, negative, singular;
  float err_sum;
};
// Spin until a peer's sequence number has reached `seq` (PeerCtl, sim.cuh).  Bounded: a rank that failed and stopped
// launching must not hang the others; after 2 s the step is flagged and fails on the host.
__device__ __forceinline__ void peer_wait(const unsigned int* flag, unsigned int seq, StepCtl* ctl) {
  // poll with plain (relaxed) loads of this GPU's own memory — L2 hits — and acquire once the number is there
  const volatile unsigned int* f = flag;
  unsigned long long t0 = 0;
  for (unsigned int spins = 0; int(*f - seq) < 0; spins++) {
    if ((spins & 4095u) == 4095u) {
      if (*reinterpret_cast<volatile unsigned int*>(&ctl->error_flags) & ERRF_PEER_TIMEOUT) return;  // somebody gave up already
      unsigned long long now;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 2000000000ull) { atomicOr(&ctl->error_flags, ERRF_PEER_TIMEOUT); return; }
    }
  }
  (void)ld_acquire_sys(flag);
}
__device__ __forceinline__ void peer_wait_halo(const PeerArgs& P, StepCtl* ctl) {
  if (!P.self || P.halo_seq == 0u) return;
  if (P.rank > 0) peer_wait(&P.self->halo_flag[0], P.halo_seq, ctl);
  if (P.rank + 1 < P.nranks) peer_wait(&P.self->halo_flag[1], P.halo_seq, ctl);
}
__device__ __forceinline__ void peer_wait_stats(const PeerArgs& P, StepCtl* ctl) {
  if (!P.self || P.stats_seq == 0u) return;
  // all ranks' numbers are requested together (one L2 round trip per polling round, not one per rank)
  const volatile unsigned int* f = P.self->stats_flag;
  unsigned long long t0 = 0;
  for (unsigned int spins = 0;; spins++) {
    unsigned int v[ASPH_MAX_RANKS];
#pragma unroll
    for (int r = 0; r < ASPH_MAX_RANKS; r++) v[r] = (r < P.nranks && r != P.rank) ? f[r] : P.stats_seq;
    bool all = true;
#pragma unroll
    for (int r = 0; r < ASPH_MAX_RANKS; r++) all = all && int(v[r] - P.stats_seq) >= 0;
    if (all) break;
    if ((spins & 1023u) == 1023u) {
      if (*reinterpret_cast<volatile unsigned int*>(&ctl->error_flags) & ERRF_PEER_TIMEOUT) return;
      unsigned long long now;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 2000000000ull) { atomicOr(&ctl->error_flags, ERRF_PEER_TIMEOUT); return; }
    }
  }
  for (int r = 0; r < P.nranks; r++)
    if (r != P.rank) { (void)ld_acquire_sys(&P.self->stats_flag[r]); break; }  // one acquire orders the mailbox reads below
}
// totals of sweep number `sweep` over all particles of all ranks (call peer_wait_stats first)
__device__ __forceinline__ SweepTotals read_totals(const StepCtl* ctl, int sweep, const PeerArgs& P) {
  const unsigned long long* acc = ctl->solver.acc[sweep % 3];
  unsigned long long w[ASPH_ACC_WORDS];
#pragma unroll
  for (int c = 0; c < ASPH_ACC_WORDS; c++) w[c] = __ldcg(acc + c);  // written by an earlier kernel; all loads in flight together
  unsigned long long nn = 0, ng = 0, sg = w[2 * ASPH_ACC_COPIES];
  long long e = 0;
#pragma unroll
  for (int c = 0; c < ASPH_ACC_COPIES; c++) {
    nn += w[2 * c] & 0xffffffffull; ng += w[2 * c] >> 32;
    e += (long long)w[2 * c + 1];
  }
  if (P.self && P.stats_seq != 0u) {  // the other ranks' parts (integers: any order gives the same sums)
    for (int r = 0; r < P.nranks; r++) {
      if (r == P.rank) continue;
      const unsigned long long* in = P.self->stats_in[r][sweep % 3];
#pragma unroll
      for (int c = 0; c < ASPH_ACC_WORDS; c++) w[c] = __ldcg(in + c);  // L2 is where the peer's stores landed
#pragma unroll
      for (int c = 0; c < ASPH_ACC_COPIES; c++) {
        nn += w[2 * c] & 0xffffffffull; ng += w[2 * c] >> 32;
        e += (long long)w[2 * c + 1];
      }
      sg += w[2 * ASPH_ACC_COPIES];
    }
  }
  SweepTotals t;
  t.normal = nn; t.negative = ng; t.singular = sg;
  t.err_sum = float(double(e) * (1.0 / 4294967296.0));
  return t;
}
// has the solver finished after sweep number `sweep` with these totals?
__device__ __forceinline__ bool sweep_stops(const SweepTotals& t, int sweep, unsigned int error_flags, float dt, float rho0, float tol,
                                            int max_iters, int density_mode) {
  const float avg = t.normal > 0 ? t.err_sum / float(t.normal) : __int_as_float(0x7fc00000);
  bool stop;
  if (density_mode) stop = (t.normal == 0) || (fabsf(avg / rho0) < tol && sweep > 1);
  else stop = (t.normal == 0) || (fabsf(avg) < tol / dt && sweep > 1);
  return stop || sweep == max_iters || (error_flags & ERRF_SOLVER_NONFINITE) != 0u;
}
// one thread: publish the evaluation of sweep number `sweep` for the host and recycle the accumulator slot of sweep + 2
__device__ __forceinline__ void record_sweep(StepCtl* ctl, const SweepTotals& t, int sweep, bool stop) {
  SolverCtl& s = ctl->solver;
  s.normal = t.normal; s.negative = t.negative; s.singular = t.singular; s.err_sum = t.err_sum;
  s.avg = t.normal > 0 ? t.err_sum / float(t.normal) : __int_as_float(0x7fc00000);
  s.max_err = dec_f(s.maxerr_enc[sweep % 3]);
  s.sweeps = sweep + 1;
  if (stop) { s.done = 1; s.k = sweep; }
  else s.k = sweep + 1;
  unsigned long long* nxt = s.acc[(sweep + 2) % 3];
  for (int c = 0; c < ASPH_ACC_WORDS; c++) nxt[c] = 0ull;
  s.maxerr_enc[(sweep + 2) % 3] = 0u;
}
// end of a batch of sweeps: evaluate the last one launched (the next batch's first kernel would do the same)
__global__ void k_solver_decide(StepCtl* ctl, int sweep, float rho0, float tol, int max_iters, int density_mode, const PeerArgs P) {
  if (ctl->solver.done || ctl->solver.sweeps > sweep) return;
  peer_wait_stats(P, ctl);
  const SweepTotals t = read_totals(ctl, sweep, P);
  record_sweep(ctl, t, sweep, sweep_stops(t, sweep, ctl->error_flags, ctl->dt, rho0, tol, max_iters, density_mode));
}

// ---- shared-memory window ------------------------------------------------------------------------------------
// Particles are sorted by strip-major cell number (sim.cuh), so almost every 2h neighbour of the kThreads consecutive
// particles of a block lies within kHalo positions of the block's range.  The block stages that window of the gathered
// pack (and of {h, m} when h is not uniform) in shared memory with asynchronous coalesced copies.  The neighbour build
// (neighbors.cu) has already split every column into the entries inside this window — stored as the byte offset of
// their slot, so a gather is one bit-field extract and one LDS.128 — and the few outside it (across a strip edge, or
// in another size level's grid), which are read from global memory in a second, short loop.  Slot t of the window
// holds particle wa + t, wa = block_first - kHalo (mod 2^32: for the first block the slots of "negative" particles
// stay unused; no column refers to them).
constexpr int kThreads = int(ASPH_PAIR_BLOCK);
constexpr uint32_t kHalo = ASPH_PAIR_HALO;
constexpr uint32_t kWin = ASPH_PAIR_WIN;      // contiguous part of the window
constexpr uint32_t kFar = ASPH_PAIR_FAR;      // slots of the tile's far table, staged behind it
constexpr uint32_t kSlots = ASPH_PAIR_SLOTS;
static_assert(kSlots * 16u <= 65536u, "window byte offsets are stored in 16 bits");
static_assert(kFar <= ASPH_PAIR_BLOCK, "one thread stages one far-table slot");

// Per-block context of a pair pass.  Order inside a kernel: issue() the asynchronous window copy, construct the
// thread's PairCol (slice header, counts, first two index chunks) and load the thread's own values, then wait() —
// so the three kinds of global-memory latency overlap instead of following one another.
struct PairWindow {
  const float4* wp;  // the staged arrays (shared memory)
  const float2* wh;
  const float* wa;
  template <bool AUX>
  __device__ __forceinline__ void issue(uint32_t n, bool uni, const float4* __restrict__ pack, const float2* __restrict__ hm,
                                        const float* __restrict__ aux, float4 (&s_pack)[kSlots], float2 (&s_hm)[kSlots], float* s_aux,
                                        const NbLists& L) {
    wp = s_pack; wh = s_hm; wa = s_aux;
    const uint32_t sp = uint32_t(__cvta_generic_to_shared(&s_pack[0]));
    const uint32_t sh = uint32_t(__cvta_generic_to_shared(&s_hm[0]));
    const uint32_t sa = AUX ? uint32_t(__cvta_generic_to_shared(s_aux)) : 0u;
    const uint32_t w0 = blockIdx.x * kThreads - kHalo;
    for (uint32_t t = threadIdx.x; t < kWin; t += kThreads) {
      const uint32_t g = w0 + t;
      if (g < n) {
        cp_async16(sp + t * 16u, pack + g);
        if (!uni) cp_async8(sh + t * 8u, hm + g);
        if (AUX) cp_async4(sa + t * 4u, aux + g);
      }
    }
    // the tile's far table (lists.cuh): particles outside the contiguous window, staged behind it
    if (threadIdx.x < min(__ldg(&L.far_cnt[blockIdx.x]), kFar)) {
      const uint32_t g = min(__ldg(&L.far_idx[blockIdx.x * kFar + threadIdx.x]), n - 1u);
      const uint32_t t = kWin + threadIdx.x;
      cp_async16(sp + t * 16u, pack + g);
      if (!uni) cp_async8(sh + t * 8u, hm + g);
      if (AUX) cp_async4(sa + t * 4u, aux + g);
    }
  }
  __device__ __forceinline__ void wait() const {
    cp_async_wait_all();
    __syncthreads();
  }
};

// where a pair pass finds {h_j, m_j}: nowhere (h is uniform, m_j = m_i), in the staged window, or in global memory
enum { HM_UNI = 0, HM_WIN = 1, HM_GLOBAL = 2 };

struct PairCol {  // a thread's column header plus its first 16 window rows, requested before the window barrier
  NbCol col;
  uint4 c0, c1;
  __device__ __forceinline__ PairCol(const NbCol& c, const uint4& a, const uint4& b) : col(c), c0(a), c1(b) {}
  __device__ __forceinline__ PairCol(const NbLists& L, uint32_t i, bool active) : c0(make_uint4(0, 0, 0, 0)), c1(make_uint4(0, 0, 0, 0)) {
    if (active) {
      col = NbCol(L, i);
      if (col.cw > 0u) c0 = col.raw8(0);
      if (col.cw > 8u) c1 = col.raw8(8);
    }
  }
};

// Σ over N_2(i) of f(pack[j], x_ij, c_ij, h_ij, aux[j]); the caller multiplies its sums by the returned scale:
//   uniform h (UNI):  c_ij = w'(q)/r un-normalised,   scale = m * 40 / (7 pi (2h)^3)
//   otherwise:        c_ij = m_j * dW/dr / r,         scale = 1
// (every f is linear in c).  The window segment of a column is consumed 8 rows at a time — the gathers, then the
// arithmetic; padding rows point at the particle itself (zero distance => c = 0), so there are no per-entry checks.
// The index chunk two iterations ahead is always in flight.  The far segment follows 4 rows at a time from global memory.
template <int HM, bool AUX, bool R4 = false, class F>
__device__ __forceinline__ void pair_apply(const float4& o, const float2& t, float a, float xi, float yi, float hi, const PairShape& shape, F& f) {
  constexpr bool UNI = HM == HM_UNI;
  const float dx = xi - o.x, dy = yi - o.y;
  const float d2 = dx * dx + dy * dy;
  if (UNI) {
    f(o, dx, dy, shape(d2), hi, a);
  } else if (R4) {  // the sweep bodies use neither h_ij nor the aux value
    f(o, dx, dy, t.y * pair_g_sum(d2, hi + t.x), 0.f, a);
  } else {
    const float hij = (hi + t.x) * 0.5f;
    f(o, dx, dy, t.y * pair_g(d2, hij), hij, a);
  }
}
// R4 (experiment, ASPH_ROWS4=1): the neighbour pass wrote the particle's own row last in the W segment; it contributes
// nothing to any gradient sum, so the loop stops one row early, and it runs in steps of 4 rows instead of 8 — on the resting
// lattice (12 neighbours + self) that is 12 rows per particle instead of 16.
template <int HM, bool AUX, bool R4 = false, class F>
__device__ __forceinline__ float for_each_pair(const PairCol& P, const PairWindow& W, const float4* __restrict__ pack,
                                               const float2* __restrict__ hm, const float* __restrict__ aux, float xi, float yi, float hi,
                                               float mi, F f) {
  constexpr bool UNI = HM == HM_UNI;
  const NbCol& col = P.col;
  const float inv2h = fast_rcp(2.f * hi);
  const PairShape shape(inv2h);
  // ---- window segment: the two prefetched chunks, then (long columns only) chunks fetched on demand
  auto chunk = [&](const uint4& v, uint32_t rows_left) {
    const uint32_t off[8] = {v.x & 0xffffu, v.x >> 16, v.y & 0xffffu, v.y >> 16, v.z & 0xffffu, v.z >> 16, v.w & 0xffffu, v.w >> 16};
#pragma unroll
    for (int half = 0; half < 2; half++) {
      if (R4 && half == 1 && rows_left <= 4u) break;
      float4 o[4];
      float2 t[4];
      float a[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const uint32_t of = off[half * 4 + u];
        o[u] = *reinterpret_cast<const float4*>(reinterpret_cast<const char*>(W.wp) + of);
        if (HM == HM_WIN) t[u] = *reinterpret_cast<const float2*>(reinterpret_cast<const char*>(W.wh) + (of >> 1));
        else if (HM == HM_GLOBAL) t[u] = __ldg(hm + ((of >> 4) < kWin ? nb_win0(col.i) + (of >> 4) : __ldg(col.far_idx + (col.i / kThreads) * kFar + ((of >> 4) - kWin))));
        else t[u] = make_float2(0.f, 0.f);
        a[u] = AUX ? *reinterpret_cast<const float*>(reinterpret_cast<const char*>(W.wa) + (of >> 2)) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; u++) pair_apply<HM, AUX, R4>(o[u], t[u], a[u], xi, yi, hi, shape, f);
    }
  };
  const uint32_t cw = R4 ? (col.cw > 0u ? col.cw - 1u : 0u) : col.cw;  // rows to process
  if (cw > 0u) chunk(P.c0, cw);
  if (cw > 8u) chunk(P.c1, cw - 8u);
  if (cw > 16u) {
    uint4 nxt = col.raw8(16u);
    for (uint32_t k0 = 16u; k0 < cw; k0 += 8u) {
      const uint4 v = nxt;
      if (k0 + 8u < cw) nxt = col.raw8(k0 + 8u);
      chunk(v, cw - k0);
    }
  }
  // ---- far segment (divergent: most threads have none)
  for (uint32_t k0 = 0; k0 < col.cf; k0 += 4u) {
    uint32_t j[4];
    col.far4(k0, j);
    float4 o[4];
    float2 t[4];
    float a[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      o[u] = __ldg(pack + j[u]);
      t[u] = UNI ? make_float2(0.f, 0.f) : __ldg(hm + j[u]);
      a[u] = AUX ? __ldg(aux + j[u]) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) pair_apply<HM, AUX, R4>(o[u], t[u], a[u], xi, yi, hi, shape, f);
  }
  return UNI ? mi * (ASPH_KNORM * inv2h * inv2h * inv2h) : 1.f;
}

// ---------------------------------------------------------------------------------------------- K12
template <bool UNI>
__device__ __forceinline__ void viscosity_body(uint32_t i, const PairCol& C, const PairWindow& W, const float4& me, const float2& own, float rho_i,
                                               const float4* __restrict__ xv_in, const float2* __restrict__ hm, const float* __restrict__ rho,
                                               const PackedParams& P, float dt, float4* __restrict__ xv_out) {
  float ax = 0.f, ay = 0.f;
  if (P.viscosity_type != ASPH_VISC_XSPH && P.viscosity != 0.f) {
    const float scale = for_each_pair<UNI ? HM_UNI : HM_WIN, true>(C, W, xv_in, hm, rho, me.x, me.y, own.x, own.y,
                             [&](const float4& o, float dx, float dy, float c, float hij, float rho_j) {
      const float est = dx * (me.z - o.z) + dy * (me.w - o.w);
      if (est < 0.f) {
        const float d2 = dx * dx + dy * dy;
        float f;
        if (P.viscosity_type == ASPH_VISC_APPROX_LAPLACE) {
          f = P.viscosity * (8.f * est / ((rho_i + rho_j) * 0.5f * (d2 + 0.01f * hij * hij)));  // 2(D+2), D = 2
        } else {  // WCSPH, speed of sound 88
          f = (2.f * P.viscosity * hij * 88.f / (rho_i + rho_j)) * est / (d2 + 0.001f * hij * hij);
        }
        ax += f * c * dx; ay += f * c * dy;
      }
    });
    ax *= scale; ay *= scale;
  }
  ay += P.gravity;
  if (P.has_pull) {
    const float px = P.pull_x - me.x, py = P.pull_y - me.y;
    const float pn = sqrtf(px * px + py * py);
    ax += px / pn * 13.f; ay += py / pn * 13.f;
  }
  xv_out[i] = make_float4(me.x, me.y, me.z + dt * ax, me.w + dt * ay);
}
__global__ void __launch_bounds__(kThreads)
k_viscosity(uint32_t n, NbLists L, const float4* __restrict__ xv_in, const float2* __restrict__ hm, const float* __restrict__ rho,
            const PackedParams P, const StepCtl* __restrict__ ctl, float4* __restrict__ xv_out) {
  __shared__ float4 s_pack[kSlots];
  __shared__ float2 s_hm[kSlots];
  __shared__ float s_aux[kSlots];
  const bool uni = ctl->hmin == ctl->hmax;
  PairWindow W;
  W.issue<true>(n, uni, xv_in, hm, rho, s_pack, s_hm, s_aux, L);
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i < n;
  const PairCol C(L, i, active);
  float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
  float2 own = make_float2(1.f, 0.f);
  float rho_i = 1.f;
  if (active) { me = xv_in[i]; own = hm[i]; rho_i = rho[i]; }
  W.wait();
  if (!active) return;
  if (uni) viscosity_body<true>(i, C, W, me, own, rho_i, xv_in, hm, rho, P, ctl->dt, xv_out);
  else viscosity_body<false>(i, C, W, me, own, rho_i, xv_in, hm, rho, P, ctl->dt, xv_out);
}

// ---------------------------------------------------------------------------------------------- K13
// kind 0: -div(v)/dt; 1: -(rho0-rho)/(rho dt^2); 2: both.  Also p <- 0 (packP[0]) and a^p <- 0 (what K14 gives for p = 0,
// so the first sweep does not launch its pressure-acceleration pass).
__global__ void __launch_bounds__(kThreads)
k_source(uint32_t n, NbLists L, const float4* __restrict__ xv, const float2* __restrict__ hm, const float* __restrict__ rho,
         float4* __restrict__ pconst, float4* __restrict__ packP0, float4* __restrict__ packA, const StepCtl* __restrict__ ctl,
         float rho0, int kind, int w2020, const float* __restrict__ omega) {
  __shared__ float4 s_pack[kSlots];
  __shared__ float2 s_hm[kSlots];
  // Winchenbach2020 operator: `hm` is {h, m / rho} (k_aii_w2020), every pair carries its own weight m_j / rho_j, and
  // neither the pair sum nor the boundary term (pconst.xy = G) is divided by rho_i (simulation.rs:1571-1575)
  const bool uni = ctl->hmin == ctl->hmax && !w2020;
  const bool pairs = kind != 1;
  PairWindow W;
  if (pairs) W.issue<false>(n, uni, xv, hm, nullptr, s_pack, s_hm, nullptr, L);
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i < n;
  const PairCol C(L, i, active && pairs);
  float4 me = make_float4(0.f, 0.f, 0.f, 0.f), pc = me;
  float2 own = make_float2(1.f, 0.f);
  float rho_i = 1.f;
  if (active) { me = xv[i]; pc = pconst[i]; rho_i = rho[i]; own = hm[i]; }
  if (pairs) W.wait();
  if (!active) return;
  const float dt = ctl->dt;
  float s = 0.f;
  if (pairs) {
    float sum = 0.f;
    auto body = [&](const float4& o, float dx, float dy, float c, float, float) { sum += c * ((o.z - me.z) * dx + (o.w - me.w) * dy); };
    const float scale = uni ? for_each_pair<HM_UNI, false>(C, W, xv, hm, nullptr, me.x, me.y, own.x, own.y, body)
                            : for_each_pair<HM_WIN, false>(C, W, xv, hm, nullptr, me.x, me.y, own.x, own.y, body);
    sum *= scale;
    const float div = (w2020 ? sum : sum / rho_i) - (me.z * pc.x + me.w * pc.y);
    s = omega ? -div / (dt * omega[i]) : -div / dt;  // calculate_source_term_full_with_omega, simulation.rs:1678-1710
  }
  if (kind != 0) s += -(rho0 - rho_i) / (((w2020 || omega) ? rho0 : rho_i) * dt * dt);  // next_density_estimate, simulation.rs:1633-1748
  pc.w = s;
  pconst[i] = pc;
  packP0[i] = make_float4(me.x, me.y, 0.f, 0.f);
  packA[i] = make_float4(me.x, me.y, 0.f, 0.f);  // a^p of the first sweep: p = 0 everywhere => exactly zero (simulation.rs:1792-1807)
}

// ---------------------------------------------------------------------------------------------- IISPH2
// omega_i = clamp(1 + H_i / (3 rho_i) * sum_j m_j dW/dH(|x_ij|, H_ij), 0.125, 2.5), H = 2h the support radius; a particle
// classified Large by the last resampling phase uses its own term only (simulation.rs:2263-2311).
__device__ __forceinline__ float dwdh(float d, float H) {
  const float q = d / H;
  const float cd = 40.f / (7.f * ASPH_PI_F);
  return cd * -2.f / (H * H * H) * cubic_w(q) + cd / (H * H) * cubic_dw(q) * (-d / (H * H));
}
__global__ void __launch_bounds__(kThreads)
k_omega(uint32_t n, NbLists L, const float4* __restrict__ xyhm, const float* __restrict__ rho, const uint8_t* __restrict__ cls,
        float* __restrict__ omega) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 me = xyhm[i];
  const float Hi = me.z * 2.f;
  const float pre = Hi / (3.f * rho[i]);
  float om = 1.f;
  if (cls && cls[i] == ASPH_CLASS_LARGE) {
    om += pre * me.w * dwdh(0.f, me.z * 2.f);
  } else {
    const NbCol col(L, i);
    for (uint32_t k = 0; k < col.cn; k++) {
      const float4 o = __ldg(&xyhm[col.get(k)]);
      const float dx = me.x - o.x, dy = me.y - o.y;
      om += pre * o.w * dwdh(sqrtf(dx * dx + dy * dy), ((me.z + o.z) * 0.5f) * 2.f);
    }
  }
  omega[i] = fminf(2.5f, fmaxf(om, 0.125f));
}
// p /= sqrt(omega) on the pack that holds the solve's result (chosen like in k_accel)
__global__ void __launch_bounds__(kThreads)
k_scale_pressure(uint32_t n, float4* __restrict__ P0, float4* __restrict__ P1, const StepCtl* __restrict__ ctl, const float* __restrict__ omega) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4* pack = (ctl->solver.sweeps & 1) ? P1 : P0;
  float4 v = pack[i];
  const float s = sqrtf(omega[i]);
  v.z /= s; v.w /= s;
  pack[i] = v;
}

// ---------------------------------------------------------------------------------------------- K14
// a^p_i = -Σ m_j (P_i + P_j) gradW_ij - p_i * Bc * G_i,  P = p / rho^2.
// The pass after a solve (the passes inside the sweeps are k_sweep<0> below).  MODE 1: v += dt a^p into the xv pack;
// 2: + HybridDFSPH integration (simulation.rs:2622-2669); 3: + IISPH integration (simulation.rs:2433-2444); 4: a^p only.
// P0 / P1: the two pressure packs; the current one is chosen by the parity of the sweeps executed so far, read from
// the control block, so the launch sequence does not depend on when the solver stops.
template <int MODE>
__global__ void __launch_bounds__(kThreads)
k_accel(uint32_t n, NbLists L, const float4* __restrict__ P0, const float4* __restrict__ P1, const float2* __restrict__ hm,
        const float2* __restrict__ gB, float4* __restrict__ packA, StepCtl* ctl, float4* __restrict__ xv, float2* __restrict__ pos,
        float2* __restrict__ vel, float hybrid_factor, const uint32_t* __restrict__ gid) {
  __shared__ float4 s_pack[kSlots];
  __shared__ float2 s_hm[kSlots];
  const int parity = ctl->solver.sweeps & 1;  // runs after k_solver_decide
  const float4* __restrict__ packP = parity ? P1 : P0;
  // After a sweep that left no particle with positive pressure (normal == 0: every p' was clamped to 0 or was
  // singular) the whole pressure field is zero and so is a^p; skip the pair sum.
  const bool pairs = ctl->solver.normal != 0;
  const bool uni = ctl->hmin == ctl->hmax;
  PairWindow W;
  if (pairs) W.issue<false>(n, uni, packP, hm, nullptr, s_pack, s_hm, nullptr, L);
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i < n;
  const PairCol C(L, i, active && pairs);
  float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
  float2 own = make_float2(1.f, 0.f), g = make_float2(0.f, 0.f);
  if (active) { me = packP[i]; own = hm[i]; g = gB[i]; }
  if (pairs) W.wait();
  if (!active) return;
  float ax = 0.f, ay = 0.f;
  if (pairs) {
    auto body = [&](const float4& o, float dx, float dy, float c, float, float) {
      const float f = c * (me.z + o.z);
      ax -= f * dx; ay -= f * dy;
    };
    const float scale = uni ? for_each_pair<HM_UNI, false>(C, W, packP, hm, nullptr, me.x, me.y, own.x, own.y, body)
                            : for_each_pair<HM_WIN, false>(C, W, packP, hm, nullptr, me.x, me.y, own.x, own.y, body);
    ax *= scale; ay *= scale;
    const float2 g = gB[i];
    ax -= me.w * g.x;
    ay -= me.w * g.y;
  }
  packA[i] = make_float4(me.x, me.y, ax, ay);
  if (MODE == 1) {
    const float dt = ctl->dt;
    float4 v = xv[i];
    v.z += dt * ax; v.w += dt * ay;
    xv[i] = v;
  } else if (MODE == 2) {
    const float dt = ctl->dt;
    const float4 v = xv[i];
    const float fac = fminf(dt * hybrid_factor, 1.f);
    const float2 x = make_float2(me.x + (dt * v.z + (dt * dt) * ax), me.y + (dt * v.w + (dt * dt) * ay));
    const float2 vn = make_float2(v.z + (dt * ax) * fac, v.w + (dt * ay) * fac);
    pos[i] = x; vel[i] = vn;
    if (!(isfinite(x.x) && isfinite(x.y) && isfinite(vn.x) && isfinite(vn.y)) && !(gid && (gid[i] & ASPH_GHOST_BIT)))
      atomicOr(&ctl->error_flags, ERRF_NONFINITE);
  } else if (MODE == 3) {
    const float dt = ctl->dt;
    const float4 v = xv[i];
    const float2 vn = make_float2(v.z + dt * ax, v.w + dt * ay);
    const float2 x = make_float2(me.x + dt * vn.x, me.y + dt * vn.y);
    pos[i] = x; vel[i] = vn;
    if (!(isfinite(x.x) && isfinite(x.y) && isfinite(vn.x) && isfinite(vn.y)) && !(gid && (gid[i] & ASPH_GHOST_BIT)))
      atomicOr(&ctl->error_flags, ERRF_NONFINITE);
  }
}

// ---------------------------------------------------------------------------------------------- K14 + K15, sweep form
// The two passes of a Jacobi sweep as PERSISTENT, software-pipelined kernels: a block walks over tiles of kThreads
// consecutive particles (tile = blockIdx.x, + gridDim.x, ...) and, while it computes tile t out of one shared-memory
// stage, the asynchronous copies (LDGSTS) of tile t + 1 — the gather window, the threads' own per-particle inputs and
// the first two index chunks of every column — are already in flight into the other stage; the column headers
// (counts, slice base) of tile t + 2 travel in registers.  No thread waits for a global load in the steady state except
// in the short far-segment loop and for columns longer than 16 window rows.
//   PASS 0 (K14): a^p from p.     own inputs: gB.                    output: packA
//   PASS 1 (K15): p' from a^p.    own inputs: pconst, rho, p_old.    output: the other pressure pack + statistics
// HMWIN: the stages have room for the {h, m} window (adaptive h).  The host picks it from the last control block it
// has seen; if that guess was "uniform" and the step turns out not to be, {h, m} are gathered from global memory.
template <bool HMWIN>
struct SweepStage {
  float4 win[kSlots];
  uint4 chunk[2][kThreads];
  float4 own4[kThreads];
  float2 own2[kThreads];
  float own1[2][kThreads];
  float2 hmw[HMWIN ? kSlots : 1];
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

struct SweepArgs {
  uint32_t n;
  NbLists L;
  const float4* P0; const float4* P1;  // pressure packs {x, y, p/rho^2, p}
  float4* P0w; float4* P1w;
  float4* packA;
  const float2* hm;
  const float2* gB;
  const float4* pconst;
  const float* rho;
  StepCtl* ctl;
  const uint32_t* gid;
  float omega, rho0, tol;
  int sweep, density_mode, max_iters;
  PeerArgs peer;  // multi-GPU peer-memory path: what this pass has to wait for
};

// W2020 (update pass under the Winchenbach2020 operator): `hm` is {h, m / rho} (k_aii_w2020), so
// every pair carries its own weight m_j / rho_j, and the pair sum is not divided by rho_i (simulation.rs:1571-1575).
#define SWEEP_BULK 0
#define SWEEP_KERNEL_NAME k_sweep
#include "sweep_kernel.inc"
#undef SWEEP_BULK
#undef SWEEP_KERNEL_NAME

// ---- experiment ASPH_BULK=1: the stage fill of an interior tile as bulk asynchronous copies.  Everything a
// tile stages except the far-table slots, gB and the own particles' p is a contiguous range of global memory, so one
// thread (and lane 0 of every warp for its list chunks) issues cp.async.bulk copies that complete on one mbarrier per
// stage, in place of ~8 LDGSTS per thread with their address arithmetic.  Tiles clipped by either end of the particle
// range keep the per-thread copies.  The kernel body lives in sweep_kernel.inc and is compiled twice.
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0u;
}
// bounded (about 4 s): a byte count that never arrives must end in an error flag, not in a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, StepCtl* ctl) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (*reinterpret_cast<volatile unsigned int*>(&ctl->error_flags) & ERRF_PEER_TIMEOUT) return;  // somebody gave up already: do not wait again
    if (clock64() - t0 > 8000000000ll) { atomicOr(&ctl->error_flags, ERRF_PEER_TIMEOUT); return; }
  }
}

#define SWEEP_BULK 1
#define SWEEP_KERNEL_NAME k_sweep_bulk
#include "sweep_kernel.inc"
#undef SWEEP_BULK
#undef SWEEP_KERNEL_NAME


// grid of a persistent pair pass: every block gets the same number of tiles (the last one possibly fewer)
inline uint32_t sweep_grid(uint32_t n, int sm_count, int blocks_per_sm) {
  const uint32_t ntiles = (n + kThreads - 1) / kThreads;
  const uint32_t slots = uint32_t(std::max(1, sm_count * blocks_per_sm));
  const uint32_t per = (ntiles + slots - 1) / slots;
  return (ntiles + per - 1) / per;
}

NbLists lists_of(asph_sim* sim) {
  NbLists L;
  L.pool = sim->nbpool.p; L.slice_base = sim->slice_base.p; L.cnt = sim->cnt.p; L.cnt_ext = sim->cnt_ext.p; L.far_idx = sim->far_idx.p; L.far_cnt = sim->far_cnt.p;
  return L;
}

}  // namespace

int launch_viscosity(asph_sim* sim) {
  const uint32_t n = sim->n;
  if (n == 0) return ASPH_OK;
  const uint32_t blocks = (n + kThreads - 1) / kThreads;
  const int a = sim->xv_cur;
  k_viscosity<<<blocks, kThreads, 0, sim->stream>>>(n, lists_of(sim), sim->xv[a].p, sim->hm.p, sim->rho.p, sim->pp, sim->ctl, sim->xv[1 - a].p);
  LAUNCH_CHECK();
  sim->xv_cur = 1 - a;
  if (sim->dist) TRY(dist_halo(sim, sim->xv[sim->xv_cur].p, 16));  // neighbours read the new velocities (K13)
  return ASPH_OK;
}

int launch_omega(asph_sim* sim) {
  const uint32_t n = sim->n;
  if (n == 0) return ASPH_OK;
  CUDA_TRY(sim->omega.ensure(sim->cap));
  k_omega<<<(n + kThreads - 1) / kThreads, kThreads, 0, sim->stream>>>(n, lists_of(sim), sim->xyhm.p, sim->rho.p,
                                                                       sim->cls_valid ? sim->cls[sim->cur].p : nullptr, sim->omega.p);
  LAUNCH_CHECK();
  return ASPH_OK;
}

int launch_scale_pressure(asph_sim* sim) {
  const uint32_t n = sim->n;
  if (n == 0) return ASPH_OK;
  k_scale_pressure<<<(n + kThreads - 1) / kThreads, kThreads, 0, sim->stream>>>(n, sim->packP[0].p, sim->packP[1].p, sim->ctl, sim->omega.p);
  LAUNCH_CHECK();
  return ASPH_OK;
}

int launch_source(asph_sim* sim, int kind) {
  const uint32_t n = sim->n;
  if (n == 0) return ASPH_OK;
  const uint32_t blocks = (n + kThreads - 1) / kThreads;
  k_solver_reset<<<1, 64, 0, sim->stream>>>(sim->ctl);
  LAUNCH_CHECK();
  const bool w2020 = op_w2020(sim);
  const float* omega = kind == 3 ? sim->omega.p : nullptr;
  if (kind == 3) kind = 2;
  k_source<<<blocks, kThreads, 0, sim->stream>>>(n, lists_of(sim), sim->xv[sim->xv_cur].p, w2020 ? sim->hv.p : sim->hm.p, sim->rho.p, sim->pconst.p,
                                                 sim->packP[0].p, sim->packA.p, sim->ctl, sim->pp.rest_density, kind, w2020 ? 1 : 0, omega);
  LAUNCH_CHECK();
  sim->p_cur = 0;
  return ASPH_OK;
}

// iisph_pressure_iterations (simulation.rs:1378-1516).  Sweeps are enqueued in batches; each kernel returns
// immediately once the device-side stop rule has fired, and the host looks at the control block once per batch.
// The first batch is sized from the sweep count of the same solve in the previous step.
int launch_solver(asph_sim* sim, bool density_mode, float max_avg_error, int* iters_out, int* sweeps_out, double* avg_out) {
  const uint32_t n = sim->n;
  *iters_out = 0; *sweeps_out = 0; *avg_out = 0;
  if (n == 0) return ASPH_OK;
  cudaStream_t st = sim->stream;
  const NbLists L = lists_of(sim);
  const uint32_t* gid = sim->dist ? sim->refid[sim->cur].p : nullptr;
  int launched = 0;
  int& predicted = sim->predicted_sweeps[density_mode ? 1 : 0];
  int batch = std::max(1, std::min(predicted, 256));
  const int max_sweeps = sim->pp.max_iters + 1;
  // stage layout of the persistent sweep kernels: with room for the {h, m} window unless the last control block the
  // host has seen says h is uniform (a wrong guess only costs speed, see k_sweep)
  const bool w2020 = op_w2020(sim);  // its update pass always gathers {h, m / rho}
  const bool hmwin = w2020 || !(sim->ctl_seen && sim->ctl_host->hmin == sim->ctl_host->hmax);
  const size_t smem = 2 * (hmwin ? sizeof(SweepStage<true>) : sizeof(SweepStage<false>));
  // blocks per SM as in the kernels' launch bounds: 3 with the {h, m} window, else 4
  uint32_t grid = sweep_grid(n, sim->sm_count, hmwin ? 3 : 4);
  if (const char* e = getenv("ASPH_SWEEP_GRID")) grid = std::max(1u, std::min(grid, uint32_t(atoi(e))));  // test hook: few blocks => many tiles per block
  if (!sim->sweep_attr_done) {  // per handle: function attributes belong to the device the handle lives on
    const int big = int(2 * sizeof(SweepStage<true>)), small = int(2 * sizeof(SweepStage<false>));
#define SWEEP_ATTR(K, BYTES) CUDA_TRY((cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTES)))
    // the 4-row kernels of the default operators: single GPU (bulk-copy stage fill, or per-thread copies) and peer memory
    SWEEP_ATTR((k_sweep_bulk<0, true, false, false, true>), big + 16); SWEEP_ATTR((k_sweep_bulk<1, true, false, false, true>), big + 16);
    SWEEP_ATTR((k_sweep_bulk<0, false, false, false, true>), small + 16); SWEEP_ATTR((k_sweep_bulk<1, false, false, false, true>), small + 16);
    SWEEP_ATTR((k_sweep<0, true, false, false, true>), big); SWEEP_ATTR((k_sweep<1, true, false, false, true>), big);
    SWEEP_ATTR((k_sweep<0, false, false, false, true>), small); SWEEP_ATTR((k_sweep<1, false, false, false, true>), small);
    SWEEP_ATTR((k_sweep<0, true, true, false, true>), big); SWEEP_ATTR((k_sweep<1, true, true, false, true>), big);
    SWEEP_ATTR((k_sweep<0, false, true, false, true>), small); SWEEP_ATTR((k_sweep<1, false, true, false, true>), small);
    // Winchenbach2020 operator: 8-row kernels with the {h, m / rho} window
    SWEEP_ATTR((k_sweep<0, true, false>), big); SWEEP_ATTR((k_sweep<0, true, true>), big);
    SWEEP_ATTR((k_sweep<1, true, false, true>), big); SWEEP_ATTR((k_sweep<1, true, true, true>), big);
#undef SWEEP_ATTR
    sim->sweep_attr_done = true;
  }
  const bool p2p = dist_p2p(sim);
  // Single GPU, default operators: the stage fill of interior tiles by bulk copies (k_sweep_bulk); ASPH_BULK=0 keeps the
  // per-thread copies (the kernels the multi-GPU NCCL path uses as well) for A/B runs.
  const bool bulk = sim->bulk && !sim->dist && !w2020;
  SweepArgs A;
  A.n = n; A.L = L; A.P0 = sim->packP[0].p; A.P1 = sim->packP[1].p; A.P0w = sim->packP[0].p; A.P1w = sim->packP[1].p;
  A.packA = sim->packA.p; A.hm = sim->hm.p; A.gB = sim->gB.p; A.pconst = sim->pconst.p; A.rho = sim->rho.p; A.ctl = sim->ctl; A.gid = gid;
  A.omega = sim->pp.jacobi_omega; A.rho0 = sim->pp.rest_density; A.tol = max_avg_error;
  A.density_mode = density_mode ? 1 : 0; A.max_iters = sim->pp.max_iters;
  struct Timed { int sweep; cudaEvent_t e0, e1, e1b, e2; };  // e0..e1: K14; e1b..e2: K15 (halo exchanges excluded)
  std::vector<Timed> timed;
  for (;;) {
    for (int b = 0; b < batch && launched < max_sweeps; b++, launched++) {
      const bool time_it = sim->kt_every > 0 && (launched % sim->kt_every) == 0;
      Timed tm{launched, nullptr, nullptr, nullptr, nullptr};
      if (time_it) { tm.e0 = kt_event(sim); tm.e1 = kt_event(sim); tm.e1b = kt_event(sim); tm.e2 = kt_event(sim); cudaEventRecord(tm.e0, st); }
      A.sweep = launched;
      if (launched > 0) {  // sweep 0: a^p = 0 was written by k_source
        A.hm = sim->hm.p;
        A.peer = dist_peer_args(sim, true, !p2p, 0, false);  // waits for the previous sweep's p' ghosts; publishes a^p
        if (w2020) { if (p2p) k_sweep<0, true, true><<<grid, kThreads, smem, st>>>(A); else k_sweep<0, true, false><<<grid, kThreads, smem, st>>>(A); }
        else if (bulk) { if (hmwin) k_sweep_bulk<0, true, false, false, true><<<grid, kThreads, smem + 16, st>>>(A); else k_sweep_bulk<0, false, false, false, true><<<grid, kThreads, smem + 16, st>>>(A); }
        else if (p2p) { if (hmwin) k_sweep<0, true, true, false, true><<<grid, kThreads, smem, st>>>(A); else k_sweep<0, false, true, false, true><<<grid, kThreads, smem, st>>>(A); }
        else { if (hmwin) k_sweep<0, true, false, false, true><<<grid, kThreads, smem, st>>>(A); else k_sweep<0, false, false, false, true><<<grid, kThreads, smem, st>>>(A); }
        LAUNCH_CHECK();
        if (time_it) cudaEventRecord(tm.e1, st);
        if (!p2p && sim->dist) TRY(dist_halo(sim, sim->packA.p, 16));
      } else if (time_it) {
        cudaEventRecord(tm.e1, st);
      }
      if (time_it) cudaEventRecord(tm.e1b, st);
      A.hm = w2020 ? sim->hv.p : sim->hm.p;
      A.peer = dist_peer_args(sim, launched > 0, p2p && launched > 0, 1 + ((launched + 1) & 1), true);  // waits for the a^p ghosts and the previous sweep's totals; publishes p' and its own
      if (w2020) { if (p2p) k_sweep<1, true, true, true><<<grid, kThreads, smem, st>>>(A); else k_sweep<1, true, false, true><<<grid, kThreads, smem, st>>>(A); }
      else if (bulk) { if (hmwin) k_sweep_bulk<1, true, false, false, true><<<grid, kThreads, smem + 16, st>>>(A); else k_sweep_bulk<1, false, false, false, true><<<grid, kThreads, smem + 16, st>>>(A); }
      else if (p2p) { if (hmwin) k_sweep<1, true, true, false, true><<<grid, kThreads, smem, st>>>(A); else k_sweep<1, false, true, false, true><<<grid, kThreads, smem, st>>>(A); }
      else { if (hmwin) k_sweep<1, true, false, false, true><<<grid, kThreads, smem, st>>>(A); else k_sweep<1, false, false, false, true><<<grid, kThreads, smem, st>>>(A); }
      LAUNCH_CHECK();
      if (time_it) { cudaEventRecord(tm.e2, st); timed.push_back(tm); }
      if (!p2p && sim->dist) {
        TRY(dist_solver_reduce(sim, launched % 3));                       // every rank sees the totals of all ranks
        TRY(dist_halo(sim, sim->packP[(launched + 1) & 1].p, 16));        // p' of the border particles to their ghosts
      }
    }
    if (launched > 0) {  // evaluate the last sweep of the batch (the sweeps before it were evaluated by their successors)
      k_solver_decide<<<1, 1, 0, st>>>(sim->ctl, launched - 1, sim->pp.rest_density, max_avg_error, sim->pp.max_iters, density_mode ? 1 : 0,
                                       dist_peer_args(sim, false, true, -1, false));
      LAUNCH_CHECK();
    }
    TRY(sync_ctl(sim));
    const SolverCtl& s = sim->ctl_host->solver;
    for (const Timed& tm : timed) {
      float a = 0.f, b2 = 0.f;
      if (tm.sweep < s.sweeps && cudaEventElapsedTime(&a, tm.e0, tm.e1) == cudaSuccess && cudaEventElapsedTime(&b2, tm.e1b, tm.e2) == cudaSuccess) {
        if (tm.sweep > 0) { sim->kt_ms[ASPH_KT_ACCEL_SWEEP] += a; sim->kt_samples[ASPH_KT_ACCEL_SWEEP]++; }  // sweep 0 has no K14 launch
        sim->kt_ms[ASPH_KT_JACOBI_SWEEP] += b2; sim->kt_samples[ASPH_KT_JACOBI_SWEEP]++;
        static const bool dbg = getenv("ASPH_DEBUG_PUSH") != nullptr;
        if (dbg) {
          static double push_ms = 0; static int push_n = 0;
          float c3 = 0.f;
          if (tm.sweep > 0 && cudaEventElapsedTime(&c3, tm.e1, tm.e1b) == cudaSuccess) { push_ms += c3; push_n++; }
          if (push_n == 50) { fprintf(stderr, "[asph] push(packA) interval avg %.2f us over %d samples; accel %.2f us jacobi %.2f us\n", push_ms / push_n * 1e3, push_n, a * 1e3, b2 * 1e3); push_ms = 0; push_n = 0; }
        }
      }
      kt_release(sim, tm.e0); kt_release(sim, tm.e1); kt_release(sim, tm.e1b); kt_release(sim, tm.e2);
    }
    timed.clear();
    if (sim->ctl_host->error_flags & ERRF_LIST_CAPACITY) return ASPH_RETRY_LISTS;
    if (sim->ctl_host->error_flags & ERRF_SOLVER_NONFINITE) {
      sim->last_error = "'!a_p.is_finite()' failed. Pressure values probably have exploded!";
      return ASPH_ERR_NONFINITE;
    }
    if (s.done || launched >= max_sweeps) break;
    // not done: continue with small batches that grow (each synchronisation costs about as much as two idle launches)
    batch = (launched <= predicted) ? 2 : std::min(batch * 2, 64);
  }
  const SolverCtl& s = sim->ctl_host->solver;
  predicted = s.sweeps;
  sim->p_cur = s.sweeps & 1;
  *iters_out = s.k;
  *sweeps_out = s.sweeps;
  *avg_out = double(s.avg);
  return ASPH_OK;
}

int launch_final_accel(asph_sim* sim, int mode) {
  const uint32_t n = sim->n;
  if (n == 0) return ASPH_OK;
  const uint32_t blocks = (n + kThreads - 1) / kThreads;
  const NbLists L = lists_of(sim);
  cudaStream_t st = sim->stream;
  const float4* P0 = sim->packP[0].p;
  const float4* P1 = sim->packP[1].p;
  float4* xv = sim->xv[sim->xv_cur].p;
  float2* pos = sim->pos[sim->cur].p;
  float2* vel = sim->vel[sim->cur].p;
  const uint32_t* gid = sim->dist ? sim->refid[sim->cur].p : nullptr;
  switch (mode) {
    case 1: k_accel<1><<<blocks, kThreads, 0, st>>>(n, L, P0, P1, sim->hm.p, sim->gB.p, sim->packA.p, sim->ctl, xv, pos, vel, 0.f, gid); break;
    case 2: k_accel<2><<<blocks, kThreads, 0, st>>>(n, L, P0, P1, sim->hm.p, sim->gB.p, sim->packA.p, sim->ctl, xv, pos, vel, sim->pp.hybrid_factor, gid); break;
    case 3: k_accel<3><<<blocks, kThreads, 0, st>>>(n, L, P0, P1, sim->hm.p, sim->gB.p, sim->packA.p, sim->ctl, xv, pos, vel, 0.f, gid); break;
    default: k_accel<4><<<blocks, kThreads, 0, st>>>(n, L, P0, P1, sim->hm.p, sim->gB.p, sim->packA.p, sim->ctl, xv, pos, vel, 0.f, gid); break;
  }
  LAUNCH_CHECK();
  if (sim->dist && mode == 1) TRY(dist_halo(sim, xv, 16));  // the density solve's source term reads the neighbours' new velocities
  return ASPH_OK;
}
