// solver.cu — the pair-sum passes of the step that run on the stored neighbour lists: non-pressure acceleration
// (K12), PPE source terms (K13), the relaxed-Jacobi pressure sweeps (K14 + K15) with their device-side stop rule,
// and the integrators (K16).
//
// Every pass is one thread per particle.  A warp reads row k of its slice's 16-bit index column as one coalesced
// 64 B line, gathers ONE float4 per neighbour (plus {h, m} when h is not uniform) and recomputes
//   m_j * gradW_ij = m_j * pair_g(|x_ij|^2, h_ij) * x_ij
// in registers (lists.cuh), so a sweep streams 2 B per neighbour from HBM instead of 8.
//
// Reference: simulation.rs:931-1005 (non-pressure accel), :1552-1592 (divergence operator), :1633-1748 (sources),
// :1751-1808 (pressure accel), :1207-1322 (Jacobi sweep + PressureSolverStatistics), :1378-1516 (loop control),
// :2389-2446 / :2502-2670 (IISPH / HybridDFSPH step orders); boundary terms boundary_winchenbach2020.rs:164-223.
#include "lists.cuh"

namespace {

constexpr int kThreads = 256;

__global__ void k_solver_reset(StepCtl* ctl) {
  SolverCtl s;
  s.k = 0; s.done = 0; s.sweeps = 0; s.normal = 0; s.singular = 0; s.negative = 0;
  s.err_sum = 0.f; s.max_err = 0.f; s.avg = 0.f; s.ticket = 0;
  ctl->solver = s;
}

// ---- shared-memory window ------------------------------------------------------------------------------------
// Particles are sorted by strip-major cell number (sim.cuh), so almost every 2h neighbour of the kThreads consecutive
// particles of a block lies within kHalo positions of the block's range.  The block stages that window of the gathered
// pack (and of {h, m} when h is not uniform) in shared memory with coalesced loads; a neighbour outside the window
// (across a strip edge, or in another size level's grid) is read from global memory instead.  Correctness never
// depends on the window, only the speed does.  Slot t of the window holds particle wa + t, wa = block_first - kHalo
// (mod 2^32: for the first block the slots of "negative" particles stay unused).
static_assert(kThreads == int(ASPH_PAIR_BLOCK), "the neighbour index bias is aligned to the pair-pass block size");
constexpr uint32_t kHalo = 192;
constexpr uint32_t kWin = kThreads + 2 * kHalo;

template <bool AUX>
__device__ __forceinline__ void stage_window(uint32_t n, bool uni, const float4* __restrict__ pack, const float2* __restrict__ hm,
                                             const float* __restrict__ aux, float4 (&s_pack)[kWin], float2 (&s_hm)[kWin], float* s_aux) {
  const uint32_t wa = blockIdx.x * kThreads - kHalo;
  for (uint32_t t = threadIdx.x; t < kWin; t += kThreads) {
    const uint32_t g = wa + t;
    if (g < n) {
      s_pack[t] = __ldg(&pack[g]);
      if (!uni) s_hm[t] = __ldg(&hm[g]);
      if (AUX) s_aux[t] = __ldg(&aux[g]);
    }
  }
  __syncthreads();
}

// Σ over N_2(i) of f(pack[j], x_ij, c_ij, h_ij, aux[j]); the caller multiplies its sums by the returned scale:
//   uniform h (UNI):  c_ij = w'(q)/r un-normalised,   scale = m * 40 / (7 pi (2h)^3)
//   otherwise:        c_ij = m_j * dW/dr / r,         scale = 1
// (every f is linear in c).  A column is consumed 8 rows at a time: one vector load of indices, then the gathers
// (shared memory, or global for the few outside the window), then the arithmetic.  Padding rows point at the particle
// itself (zero distance => c = 0), so there are no per-entry bounds checks.
template <bool UNI, bool AUX, bool WIDE, class F>
__device__ __forceinline__ void pair_rows(const NbCol& col, uint32_t sp, uint32_t sh, uint32_t sa, const float4* __restrict__ pack,
                                          const float2* __restrict__ hm, const float* __restrict__ aux, float xi, float yi, float hi,
                                          const PairShape& shape, F& f) {
  const uint32_t wa = blockIdx.x * kThreads - kHalo;
  for (uint32_t k0 = 0; k0 < col.cn; k0 += 8u) {
    uint32_t off[8];
    col.get8_off<WIDE>(k0, kHalo, off);
#pragma unroll
    for (int half = 0; half < 2; half++) {
      float4 o[4];
      float2 t[4];
      float a[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const uint32_t of = off[half * 4 + u];
        if (of < kWin) {
          o[u] = lds_f4(sp + of * 16u);
          if (!UNI) t[u] = lds_f2(sh + of * 8u);
          if (AUX) a[u] = lds_f1(sa + of * 4u);
        } else {
          const uint32_t jj = wa + of;
          o[u] = __ldg(pack + jj);
          if (!UNI) t[u] = __ldg(hm + jj);
          if (AUX) a[u] = __ldg(aux + jj);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const float dx = xi - o[u].x, dy = yi - o[u].y;
        const float d2 = dx * dx + dy * dy;
        if (UNI) {
          f(o[u], dx, dy, shape(d2), hi, AUX ? a[u] : 0.f);
        } else {
          const float hij = (hi + t[u].x) * 0.5f;
          f(o[u], dx, dy, t[u].y * pair_g(d2, hij), hij, AUX ? a[u] : 0.f);
        }
      }
    }
  }
}
template <bool UNI, bool AUX, class F>
__device__ __forceinline__ float for_each_pair(const NbLists& L, uint32_t i, const float4 (&s_pack)[kWin], const float2 (&s_hm)[kWin],
                                               const float* s_aux, const float4* __restrict__ pack, const float2* __restrict__ hm,
                                               const float* __restrict__ aux, float xi, float yi, float hi, float mi, F f) {
  const NbCol col(L, i);
  const float inv2h = fast_rcp(2.f * hi);
  const PairShape shape(inv2h);
  // shared-window base addresses, pinned in registers (the compiler would otherwise re-derive them from the CTA id
  // special register at every gather)
  uint32_t sp = uint32_t(__cvta_generic_to_shared(&s_pack[0]));
  uint32_t sh = uint32_t(__cvta_generic_to_shared(&s_hm[0]));
  uint32_t sa = AUX ? uint32_t(__cvta_generic_to_shared(s_aux)) : 0u;
  asm volatile("" : "+r"(sp), "+r"(sh), "+r"(sa));
  if (!col.wide) pair_rows<UNI, AUX, false>(col, sp, sh, sa, pack, hm, aux, xi, yi, hi, shape, f);
  else pair_rows<UNI, AUX, true>(col, sp, sh, sa, pack, hm, aux, xi, yi, hi, shape, f);
  return UNI ? mi * (ASPH_KNORM * inv2h * inv2h * inv2h) : 1.f;
}

// ---------------------------------------------------------------------------------------------- K12
template <bool UNI>
__device__ __forceinline__ void viscosity_body(uint32_t i, const NbLists& L, const float4 (&s_pack)[kWin], const float2 (&s_hm)[kWin],
                                               const float* s_aux, const float4* __restrict__ xv_in, const float2* __restrict__ hm,
                                               const float* __restrict__ rho, const PackedParams& P, float dt, float4* __restrict__ xv_out) {
  const float4 me = xv_in[i];
  const float2 own = hm[i];
  const float rho_i = rho[i];
  float ax = 0.f, ay = 0.f;
  if (P.viscosity_type != ASPH_VISC_XSPH && P.viscosity != 0.f) {
    const float scale = for_each_pair<UNI, true>(L, i, s_pack, s_hm, s_aux, xv_in, hm, rho, me.x, me.y, own.x, own.y,
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
  __shared__ float4 s_pack[kWin];
  __shared__ float2 s_hm[kWin];
  __shared__ float s_aux[kWin];
  const bool uni = ctl->hmin == ctl->hmax;
  stage_window<true>(n, uni, xv_in, hm, rho, s_pack, s_hm, s_aux);
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (uni) viscosity_body<true>(i, L, s_pack, s_hm, s_aux, xv_in, hm, rho, P, ctl->dt, xv_out);
  else viscosity_body<false>(i, L, s_pack, s_hm, s_aux, xv_in, hm, rho, P, ctl->dt, xv_out);
}

// ---------------------------------------------------------------------------------------------- K13
// kind 0: -div(v)/dt; 1: -(rho0-rho)/(rho dt^2); 2: both.  Also p <- 0 (packP[0]) and a^p <- 0 (what K14 gives for p = 0,
// so the first sweep does not launch its pressure-acceleration pass).
__global__ void __launch_bounds__(kThreads)
k_source(uint32_t n, NbLists L, const float4* __restrict__ xv, const float2* __restrict__ hm, const float* __restrict__ rho,
         float4* __restrict__ pconst, float4* __restrict__ packP0, float4* __restrict__ packA, const StepCtl* __restrict__ ctl,
         float rho0, int kind) {
  __shared__ float4 s_pack[kWin];
  __shared__ float2 s_hm[kWin];
  const bool uni = ctl->hmin == ctl->hmax;
  if (kind != 1) stage_window<false>(n, uni, xv, hm, nullptr, s_pack, s_hm, nullptr);
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float dt = ctl->dt;
  const float4 me = xv[i];
  float4 pc = pconst[i];
  const float rho_i = rho[i];
  float s = 0.f;
  if (kind != 1) {
    const float2 own = hm[i];
    float sum = 0.f;
    auto body = [&](const float4& o, float dx, float dy, float c, float, float) { sum += c * ((o.z - me.z) * dx + (o.w - me.w) * dy); };
    const float scale = uni ? for_each_pair<true, false>(L, i, s_pack, s_hm, nullptr, xv, hm, nullptr, me.x, me.y, own.x, own.y, body)
                            : for_each_pair<false, false>(L, i, s_pack, s_hm, nullptr, xv, hm, nullptr, me.x, me.y, own.x, own.y, body);
    sum *= scale;
    const float div = sum / rho_i - (me.z * pc.x + me.w * pc.y);
    s = -div / dt;
  }
  if (kind != 0) s += -(rho0 - rho_i) / (rho_i * dt * dt);
  pc.w = s;
  pconst[i] = pc;
  packP0[i] = make_float4(me.x, me.y, 0.f, 0.f);
  packA[i] = make_float4(me.x, me.y, 0.f, 0.f);  // a^p of the first sweep: p = 0 everywhere => exactly zero (simulation.rs:1792-1807)
}

// ---------------------------------------------------------------------------------------------- K14
// a^p_i = -Σ m_j (P_i + P_j) gradW_ij - p_i * Bc * G_i,  P = p / rho^2.
// MODE 0: sweep (skipped once the solver is done); 1: final, v += dt a^p into the xv pack; 2: final + HybridDFSPH
// integration (simulation.rs:2622-2669); 3: final + IISPH integration (simulation.rs:2433-2444); 4: final only.
// P0 / P1: the two pressure packs; the current one is chosen by the parity of the sweeps executed so far, read from
// the control block, so the launch sequence does not depend on when the solver stops.
template <int MODE>
__global__ void __launch_bounds__(kThreads)
k_accel(uint32_t n, NbLists L, const float4* __restrict__ P0, const float4* __restrict__ P1, const float2* __restrict__ hm,
        const float2* __restrict__ gB, float4* __restrict__ packA, StepCtl* ctl, float4* __restrict__ xv, float2* __restrict__ pos,
        float2* __restrict__ vel, float hybrid_factor, const uint32_t* __restrict__ gid) {
  if (MODE == 0 && ctl->solver.done) return;
  __shared__ float4 s_pack[kWin];
  __shared__ float2 s_hm[kWin];
  const float4* __restrict__ packP = (ctl->solver.sweeps & 1) ? P1 : P0;
  // After a sweep that left no particle with positive pressure (normal == 0: every p' was clamped to 0 or was
  // singular) the whole pressure field is zero and so is a^p; skip the pair sum.
  const bool pairs = MODE == 0 || ctl->solver.normal != 0;
  const bool uni = ctl->hmin == ctl->hmax;
  if (pairs) stage_window<false>(n, uni, packP, hm, nullptr, s_pack, s_hm, nullptr);
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 me = packP[i];
  float ax = 0.f, ay = 0.f;
  if (pairs) {
    const float2 own = hm[i];
    auto body = [&](const float4& o, float dx, float dy, float c, float, float) {
      const float f = c * (me.z + o.z);
      ax -= f * dx; ay -= f * dy;
    };
    const float scale = uni ? for_each_pair<true, false>(L, i, s_pack, s_hm, nullptr, packP, hm, nullptr, me.x, me.y, own.x, own.y, body)
                            : for_each_pair<false, false>(L, i, s_pack, s_hm, nullptr, packP, hm, nullptr, me.x, me.y, own.x, own.y, body);
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

// loop control of iisph_pressure_iterations after a sweep (simulation.rs:1453-1477)
__device__ __forceinline__ void solver_decide(SolverCtl& s, unsigned int error_flags, float dt, float rho0, float tol, int max_iters,
                                              int density_mode) {
  s.sweeps += 1;
  const float avg = s.normal > 0 ? s.err_sum / float(s.normal) : __int_as_float(0x7fc00000);
  s.avg = avg;
  bool stop;
  if (density_mode) stop = (s.normal == 0) || (fabsf(avg / rho0) < tol && s.k > 1);
  else stop = (s.normal == 0) || (fabsf(avg) < tol / dt && s.k > 1);
  if (stop || s.k == max_iters || (error_flags & ERRF_SOLVER_NONFINITE)) s.done = 1;
  else s.k += 1;
}

// ---------------------------------------------------------------------------------------------- K15
// (Ap)_i = div(a^p)_i; p' = p + ω (s - Ap) / a_ii, clamped at 0; PressureSolverStatistics reduced per block,
// then by the last block to finish (fixed order => deterministic), which also evaluates the stop rule.
__global__ void __launch_bounds__(kThreads)
k_jacobi(uint32_t n, NbLists L, const float4* __restrict__ packA, float4* __restrict__ P0, float4* __restrict__ P1,
         const float2* __restrict__ hm, const float4* __restrict__ pconst, const float* __restrict__ rho, StepCtl* ctl,
         float* __restrict__ blockstats, float omega, float rho0, float tol, int max_iters, int density_mode,
         const uint32_t* __restrict__ gid) {
  if (ctl->solver.done) return;
  __shared__ float4 s_pack[kWin];
  __shared__ float2 s_hm[kWin];
  const bool uni = ctl->hmin == ctl->hmax;
  stage_window<false>(n, uni, packA, hm, nullptr, s_pack, s_hm, nullptr);
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const float dt = ctl->dt;
  const bool odd = (ctl->solver.sweeps & 1) != 0;
  const float4* __restrict__ packP = odd ? P1 : P0;
  float4* __restrict__ packP_next = odd ? P0 : P1;
  uint32_t c_normal = 0, c_sing = 0, c_neg = 0;
  float e_sum = 0.f, e_max = 0.f;
  bool bad = false;
  // gid != nullptr: multi-GPU.  Ghost particles are skipped (their p' arrives from the owner rank) and the stop rule is
  // applied by k_solver_decide once the statistics of all ranks are summed.
  if (i < n && !(gid && (gid[i] & ASPH_GHOST_BIT))) {
    const float4 me = packA[i];
    const float4 pc = pconst[i];
    const float p_old = packP[i].w;
    float pn = 0.f;
    const float rho_i = rho[i];
    if (fabsf(pc.z) < 10e-4f) {
      c_sing = 1;
    } else {
      const float2 own = hm[i];
      float sum = 0.f;
      auto body = [&](const float4& o, float dx, float dy, float c, float, float) { sum += c * ((o.z - me.z) * dx + (o.w - me.w) * dy); };
      const float scale = uni ? for_each_pair<true, false>(L, i, s_pack, s_hm, nullptr, packA, hm, nullptr, me.x, me.y, own.x, own.y, body)
                              : for_each_pair<false, false>(L, i, s_pack, s_hm, nullptr, packA, hm, nullptr, me.x, me.y, own.x, own.y, body);
      sum *= scale;
      const float Ap = sum / rho_i - (me.z * pc.x + me.w * pc.y);
      const float resid = pc.w - Ap;
      pn = p_old + omega * resid / pc.z;
      if (!isfinite(Ap) || !isfinite(pn)) bad = true;
      const float perr = density_mode ? rho_i * dt * dt * resid : dt * resid;
      if (pn <= 0.f) { pn = 0.f; c_neg = 1; }
      else { c_normal = 1; e_sum = perr; e_max = fabsf(perr); }
    }
    packP_next[i] = make_float4(me.x, me.y, pn / (rho_i * rho_i), pn);
  }
  if (bad) atomicOr(&ctl->error_flags, ERRF_SOLVER_NONFINITE);
  // block reduction
  for (int o = 16; o > 0; o >>= 1) {
    c_normal += __shfl_xor_sync(0xffffffffu, c_normal, o);
    c_sing += __shfl_xor_sync(0xffffffffu, c_sing, o);
    c_neg += __shfl_xor_sync(0xffffffffu, c_neg, o);
    e_sum += __shfl_xor_sync(0xffffffffu, e_sum, o);
    e_max = fmaxf(e_max, __shfl_xor_sync(0xffffffffu, e_max, o));
  }
  __shared__ float sh[5][kThreads];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sh[0][w] = float(c_normal); sh[1][w] = float(c_sing); sh[2][w] = float(c_neg); sh[3][w] = e_sum; sh[4][w] = e_max; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f, c = 0.f, d = 0.f, e = 0.f;
    for (int k = 0; k < kThreads / 32; k++) { a += sh[0][k]; b += sh[1][k]; c += sh[2][k]; d += sh[3][k]; e = fmaxf(e, sh[4][k]); }
    float* out = blockstats + 5 * size_t(blockIdx.x);
    out[0] = a; out[1] = b; out[2] = c; out[3] = d; out[4] = e;  // counts <= 256 are exact in fp32
    __threadfence();
    const unsigned int t = atomicAdd(&ctl->solver.ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // final reduction in a fixed order: thread t takes blocks t, t+256, ...; then a fixed tree
  unsigned long long tn = 0, ts = 0, tg = 0;
  float td = 0.f, te = 0.f;
  for (uint32_t b = threadIdx.x; b < gridDim.x; b += kThreads) {
    const float* in = blockstats + 5 * size_t(b);
    tn += (unsigned long long)__ldcg(in + 0); ts += (unsigned long long)__ldcg(in + 1); tg += (unsigned long long)__ldcg(in + 2);
    td += __ldcg(in + 3); te = fmaxf(te, __ldcg(in + 4));
  }
  __shared__ unsigned long long shn[kThreads], shs[kThreads], shg[kThreads];
  shn[threadIdx.x] = tn; shs[threadIdx.x] = ts; shg[threadIdx.x] = tg; sh[3][threadIdx.x] = td; sh[4][threadIdx.x] = te;
  __syncthreads();
  for (int s = kThreads / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      shn[threadIdx.x] += shn[threadIdx.x + s]; shs[threadIdx.x] += shs[threadIdx.x + s]; shg[threadIdx.x] += shg[threadIdx.x + s];
      sh[3][threadIdx.x] += sh[3][threadIdx.x + s]; sh[4][threadIdx.x] = fmaxf(sh[4][threadIdx.x], sh[4][threadIdx.x + s]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (gid) {
      ctl->solver.partial[0] = double(shn[0]); ctl->solver.partial[1] = double(shs[0]); ctl->solver.partial[2] = double(shg[0]);
      ctl->solver.partial[3] = double(sh[3][0]);
      ctl->solver.max_err = sh[4][0];
      ctl->solver.ticket = 0;
      return;
    }
    SolverCtl s = ctl->solver;
    s.normal = shn[0]; s.singular = shs[0]; s.negative = shg[0]; s.err_sum = sh[3][0]; s.max_err = sh[4][0];
    s.ticket = 0;
    solver_decide(s, ctl->error_flags, dt, rho0, tol, max_iters, density_mode);
    ctl->solver = s;
  }
}

// multi-GPU: the stop rule on the statistics summed over all ranks (every rank computes the same decision)
__global__ void k_solver_decide(StepCtl* ctl, float rho0, float tol, int max_iters, int density_mode) {
  SolverCtl s = ctl->solver;
  if (s.done) return;
  s.normal = (unsigned long long)s.partial[0]; s.singular = (unsigned long long)s.partial[1]; s.negative = (unsigned long long)s.partial[2];
  s.err_sum = float(s.partial[3]);
  solver_decide(s, ctl->error_flags, ctl->dt, rho0, tol, max_iters, density_mode);
  ctl->solver = s;
}

NbLists lists_of(asph_sim* sim) {
  NbLists L;
  L.pool = sim->nbpool.p; L.slice_base = sim->slice_base.p; L.cnt = sim->cnt.p;
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

int launch_source(asph_sim* sim, int kind) {
  const uint32_t n = sim->n;
  if (n == 0) return ASPH_OK;
  const uint32_t blocks = (n + kThreads - 1) / kThreads;
  k_solver_reset<<<1, 1, 0, sim->stream>>>(sim->ctl);
  LAUNCH_CHECK();
  k_source<<<blocks, kThreads, 0, sim->stream>>>(n, lists_of(sim), sim->xv[sim->xv_cur].p, sim->hm.p, sim->rho.p, sim->pconst.p,
                                                 sim->packP[0].p, sim->packA.p, sim->ctl, sim->pp.rest_density, kind);
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
  const uint32_t blocks = (n + kThreads - 1) / kThreads;
  CUDA_TRY(sim->blockstats.ensure(size_t(blocks) * 5 + 8));
  cudaStream_t st = sim->stream;
  const NbLists L = lists_of(sim);
  const uint32_t* gid = sim->dist ? sim->refid[sim->cur].p : nullptr;
  int launched = 0;
  int& predicted = sim->predicted_sweeps[density_mode ? 1 : 0];
  int batch = std::max(1, std::min(predicted, 256));
  const int max_sweeps = sim->pp.max_iters + 1;
  struct Timed { int sweep; cudaEvent_t e0, e1, e1b, e2; };  // e0..e1: K14; e1b..e2: K15 (halo exchanges excluded)
  std::vector<Timed> timed;
  for (;;) {
    for (int b = 0; b < batch && launched < max_sweeps; b++, launched++) {
      const bool time_it = sim->kt_every > 0 && (launched % sim->kt_every) == 0;
      Timed tm{launched, nullptr, nullptr, nullptr, nullptr};
      if (time_it) { tm.e0 = kt_event(sim); tm.e1 = kt_event(sim); tm.e1b = kt_event(sim); tm.e2 = kt_event(sim); cudaEventRecord(tm.e0, st); }
      if (launched > 0) {  // sweep 0: a^p = 0 was written by k_source
        k_accel<0><<<blocks, kThreads, 0, st>>>(n, L, sim->packP[0].p, sim->packP[1].p, sim->hm.p, sim->gB.p, sim->packA.p, sim->ctl, nullptr,
                                                nullptr, nullptr, 0.f, gid);
        LAUNCH_CHECK();
        if (time_it) cudaEventRecord(tm.e1, st);
        if (sim->dist) TRY(dist_halo(sim, sim->packA.p, 16));
      } else if (time_it) {
        cudaEventRecord(tm.e1, st);
      }
      if (time_it) cudaEventRecord(tm.e1b, st);
      k_jacobi<<<blocks, kThreads, 0, st>>>(n, L, sim->packA.p, sim->packP[0].p, sim->packP[1].p, sim->hm.p, sim->pconst.p, sim->rho.p,
                                            sim->ctl, sim->blockstats.p, sim->pp.jacobi_omega, sim->pp.rest_density, max_avg_error,
                                            sim->pp.max_iters, density_mode ? 1 : 0, gid);
      LAUNCH_CHECK();
      if (time_it) { cudaEventRecord(tm.e2, st); timed.push_back(tm); }
      if (sim->dist) {
        TRY(dist_solver_reduce(sim));
        k_solver_decide<<<1, 1, 0, st>>>(sim->ctl, sim->pp.rest_density, max_avg_error, sim->pp.max_iters, density_mode ? 1 : 0);
        LAUNCH_CHECK();
        TRY(dist_halo_pressure(sim));
      }
    }
    TRY(sync_ctl(sim));
    const SolverCtl& s = sim->ctl_host->solver;
    for (const Timed& tm : timed) {
      float a = 0.f, b2 = 0.f;
      if (tm.sweep < s.sweeps && cudaEventElapsedTime(&a, tm.e0, tm.e1) == cudaSuccess && cudaEventElapsedTime(&b2, tm.e1b, tm.e2) == cudaSuccess) {
        if (tm.sweep > 0) { sim->kt_ms[ASPH_KT_ACCEL_SWEEP] += a; sim->kt_samples[ASPH_KT_ACCEL_SWEEP]++; }  // sweep 0 has no K14 launch
        sim->kt_ms[ASPH_KT_JACOBI_SWEEP] += b2; sim->kt_samples[ASPH_KT_JACOBI_SWEEP]++;
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
