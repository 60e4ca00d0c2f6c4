// neighbors.cu — neighbour lists (sliced ELL, lists.cuh) from the multi-resolution cell grid, fused with every per-particle
// quantity that only needs the pre-advection snapshot {x, h, m}: boundary terms (K7), density (K9), the a_ii
// diagonal (K11) and the un-normalised surface normal of the level-set detector (first half of K3).
//
// Neighbour set (bit exact): N_f(i) = { j : |x_ij|^2 < ((h_i+h_j)*0.5*f)^2 }, fp32, no FMA contraction —
// the final set of build_neighborhood_list_rstar after its symmetrize pass (neighborhood_search.rs:123-185, checked
// by the reference's own brute-force test at :216-237 and simulation.rs:1810-1863).  The W and F segments of a
// particle's ELL column (lists.cuh) are N_2 (what NeighborhoodCache::filter_down leaves, neighborhood_search.rs:56-70),
// the E segment the rest of N_{f_ext} used only by the level-set estimation.
#include "lists.cuh"

namespace {

constexpr int kThreads = 128;
constexpr uint32_t kStageCap = 48;  // hits a thread stages in shared memory between its scan and the writing of its column

// LookupTable1D::get over [-1, 1], 10000 steps (lookup_table.rs:32-48), same operation order
__device__ __forceinline__ float lut_get(const float* __restrict__ data, float x) {
  float fidx = __fmul_rn(__fmul_rn(__fsub_rn(x, -1.f), 0.5f), 10000.f);
  float fl = floorf(fidx);
  float t = __fsub_rn(fidx, fl);
  int idx = int(fl);
  if (idx + 1 >= 10001) return data[idx];
  return __fadd_rn(__fmul_rn(data[idx], __fsub_rn(1.f, t)), __fmul_rn(data[idx + 1], t));
}

// BoundaryWinchenbach2020::update_after_advect (boundary_winchenbach2020.rs:58-152): Σ λ·penalty and Σ ∇(λ·penalty)
__device__ __forceinline__ void boundary_terms(const PackedParams& P, const float* __restrict__ lut, float x, float y, float h,
                                               float& lam_sum, float& gx_sum, float& gy_sum) {
  lam_sum = 0.f; gx_sum = 0.f; gy_sum = 0.f;
  const float sr = __fmul_rn(h, 2.f);
  const float eps = P.sdf_gradient_eps;
  const float inv_2eps = __fdiv_rn(1.f, __fmul_rn(2.f, eps));
  const int ns = sdf_count(P);
  for (int s = 0; s < ns; s++) {
    float d = __fdiv_rn(sdf_probe(P, s, x, y), sr);
    if (!(d < 1.f)) continue;
    // finite_diff_gradient sdf.rs:50-62
    float gx = __fmul_rn(__fsub_rn(sdf_probe(P, s, __fadd_rn(x, eps), y), sdf_probe(P, s, __fsub_rn(x, eps), y)), inv_2eps);
    float gy = __fmul_rn(__fsub_rn(sdf_probe(P, s, x, __fadd_rn(y, eps)), sdf_probe(P, s, x, __fsub_rn(y, eps))), inv_2eps);
    float gn = __fsqrt_rn(dist_sq_exact(gx, gy));
    if (gn < 0.00001f) continue;
    gx = __fdiv_rn(gx, gn); gy = __fdiv_rn(gy, gn);
    float pen, pder;
    switch (P.penalty) {
      case ASPH_PENALTY_NONE: pen = 1.f; pder = 0.f; break;
      case ASPH_PENALTY_LINEAR: pen = __fsub_rn(1.f, d); pder = -1.f; break;
      case ASPH_PENALTY_QUADRATIC1:
        if (d > 0.f) { pen = 1.f; pder = 0.f; }
        else if (d > -1.f) { pen = __fadd_rn(__fmul_rn(__fmul_rn(0.5f, d), d), 1.f); pder = d; }
        else { pen = __fsub_rn(0.5f, d); pder = -1.f; }
        break;
      default:
        if (d > 0.f) { pen = 1.f; pder = 0.f; }
        else if (d > -0.5f) { pen = __fadd_rn(__fmul_rn(d, d), 1.f); pder = __fmul_rn(2.f, d); }
        else { pen = __fsub_rn(0.75f, d); pder = -1.f; }
        break;
    }
    float lam, lamd;
    if (d <= -1.f) { lam = 1.f; lamd = 0.f; }
    else { lam = lut_get(lut, d); lamd = lut_get(lut + 10001, d); }
    float scale = __fadd_rn(__fmul_rn(pder, lam), __fmul_rn(pen, lamd));
    lam_sum = __fadd_rn(lam_sum, __fmul_rn(lam, pen));
    gx_sum = __fadd_rn(gx_sum, __fmul_rn(__fdiv_rn(gx, sr), scale));
    gy_sum = __fadd_rn(gy_sum, __fmul_rn(__fdiv_rn(gy, sr), scale));
  }
}

// Visit every particle j stored in a grid cell that overlaps the search box of particle i, level by level.
// f(j, xyhm[j]): the candidates' records are requested four at a time before the first of them is looked at (the scan is
// a chain load -> test otherwise: a third of the kernel's stall samples sat on the first use of the loaded record)
template <class F>
__device__ __forceinline__ void for_each_candidate(float xi, float yi, float hi, const StepCtl* __restrict__ ctl,
                                                   const uint32_t* __restrict__ cellstart, const float4* __restrict__ xyhm, float f_search, F f) {
  const int nl = ctl->nlevels;
  const float ox = ctl->origin_x, oy = ctl->origin_y;
  for (int b = 0; b < nl; b++) {
    const GridLevel g = ctl->lv[b];
    const float r = f_search * (hi + g.hmax) * 0.5f * ASPH_SLACK;
    const int cx0 = max(0, int(floorf((xi - r - ox) * g.inv_cell)));
    const int cx1 = min(g.nx - 1, int(floorf((xi + r - ox) * g.inv_cell)));
    const int cy0 = max(0, int(floorf((yi - r - oy) * g.inv_cell)));
    const int cy1 = min(g.ny - 1, int(floorf((yi + r - oy) * g.inv_cell)));
    if (cx0 > cx1) continue;
    // cells are stored strip-major (grid.cu): a row's cells are contiguous inside one strip of 2^strip_log2 columns
    const int sl = g.strip_log2;
    for (int cy = cy0; cy <= cy1; cy++) {
      for (int st = cx0 >> sl; st <= (cx1 >> sl); st++) {
        const int ca = max(cx0, st << sl), cb = min(cx1, (st << sl) + (1 << sl) - 1);
        const uint32_t s = cellstart[cell_index(g, ca, cy)], e = cellstart[cell_index(g, cb, cy) + 1];
        for (uint32_t j = s; j < e; j += 4u) {
          float4 o[4];
#pragma unroll
          for (uint32_t u = 0; u < 4u; u++) o[u] = __ldg(&xyhm[min(j + u, e - 1u)]);
#pragma unroll
          for (uint32_t u = 0; u < 4u; u++) if (j + u < e) f(j + u, o[u]);
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads)
k_neighbors(uint32_t n, const float4* __restrict__ xyhm, const StepCtl* __restrict__ ctl_in, StepCtl* ctl,
            const uint32_t* __restrict__ cellstart, const PackedParams P, const float* __restrict__ lut, float f_ext, float f_near,
            uint32_t pool_cap64, uint16_t* __restrict__ pool, uint32_t* __restrict__ slice_base, uint32_t* __restrict__ cnt, uint32_t* __restrict__ cnt_ext,
            uint32_t* __restrict__ far_idx, uint32_t* __restrict__ far_cnt,
            float* __restrict__ rho_out, float2* __restrict__ gB_out, float4* __restrict__ pconst, float* __restrict__ lam_sum_out,
            float2* __restrict__ lam_grad_out, float2* __restrict__ nrm_out, const uint32_t* __restrict__ gid) {
  __shared__ uint32_t s_stage[kStageCap * kThreads];  // [hit][thread]: conflict-free
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t i0 = i & ~(ASPH_PAIR_BLOCK - 1u);  // the index bias is per 256-particle block (lists.cuh)
  if (ctl_in->error_flags & ERRF_CELL_BUDGET) {  // no usable grid (k_make_levels): empty columns, the host reports the failure
    if (i < n) { cnt[i] = 0u; cnt_ext[i] = 0u; if (lane == 0) slice_base[i >> 5] = 0u; }
    return;
  }
  const bool active = i < n;
  float4 me = make_float4(0.f, 0.f, 1.f, 0.f);
  if (active) me = xyhm[i];
  const float xi = me.x, yi = me.y, hi = me.z;

  // pass 1, the ONE scan of the candidates: counts — 2h neighbours inside / outside the pair passes' shared-memory window
  // (lists.cuh), the extended range, and whether every index stored outside the window segment fits 16 bits around the
  // block —, the pair sums over N_2(i) (density, a_ii and the surface normal need nothing but the snapshot), and the hits
  // themselves, staged in shared memory (index | 2h flag) so that writing the lists needs no second scan.  A thread with
  // more than kStageCap hits (a coarse particle next to fine ones) scans again in pass 2.
  const uint32_t win0 = i0 - ASPH_PAIR_HALO;
  uint32_t cw = 0, cf = 0, ce = 0;
  bool fits = true, fits_far = true, fits_ext = true;
  float rho = 0.f, Sx = 0.f, Sy = 0.f, Q = 0.f, Nx = 0.f, Ny = 0.f;
  if (active) {
    for_each_candidate(xi, yi, hi, ctl_in, cellstart, xyhm, f_ext, [&](uint32_t j, const float4& o) {
      const float ddx = __fsub_rn(xi, o.x), ddy = __fsub_rn(yi, o.y);
      const float d2 = dist_sq_exact(ddx, ddy);
      if (d2 < support_sq_exact(hi, o.z, f_ext)) {
        const bool near = d2 < support_sq_exact(hi, o.z, f_near);
        if (ce < kStageCap) s_stage[ce * kThreads + threadIdx.x] = j | (near ? 0x80000000u : 0u);
        ce++;
        if (near) {
          if (j - win0 < ASPH_PAIR_WIN) cw++;
          else { cf++; if (j - i0 + 32768u > 65535u) fits_far = false; }
          float w, g;
          pair_wg(d2, (hi + o.z) * 0.5f, w, g);
          rho += o.w * w;
          const float c = o.w * g;
          Sx += c * ddx; Sy += c * ddy;
          Q += c * g * d2;  // m_j |gradW|^2
          Nx += g * ddx; Ny += g * ddy;
        } else if (j - i0 + 32768u > 65535u) {
          fits_ext = false;
        }
      }
    });
  }
  // far 2h neighbours become window rows when the tile's far table has room for all of them (lists.cuh)
  uint32_t far_base = 0xffffffffu;
  if (cf > 0u) {
    const uint32_t tile = i / ASPH_PAIR_BLOCK;
    const uint32_t b = atomicAdd(&far_cnt[tile], cf);
    if (b + cf <= ASPH_PAIR_FAR) {
      far_base = b;
    } else {  // no room: keep them as F entries; slots claimed past the end of the table are never staged, the ones
              // inside it must still name a valid particle
      for (uint32_t s = b; s < ASPH_PAIR_FAR; s++) far_idx[tile * ASPH_PAIR_FAR + s] = i;
    }
  }
  const bool in_table = far_base != 0xffffffffu;
  if (in_table) { cw += cf; cf = 0u; }
  // the extended-range entries and the F entries that stayed decide whether the slice needs 32-bit entries
  if (in_table && !fits_ext) fits = false;
  else if (!in_table && !(fits_ext && fits_far)) fits = false;
  const uint32_t cn = cw + cf;
  const bool wide = !__all_sync(0xffffffffu, fits);
  uint32_t chunks = nb_col_chunks(cw, cf, ce, wide), ce_max = ce;
  for (int o = 16; o > 0; o >>= 1) {
    chunks = max(chunks, __shfl_xor_sync(0xffffffffu, chunks, o));
    ce_max = max(ce_max, __shfl_xor_sync(0xffffffffu, ce_max, o));
  }
  // a chunk = 32 lanes x 16 B = 4 pool units of 64 uint16.  Every slice owns at least two chunks: the sweep kernels stage
  // the first two of every warp with one bulk copy each, whatever the column lengths (sweep_kernel.inc).
  const uint32_t units = 4u * max(chunks, 2u);
  uint32_t base64 = 0;
  bool fits_pool = true;
  if (lane == 0) {
    base64 = atomicAdd(&ctl->list_used, units);
    atomicMax(&ctl->max_count, ce_max);
    if (ce_max > 20000u) atomicOr(&ctl->error_flags, ERRF_NEIGHBOR_OVERFLOW);  // MAX_NEIGHBOR_COUNT neighborhood_search.rs:3,148-150
    fits_pool = !(base64 + units > pool_cap64 || base64 + units < base64);
    if (!fits_pool) atomicOr(&ctl->error_flags, ERRF_LIST_CAPACITY);
    // a slice that found no room points at the start of the pool (always at least two chunks long): whatever a later pass
    // of this void attempt reads through it stays in bounds
    slice_base[i >> 5] = fits_pool ? (base64 | (wide ? 0x80000000u : 0u)) : 0u;
  }
  base64 = __shfl_sync(0xffffffffu, base64, 0);
  fits_pool = __shfl_sync(0xffffffffu, fits_pool ? 1 : 0, 0) != 0;
  if (!active) return;
  if (!fits_pool) { cnt[i] = 0u; cnt_ext[i] = 0u; return; }  // empty column: later passes stay in bounds
  cnt[i] = cw | (min(cf, 0x7ffffu) << 12) | ((gid && (gid[i] & ASPH_GHOST_BIT)) ? 0x80000000u : 0u);
  cnt_ext[i] = ce;

  // pass 2: write the entries (window segment, far 2h segment, extended-range rest), in the order of the scan
  uint16_t* slice = pool + size_t(base64) * 64u;
  const uint32_t bias = i0 - 32768u;
  {
    uint32_t kw = 0, kf = 0, ke = nb_pad4(cf), kt = 0;
    auto place = [&](uint32_t j, bool near) {
      if (near) {
        if (j - win0 < ASPH_PAIR_WIN) {
          if (!(P.self_last && j == i)) nb_store_w(slice, lane, kw++, (j - win0) * 16u);
        } else if (in_table) {
          far_idx[(i / ASPH_PAIR_BLOCK) * ASPH_PAIR_FAR + far_base + kt] = j;
          nb_store_w(slice, lane, kw++, (ASPH_PAIR_WIN + far_base + kt) * 16u);
          kt++;
        } else {
          nb_store_fe(slice, wide, lane, cw, kf++, j, bias);
        }
      } else {
        nb_store_fe(slice, wide, lane, cw, ke++, j, bias);
      }
    };
    if (ce <= kStageCap) {
      for (uint32_t k = 0; k < ce; k++) {
        const uint32_t e = s_stage[k * kThreads + threadIdx.x];
        place(e & 0x7fffffffu, (e >> 31) != 0u);
      }
    } else {
      for_each_candidate(xi, yi, hi, ctl_in, cellstart, xyhm, f_ext, [&](uint32_t j, const float4& o) {
        const float d2 = dist_sq_exact(__fsub_rn(xi, o.x), __fsub_rn(yi, o.y));
        if (d2 < support_sq_exact(hi, o.z, f_near)) place(j, true);
        else if (d2 < support_sq_exact(hi, o.z, f_ext)) place(j, false);
      });
    }
    if (P.self_last) nb_store_w(slice, lane, kw++, (i - win0) * 16u);  // the particle's own row closes the W segment (solver.cu, R4)
    for (uint32_t r = cw; r < nb_pad8(cw); r++) nb_store_w(slice, lane, r, (i - win0) * 16u);  // padding: the particle itself
    for (uint32_t k = cf; k < nb_pad4(cf); k++) nb_store_fe(slice, wide, lane, cw, k, i, bias);
  }

  nrm_out[i] = make_float2(Nx, Ny);  // Σ_j ∇W_ij; K3 scales it by -(m_i/ρ0)
  if (rho_out == nullptr) return;    // lists (and the surface normal) only: the second pass of level_estimation_after_advection

  // boundary terms
  float lam, Gx, Gy;
  boundary_terms(P, lut, xi, yi, hi, lam, Gx, Gy);

  // density, simulation.rs:1018-1047
  rho += lam;
  unsigned int err = 0;
  if (!isfinite(rho)) err |= ERRF_NONFINITE;
  else if (!(rho > 0.0001f)) err |= ERRF_DENSITY;
  // a_ii, iisph_aii boundary_winchenbach2020.rs:270-304 (Consistent{Simple,Symmetric}Gradient)
  const float rho0 = P.rest_density;
  const float rho2 = rho * rho;
  const float Bc = (P.opdisc == ASPH_OP_CONSISTENT_SYMMETRIC_GRADIENT) ? rho0 * (1.f / rho2 + 1.f / (rho0 * rho0)) : rho0 / rho2;
  const float ax = Sx / rho2 + Bc * Gx, ay = Sy / rho2 + Bc * Gy;
  const float bx = Sx / rho + rho0 * Gx / rho, by = Sy / rho + rho0 * Gy / rho;
  const float aii = (ax * bx + ay * by) + (me.w * Q) / (rho2 * rho);
  if (P.opdisc != ASPH_OP_WINCHENBACH2020) {  // that operator's diagonal needs the neighbours' densities: k_aii_w2020
    if (!isfinite(aii)) err |= ERRF_NONFINITE;
    else if (aii < 0.f) err |= ERRF_NEG_AII;
  }
  if (err && !(gid && (gid[i] & ASPH_GHOST_BIT))) atomicOr(&ctl->error_flags, err);  // a ghost's neighbourhood is incomplete by design
  rho_out[i] = rho;
  gB_out[i] = make_float2(Bc * Gx, Bc * Gy);
  pconst[i] = make_float4(rho0 * Gx / rho, rho0 * Gy / rho, aii, 0.f);
  lam_sum_out[i] = lam;
  lam_grad_out[i] = make_float2(Gx, Gy);
}

// Winchenbach2020 operator.  Its divergence weights every neighbour with m_j / rho_j instead of
// m_j / rho_i (simulation.rs:1571-1575) and drops rho_b / rho_i from the boundary term (boundary_winchenbach2020.rs:
// 207-213), so the diagonal (boundary_winchenbach2020.rs:236-269)
//   a_ii = (S / rho_i^2 + rho_b G / rho_i^2) . (T + G) + m_i U / rho_i^2,
//   S = sum m_j gradW,  T = sum (m_j / rho_j) gradW,  U = sum (m_j / rho_j) |gradW|^2
// needs the neighbours' densities: a second pass over the 2h columns after the density pass.  It also writes
// hv = {h, m / rho}, which the divergence passes (K13, K15) gather in place of {h, m}, and pconst.xy = G.
__global__ void __launch_bounds__(kThreads)
k_aii_w2020(uint32_t n, NbLists L, const float4* __restrict__ xyhm, const float* __restrict__ rho, const float2* __restrict__ lam_grad,
            float rho0, float4* __restrict__ pconst, float2* __restrict__ hv, StepCtl* ctl, const uint32_t* __restrict__ gid) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // The neighbour pass gave up on some slices (pool too small: the step is redone from that pass after neighbors_grow)
  // or had no grid at all: their densities were never written, and flagging "non-finite" for them would outlive the retry.
  if (ctl->error_flags & (ERRF_LIST_CAPACITY | ERRF_CELL_BUDGET)) return;
  const float4 me = xyhm[i];
  const float rho_i = rho[i];
  hv[i] = make_float2(me.z, me.w / rho_i);
  const NbCol col(L, i);
  float Sx = 0.f, Sy = 0.f, Tx = 0.f, Ty = 0.f, U = 0.f;
  for (uint32_t k = 0; k < col.cn; k++) {
    const uint32_t j = col.get(k);
    const float4 o = __ldg(&xyhm[j]);
    const float dx = me.x - o.x, dy = me.y - o.y;
    const float d2 = dx * dx + dy * dy;
    float w, g;
    pair_wg(d2, (me.z + o.z) * 0.5f, w, g);
    const float c = o.w * g, v = o.w / __ldg(&rho[j]);
    Sx += c * dx; Sy += c * dy;
    Tx += (v * g) * dx; Ty += (v * g) * dy;
    U += v * (g * g * d2);
  }
  const float2 G = lam_grad[i];
  const float rho2 = rho_i * rho_i;
  const float ax = Sx / rho2 + (rho0 / rho2) * G.x, ay = Sy / rho2 + (rho0 / rho2) * G.y;
  const float aii = (ax * (Tx + G.x) + ay * (Ty + G.y)) + (me.w * U) / rho2;
  unsigned int err = 0;
  if (!isfinite(aii)) err |= ERRF_NONFINITE;
  else if (aii < 0.f) err |= ERRF_NEG_AII;
  if (err && !(gid && (gid[i] & ASPH_GHOST_BIT))) atomicOr(&ctl->error_flags, err);
  pconst[i] = make_float4(G.x, G.y, aii, 0.f);
}

// Support length from the particle distribution (simulation.rs:1873-1971; "Constrained Neighbor Lists for SPH-based Fluid
// Simulations" eq. 3-4), after the 2h lists of the step exist and before the boundary terms of the step replace the
// previous ones (simulation.rs:2090-2143, 2179):
//   FromDistribution / Clamped1 / Clamped2:  V = (1 - min(Lambda_prev, 0.5)) / sum_j W(x_ij, h_ij)
//   FromDistribution2:                       V = (m_i / rho0) / (sum_j (m_j / rho0) W(x_ij, h_ij) + Lambda_prev)
//   h_next = 0.5 * ETA * sqrt(V / pi) + 0.5 * h_i   [clamped to 1x / 2x the length from the mass]
// W in the reference's operation order (sph_kernels.rs:49-52) so that the only difference to the CPU is the summation order.
__global__ void __launch_bounds__(kThreads)
k_estimate_h(uint32_t n, NbLists L, const float4* __restrict__ xyhm, const float* __restrict__ lam_sum, int mode, float rho0,
             float* __restrict__ hnext, float* __restrict__ lamprev, StepCtl* ctl) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (ctl->error_flags & (ERRF_LIST_CAPACITY | ERRF_CELL_BUDGET)) return;  // void attempt: the step is redone / fails (hnext untouched)
  const float4 me = xyhm[i];
  const NbCol col(L, i);
  float sum = 0.f;
  for (uint32_t k = 0; k < col.cn; k++) {
    const float4 o = __ldg(&xyhm[col.get(k)]);
    const float r = __fsqrt_rn(dist_sq_exact(__fsub_rn(me.x, o.x), __fsub_rn(me.y, o.y)));
    const float hij = __fmul_rn(__fadd_rn(me.z, o.z), 0.5f);
    // cubic_kernel_2d: (10 / (7 pi (h h))) * w(r / (2 h))
    const float q = __fdiv_rn(r, __fmul_rn(2.f, hij));
    float w;
    if (q < 0.5f) w = __fadd_rn(__fmul_rn(6.f, __fsub_rn(__fmul_rn(__fmul_rn(q, q), q), __fmul_rn(q, q))), 1.f);
    else if (q < 1.f) { const float v = __fsub_rn(1.f, q); w = __fmul_rn(2.f, __fmul_rn(__fmul_rn(v, v), v)); }
    else w = 0.f;
    const float W = __fmul_rn(__fdiv_rn(10.f, __fmul_rn(__fmul_rn(7.f, ASPH_PI_F), __fmul_rn(hij, hij))), w);
    sum = __fadd_rn(sum, mode == ASPH_H_FROM_DISTRIBUTION2 ? __fmul_rn(__fdiv_rn(o.w, rho0), W) : W);
  }
  const float lp = lamprev[i];
  float vol;
  if (mode == ASPH_H_FROM_DISTRIBUTION2) vol = __fdiv_rn(__fdiv_rn(me.w, rho0), __fadd_rn(sum, lp));
  else vol = __fdiv_rn(__fsub_rn(1.f, fminf(lp, 0.5f)), sum);
  if (!(vol >= 0.f)) { atomicOr(&ctl->error_flags, ERRF_NONFINITE); return; }  // assert!(volume_estimate >= 0.)
  const float h_new = __fmul_rn(1.9f, __fsqrt_rn(__fmul_rn(vol, ASPH_FRAC_1_PI_F)));
  float hn = __fadd_rn(__fmul_rn(0.5f, h_new), __fmul_rn(0.5f, me.z));
  if (mode == ASPH_H_FROM_DISTRIBUTION_CLAMPED1) hn = fminf(hn, __fmul_rn(1.f, h_from_mass(me.w, rho0)));
  if (mode == ASPH_H_FROM_DISTRIBUTION_CLAMPED2) hn = fminf(hn, __fmul_rn(2.f, h_from_mass(me.w, rho0)));
  hnext[i] = hn;
  lamprev[i] = lam_sum[i];  // this step's boundary terms (k_neighbors) are what the next step's estimate sees
}

// ---- constrain_neighborhood_count (simulation.rs:2145-2177) ---------------------------------------------------------
// A particle with more than optimal_neighbor_number + 5 = 19 neighbours in N_2 takes, as its new smoothing length, the
// (count - 19)-th largest of the "fringe" values 2 |x_ij| - 2 h_j of its neighbours (all read from this step's h: the new
// lengths go to a second array first).  The operations follow the reference one by one: the neighbour predicate of
// everything after depends on h.  The reference keeps its lists and lets the pairs that fall out of the shrunken supports
// contribute zeros; here the lists are rebuilt from the new h instead (launch_neighbors after this), which leaves the
// same non-zero pairs — the pair passes assume that every stored pair lies inside the support.
constexpr uint32_t kConstrainTarget = 19u;  // (ETA * 2)^2 = 14.44 as usize, + 5 (simulation.rs:386, 2147)
__global__ void __launch_bounds__(kThreads)
k_constrain(uint32_t n, NbLists L, const float4* __restrict__ xyhm, float* __restrict__ h_next, StepCtl* ctl) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 me = xyhm[i];
  const NbCol col(L, i);
  const uint32_t cn = col.cn;
  float hn = me.z;
  if (cn > kConstrainTarget) {
    auto fringe = [&](uint32_t k) {
      const float4 o = __ldg(&xyhm[col.get(k)]);
      const float d = __fsqrt_rn(dist_sq_exact(__fsub_rn(me.x, o.x), __fsub_rn(me.y, o.y)));
      return __fsub_rn(__fmul_rn(2.f, d), __fmul_rn(o.z, 2.f));
    };
    // the value of rank r in descending order, by counting (a column has some thirty entries; the mode is not a hot path)
    const uint32_t r = cn - kConstrainTarget;
    bool found = false;
    for (uint32_t k = 0; k < cn && !found; k++) {
      const float v = fringe(k);
      uint32_t greater = 0, equal = 0;
      for (uint32_t m = 0; m < cn; m++) {
        const float w = fringe(m);
        greater += w > v ? 1u : 0u;
        equal += w == v ? 1u : 0u;
      }
      if (greater <= r && r < greater + equal) { hn = v; found = true; }
    }
    if (!found || !(hn < me.z) || !(hn >= 0.f)) { atomicOr(&ctl->error_flags, ERRF_CONSTRAIN); hn = me.z; }
  }
  h_next[i] = hn;
}
// the new lengths replace the old ones wherever a later kernel reads them; h range and CFL minimum of the new lengths
// ((2h)^2 / (v.v + 0.01) with the new h, simulation.rs:2182-2191: the reference computes dt after this)
__global__ void __launch_bounds__(kThreads)
k_apply_h(uint32_t n, const float* __restrict__ h_next, const float4* __restrict__ xv, float4* __restrict__ xyhm, float2* __restrict__ hm, StepCtl* ctl) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const float inf = __int_as_float(0x7f800000);
  float hmin = inf, hmax = -inf, cfl = inf;
  if (i < n) {
    const float h = h_next[i];
    float4 r = xyhm[i]; r.z = h; xyhm[i] = r;
    float2 q = hm[i]; q.x = h; hm[i] = q;
    const float4 v = xv[i];
    const float sr = __fmul_rn(h, 2.f);
    cfl = __fdiv_rn(__fmul_rn(sr, sr), __fadd_rn(dist_sq_exact(v.z, v.w), 0.01f));
    hmin = hmax = h;
  }
  for (int o = 16; o > 0; o >>= 1) {
    hmin = fminf(hmin, __shfl_xor_sync(0xffffffffu, hmin, o)); hmax = fmaxf(hmax, __shfl_xor_sync(0xffffffffu, hmax, o));
    cfl = fminf(cfl, __shfl_xor_sync(0xffffffffu, cfl, o));
  }
  if ((threadIdx.x & 31u) == 0u && hmin <= hmax) {
    atomicMin(&ctl->hmin_enc, enc_f(hmin)); atomicMax(&ctl->hmax_enc, enc_f(hmax)); atomicMin(&ctl->cfl_enc, enc_f(cfl));
  }
}
__global__ void k_constrain_begin(StepCtl* ctl) { ctl->hmin_enc = 0xFFFFFFFFu; ctl->hmax_enc = 0u; ctl->cfl_enc = 0xFFFFFFFFu; }
// (the grid keeps its levels: their h_max still bound every particle's h from above, which is all the candidate scan needs)
__global__ void k_constrain_end(StepCtl* ctl, float max_dt, float cfl_factor) {
  ctl->hmin = dec_f(ctl->hmin_enc); ctl->hmax = dec_f(ctl->hmax_enc);
  ctl->dt = fminf(max_dt, __fmul_rn(cfl_factor, __fsqrt_rn(dec_f(ctl->cfl_enc))));
  ctl->list_used = 0; ctl->max_count = 0;
}

}  // namespace

int launch_constrain_neighborhood(asph_sim* sim) {
  const uint32_t n = sim->n;
  if (n == 0) return ASPH_OK;
  const uint32_t blocks = (n + kThreads - 1) / kThreads;
  cudaStream_t st = sim->stream;
  NbLists L;
  L.pool = sim->nbpool.p; L.slice_base = sim->slice_base.p; L.cnt = sim->cnt.p; L.cnt_ext = sim->cnt_ext.p; L.far_idx = sim->far_idx.p; L.far_cnt = sim->far_cnt.p;
  k_constrain<<<blocks, kThreads, 0, st>>>(n, L, sim->xyhm.p, sim->scratch_f.p, sim->ctl);
  LAUNCH_CHECK();
  k_constrain_begin<<<1, 1, 0, st>>>(sim->ctl);
  LAUNCH_CHECK();
  k_apply_h<<<blocks, kThreads, 0, st>>>(n, sim->scratch_f.p, sim->xv[sim->xv_cur].p, sim->xyhm.p, sim->hm.p, sim->ctl);
  LAUNCH_CHECK();
  k_constrain_end<<<1, 1, 0, st>>>(sim->ctl, sim->pp.max_dt, sim->pp.cfl_factor);
  LAUNCH_CHECK();
  return ASPH_OK;
}

int launch_neighbors(asph_sim* sim, float f_ext, float f_near, bool lists_only) {
  const uint32_t n = sim->n;
  sim->lists_valid = false;
  if (n == 0) { sim->lists_valid = true; return ASPH_OK; }
  const uint32_t blocks = (n + kThreads - 1) / kThreads;
  if (sim->nbpool.cap == 0) {
    // first guess: 24 (2h) / 48 (extended) 16-bit entries per particle
    const size_t per = (f_ext > f_near) ? 48 : 24;
    CUDA_TRY(sim->nbpool.ensure(size_t(sim->cap) * per + 8192));  // never shorter than the two chunks (512 uint16) a void slice points at
  }
  const uint32_t cap64 = uint32_t(std::min<size_t>(sim->nbpool.cap / 64, 0x7FFFFFF0u));
  CUDA_TRY(cudaMemsetAsync(sim->far_cnt.p, 0, (size_t(n + ASPH_PAIR_BLOCK - 1) / ASPH_PAIR_BLOCK) * sizeof(uint32_t), sim->stream));
  k_neighbors<<<blocks, kThreads, 0, sim->stream>>>(n, sim->xyhm.p, sim->ctl, sim->ctl, sim->cellstart.p, sim->pp, sim->lut.p, f_ext, f_near,
                                                    cap64, sim->nbpool.p, sim->slice_base.p, sim->cnt.p, sim->cnt_ext.p, sim->far_idx.p, sim->far_cnt.p, lists_only ? nullptr : sim->rho.p, sim->gB.p,
                                                    sim->pconst.p, sim->lam_sum.p, sim->lam_grad.p, sim->nrm.p,
                                                    sim->dist ? sim->refid[sim->cur].p : nullptr);
  LAUNCH_CHECK();
  if (lists_only) { sim->lists_valid = true; return ASPH_OK; }
  if (sim->dist) TRY(dist_halo(sim, sim->rho.p, 4));  // K12 / K17 read the neighbours' densities
  if (h_from_distribution(sim)) {
    NbLists L;
    L.pool = sim->nbpool.p; L.slice_base = sim->slice_base.p; L.cnt = sim->cnt.p; L.cnt_ext = sim->cnt_ext.p; L.far_idx = sim->far_idx.p; L.far_cnt = sim->far_cnt.p;
    k_estimate_h<<<blocks, kThreads, 0, sim->stream>>>(n, L, sim->xyhm.p, sim->lam_sum.p, sim->pp.h_mode, sim->pp.rest_density,
                                                       sim->hnext[sim->cur].p, sim->lamprev[sim->cur].p, sim->ctl);
    LAUNCH_CHECK();
  }
  if (op_w2020(sim)) {
    CUDA_TRY(sim->hv.ensure(sim->cap));
    NbLists L;
    L.pool = sim->nbpool.p; L.slice_base = sim->slice_base.p; L.cnt = sim->cnt.p; L.cnt_ext = sim->cnt_ext.p; L.far_idx = sim->far_idx.p; L.far_cnt = sim->far_cnt.p;
    k_aii_w2020<<<blocks, kThreads, 0, sim->stream>>>(n, L, sim->xyhm.p, sim->rho.p, sim->lam_grad.p, sim->pp.rest_density, sim->pconst.p,
                                                      sim->hv.p, sim->ctl, sim->dist ? sim->refid[sim->cur].p : nullptr);
    LAUNCH_CHECK();
  }
  sim->lists_valid = true;  // provisional: the caller checks ERRF_LIST_CAPACITY at its next synchronisation (neighbors_grow)
  return ASPH_OK;
}

// Called after a synchronisation showed ERRF_LIST_CAPACITY: grow the pool to what the step asked for and clear the
// flag and counters so the neighbour pass can be launched again (nothing irreversible has happened yet).
int neighbors_grow(asph_sim* sim) {
  const StepCtl& c = *sim->ctl_host;
  // multi-GPU: the flag is shared by all ranks, the pool only grows on the ranks whose own request did not fit
  const bool overflowed = size_t(c.list_used) * 64 > sim->nbpool.cap || !sim->dist;
  if (overflowed) {
    size_t want = (size_t(c.list_used) + size_t(c.list_used) / 4 + 128) * 64;
    if (want <= sim->nbpool.cap) want = sim->nbpool.cap * 2;
    if (want / 64 > 0x7FFFFFF0u) { sim->last_error = "neighbour list pool exceeds its addressable size"; return ASPH_ERR_CAPACITY; }
    CUDA_TRY(sim->nbpool.ensure(want));
  }
  StepCtl patch = c;
  patch.list_used = 0; patch.max_count = 0; patch.error_flags = c.error_flags & ~ERRF_LIST_CAPACITY;
  CUDA_TRY(cudaMemcpyAsync(sim->ctl, &patch, sizeof(StepCtl), cudaMemcpyHostToDevice, sim->stream));
  CUDA_TRY(cudaStreamSynchronize(sim->stream));
  return ASPH_OK;
}
