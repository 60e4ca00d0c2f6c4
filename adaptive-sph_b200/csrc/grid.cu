// grid.cu — per-step preparation: h from mass, CFL minimum, bounding box, multi-resolution cell grid,
// deterministic counting sort of the particles by (size level, cell) and the physical reorder of the SoA.
//
// Replaces (semantically) the R*-tree bulk load of neighborhood_search.rs:113-119 and h_next_from_mass
// (simulation.rs:1865-1871), plus the CFL reduce of simulation.rs:2182-2191.
//
// Multi-resolution grid: level L holds particles with h in [hmin*2^L, hmin*2^(L+1)); its cells have edge
// f_search * hmax_L * SLACK with hmax_L = min(hmin*2^(L+1), hmax).  A pair (i in level a, j in level b) has
// support f*(h_i+h_j)/2 <= f*(h_i + hmax_b)/2, so i finds all its level-b neighbours in the cells of grid b that
// overlap the box of that radius around x_i (neighbors.cu).  With uniform h there is one level whose cell is
// exactly the support radius.
#include "lists.cuh"

namespace {

constexpr int kThreads = 256;

__global__ void k_ctl_reset(StepCtl* ctl) {
  ctl->hmin_enc = 0xFFFFFFFFu; ctl->minx_enc = 0xFFFFFFFFu; ctl->miny_enc = 0xFFFFFFFFu; ctl->cfl_enc = 0xFFFFFFFFu;
  ctl->hmax_enc = 0u; ctl->maxx_enc = 0u; ctl->maxy_enc = 0u;
  ctl->list_used = 0; ctl->max_count = 0;
  ctl->error_flags = 0;
}

__device__ __forceinline__ float warp_min(float v) {
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// K1 + K8 + bounding box.  h is bit-exact (neighbour predicate); the CFL term follows simulation.rs:2184-2186
// operation by operation ((2h)^2 / (v.v + 0.01)), and min is order independent, so dt is bit-exact too.
__global__ void k_prepare(uint32_t n, const float2* __restrict__ pos, const float2* __restrict__ vel,
                          const float* __restrict__ mass, float rho0, const float* __restrict__ h_in, float* __restrict__ h_out,
                          StepCtl* ctl) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const float inf = __int_as_float(0x7f800000);
  float hmin = inf, hmax = -inf, minx = inf, miny = inf, maxx = -inf, maxy = -inf, cfl = inf;
  if (i < n) {
    float2 x = pos[i], v = vel[i];
    // FromMass (simulation.rs:1865-1871), or the length the previous step left in h2_next (simulation.rs:2005-2014)
    float h = h_in ? h_in[i] : h_from_mass(mass[i], rho0);
    h_out[i] = h;
    hmin = hmax = h;
    minx = maxx = x.x; miny = maxy = x.y;
    float sr = __fmul_rn(h, 2.f);
    cfl = __fdiv_rn(__fmul_rn(sr, sr), __fadd_rn(dist_sq_exact(v.x, v.y), 0.01f));
  }
  hmin = warp_min(hmin); hmax = warp_max(hmax); minx = warp_min(minx); miny = warp_min(miny);
  maxx = warp_max(maxx); maxy = warp_max(maxy); cfl = warp_min(cfl);
  __shared__ float s[7][kThreads / 32];
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { s[0][w] = hmin; s[1][w] = hmax; s[2][w] = minx; s[3][w] = miny; s[4][w] = maxx; s[5][w] = maxy; s[6][w] = cfl; }
  __syncthreads();
  if (w == 0) {
    const int nw = kThreads / 32;
    hmin = l < nw ? s[0][l] : inf; hmax = l < nw ? s[1][l] : -inf; minx = l < nw ? s[2][l] : inf;
    miny = l < nw ? s[3][l] : inf; maxx = l < nw ? s[4][l] : -inf; maxy = l < nw ? s[5][l] : -inf;
    cfl = l < nw ? s[6][l] : inf;
    hmin = warp_min(hmin); hmax = warp_max(hmax); minx = warp_min(minx); miny = warp_min(miny);
    maxx = warp_max(maxx); maxy = warp_max(maxy); cfl = warp_min(cfl);
    if (l == 0) {
      atomicMin(&ctl->hmin_enc, enc_f(hmin)); atomicMax(&ctl->hmax_enc, enc_f(hmax));
      atomicMin(&ctl->minx_enc, enc_f(minx)); atomicMin(&ctl->miny_enc, enc_f(miny));
      atomicMax(&ctl->maxx_enc, enc_f(maxx)); atomicMax(&ctl->maxy_enc, enc_f(maxy));
      atomicMin(&ctl->cfl_enc, enc_f(cfl));
    }
  }
}

// Single thread: decide levels, cell sizes and grid dimensions; dt = min(max_dt, cfl * sqrt(min)) (simulation.rs:2188-2191).
__global__ void k_make_levels(StepCtl* ctl, float f_search, uint32_t cells_budget, float max_dt, float cfl_factor, float cell_scale) {
  float hmin = dec_f(ctl->hmin_enc), hmax = dec_f(ctl->hmax_enc);
  float minx = dec_f(ctl->minx_enc), miny = dec_f(ctl->miny_enc), maxx = dec_f(ctl->maxx_enc), maxy = dec_f(ctl->maxy_enc);
  ctl->hmin = hmin; ctl->hmax = hmax; ctl->origin_x = minx; ctl->origin_y = miny;
  ctl->dt = fminf(max_dt, __fmul_rn(cfl_factor, __fsqrt_rn(dec_f(ctl->cfl_enc))));
  int nl = 1;
  {
    float bound = hmin * 2.f;
    while (hmax >= bound && nl < ASPH_MAX_LEVELS) { nl++; bound *= 2.f; }
  }
  ctl->nlevels = nl;
  // A grid that does not fit the cell budget is retried with larger cells (a few stray particles far from the bulk);
  // beyond ~11x the natural cell size the candidate scans would degenerate to O(N^2), and positions that far apart mean
  // the simulation has exploded anyway: flag it and leave a 1 x 1 grid that keeps every later kernel in bounds
  // (k_sort_cells and k_neighbors return at once when the flag is set).
  // Cell edge in units of the level's largest search radius.  A level spans a factor 2 in h and the search radius of a pair
  // is f (h_i + h_max) / 2, so with cells of the full radius the finest particles of a level — most of them — test 8 x more
  // candidates than they have neighbours.  With several levels, or with the extended range of the level set, cells of 0.35
  // radii (boxes of up to 7 x 7 cells) measured best on the 16 M adaptive dam break: candidate scan -20 %, cell sort -40 %,
  // and the pair passes -20 % (finer cells sort the particles into shorter strips: more neighbours inside the window).
  // Uniform h without level set keeps cell = support (3 x 3 cells; smaller cells measured no gain there).
  float scale = cell_scale > 0.f ? cell_scale : ((nl > 1 || f_search > 2.05f) ? 0.35f : 1.f);
  for (int attempt = 0; attempt < 9; attempt++) {
    unsigned long long total = 0;
    float upper = hmin * 2.f;
    bool ok = true;
    for (int L = 0; L < nl; L++) {
      float hm = (L == nl - 1) ? hmax : fminf(upper, hmax);
      float cell = f_search * hm * ASPH_SLACK * scale;
      GridLevel g;
      g.hmax = hm; g.cell = cell; g.inv_cell = 1.f / cell;
      const float fx = floorf((maxx - minx) * g.inv_cell), fy = floorf((maxy - miny) * g.inv_cell);
      if (!(fx < 4.0e6f) || !(fy < 4.0e6f)) { ok = false; break; }  // also catches inf / NaN extents
      g.nx = int(fx) + 1;
      g.ny = int(fy) + 1;
      // a cell of edge f * h holds about 1.15 f^2 particles of smoothing length h at rest density (h = 1.9 sqrt(A / pi));
      // one strip row plus 1.5 cells, with 10 % slack, should fit the halo of the pair passes' window (lists.cuh)
      const float per_cell = 1.15f * f_search * f_search * scale * scale;
      g.strip_log2 = 5;
      while (g.strip_log2 > 2 && (float(1 << g.strip_log2) + 1.5f) * per_cell * 1.1f > float(ASPH_PAIR_HALO)) g.strip_log2--;
      g.base = uint32_t(total);
      ctl->lv[L] = g;
      total += level_cells(g.nx, g.ny, g.strip_log2);
      upper *= 2.f;
    }
    if (ok && total <= cells_budget) { ctl->total_cells = uint32_t(total); return; }
    scale *= 1.5f;
  }
  GridLevel g;
  g.hmax = hmax; g.cell = 1.f; g.inv_cell = 0.f; g.nx = 1; g.ny = 1; g.strip_log2 = 2; g.base = 0;
  ctl->lv[0] = g;
  ctl->nlevels = 1;
  ctl->total_cells = uint32_t(level_cells(1, 1, 2));
  atomicOr(&ctl->error_flags, ERRF_CELL_BUDGET);
}

__global__ void k_zero_cells(const StepCtl* __restrict__ ctl, uint32_t* __restrict__ cellcount) {
  uint32_t total = ctl->total_cells + 1;
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < total; c += gridDim.x * blockDim.x) cellcount[c] = 0;
}

__device__ __forceinline__ int level_of(float h, float hmin, int nl) {
  int L = 0;
  float bound = hmin * 2.f;
  while (L < nl - 1 && h >= bound) { L++; bound *= 2.f; }
  return L;
}

__global__ void k_bin(uint32_t n, const float2* __restrict__ pos, const float* __restrict__ h, const StepCtl* __restrict__ ctl,
                      uint32_t* __restrict__ key, uint32_t* __restrict__ cellcount) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int L = level_of(h[i], ctl->hmin, ctl->nlevels);
  GridLevel g = ctl->lv[L];
  float2 x = pos[i];
  int cx = min(g.nx - 1, max(0, int(floorf((x.x - ctl->origin_x) * g.inv_cell))));
  int cy = min(g.ny - 1, max(0, int(floorf((x.y - ctl->origin_y) * g.inv_cell))));
  uint32_t k = cell_index(g, cx, cy);
  key[i] = k;
  atomicAdd(&cellcount[k], 1u);
}

// ---- exclusive scan over a device-side length (3 kernels; tiles of 2048) ---------------------------------
constexpr int kScanTile = 2048;  // 256 threads * 8

__global__ void k_scan_tiles(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t* __restrict__ sums,
                             const uint32_t* __restrict__ n_dev, uint32_t n_add) {
  const uint32_t n = *n_dev + n_add;
  const uint32_t tile0 = blockIdx.x * kScanTile;
  if (tile0 >= n) return;
  uint32_t v[8], local = 0;
  const uint32_t base = tile0 + threadIdx.x * 8;
#pragma unroll
  for (int k = 0; k < 8; k++) { v[k] = (base + k < n) ? in[base + k] : 0u; local += v[k]; }
  // block exclusive scan of `local`
  __shared__ uint32_t ws[kThreads / 32];
  uint32_t incl = local;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) ws[w] = incl;
  __syncthreads();
  if (w == 0) {
    uint32_t x = lane < kThreads / 32 ? ws[lane] : 0u, xi = x;
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, xi, o); if (lane >= o) xi += t; }
    if (lane < kThreads / 32) ws[lane] = xi - x;
    if (lane == kThreads / 32 - 1) sums[blockIdx.x] = xi;
  }
  __syncthreads();
  uint32_t run = ws[w] + incl - local;
#pragma unroll
  for (int k = 0; k < 8; k++) { if (base + k < n) out[base + k] = run; run += v[k]; }
}
__global__ void k_scan_sums(uint32_t* __restrict__ sums, const uint32_t* __restrict__ n_dev, uint32_t n_add) {
  const uint32_t n = *n_dev + n_add;
  const uint32_t nt = (n + kScanTile - 1) / kScanTile;
  __shared__ uint32_t ws[32];
  __shared__ uint32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (uint32_t t0 = 0; t0 < nt; t0 += blockDim.x) {
    uint32_t t = t0 + threadIdx.x;
    uint32_t x = t < nt ? sums[t] : 0u, incl = x;
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    if (lane == 31) ws[w] = incl;
    __syncthreads();
    if (w == 0) {
      uint32_t a = lane < (blockDim.x >> 5) ? ws[lane] : 0u, ai = a;
      for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, ai, o); if (lane >= o) ai += y; }
      ws[lane] = ai - a;
    }
    __syncthreads();
    uint32_t carry = carry_s;
    uint32_t excl = carry + ws[w] + incl - x;
    if (t < nt) sums[t] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = excl + x;
    __syncthreads();
  }
}
__global__ void k_scan_add(uint32_t* __restrict__ out, const uint32_t* __restrict__ sums, const uint32_t* __restrict__ n_dev,
                           uint32_t n_add) {
  const uint32_t n = *n_dev + n_add;
  const uint32_t tile0 = blockIdx.x * kScanTile;
  if (tile0 >= n) return;
  const uint32_t add = sums[blockIdx.x];
  const uint32_t base = tile0 + threadIdx.x * 8;
#pragma unroll
  for (int k = 0; k < 8; k++) if (base + k < n) out[base + k] += add;
}

// slot = cellstart + (count-- - 1): leaves cellcount all zero again
__global__ void k_scatter(uint32_t n, const uint32_t* __restrict__ key, const uint32_t* __restrict__ cellstart,
                          uint32_t* __restrict__ cellcount, uint32_t* __restrict__ order) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t k = key[i];
  uint32_t c = atomicSub(&cellcount[k], 1u);
  order[cellstart[k] + c - 1u] = i;
}
// make the counting sort stable (deterministic): ascending source index inside each cell
// gid != nullptr (multi-GPU): ascending global particle index instead — owned particles and ghosts arrive in arbitrary
// order, the global index makes the sorted order (and with it every fp32 sum) reproducible
__global__ void k_sort_cells(const StepCtl* __restrict__ ctl, const uint32_t* __restrict__ cellstart, uint32_t* __restrict__ order,
                             const uint32_t* __restrict__ gid) {
  if (ctl->error_flags & ERRF_CELL_BUDGET) return;  // degenerate 1 x 1 grid (k_make_levels): the step is failing anyway
  const uint32_t total = ctl->total_cells;
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < total; c += gridDim.x * blockDim.x) {
    uint32_t s = cellstart[c], e = cellstart[c + 1];
    if (e - s <= 1u) continue;
    if (e - s <= 24u) {  // the usual cell: all its entries (and keys) requested at once, sorted in local memory, written back
      uint32_t v[24], k[24];
      const uint32_t m = e - s;
#pragma unroll
      for (uint32_t a = 0; a < 24u; a++) v[a] = a < m ? order[s + a] : 0u;
#pragma unroll
      for (uint32_t a = 0; a < 24u; a++) k[a] = a < m ? (gid ? (gid[v[a]] & ~ASPH_GHOST_BIT) : v[a]) : 0xffffffffu;
      for (uint32_t a = 1; a < m; a++) {
        const uint32_t kv = k[a], vv = v[a];
        uint32_t b = a;
        while (b > 0u && k[b - 1] > kv) { k[b] = k[b - 1]; v[b] = v[b - 1]; b--; }
        k[b] = kv; v[b] = vv;
      }
      for (uint32_t a = 0; a < m; a++) order[s + a] = v[a];
      continue;
    }
    for (uint32_t a = s + 1; a < e; a++) {
      uint32_t v = order[a];
      uint32_t b = a;
      if (gid) {
        const uint32_t kv = gid[v] & ~ASPH_GHOST_BIT;
        while (b > s && (gid[order[b - 1]] & ~ASPH_GHOST_BIT) > kv) { order[b] = order[b - 1]; b--; }
      } else {
        while (b > s && order[b - 1] > v) { order[b] = order[b - 1]; b--; }
      }
      order[b] = v;
    }
  }
}

__global__ void k_reorder(uint32_t n, const uint32_t* __restrict__ order, const float2* __restrict__ pos, const float2* __restrict__ vel,
                          const float* __restrict__ mass, const uint32_t* __restrict__ refid, const float* __restrict__ level,
                          const float* __restrict__ h, float2* __restrict__ pos_o, float2* __restrict__ vel_o,
                          float* __restrict__ mass_o, uint32_t* __restrict__ refid_o, float* __restrict__ level_o,
                          float4* __restrict__ xyhm, float4* __restrict__ xv, float2* __restrict__ hm) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  uint32_t src = order[s];
  float2 x = pos[src], v = vel[src];
  float m = mass[src];
  pos_o[s] = x; vel_o[s] = v; mass_o[s] = m; refid_o[s] = refid[src]; level_o[s] = level[src];
  xyhm[s] = make_float4(x.x, x.y, h[src], m);
  xv[s] = make_float4(x.x, x.y, v.x, v.y);
  hm[s] = make_float2(h[src], m);
}

// support_length_estimation != FromMass: the per-particle state besides x, v, m (sim.cuh)
__global__ void k_hdist_init(uint32_t n, const float* __restrict__ mass, float* __restrict__ hnext, float* __restrict__ lamprev) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  hnext[i] = h_from_mass(mass[i], 1.0f);  // INIT_REST_DENSITY, simulation.rs:505-520
  lamprev[i] = 0.f;                       // the boundary handler starts with empty lambda lists
}
__global__ void k_reorder_extra(uint32_t n, const uint32_t* __restrict__ order, const float* __restrict__ hnext, const float* __restrict__ lamprev,
                                float* __restrict__ hnext_o, float* __restrict__ lamprev_o) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const uint32_t src = order[s];
  hnext_o[s] = hnext[src]; lamprev_o[s] = lamprev[src];
}

__global__ void k_reorder_cls(uint32_t n, const uint32_t* __restrict__ order, const uint8_t* __restrict__ cls, uint8_t* __restrict__ cls_o) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) cls_o[s] = cls[order[s]];
}

}  // namespace

int sync_ctl(asph_sim* sim) {
  if (sim->dist) TRY(dist_reduce_flags(sim, false));  // every rank sees the same error flags => the same control flow
  CUDA_TRY(cudaMemcpyAsync(sim->ctl_host, sim->ctl, sizeof(StepCtl), cudaMemcpyDeviceToHost, sim->stream));
  CUDA_TRY(cudaStreamSynchronize(sim->stream));
  sim->ctl_seen = true;
  return ASPH_OK;
}

int launch_exclusive_scan(asph_sim* sim, const uint32_t* in, uint32_t* data, const uint32_t* n_dev, uint32_t n_add,
                          uint32_t n_max) {
  const uint32_t tiles = (n_max + kScanTile - 1) / kScanTile;
  CUDA_TRY(sim->scan_sums.ensure(tiles + 1));
  k_scan_tiles<<<tiles, kThreads, 0, sim->stream>>>(in, data, sim->scan_sums.p, n_dev, n_add);
  LAUNCH_CHECK();
  k_scan_sums<<<1, 1024, 0, sim->stream>>>(sim->scan_sums.p, n_dev, n_add);
  LAUNCH_CHECK();
  k_scan_add<<<tiles, kThreads, 0, sim->stream>>>(data, sim->scan_sums.p, n_dev, n_add);
  LAUNCH_CHECK();
  return ASPH_OK;
}

int ensure_capacity(asph_sim* sim, uint32_t want) {
  if (want <= sim->cap && sim->cap > 0) return ASPH_OK;
  // growing: persistent arrays must be preserved
  uint32_t newcap = want;
  const uint32_t old_n = sim->n;
  for (int b = 0; b < 2; b++) {
    if (b == sim->cur && sim->cap > 0 && old_n > 0) {
      DevBuf<float2> p2, v2; DevBuf<float> m2, l2; DevBuf<uint32_t> r2;
      CUDA_TRY(p2.ensure(newcap)); CUDA_TRY(v2.ensure(newcap)); CUDA_TRY(m2.ensure(newcap)); CUDA_TRY(l2.ensure(newcap));
      CUDA_TRY(r2.ensure(newcap));
      CUDA_TRY(cudaMemcpyAsync(p2.p, sim->pos[b].p, old_n * sizeof(float2), cudaMemcpyDeviceToDevice, sim->stream));
      CUDA_TRY(cudaMemcpyAsync(v2.p, sim->vel[b].p, old_n * sizeof(float2), cudaMemcpyDeviceToDevice, sim->stream));
      CUDA_TRY(cudaMemcpyAsync(m2.p, sim->mass[b].p, old_n * sizeof(float), cudaMemcpyDeviceToDevice, sim->stream));
      CUDA_TRY(cudaMemcpyAsync(l2.p, sim->level[b].p, old_n * sizeof(float), cudaMemcpyDeviceToDevice, sim->stream));
      CUDA_TRY(cudaMemcpyAsync(r2.p, sim->refid[b].p, old_n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, sim->stream));
      CUDA_TRY(cudaStreamSynchronize(sim->stream));
      sim->pos[b].release(); sim->vel[b].release(); sim->mass[b].release(); sim->level[b].release(); sim->refid[b].release();
      sim->pos[b] = p2; sim->vel[b] = v2; sim->mass[b] = m2; sim->level[b] = l2; sim->refid[b] = r2;
    } else {
      CUDA_TRY(sim->pos[b].ensure(newcap)); CUDA_TRY(sim->vel[b].ensure(newcap)); CUDA_TRY(sim->mass[b].ensure(newcap));
      CUDA_TRY(sim->level[b].ensure(newcap)); CUDA_TRY(sim->refid[b].ensure(newcap));
    }
  }
  if (sim->hdist_valid) {  // support_length_estimation != FromMass: two more persistent arrays
    for (int b = 0; b < 2; b++) {
      if (b == sim->cur && old_n > 0) {
        DevBuf<float> h2, l2;
        CUDA_TRY(h2.ensure(newcap)); CUDA_TRY(l2.ensure(newcap));
        CUDA_TRY(cudaMemcpyAsync(h2.p, sim->hnext[b].p, old_n * sizeof(float), cudaMemcpyDeviceToDevice, sim->stream));
        CUDA_TRY(cudaMemcpyAsync(l2.p, sim->lamprev[b].p, old_n * sizeof(float), cudaMemcpyDeviceToDevice, sim->stream));
        CUDA_TRY(cudaStreamSynchronize(sim->stream));
        sim->hnext[b].release(); sim->lamprev[b].release();
        sim->hnext[b] = h2; sim->lamprev[b] = l2;
      } else {
        CUDA_TRY(sim->hnext[b].ensure(newcap)); CUDA_TRY(sim->lamprev[b].ensure(newcap));
      }
    }
  }
  if (sim->cls_valid) {
    for (int b = 0; b < 2; b++) {
      if (b == sim->cur && old_n > 0) {
        DevBuf<uint8_t> c2;
        CUDA_TRY(c2.ensure(newcap));
        CUDA_TRY(cudaMemcpyAsync(c2.p, sim->cls[b].p, old_n, cudaMemcpyDeviceToDevice, sim->stream));
        CUDA_TRY(cudaStreamSynchronize(sim->stream));
        sim->cls[b].release();
        sim->cls[b] = c2;
      } else {
        CUDA_TRY(sim->cls[b].ensure(newcap));
      }
    }
  }
  CUDA_TRY(sim->xyhm.ensure(newcap)); CUDA_TRY(sim->xv[0].ensure(newcap)); CUDA_TRY(sim->xv[1].ensure(newcap));
  CUDA_TRY(sim->packP[0].ensure(newcap)); CUDA_TRY(sim->packP[1].ensure(newcap)); CUDA_TRY(sim->packA.ensure(newcap));
  CUDA_TRY(sim->pconst.ensure(newcap));
  CUDA_TRY(sim->h_tmp.ensure(newcap)); CUDA_TRY(sim->rho.ensure(newcap)); CUDA_TRY(sim->lam_sum.ensure(newcap));
  CUDA_TRY(sim->nrm.ensure(newcap)); CUDA_TRY(sim->gB.ensure(newcap)); CUDA_TRY(sim->lam_grad.ensure(newcap));
  CUDA_TRY(sim->key.ensure(newcap)); CUDA_TRY(sim->order.ensure(newcap));
  CUDA_TRY(sim->cnt.ensure(newcap)); CUDA_TRY(sim->cnt_ext.ensure(newcap));
  {
    const size_t tiles = (size_t(newcap) + ASPH_PAIR_BLOCK - 1) / ASPH_PAIR_BLOCK + 1;
    CUDA_TRY(sim->far_cnt.ensure(tiles)); CUDA_TRY(sim->far_idx.ensure(tiles * ASPH_PAIR_FAR));
  }
  const uint32_t nslices = (newcap + 31) / 32 + 1;
  CUDA_TRY(sim->slice_base.ensure(nslices)); CUDA_TRY(sim->hm.ensure(newcap));
  CUDA_TRY(sim->size_class.ensure(newcap)); CUDA_TRY(sim->flags.ensure(newcap));
  CUDA_TRY(sim->merge_partner.ensure(newcap)); CUDA_TRY(sim->merge_counter.ensure(newcap));
  CUDA_TRY(sim->front[0].ensure(newcap)); CUDA_TRY(sim->front[1].ensure(newcap)); CUDA_TRY(sim->cand.ensure(newcap));
  CUDA_TRY(sim->work[0].ensure(newcap)); CUDA_TRY(sim->work[1].ensure(newcap));
  for (int k = 0; k < 4; k++) CUDA_TRY(sim->scratch_u[k].ensure(size_t(newcap) + 1));
  CUDA_TRY(sim->stamp.ensure(newcap));
  CUDA_TRY(sim->scratch_f.ensure(size_t(newcap) * 2));
  sim->cells_budget = 2u * newcap + 65536u;
  CUDA_TRY(sim->cellcount.ensure(size_t(sim->cells_budget) + 2));
  CUDA_TRY(sim->cellstart.ensure(size_t(sim->cells_budget) + 2));
  CUDA_TRY(cudaMemsetAsync(sim->cellcount.p, 0, (size_t(sim->cells_budget) + 2) * sizeof(uint32_t), sim->stream));
  sim->cap = newcap;
  sim->lists_valid = false;
  return ASPH_OK;
}

// h, bbox, CFL/dt, levels, counting sort, reorder.  Afterwards the persistent arrays live in buffer `cur`
// in sorted order and xyhm / xv[xv_cur] hold the step's snapshot.
int launch_sort_and_grid(asph_sim* sim, float f_search) {
  const uint32_t n = sim->n;
  cudaStream_t st = sim->stream;
  k_ctl_reset<<<1, 1, 0, st>>>(sim->ctl);
  LAUNCH_CHECK();
  if (n == 0) return ASPH_OK;
  const uint32_t blocks = (n + kThreads - 1) / kThreads;
  const int c = sim->cur;
  const bool hdist = h_from_distribution(sim);
  if (hdist && !sim->hdist_valid) {  // first step in this mode (or after asph_set_state): h2_next = h of the masses
    for (int b = 0; b < 2; b++) { CUDA_TRY(sim->hnext[b].ensure(sim->cap)); CUDA_TRY(sim->lamprev[b].ensure(sim->cap)); }
    k_hdist_init<<<blocks, kThreads, 0, st>>>(n, sim->mass[c].p, sim->hnext[c].p, sim->lamprev[c].p);
    LAUNCH_CHECK();
    sim->hdist_valid = true;
  }
  if (!hdist) sim->hdist_valid = false;  // a FromMass step does not maintain h2_next
  k_prepare<<<blocks, kThreads, 0, st>>>(n, sim->pos[c].p, sim->vel[c].p, sim->mass[c].p, sim->pp.rest_density,
                                         hdist ? sim->hnext[c].p : nullptr, sim->h_tmp.p, sim->ctl);
  LAUNCH_CHECK();
  if (sim->dist) TRY(dist_allreduce_cfl(sim));  // dt is global: min over all ranks
  k_make_levels<<<1, 1, 0, st>>>(sim->ctl, f_search, sim->cells_budget, sim->pp.max_dt, sim->pp.cfl_factor, sim->cell_scale);
  LAUNCH_CHECK();
  const int wide = sim->sm_count * 8;
  k_zero_cells<<<wide, kThreads, 0, st>>>(sim->ctl, sim->cellcount.p);
  LAUNCH_CHECK();
  k_bin<<<blocks, kThreads, 0, st>>>(n, sim->pos[c].p, sim->h_tmp.p, sim->ctl, sim->key.p, sim->cellcount.p);
  LAUNCH_CHECK();
  // cellstart = exclusive scan of cellcount over total_cells + 1 entries
  TRY(launch_exclusive_scan(sim, sim->cellcount.p, sim->cellstart.p, &sim->ctl->total_cells, 1, sim->cells_budget + 1));
  k_scatter<<<blocks, kThreads, 0, st>>>(n, sim->key.p, sim->cellstart.p, sim->cellcount.p, sim->order.p);
  LAUNCH_CHECK();
  k_sort_cells<<<wide, kThreads, 0, st>>>(sim->ctl, sim->cellstart.p, sim->order.p, sim->dist ? sim->refid[c].p : nullptr);
  LAUNCH_CHECK();
  sim->xv_cur = 0;
  k_reorder<<<blocks, kThreads, 0, st>>>(n, sim->order.p, sim->pos[c].p, sim->vel[c].p, sim->mass[c].p, sim->refid[c].p,
                                         sim->level[c].p, sim->h_tmp.p, sim->pos[1 - c].p, sim->vel[1 - c].p, sim->mass[1 - c].p,
                                         sim->refid[1 - c].p, sim->level[1 - c].p, sim->xyhm.p, sim->xv[0].p, sim->hm.p);
  LAUNCH_CHECK();
  if (hdist) {
    k_reorder_extra<<<blocks, kThreads, 0, st>>>(n, sim->order.p, sim->hnext[c].p, sim->lamprev[c].p, sim->hnext[1 - c].p, sim->lamprev[1 - c].p);
    LAUNCH_CHECK();
  }
  if (sim->cls_valid) {  // IISPH2: the size classes of the last resampling phase follow the particles
    k_reorder_cls<<<blocks, kThreads, 0, st>>>(n, sim->order.p, sim->cls[c].p, sim->cls[1 - c].p);
    LAUNCH_CHECK();
  }
  sim->cur = 1 - c;
  if (sim->dist) TRY(dist_after_sort(sim));
  return ASPH_OK;
}
