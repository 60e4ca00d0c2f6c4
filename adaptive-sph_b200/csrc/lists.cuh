// lists.cuh — the per-step neighbour lists: sliced, chunked ELL of 16-bit entries in three segments per column.
//
// A slice is 32 consecutive particles (one warp).  A particle's column is cut into chunks of 8 entries; chunk c of
// lane l is the 16 bytes at  pool[off + c*256 + l*8 .. +8)  (uint16 units), so ONE 16-byte vector load per lane brings
// 8 entries and the warp's loads of a chunk are one contiguous 512 B run.
// Particles are sorted by grid cell, so a neighbour index j is close to i itself.  The pair passes (solver.cu) run one
// block per ASPH_PAIR_BLOCK consecutive particles and stage the WINDOW of particles [b0 - ASPH_PAIR_HALO,
// b0 + ASPH_PAIR_BLOCK + ASPH_PAIR_HALO), b0 = i rounded down to a multiple of ASPH_PAIR_BLOCK, in shared memory.
// A column therefore holds
//   W  rows [0, cw):  the 2h neighbours that lie inside the window, stored as the BYTE OFFSET of their 16-byte window
//      slot, (j - (b0 - HALO)) * 16, always 16 bits wide — the gather address of the hot loops is one extract away;
//      padded with the particle's own slot up to a multiple of 8 (a zero-distance pair contributes nothing to any
//      gradient sum, so the loops run whole chunks without per-entry checks);
//      the few 2h neighbours outside the contiguous window (across a strip edge, or in another size level's grid) are
//      W rows as well: each tile of ASPH_PAIR_BLOCK particles owns a FAR TABLE of up to ASPH_PAIR_FAR particle indices
//      (far_idx[tile * ASPH_PAIR_FAR + s], far_cnt[tile] of them), which the pair passes stage right behind the
//      contiguous window — slot ASPH_PAIR_WIN + s — by indirect copies;
//   F  the cf 2h neighbours that found no room in the tile's far table (across a strip edge, or in another size level's grid), starting at the
//      chunk after W, padded with the particle itself up to a multiple of 4; the pair passes read these from global
//      memory;
//   E  the ce - cn remaining neighbours of the extended range used by the level set, right after F.
// F and E entries are uint16(j - b0 + 32768) in chunks of 8 — unless some F / E neighbour of the slice is further than
// +-32767 positions away; such a slice is "wide": its F / E entries are plain 32-bit indices in chunks of 4 (the W
// segment keeps its 16-bit form).  slice_base[s] = offset in units of 64 uint16 | wide << 31.
// W + F = N_2(i) (what NeighborhoodCache::filter_down keeps, neighborhood_search.rs:56-70), cn = cw + cf.
// cnt[i] = cw | cf << 12 | ghost << 31 (cw <= window size < 4096; ghost: the particle is a copy owned by another GPU);
// cnt_ext[i] = ce.  Counts are of real neighbours only.
// No per-pair coefficient is stored: every pass recomputes dW/dr / r from the gathered positions (PairShape below).
#pragma once
#include "sim.cuh"

#define ASPH_PAIR_BLOCK 256u  // threads per block of the pair passes == alignment of the index bias
#define ASPH_PAIR_HALO 192u   // window slots before / after the block's own particles
#define ASPH_PAIR_WIN (ASPH_PAIR_BLOCK + 2u * ASPH_PAIR_HALO)
#define ASPH_PAIR_FAR 96u     // far-table slots per tile (staged one per thread; sized so that four blocks of the sweep kernels fit an SM)
#define ASPH_PAIR_SLOTS (ASPH_PAIR_WIN + ASPH_PAIR_FAR)

struct NbLists {
  const uint16_t* __restrict__ pool;
  const uint32_t* __restrict__ slice_base;
  const uint32_t* __restrict__ cnt;
  const uint32_t* __restrict__ cnt_ext;
  const uint32_t* __restrict__ far_idx;
  const uint32_t* __restrict__ far_cnt;
};

#ifdef __CUDACC__
#define ASPH_KNORM 1.81891363533f  // 40 / (7 pi)
__host__ __device__ __forceinline__ uint32_t nb_block0(uint32_t i) { return i & ~(ASPH_PAIR_BLOCK - 1u); }
__host__ __device__ __forceinline__ uint32_t nb_bias(uint32_t i) { return nb_block0(i) - 32768u; }
__host__ __device__ __forceinline__ uint32_t nb_win0(uint32_t i) { return nb_block0(i) - ASPH_PAIR_HALO; }  // mod 2^32
__host__ __device__ __forceinline__ uint32_t nb_cw(uint32_t c) { return c & 0xfffu; }
__host__ __device__ __forceinline__ uint32_t nb_cf(uint32_t c) { return (c >> 12) & 0x7ffffu; }
__host__ __device__ __forceinline__ bool nb_ghost(uint32_t c) { return (c >> 31) != 0u; }
__host__ __device__ __forceinline__ uint32_t nb_cn(uint32_t c) { return nb_cw(c) + nb_cf(c); }
__host__ __device__ __forceinline__ uint32_t nb_pad8(uint32_t x) { return (x + 7u) & ~7u; }
__host__ __device__ __forceinline__ uint32_t nb_pad4(uint32_t x) { return (x + 3u) & ~3u; }
// 512-byte chunks a column occupies: W in 16-bit form, then F (padded to 4) and E in 16- or 32-bit form
__host__ __device__ __forceinline__ uint32_t nb_col_chunks(uint32_t cw, uint32_t cf, uint32_t ce, bool wide) {
  const uint32_t fe = nb_pad4(cf) + (ce - cw - cf);
  return nb_pad8(cw) / 8u + (wide ? (fe + 3u) / 4u : (fe + 7u) / 8u);
}
// position of W row r of lane `lane` (uint16 units from the slice start)
__host__ __device__ __forceinline__ uint32_t nb_pos_w(uint32_t lane, uint32_t r) { return (r >> 3) * 256u + lane * 8u + (r & 7u); }
// position of entry k of the F / E region of a column with cw window rows: uint16 units (narrow) / uint32 units (wide)
__host__ __device__ __forceinline__ uint32_t nb_pos_fe(uint32_t lane, uint32_t cw, bool wide, uint32_t k) {
  const uint32_t c0 = nb_pad8(cw) >> 3;
  return wide ? (c0 + (k >> 2)) * 128u + lane * 4u + (k & 3u) : (c0 + (k >> 3)) * 256u + lane * 8u + (k & 7u);
}
// k-th real neighbour (k < ce) of particle i (host and device; `slice` = start of the slice in the pool)
__host__ __device__ __forceinline__ uint32_t nb_get(const uint16_t* slice, const uint32_t* far_idx, bool wide, uint32_t i, uint32_t k,
                                                    uint32_t cw, uint32_t cf) {
  const uint32_t lane = i & 31u;
  if (k < cw) {
    const uint32_t slot = uint32_t(slice[nb_pos_w(lane, k)]) >> 4;
    return slot < ASPH_PAIR_WIN ? nb_win0(i) + slot : far_idx[(i / ASPH_PAIR_BLOCK) * ASPH_PAIR_FAR + (slot - ASPH_PAIR_WIN)];
  }
  const uint32_t kk = k < cw + cf ? k - cw : nb_pad4(cf) + (k - cw - cf);
  const uint32_t pos = nb_pos_fe(lane, cw, wide, kk);
  return wide ? reinterpret_cast<const uint32_t*>(slice)[pos] : nb_bias(i) + uint32_t(slice[pos]);
}
// filling a slice: W rows take a window byte offset, F / E entries an index
__device__ __forceinline__ void nb_store_w(uint16_t* slice, uint32_t lane, uint32_t r, uint32_t byte_off) {
  slice[nb_pos_w(lane, r)] = uint16_t(byte_off);
}
__device__ __forceinline__ void nb_store_fe(uint16_t* slice, bool wide, uint32_t lane, uint32_t cw, uint32_t k, uint32_t j, uint32_t bias) {
  const uint32_t pos = nb_pos_fe(lane, cw, wide, k);
  if (wide) reinterpret_cast<uint32_t*>(slice)[pos] = j;
  else slice[pos] = uint16_t(j - bias);
}

struct NbCol {
  const uint16_t* slice;  // start of the slice in the pool
  const uint32_t* far_idx;
  uint32_t i;
  uint32_t cw, cf, cn;    // window / far 2h neighbours, cn = cw + cf
  bool wide;
  __device__ __forceinline__ NbCol() : slice(nullptr), far_idx(nullptr), i(0), cw(0), cf(0), cn(0), wide(false) {}  // empty column
  __device__ __forceinline__ NbCol(const NbLists& L, uint32_t i_) : far_idx(L.far_idx), i(i_) {
    const uint32_t sb = __ldg(&L.slice_base[i >> 5]);
    const uint32_t c = __ldg(&L.cnt[i]);
    cw = nb_cw(c); cf = nb_cf(c); cn = cw + cf;
    wide = (sb >> 31) != 0u;
    slice = L.pool + size_t(sb & 0x7fffffffu) * 64u;
  }
  // k-th real neighbour, k < ce (cold paths: level set, resampling)
  __device__ __forceinline__ uint32_t get(uint32_t k) const { return nb_get(slice, far_idx, wide, i, k, cw, cf); }
  // raw W rows [k0, k0 + 8), k0 a multiple of 8 (8 x uint16 window byte offsets).  Streaming loads: list entries are
  // read once per pass and should not displace the gathered packs in L2.
  __device__ __forceinline__ uint4 raw8(uint32_t k0) const {
    return __ldcs(reinterpret_cast<const uint4*>(slice + (k0 >> 3) * 256u + (i & 31u) * 8u));
  }
  // F entries [k0, k0 + 4), k0 a multiple of 4, as particle indices
  __device__ __forceinline__ void far4(uint32_t k0, uint32_t (&j)[4]) const {
    const uint32_t pos = nb_pos_fe(i & 31u, cw, wide, k0);
    if (!wide) {
      const uint2 v = __ldcs(reinterpret_cast<const uint2*>(slice + pos));
      const uint32_t bias = nb_bias(i);
      j[0] = bias + (v.x & 0xffffu); j[1] = bias + (v.x >> 16); j[2] = bias + (v.y & 0xffffu); j[3] = bias + (v.y >> 16);
    } else {
      const uint4 v = __ldcs(reinterpret_cast<const uint4*>(reinterpret_cast<const uint32_t*>(slice) + pos));
      j[0] = v.x; j[1] = v.y; j[2] = v.z; j[3] = v.w;
    }
  }
};

// Single-instruction SFU approximations (2 ulp); flush-to-zero is harmless here: squared distances and smoothing
// lengths of a simulation are many orders of magnitude above the denormal range.
__device__ __forceinline__ float fast_rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// dW/dr / r for the cubic spline (sph_kernels.rs:61-71) from the squared distance; gradW_ij = pair_g * x_ij.
//   q < 1/2:  w'(q) / r = (18 q - 12) q / r = (18 q - 12) / (2h)          (q / r = 1 / (2h))
//   q < 1  :  w'(q) / r = -6 (1 - q)^2 / r
// times norm / (2h) = 10 / (7 pi h^2) / (2h) = (40 / (7 pi)) / (2h)^3.  Every list entry is a neighbour (q < 1 up to
// rounding), so there is no q >= 1 branch.  Zero for q <= 1e-5 (the reference's guard); for r = 0 (the particle
// itself, padding rows) rsqrt gives inf, q = 0 * inf = NaN, and the comparison is false.
// pair_g_shape: w'(q) / r without the normalisation (the uniform-h passes apply it once per particle), as two FMAs
// per branch with i = 1/(2h):   q < 1/2: 18 i q - 12 i        q >= 1/2: -6/r + 12 i - 6 i q   (= -6 (1-q)^2 / r)
struct PairShape {
  float i, a, b, c, d;  // 1/(2h), 18 i, -12 i, -6 i, 12 i
  __device__ __forceinline__ explicit PairShape(float inv2h) : i(inv2h), a(18.f * inv2h), b(-12.f * inv2h), c(-6.f * inv2h), d(12.f * inv2h) {
    asm volatile("" : "+f"(a), "+f"(b), "+f"(c), "+f"(d));  // keep the four products in registers across the pair loop
  }
  __device__ __forceinline__ float operator()(float d2) const {
    const float inv_r = fast_rsqrt(d2);
    const float q = (d2 * inv_r) * i;
    const float g1 = fmaf(q, a, b);
    const float g2 = fmaf(inv_r, -6.f, fmaf(q, c, d));
    return q > 1.0e-5f ? (q < 0.5f ? g1 : g2) : 0.f;
  }
};
__device__ __forceinline__ float pair_g(float d2, float hij) {
  const float inv2h = fast_rcp(2.f * hij);
  return (ASPH_KNORM * inv2h * inv2h * inv2h) * PairShape(inv2h)(d2);
}
// The same value from h_i + h_j = 2 h_ij with the four per-pair constants of PairShape folded away (the experimental 4-row
// sweep kernels, solver.cu R4): i = 1 / (h_i + h_j);  q < 1/2: i (18 q - 12);  q >= 1/2: i (12 - 6 q) - 6 / r.
__device__ __forceinline__ float pair_g_sum(float d2, float hsum) {
  const float i = fast_rcp(hsum);
  const float inv_r = fast_rsqrt(d2);
  const float q = (d2 * inv_r) * i;
  const float s = q < 0.5f ? fmaf(q, 18.f, -12.f) : fmaf(q, -6.f, 12.f);
  float g = s * i;
  if (!(q < 0.5f)) g = fmaf(inv_r, -6.f, g);
  g = q > 1.0e-5f ? g : 0.f;
  return ((ASPH_KNORM * i) * (i * i)) * g;
}
// 16-byte / 8-byte / 4-byte shared-memory loads by 32-bit shared address (one LEA + LDS per gather)
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
// asynchronous global -> shared copies (LDGSTS): no register staging, completion awaited by cp_async_wait_all
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g) { asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(saddr), "l"(g) : "memory"); }
__device__ __forceinline__ void cp_async8(uint32_t saddr, const void* g) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(saddr), "l"(g) : "memory"); }
__device__ __forceinline__ void cp_async4(uint32_t saddr, const void* g) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(saddr), "l"(g) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ float lds_f1(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
// W_ij and dW/dr / r together (neighbour build: density and a_ii)
__device__ __forceinline__ void pair_wg(float d2, float hij, float& w, float& g) {
  const float inv_r = rsqrtf(d2);
  const float inv2h = __frcp_rn(2.f * hij);
  const float r = d2 > 0.f ? d2 * inv_r : 0.f;
  const float q = r * inv2h;
  const float v = fmaxf(1.f - q, 0.f);
  const float nf = ASPH_KNORM * inv2h * inv2h;  // 10 / (7 pi h^2)
  w = nf * (q < 0.5f ? 6.f * (q * q * q - q * q) + 1.f : 2.f * (v * v * v));
  const float dw = q < 0.5f ? (18.f * q - 12.f) * q : -6.f * v * v;
  g = q > 1.0e-5f ? nf * inv2h * dw * inv_r : 0.f;
}
#endif
