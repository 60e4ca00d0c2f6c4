// lists.cuh — the per-step neighbour lists: sliced, chunked ELL of 16-bit relative indices.
//
// A slice is 32 consecutive particles (one warp).  A particle's column is cut into chunks of 8 entries; chunk c of
// lane l is the 16 bytes at  pool[off + c*256 + l*8 .. +8)  (uint16 units), so ONE 16-byte vector load per lane brings
// 8 neighbour indices and the warp's loads of a chunk are one contiguous 512 B run.  A whole 2h column (about 14
// entries) is therefore two load instructions, after which all gathers of the column can be in flight at once.
// Particles are sorted by grid cell, so a neighbour index j is close to i itself: it is stored as
// uint16(j - b0 + 32768) with b0 = i rounded down to a multiple of ASPH_PAIR_BLOCK — the first particle of the thread
// block that consumes the column in the pair passes, so that a stored index minus a compile-time constant IS the
// position in that block's shared-memory window (solver.cu).  A slice in which some neighbour is further than +-32767 positions away (neighbours in
// another size level's grid) is stored "wide": plain 32-bit indices, chunks of 4.
// slice_base[s] = offset in units of 64 uint16 | wide << 31.
// A column holds first N_2(i) (2h range, what NeighborhoodCache::filter_down keeps, neighborhood_search.rs:56-70):
// rows 0 .. cnt_near-1, padded with the particle's own index up to the next multiple of 8 (a zero-distance pair
// contributes nothing to any gradient sum, so the pair loops run whole chunks without per-entry bounds checks); the
// rest of the extended range used by the level set follows from row pad8(cnt_near) on.  cnt[i] = cnt_near | cnt_ext << 16,
// both counting real neighbours only (NbCol::get(k), k < cnt_ext, skips the padding).
// No per-pair coefficient is stored: every pass recomputes dW/dr / r from the gathered positions (pair_g below).
#pragma once
#include "sim.cuh"

#define ASPH_PAIR_BLOCK 256u  // threads per block of the pair passes == alignment of the index bias

struct NbLists {
  const uint16_t* __restrict__ pool;
  const uint32_t* __restrict__ slice_base;
  const uint32_t* __restrict__ cnt;
};

#ifdef __CUDACC__
#define ASPH_KNORM 1.81891363533f  // 40 / (7 pi)
__host__ __device__ __forceinline__ uint32_t nb_bias(uint32_t i) { return (i & ~(ASPH_PAIR_BLOCK - 1u)) - 32768u; }
struct NbCol {
  const uint16_t* p16;  // lane's first chunk, narrow view
  const uint32_t* p32;  // lane's first chunk, wide view
  uint32_t bias;        // i0 - 32768 (mod 2^32)
  uint32_t cn, ce;      // real neighbours in the 2h range / in the extended range
  bool wide;
  __device__ __forceinline__ NbCol() : p16(nullptr), p32(nullptr), bias(0), cn(0), ce(0), wide(false) {}  // empty column
  __device__ __forceinline__ NbCol(const NbLists& L, uint32_t i) {
    const uint32_t sb = __ldg(&L.slice_base[i >> 5]);
    const uint32_t c = __ldg(&L.cnt[i]);
    cn = c & 0xffffu; ce = c >> 16;
    wide = (sb >> 31) != 0u;
    const uint16_t* base = L.pool + size_t(sb & 0x7fffffffu) * 64u;
    p16 = base + (i & 31u) * 8u;
    p32 = reinterpret_cast<const uint32_t*>(base) + (i & 31u) * 4u;
    bias = nb_bias(i);
  }
  __device__ __forceinline__ uint32_t row(uint32_t r) const {  // stored row r of the column
    return wide ? p32[(r >> 2) * 128u + (r & 3u)] : bias + uint32_t(p16[(r >> 3) * 256u + (r & 7u)]);
  }
  // k-th real neighbour, k < ce (cold paths: level set, resampling)
  __device__ __forceinline__ uint32_t get(uint32_t k) const { return row(k < cn ? k : k - cn + ((cn + 7u) & ~7u)); }
  // rows [k0, k0 + 8) of the 2h part (k0 a multiple of 8, k0 < cn): real neighbours, then the particle itself as padding.
  // Streaming loads: list entries are read once per pass and should not displace the gathered packs in L2.
  __device__ __forceinline__ void get8(uint32_t k0, uint32_t (&j)[8]) const {
    if (!wide) {
      const uint4 v = __ldcs(reinterpret_cast<const uint4*>(p16 + (k0 >> 3) * 256u));
      j[0] = bias + (v.x & 0xffffu); j[1] = bias + (v.x >> 16);
      j[2] = bias + (v.y & 0xffffu); j[3] = bias + (v.y >> 16);
      j[4] = bias + (v.z & 0xffffu); j[5] = bias + (v.z >> 16);
      j[6] = bias + (v.w & 0xffffu); j[7] = bias + (v.w >> 16);
    } else {
      const uint4 a = __ldcs(reinterpret_cast<const uint4*>(p32 + (k0 >> 2) * 128u));
      const uint4 b = __ldcs(reinterpret_cast<const uint4*>(p32 + ((k0 >> 2) + 1u) * 128u));
      j[0] = a.x; j[1] = a.y; j[2] = a.z; j[3] = a.w;
      j[4] = b.x; j[5] = b.y; j[6] = b.z; j[7] = b.w;
    }
  }
  // raw rows [k0, k0 + 8) of a narrow column (8 x uint16), and their decoding into window offsets
  __device__ __forceinline__ uint4 raw8(uint32_t k0) const { return __ldcs(reinterpret_cast<const uint4*>(p16 + (k0 >> 3) * 256u)); }
  static __device__ __forceinline__ void decode8(const uint4& v, uint32_t halo, uint32_t (&off)[8]) {
    const uint32_t K = 32768u - halo;
    off[0] = (v.x & 0xffffu) - K; off[1] = (v.x >> 16) - K;
    off[2] = (v.y & 0xffffu) - K; off[3] = (v.y >> 16) - K;
    off[4] = (v.z & 0xffffu) - K; off[5] = (v.z >> 16) - K;
    off[6] = (v.w & 0xffffu) - K; off[7] = (v.w >> 16) - K;
  }
  // same rows as window offsets: off = j - wa with wa = b0 - halo (mod 2^32); narrow rows need one subtraction.
  // WIDE is warp-uniform (a property of the slice), so callers branch on it once outside their loops.
  template <bool WIDE>
  __device__ __forceinline__ void get8_off(uint32_t k0, uint32_t halo, uint32_t (&off)[8]) const {
    if (!WIDE) {
      const uint32_t K = 32768u - halo;
      const uint4 v = __ldcs(reinterpret_cast<const uint4*>(p16 + (k0 >> 3) * 256u));
      off[0] = (v.x & 0xffffu) - K; off[1] = (v.x >> 16) - K;
      off[2] = (v.y & 0xffffu) - K; off[3] = (v.y >> 16) - K;
      off[4] = (v.z & 0xffffu) - K; off[5] = (v.z >> 16) - K;
      off[6] = (v.w & 0xffffu) - K; off[7] = (v.w >> 16) - K;
    } else {
      const uint32_t wa = bias + 32768u - halo;
      const uint4 a = __ldcs(reinterpret_cast<const uint4*>(p32 + (k0 >> 2) * 128u));
      const uint4 b = __ldcs(reinterpret_cast<const uint4*>(p32 + ((k0 >> 2) + 1u) * 128u));
      off[0] = a.x - wa; off[1] = a.y - wa; off[2] = a.z - wa; off[3] = a.w - wa;
      off[4] = b.x - wa; off[5] = b.y - wa; off[6] = b.z - wa; off[7] = b.w - wa;
    }
  }
};
// rows a column with cn near and ce total neighbours occupies
__device__ __forceinline__ uint32_t nb_col_rows(uint32_t cn, uint32_t ce) { return ((cn + 7u) & ~7u) + (ce - cn); }
// pool units (64 uint16) a slice of `rows` rows needs
__device__ __forceinline__ uint32_t nb_slice_units(uint32_t rows, bool wide) { return 4u * (wide ? (rows + 3u) / 4u : (rows + 7u) / 8u); }
// where entry k of lane `lane` goes when the slice is being filled
__device__ __forceinline__ void nb_store(uint16_t* slice, bool wide, uint32_t lane, uint32_t k, uint32_t j, uint32_t bias) {
  if (wide) reinterpret_cast<uint32_t*>(slice)[(k >> 2) * 128u + lane * 4u + (k & 3u)] = j;
  else slice[(k >> 3) * 256u + lane * 8u + (k & 7u)] = uint16_t(j - bias);
}

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
