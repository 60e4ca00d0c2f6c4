// lists.cuh — the per-step neighbour lists: sliced, chunked ELL of 16-bit relative indices.
//
// A slice is 32 consecutive particles (one warp).  A particle's column is cut into chunks of 8 entries; chunk c of
// lane l is the 16 bytes at  pool[off + c*256 + l*8 .. +8)  (uint16 units), so ONE 16-byte vector load per lane brings
// 8 neighbour indices and the warp's loads of a chunk are one contiguous 512 B run.  A whole 2h column (about 14
// entries) is therefore two load instructions, after which all gathers of the column can be in flight at once.
// Particles are sorted by grid cell, so a neighbour index j is close to the slice's first particle i0: it is stored as
// uint16(j - i0 + 32768).  A slice in which some neighbour is further than +-32767 positions away (neighbours in
// another size level's grid) is stored "wide": plain 32-bit indices, chunks of 4.
// slice_base[s] = offset in units of 64 uint16 | wide << 31.
// Entries 0 .. cnt_near-1 of a column are N_2(i) (2h range, what NeighborhoodCache::filter_down keeps,
// neighborhood_search.rs:56-70); entries cnt_near .. cnt_ext-1 are the rest of the extended range used by the level
// set.  cnt[i] = cnt_near | cnt_ext << 16.
// No per-pair coefficient is stored: every pass recomputes dW/dr / r from the gathered positions (pair_g below).
#pragma once
#include "sim.cuh"

struct NbLists {
  const uint16_t* __restrict__ pool;
  const uint32_t* __restrict__ slice_base;
  const uint32_t* __restrict__ cnt;
};

#ifdef __CUDACC__
#define ASPH_KNORM 1.81891363533f  // 40 / (7 pi)
struct NbCol {
  const uint16_t* p16;  // lane's first chunk, narrow view
  const uint32_t* p32;  // lane's first chunk, wide view
  uint32_t bias;        // i0 - 32768 (mod 2^32)
  uint32_t self;        // the particle itself (used for padding entries: zero distance => zero pair term)
  bool wide;
  __device__ __forceinline__ NbCol(const NbLists& L, uint32_t i) {
    const uint32_t sb = __ldg(&L.slice_base[i >> 5]);
    wide = (sb >> 31) != 0u;
    const uint16_t* base = L.pool + size_t(sb & 0x7fffffffu) * 64u;
    p16 = base + (i & 31u) * 8u;
    p32 = reinterpret_cast<const uint32_t*>(base) + (i & 31u) * 4u;
    bias = (i & ~31u) - 32768u;
    self = i;
  }
  // single entry (cold paths: level set, resampling)
  __device__ __forceinline__ uint32_t get(uint32_t k) const {
    return wide ? p32[(k >> 2) * 128u + (k & 3u)] : bias + uint32_t(p16[(k >> 3) * 256u + (k & 7u)]);
  }
  // entries [k0, k0 + 8) of the column (k0 a multiple of 8); entries at or beyond `count` are replaced by `self`.
  // Streaming loads: list entries are read once per pass and should not displace the gathered packs in L2.
  __device__ __forceinline__ void get8(uint32_t k0, uint32_t count, uint32_t (&j)[8]) const {
    if (!wide) {
      const uint4 v = __ldcs(reinterpret_cast<const uint4*>(p16 + (k0 >> 3) * 256u));
      j[0] = bias + (v.x & 0xffffu); j[1] = bias + (v.x >> 16);
      j[2] = bias + (v.y & 0xffffu); j[3] = bias + (v.y >> 16);
      j[4] = bias + (v.z & 0xffffu); j[5] = bias + (v.z >> 16);
      j[6] = bias + (v.w & 0xffffu); j[7] = bias + (v.w >> 16);
    } else {
      const uint4 a = __ldcs(reinterpret_cast<const uint4*>(p32 + (k0 >> 2) * 128u));
      j[0] = a.x; j[1] = a.y; j[2] = a.z; j[3] = a.w;
      if (k0 + 4u < count) {
        const uint4 b = __ldcs(reinterpret_cast<const uint4*>(p32 + ((k0 >> 2) + 1u) * 128u));
        j[4] = b.x; j[5] = b.y; j[6] = b.z; j[7] = b.w;
      } else {
        j[4] = j[5] = j[6] = j[7] = self;
      }
    }
#pragma unroll
    for (int t = 0; t < 8; t++) if (k0 + uint32_t(t) >= count) j[t] = self;
  }
};
// pool units (64 uint16) a slice of width `we` entries needs
__device__ __forceinline__ uint32_t nb_slice_units(uint32_t we, bool wide) { return 4u * (wide ? (we + 3u) / 4u : (we + 7u) / 8u); }
// where entry k of lane `lane` goes when the slice is being filled
__device__ __forceinline__ void nb_store(uint16_t* slice, bool wide, uint32_t lane, uint32_t k, uint32_t j, uint32_t bias) {
  if (wide) reinterpret_cast<uint32_t*>(slice)[(k >> 2) * 128u + lane * 4u + (k & 3u)] = j;
  else slice[(k >> 3) * 256u + lane * 8u + (k & 7u)] = uint16_t(j - bias);
}

// dW/dr / r for the cubic spline (sph_kernels.rs:61-71) from the squared distance; gradW_ij = pair_g * x_ij.
// Zero for q <= 1e-5 (the reference's guard), also for r = 0 (self) where rsqrt gives inf -> NaN -> comparison false.
__device__ __forceinline__ float pair_g(float d2, float hij) {
  const float inv_r = rsqrtf(d2);
  const float inv2h = __frcp_rn(2.f * hij);
  const float q = (d2 * inv_r) * inv2h;
  const float v = fmaxf(1.f - q, 0.f);
  const float dw = q < 0.5f ? (18.f * q - 12.f) * q : -6.f * v * v;
  // norm_factor / (2h) = 10 / (7 pi h^2) / (2h) = (40 / (7 pi)) * inv2h^3
  const float nfac = ASPH_KNORM * inv2h * inv2h * inv2h;
  return q > 1.0e-5f ? nfac * dw * inv_r : 0.f;
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
// uniform-h variant: inv2h and nfac are kernel constants
__device__ __forceinline__ float pair_g_uniform(float d2, float inv2h, float nfac) {
  const float inv_r = rsqrtf(d2);
  const float q = (d2 * inv_r) * inv2h;
  const float v = fmaxf(1.f - q, 0.f);
  const float dw = q < 0.5f ? (18.f * q - 12.f) * q : -6.f * v * v;
  return q > 1.0e-5f ? nfac * dw * inv_r : 0.f;
}
#endif
