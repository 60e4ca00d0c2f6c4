// dist.cu — multi-GPU: 1-D slab decomposition of the particle set along x with a one-support-radius ghost halo.
//
// One process per GPU (rank r owns the particles with bounds[r] <= x < bounds[r+1]); the processes are tied together
// by an NCCL communicator whose unique id the host harness broadcasts (asph_comm_unique_id, asph_create_distributed).
// The reference has no counterpart (single address space); what has to hold is that the N-GPU
// step computes what the 1-GPU step computes: the neighbour set of every owned particle is complete (ghost width =
// largest pair support f * h_max), and every pass that reads a neighbour field written by the previous pass sees the
// owner's value (one halo exchange per such pass: rho, v after the non-pressure forces, a^p and p per Jacobi sweep, ...).
//
// Per step (dist_begin_step):  drop last step's ghosts, hand particles that left the slab to the neighbour rank
// (migration), pick the border particles within the ghost width and send them to the neighbour (ghost exchange),
// then the ordinary single-GPU sort / grid / neighbour build runs over owned + ghost particles.  Ghosts carry
// ASPH_GHOST_BIT in refid (refid = global particle index); they are computed like any particle (their results are
// garbage near the outer edge of the halo and are overwritten by the owner's values) but never counted in reductions.
// Global scalars: dt (min of the CFL term), the 4 Jacobi statistics per sweep (sum), error flags (max).
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): a single-GPU user of the library needs no NCCL at all, and in a
// process that already loaded torch's NCCL the same copy is used.
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstring>

#include "lists.cuh"

namespace {

constexpr int kThreads = 256;
constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr int kHistBins = 8192;

// host copy of sim.cuh's dec_f (order-preserving float encoding used with atomicMin / atomicMax)
inline float host_dec_f(uint32_t e) {
  const uint32_t u = (e & 0x80000000u) ? (e & 0x7fffffffu) : ~e;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

struct Nccl {
  void* so = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

Nccl& nccl() {
  static Nccl api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    api.so = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.so) break;
  }
  if (!api.so) return api;
  bool all = true;
  auto sym = [&](const char* s) { void* p = dlsym(api.so, s); if (!p) all = false; return p; };
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
  api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
  api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  api.ok = all;
  return api;
}

// device words gathered over all ranks once per phase
enum { W_STAY = 0, W_LEFT = 1, W_RIGHT = 2, W_HMAX = 3, W_GLEFT = 4, W_GRIGHT = 5, W_MINX = 6, W_MAXX = 7, W_COUNT = 8 };

}  // namespace

struct DistState {
  int rank = 0, nranks = 1;
  ncclComm_t comm = nullptr;
  uint64_t n_global = 0;
  std::vector<float> bounds;  // nranks + 1 entries; bounds[0] = -inf, bounds[nranks] = +inf
  bool rebalance_due = true, full_migration = true;
  uint64_t steps = 0;
  float ghost_w = 0.f;
  uint32_t cap_h = 0;         // halo / migration buffer capacity in particles per side
  uint32_t n_send[2] = {0, 0}, n_recv[2] = {0, 0};  // ghost halo sizes; [0] = towards / from rank-1, [1] = rank+1
  DevBuf<uint32_t> send_idx, recv_idx;  // sorted particle indices, left part then right part
  DevBuf<uint32_t> slot[2];             // per pre-sort owned particle: position in the send list of that side or kNone
  DevBuf<float4> sendbuf, recvbuf;      // 2 float4 per particle and side
  DevBuf<uint32_t> words, gather, hist, flag_bits;
  uint32_t* gather_host = nullptr;      // pinned: nranks * W_COUNT words, or the histogram
  bool map_valid = false;
  uint64_t halo_calls = 0, halo_bytes = 0;
  // peer-memory path (sweeps): see PeerCtl in sim.cuh
  bool p2p = false;
  PeerCtl* my_ctl = nullptr;                      // device memory of this rank
  PeerCtl* peer_ctl[ASPH_MAX_RANKS] = {nullptr};  // every rank's PeerCtl mapped here ([rank] = my_ctl)
  void* mapped_src[2][4] = {{nullptr}};           // the neighbour's own pointers of packA / packP[0] / packP[1] / mailboxes last mapped ([0] = rank - 1)
  float4* peer_field[2][4] = {{nullptr}};         // ... and where they are mapped in this process
  DevBuf<uint32_t> remote_slot;                   // ghost slot on the neighbour of my k-th send-list entry (left part, right part)
  DevBuf<uint32_t> rslot[2], blocks_done;         // the same per local particle (~0: not a border particle); pass-completion counter
  DevBuf<unsigned char> tile_border;
  DevBuf<uint32_t> tile_order;                    // edge tiles first; [ntiles] = their number
  DevBuf<unsigned char> rec_dev;                  // allgather buffers of P2PRecord: mine, then one per rank
  unsigned char* rec_host = nullptr;
  DevBuf<unsigned long long> peer_ctl_dev;        // peer_ctl[] for k_stats_push
  bool peer_ctl_ready = false;
  unsigned int halo_seq = 0, stats_seq = 0;       // pushes launched so far
  // persistent cooperative kernels (CoopPeer, sim.cuh)
  DevBuf<uint2> mbox;                             // this rank's mailboxes: 2 parities x 2 sides x cap_h messages
  uint32_t mbox_cap = 0;
  DevBuf<uint32_t> owner_slot, owner_tmp;         // per local particle: a ghost's index on its owner rank (~0 otherwise)
  DevBuf<unsigned int> verdict;
  unsigned int coop_seq = 0;                      // cross-GPU barriers run so far (the same on every rank)
  // resampling: lists gathered over all ranks, arrays over the shared reference index space
  DevBuf<uint32_t> ag_send, ag_recv, ag_counts, ref_a, ref_b;
  DevBuf<unsigned long long> red;
};

namespace {

#define NCCL_TRY(x)                                                                                   \
  do {                                                                                                \
    ncclResult_t r__ = (x);                                                                           \
    if (r__ != ncclSuccess) {                                                                         \
      sim->last_error = std::string(#x) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(r__) : "nccl error"); \
      return ASPH_ERR_NCCL;                                                                           \
    }                                                                                                 \
  } while (0)

__global__ void k_words_reset(uint32_t* w) {
  w[W_STAY] = 0; w[W_LEFT] = 0; w[W_RIGHT] = 0; w[W_HMAX] = 0; w[W_GLEFT] = 0; w[W_GRIGHT] = 0;
  w[W_MINX] = 0xFFFFFFFFu; w[W_MAXX] = 0;
}

// Owned particles are partitioned into stay (compacted into the other buffer) / leave-left / leave-right (packed into
// the send buffer); last step's ghosts are dropped.  Warp-aggregated appends: the arrival order is arbitrary, the
// cell sort orders by global index afterwards.  Also h_max and the x-extent of the owned particles.
__global__ void __launch_bounds__(kThreads)
k_owner_split(uint32_t n, const float2* __restrict__ pos, const float2* __restrict__ vel, const float* __restrict__ mass,
              const float* __restrict__ level, const uint32_t* __restrict__ refid, float lo, float hi, float rho0,
              float2* __restrict__ pos_o, float2* __restrict__ vel_o, float* __restrict__ mass_o, float* __restrict__ level_o,
              uint32_t* __restrict__ refid_o, float4* __restrict__ send_l, float4* __restrict__ send_r, uint32_t cap_h,
              uint32_t* __restrict__ w) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31u;
  int dest = -1;  // -1 none (out of range or ghost), 0 stay, 1 left, 2 right
  float2 x = make_float2(0.f, 0.f), v = x;
  float m = 0.f, lv = 0.f;
  uint32_t id = 0;
  if (i < n) {
    id = refid[i];
    if (!(id & ASPH_GHOST_BIT)) {
      x = pos[i]; v = vel[i]; m = mass[i]; lv = level[i];
      dest = x.x < lo ? 1 : (x.x >= hi ? 2 : 0);
    }
  }
  float hm = dest >= 0 ? h_from_mass(m, rho0) : 0.f;
  float mnx = dest >= 0 ? x.x : __int_as_float(0x7f800000), mxx = dest >= 0 ? x.x : -__int_as_float(0x7f800000);
  for (int o = 16; o > 0; o >>= 1) {
    hm = fmaxf(hm, __shfl_xor_sync(0xffffffffu, hm, o));
    mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
    mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
  }
  if (lane == 0 && hm > 0.f) { atomicMax(&w[W_HMAX], enc_f(hm)); atomicMin(&w[W_MINX], enc_f(mnx)); atomicMax(&w[W_MAXX], enc_f(mxx)); }
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const unsigned int mask = __ballot_sync(0xffffffffu, dest == d);
    if (!mask) continue;
    uint32_t base = 0;
    const int leader = __ffs(mask) - 1;
    if (int(lane) == leader) base = atomicAdd(&w[d], uint32_t(__popc(mask)));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (dest != d) continue;
    const uint32_t k = base + uint32_t(__popc(mask & ((1u << lane) - 1u)));
    if (d == 0) {
      pos_o[k] = x; vel_o[k] = v; mass_o[k] = m; level_o[k] = lv; refid_o[k] = id;
    } else if (k < cap_h) {
      float4* dst = (d == 1 ? send_l : send_r) + 2 * size_t(k);
      dst[0] = make_float4(x.x, x.y, v.x, v.y);
      dst[1] = make_float4(m, lv, __uint_as_float(id), 0.f);
    }
  }
}

// Border particles within the ghost width of a slab face: packed for the neighbour and remembered by slot.
__global__ void __launch_bounds__(kThreads)
k_ghost_select(uint32_t n_owned, const float2* __restrict__ pos, const float2* __restrict__ vel, const float* __restrict__ mass,
               const float* __restrict__ level, const uint32_t* __restrict__ refid, float left_below, float right_from,
               uint32_t* __restrict__ slot_l, uint32_t* __restrict__ slot_r, float4* __restrict__ send_l, float4* __restrict__ send_r,
               uint32_t cap_h, uint32_t* __restrict__ w) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31u;
  bool to_l = false, to_r = false;
  float2 x = make_float2(0.f, 0.f);
  if (i < n_owned) { x = pos[i]; to_l = x.x < left_below; to_r = x.x >= right_from; }
#pragma unroll
  for (int d = 0; d < 2; d++) {
    const bool mine = d == 0 ? to_l : to_r;
    const unsigned int mask = __ballot_sync(0xffffffffu, mine);
    uint32_t k = kNone;
    if (mask) {
      uint32_t base = 0;
      const int leader = __ffs(mask) - 1;
      if (int(lane) == leader) base = atomicAdd(&w[d == 0 ? W_GLEFT : W_GRIGHT], uint32_t(__popc(mask)));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (mine) {
        k = base + uint32_t(__popc(mask & ((1u << lane) - 1u)));
        if (k < cap_h) {
          const float2 v = vel[i];
          float4* dst = (d == 0 ? send_l : send_r) + 2 * size_t(k);
          dst[0] = make_float4(x.x, x.y, v.x, v.y);
          dst[1] = make_float4(mass[i], level[i], __uint_as_float(refid[i]), 0.f);
        }
      }
    }
    if (i < n_owned) (d == 0 ? slot_l : slot_r)[i] = k;
  }
}

__global__ void k_append(uint32_t count, const float4* __restrict__ recv, uint32_t at, uint32_t idbits, float2* __restrict__ pos,
                         float2* __restrict__ vel, float* __restrict__ mass, float* __restrict__ level, uint32_t* __restrict__ refid) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const float4 a = recv[2 * size_t(k)], b = recv[2 * size_t(k) + 1];
  pos[at + k] = make_float2(a.x, a.y); vel[at + k] = make_float2(a.z, a.w);
  mass[at + k] = b.x; level[at + k] = b.y; refid[at + k] = __float_as_uint(b.z) | idbits;
}

// After the sort: where did each halo particle end up?  order[s] = pre-sort index of sorted particle s.
__global__ void k_build_maps(uint32_t n, const uint32_t* __restrict__ order, uint32_t n_owned, uint32_t recv_l,
                             const uint32_t* __restrict__ slot_l, const uint32_t* __restrict__ slot_r, uint32_t send_l,
                             uint32_t* __restrict__ send_idx, uint32_t* __restrict__ recv_idx) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const uint32_t src = order[s];
  if (src >= n_owned) { recv_idx[src - n_owned] = s; return; }  // ghosts were appended left part first
  const uint32_t a = slot_l[src], b = slot_r[src];
  if (a != kNone) send_idx[a] = s;
  if (b != kNone) send_idx[send_l + b] = s;
  (void)recv_l;
}

// field[idx[k]] -> buf[k] (words 32-bit words per element); P0/P1 + ctl: pick the pressure pack by sweep parity
__global__ void k_pack(uint32_t count, int words, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ f0,
                       const uint32_t* __restrict__ f1, const StepCtl* __restrict__ ctl, uint32_t* __restrict__ buf) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count * uint32_t(words)) return;
  const uint32_t* __restrict__ f = (ctl && (ctl->solver.sweeps & 1)) ? f1 : f0;
  const uint32_t k = t / uint32_t(words), c = t - k * uint32_t(words);
  buf[t] = f[size_t(idx[k]) * words + c];
}
__global__ void k_unpack(uint32_t count, int words, const uint32_t* __restrict__ idx, uint32_t* __restrict__ f0,
                         uint32_t* __restrict__ f1, const StepCtl* __restrict__ ctl, const uint32_t* __restrict__ buf) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count * uint32_t(words)) return;
  uint32_t* __restrict__ f = (ctl && (ctl->solver.sweeps & 1)) ? f1 : f0;
  const uint32_t k = t / uint32_t(words), c = t - k * uint32_t(words);
  f[size_t(idx[k]) * words + c] = buf[t];
}

__global__ void k_hist(uint32_t n, const float2* __restrict__ pos, const uint32_t* __restrict__ refid, float x0, float inv_w,
                       uint32_t* __restrict__ hist) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || (refid[i] & ASPH_GHOST_BIT)) return;
  const int b = min(kHistBins - 1, max(0, int((pos[i].x - x0) * inv_w)));
  atomicAdd(&hist[b], 1u);
}

__global__ void k_owned_flag(uint32_t n, const uint32_t* __restrict__ refid, uint32_t* __restrict__ flag) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = (refid[i] & ASPH_GHOST_BIT) ? 0u : 1u;
}
__global__ void k_mask_ghost_slots(uint32_t n, const uint32_t* __restrict__ refid, uint32_t* __restrict__ map) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && (refid[i] & ASPH_GHOST_BIT)) map[i] = kNone;
}
__global__ void k_global_index(uint32_t n, const uint32_t* __restrict__ refid, const uint32_t* __restrict__ map, uint32_t* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && map[i] != kNone) out[map[i]] = refid[i] & ~ASPH_GHOST_BIT;
}

int gather_words(asph_sim* sim) {  // all ranks' phase words -> D->gather_host
  DistState* D = sim->dist;
  NCCL_TRY(nccl().AllGather(D->words.p, D->gather.p, W_COUNT, ncclUint32, D->comm, sim->stream));
  CUDA_TRY(cudaMemcpyAsync(D->gather_host, D->gather.p, size_t(D->nranks) * W_COUNT * sizeof(uint32_t), cudaMemcpyDeviceToHost, sim->stream));
  CUDA_TRY(cudaStreamSynchronize(sim->stream));
  return ASPH_OK;
}

int ensure_halo_capacity(asph_sim* sim, uint32_t want) {
  DistState* D = sim->dist;
  if (want <= D->cap_h) return ASPH_OK;
  D->cap_h = want;
  CUDA_TRY(D->sendbuf.ensure(size_t(want) * 4)); CUDA_TRY(D->recvbuf.ensure(size_t(want) * 4));
  CUDA_TRY(D->send_idx.ensure(size_t(want) * 2)); CUDA_TRY(D->recv_idx.ensure(size_t(want) * 2));
  return ASPH_OK;
}

// payloads of 2 float4 per particle to / from the two neighbour ranks; receives land contiguously (left part first)
int exchange_particles(asph_sim* sim, uint32_t out_l, uint32_t out_r, uint32_t in_l, uint32_t in_r) {
  DistState* D = sim->dist;
  if (!(out_l | out_r | in_l | in_r)) return ASPH_OK;
  const Nccl& N = nccl();
  NCCL_TRY(N.GroupStart());
  if (out_l) NCCL_TRY(N.Send(D->sendbuf.p, size_t(out_l) * 8, ncclFloat, D->rank - 1, D->comm, sim->stream));
  if (out_r) NCCL_TRY(N.Send(D->sendbuf.p + 2 * size_t(D->cap_h), size_t(out_r) * 8, ncclFloat, D->rank + 1, D->comm, sim->stream));
  if (in_l) NCCL_TRY(N.Recv(D->recvbuf.p, size_t(in_l) * 8, ncclFloat, D->rank - 1, D->comm, sim->stream));
  if (in_r) NCCL_TRY(N.Recv(D->recvbuf.p + 2 * size_t(in_l), size_t(in_r) * 8, ncclFloat, D->rank + 1, D->comm, sim->stream));
  NCCL_TRY(N.GroupEnd());
  return ASPH_OK;
}

// Slab faces from the global x-histogram of the owned particles: equal particle counts per rank.
int rebalance(asph_sim* sim) {
  DistState* D = sim->dist;
  const int R = D->nranks;
  D->bounds.assign(size_t(R) + 1, 0.f);
  D->bounds[0] = -INFINITY; D->bounds[R] = INFINITY;
  D->rebalance_due = false;
  D->full_migration = true;
  if (R == 1) return ASPH_OK;
  cudaStream_t st = sim->stream;
  const int c = sim->cur;
  const uint32_t n = sim->n;
  // global x-extent: run the owner split's statistics only (a split with lo = -inf, hi = +inf moves nothing)
  k_words_reset<<<1, 1, 0, st>>>(D->words.p);
  LAUNCH_CHECK();
  if (n) {
    k_owner_split<<<(n + kThreads - 1) / kThreads, kThreads, 0, st>>>(
        n, sim->pos[c].p, sim->vel[c].p, sim->mass[c].p, sim->level[c].p, sim->refid[c].p, -INFINITY, INFINITY, sim->pp.rest_density,
        sim->pos[1 - c].p, sim->vel[1 - c].p, sim->mass[1 - c].p, sim->level[1 - c].p, sim->refid[1 - c].p, D->sendbuf.p,
        D->sendbuf.p + 2 * size_t(D->cap_h), D->cap_h, D->words.p);
    LAUNCH_CHECK();
  }
  TRY(gather_words(sim));
  float gmin = INFINITY, gmax = -INFINITY;
  uint64_t total = 0;
  for (int r = 0; r < R; r++) {
    const uint32_t* w = D->gather_host + size_t(r) * W_COUNT;
    if (w[W_STAY] == 0) continue;
    total += w[W_STAY];
    gmin = std::min(gmin, host_dec_f(w[W_MINX])); gmax = std::max(gmax, host_dec_f(w[W_MAXX]));
  }
  if (total == 0 || !(gmax > gmin)) {  // nothing to split by position: equal-width faces are as good as any
    for (int r = 1; r < R; r++) D->bounds[r] = (total == 0 ? 0.f : gmin) + float(r);
    return ASPH_OK;
  }
  const float binw = (gmax - gmin) / float(kHistBins);
  CUDA_TRY(cudaMemsetAsync(D->hist.p, 0, kHistBins * sizeof(uint32_t), st));
  if (n) {
    k_hist<<<(n + kThreads - 1) / kThreads, kThreads, 0, st>>>(n, sim->pos[c].p, sim->refid[c].p, gmin, 1.f / binw, D->hist.p);
    LAUNCH_CHECK();
  }
  NCCL_TRY(nccl().AllReduce(D->hist.p, D->hist.p, kHistBins, ncclUint32, ncclSum, D->comm, st));
  std::vector<uint32_t> hist(kHistBins);
  CUDA_TRY(cudaMemcpyAsync(hist.data(), D->hist.p, kHistBins * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  uint64_t cum = 0;
  int b = 0;
  for (int r = 1; r < R; r++) {
    const uint64_t target = total * uint64_t(r) / uint64_t(R);
    while (b < kHistBins && cum + hist[b] < target) { cum += hist[b]; b++; }
    const float frac = (b < kHistBins && hist[b] > 0) ? float(double(target - cum) / double(hist[b])) : 0.f;
    D->bounds[r] = gmin + (float(b) + frac) * binw;
  }
  return ASPH_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------
// step hooks
// ------------------------------------------------------------------------------------------------------------
int dist_begin_step(asph_sim* sim, float f_search) {
  DistState* D = sim->dist;
  const int R = D->nranks, r = D->rank;
  cudaStream_t st = sim->stream;
  D->map_valid = false;
  if (D->steps > 0 && (D->steps % 64) == 0 && R > 1) {  // imbalance check on the counts every rank already holds
    uint64_t total = 0, worst = 0;
    for (int k = 0; k < R; k++) { const uint64_t c = D->gather_host[size_t(k) * W_COUNT + W_STAY]; total += c; worst = std::max(worst, c); }
    if (total > 0 && double(worst) * R > 1.10 * double(total)) D->rebalance_due = true;
  }
  D->steps++;
  if (D->rebalance_due) TRY(rebalance(sim));
  const float lo = D->bounds[r], hi = D->bounds[r + 1];
  float hmax_g = 0.f;
  // ---- migration: one round normally (a particle moves less than a slab per step); after a rebalance as many
  //      rounds as it takes for every particle to reach its owner, one rank per round
  for (int round = 0; round < R + 1; round++) {
    const int c = sim->cur;
    const uint32_t n = sim->n;
    uint32_t out_l = 0, out_r = 0, in_l = 0, in_r = 0, n_stay = 0;
    uint64_t moved = 0;
    for (;;) {
      k_words_reset<<<1, 1, 0, st>>>(D->words.p);
      LAUNCH_CHECK();
      if (n) {
        k_owner_split<<<(n + kThreads - 1) / kThreads, kThreads, 0, st>>>(
            n, sim->pos[c].p, sim->vel[c].p, sim->mass[c].p, sim->level[c].p, sim->refid[c].p, lo, hi, sim->pp.rest_density,
            sim->pos[1 - c].p, sim->vel[1 - c].p, sim->mass[1 - c].p, sim->level[1 - c].p, sim->refid[1 - c].p, D->sendbuf.p,
            D->sendbuf.p + 2 * size_t(D->cap_h), D->cap_h, D->words.p);
        LAUNCH_CHECK();
      }
      TRY(gather_words(sim));
      uint32_t need = 0;
      moved = 0; hmax_g = 0.f;
      for (int k = 0; k < R; k++) {
        const uint32_t* w = D->gather_host + size_t(k) * W_COUNT;
        need = std::max(need, std::max(w[W_LEFT], w[W_RIGHT]));
        moved += uint64_t(w[W_LEFT]) + w[W_RIGHT];
        if (w[W_HMAX]) hmax_g = std::max(hmax_g, host_dec_f(w[W_HMAX]));
      }
      if (need <= D->cap_h) break;
      TRY(ensure_halo_capacity(sim, need + need / 2));  // same decision on every rank (all see all counts); split again
    }
    const uint32_t* me = D->gather_host + size_t(r) * W_COUNT;
    n_stay = me[W_STAY]; out_l = me[W_LEFT]; out_r = me[W_RIGHT];
    in_l = r > 0 ? D->gather_host[size_t(r - 1) * W_COUNT + W_RIGHT] : 0u;
    in_r = r < R - 1 ? D->gather_host[size_t(r + 1) * W_COUNT + W_LEFT] : 0u;
    sim->cur = 1 - c;
    sim->n = n_stay;
    const uint64_t n_new = uint64_t(n_stay) + in_l + in_r;
    if (n_new + 2 * uint64_t(D->cap_h) > sim->cap) TRY(ensure_capacity(sim, uint32_t(std::min<uint64_t>(n_new + n_new / 4 + 2 * uint64_t(D->cap_h) + 1024, 0x7FFFFFF0ull))));
    TRY(exchange_particles(sim, out_l, out_r, in_l, in_r));
    if (in_l + in_r) {
      k_append<<<(in_l + in_r + kThreads - 1) / kThreads, kThreads, 0, st>>>(in_l + in_r, D->recvbuf.p, n_stay, 0u, sim->pos[sim->cur].p,
                                                                             sim->vel[sim->cur].p, sim->mass[sim->cur].p,
                                                                             sim->level[sim->cur].p, sim->refid[sim->cur].p);
      LAUNCH_CHECK();
    }
    sim->n = uint32_t(n_new);
    sim->n_owned = sim->n;
    if (!D->full_migration || moved == 0) break;
  }
  D->full_migration = false;
  // ---- ghost exchange
  const float W = f_search * hmax_g * 1.0001f;
  D->ghost_w = W;
  for (int k = 1; k + 1 < R; k++) {
    if (!(D->bounds[k + 1] - D->bounds[k] >= W)) {
      sim->last_error = "a slab is narrower than the ghost width (too many GPUs for this particle size)";
      return ASPH_ERR_INVALID;
    }
  }
  const int c = sim->cur;
  const uint32_t no = sim->n_owned;
  CUDA_TRY(D->slot[0].ensure(sim->cap)); CUDA_TRY(D->slot[1].ensure(sim->cap));
  uint32_t gs_l = 0, gs_r = 0, gr_l = 0, gr_r = 0;
  for (;;) {
    k_words_reset<<<1, 1, 0, st>>>(D->words.p);
    LAUNCH_CHECK();
    if (no) {
      k_ghost_select<<<(no + kThreads - 1) / kThreads, kThreads, 0, st>>>(
          no, sim->pos[c].p, sim->vel[c].p, sim->mass[c].p, sim->level[c].p, sim->refid[c].p, r > 0 ? lo + W : -INFINITY,
          r < R - 1 ? hi - W : INFINITY, D->slot[0].p, D->slot[1].p, D->sendbuf.p, D->sendbuf.p + 2 * size_t(D->cap_h), D->cap_h, D->words.p);
      LAUNCH_CHECK();
    }
    TRY(gather_words(sim));
    uint32_t need = 0;
    for (int k = 0; k < R; k++) {
      const uint32_t* w = D->gather_host + size_t(k) * W_COUNT;
      need = std::max(need, std::max(w[W_GLEFT], w[W_GRIGHT]));
    }
    if (need <= D->cap_h) break;
    TRY(ensure_halo_capacity(sim, need + need / 2));
  }
  {
    const uint32_t* me = D->gather_host + size_t(r) * W_COUNT;
    gs_l = me[W_GLEFT]; gs_r = me[W_GRIGHT];
    gr_l = r > 0 ? D->gather_host[size_t(r - 1) * W_COUNT + W_GRIGHT] : 0u;
    gr_r = r < R - 1 ? D->gather_host[size_t(r + 1) * W_COUNT + W_GLEFT] : 0u;
    // keep the owned count in the gathered words for the imbalance check
    for (int k = 0; k < R; k++) D->gather_host[size_t(k) * W_COUNT + W_STAY] = (k == r) ? no : D->gather_host[size_t(k) * W_COUNT + W_STAY];
  }
  const uint64_t n_all = uint64_t(no) + gr_l + gr_r;
  if (n_all > sim->cap) TRY(ensure_capacity(sim, uint32_t(std::min<uint64_t>(n_all + n_all / 4 + 1024, 0x7FFFFFF0ull))));
  TRY(exchange_particles(sim, gs_l, gs_r, gr_l, gr_r));
  if (gr_l + gr_r) {
    k_append<<<(gr_l + gr_r + kThreads - 1) / kThreads, kThreads, 0, st>>>(gr_l + gr_r, D->recvbuf.p, no, ASPH_GHOST_BIT, sim->pos[sim->cur].p,
                                                                           sim->vel[sim->cur].p, sim->mass[sim->cur].p, sim->level[sim->cur].p,
                                                                           sim->refid[sim->cur].p);
    LAUNCH_CHECK();
  }
  sim->n = uint32_t(n_all);
  D->n_send[0] = gs_l; D->n_send[1] = gs_r; D->n_recv[0] = gr_l; D->n_recv[1] = gr_r;
  return ASPH_OK;
}

int dist_allreduce_cfl(asph_sim* sim) {
  DistState* D = sim->dist;
  if (D->nranks == 1) return ASPH_OK;
  NCCL_TRY(nccl().AllReduce(&sim->ctl->cfl_enc, &sim->ctl->cfl_enc, 1, ncclUint32, ncclMin, D->comm, sim->stream));
  return ASPH_OK;
}

static int p2p_refresh(asph_sim* sim);

int dist_after_sort(asph_sim* sim) {
  DistState* D = sim->dist;
  const uint32_t n = sim->n;
  if (n == 0 || D->nranks == 1) return ASPH_OK;
  k_build_maps<<<(n + kThreads - 1) / kThreads, kThreads, 0, sim->stream>>>(n, sim->order.p, sim->n_owned, D->n_recv[0], D->slot[0].p, D->slot[1].p,
                                                                           D->n_send[0], D->send_idx.p, D->recv_idx.p);
  LAUNCH_CHECK();
  TRY(p2p_refresh(sim));
  return ASPH_OK;
}

static int halo_impl(asph_sim* sim, void* f0, void* f1, const StepCtl* ctl, int elem_bytes) {
  DistState* D = sim->dist;
  if (D->nranks == 1) return ASPH_OK;
  const Nccl& N = nccl();
  cudaStream_t st = sim->stream;
  const int words = elem_bytes / 4;
  const uint32_t ns = D->n_send[0] + D->n_send[1], nr = D->n_recv[0] + D->n_recv[1];
  uint32_t* sb = reinterpret_cast<uint32_t*>(D->sendbuf.p);
  uint32_t* rb = reinterpret_cast<uint32_t*>(D->recvbuf.p);
  if (ns) {
    k_pack<<<(ns * words + kThreads - 1) / kThreads, kThreads, 0, st>>>(ns, words, D->send_idx.p, (const uint32_t*)f0, (const uint32_t*)f1, ctl, sb);
    LAUNCH_CHECK();
  }
  if (ns | nr) {
    NCCL_TRY(N.GroupStart());
    if (D->n_send[0]) NCCL_TRY(N.Send(sb, size_t(D->n_send[0]) * words, ncclFloat, D->rank - 1, D->comm, st));
    if (D->n_send[1]) NCCL_TRY(N.Send(sb + size_t(D->n_send[0]) * words, size_t(D->n_send[1]) * words, ncclFloat, D->rank + 1, D->comm, st));
    if (D->n_recv[0]) NCCL_TRY(N.Recv(rb, size_t(D->n_recv[0]) * words, ncclFloat, D->rank - 1, D->comm, st));
    if (D->n_recv[1]) NCCL_TRY(N.Recv(rb + size_t(D->n_recv[0]) * words, size_t(D->n_recv[1]) * words, ncclFloat, D->rank + 1, D->comm, st));
    NCCL_TRY(N.GroupEnd());
  }
  if (nr) {
    k_unpack<<<(nr * words + kThreads - 1) / kThreads, kThreads, 0, st>>>(nr, words, D->recv_idx.p, (uint32_t*)f0, (uint32_t*)f1, ctl, rb);
    LAUNCH_CHECK();
  }
  D->halo_calls++;
  D->halo_bytes += uint64_t(ns) * elem_bytes;
  return ASPH_OK;
}

int dist_halo(asph_sim* sim, void* field, int elem_bytes) { return halo_impl(sim, field, field, nullptr, elem_bytes); }

int dist_solver_reduce(asph_sim* sim, int slot) {
  DistState* D = sim->dist;
  if (D->nranks == 1) return ASPH_OK;
  unsigned long long* acc = sim->ctl->solver.acc[slot];
  NCCL_TRY(nccl().AllReduce(acc, acc, ASPH_ACC_WORDS, ncclUint64, ncclSum, D->comm, sim->stream));
  return ASPH_OK;
}

// The error flags are a bit mask and NCCL has no bitwise OR: one word per bit, maximum over the ranks, packed again.
__global__ void k_flags_expand(const StepCtl* __restrict__ ctl, uint32_t* __restrict__ bits) { bits[threadIdx.x] = (ctl->error_flags >> threadIdx.x) & 1u; }
__global__ void k_flags_pack(StepCtl* ctl, const uint32_t* __restrict__ bits) {
  const unsigned int m = __ballot_sync(0xffffffffu, bits[threadIdx.x] != 0u);
  if (threadIdx.x == 0) ctl->error_flags = m;
}
int dist_reduce_flags(asph_sim* sim, bool) {
  DistState* D = sim->dist;
  if (D->nranks == 1) return ASPH_OK;
  k_flags_expand<<<1, 32, 0, sim->stream>>>(sim->ctl, D->flag_bits.p);
  LAUNCH_CHECK();
  NCCL_TRY(nccl().AllReduce(D->flag_bits.p, D->flag_bits.p, 32, ncclUint32, ncclMax, D->comm, sim->stream));
  k_flags_pack<<<1, 32, 0, sim->stream>>>(sim->ctl, D->flag_bits.p);
  LAUNCH_CHECK();
  return ASPH_OK;
}

// ---- peer-memory path --------------------------------------------------------------------------------------------
namespace {

constexpr int kPeerFields = 4;
struct P2PRecord {  // what every rank tells the others once per step
  void* ptr[kPeerFields];                 // its packA, packP[0], packP[1], mailboxes
  void* ctl;                              // its PeerCtl
  cudaIpcMemHandle_t h[kPeerFields], hc;  // and their IPC handles
};

// tiles in processing order: the ones with border particles first (one block; a few thousand tiles)
__global__ void __launch_bounds__(1024) k_tile_order(uint32_t ntiles, const unsigned char* __restrict__ tile_border, uint32_t* __restrict__ order) {
  __shared__ uint32_t s_cnt[1024], s_total;
  const uint32_t per = (ntiles + blockDim.x - 1) / blockDim.x;
  const uint32_t a = min(ntiles, threadIdx.x * per), b = min(ntiles, a + per);
  uint32_t c = 0;
  for (uint32_t t = a; t < b; t++) c += tile_border[t] ? 1u : 0u;
  s_cnt[threadIdx.x] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (uint32_t k = 0; k < blockDim.x; k++) { const uint32_t v = s_cnt[k]; s_cnt[k] = run; run += v; }
    s_total = run;
    order[ntiles] = run;
  }
  __syncthreads();
  uint32_t e = s_cnt[threadIdx.x], in = s_total + (a - s_cnt[threadIdx.x]);  // edge tiles before a; interior tiles before a
  for (uint32_t t = a; t < b; t++) {
    if (tile_border[t]) order[e++] = t;
    else order[in++] = t;
  }
}

// per local particle: the ghost slot of its copy on the left / right neighbour (~0: none), and which tiles have any
__global__ void k_build_rslot(uint32_t ns0, uint32_t ns1, const uint32_t* __restrict__ send_idx, const uint32_t* __restrict__ remote_slot,
                              uint32_t* __restrict__ rslot_l, uint32_t* __restrict__ rslot_r, unsigned char* __restrict__ tile_border) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= ns0 + ns1) return;
  const uint32_t i = send_idx[k];
  (k < ns0 ? rslot_l : rslot_r)[i] = remote_slot[k];
  tile_border[i / ASPH_PAIR_BLOCK] = 1;
}

// owner_slot[ghost] = its index on the owning rank | side << 31 (side 1: the owner is rank + 1); ghosts arrive left part first
__global__ void k_scatter_owner(uint32_t count, uint32_t n_left, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ src, uint32_t* __restrict__ dst) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < count) dst[idx[k]] = src[k] | (k < n_left ? 0u : 0x80000000u);
}

}  // namespace

bool dist_p2p(asph_sim* sim) { return sim->dist && sim->dist->p2p && sim->dist->nranks > 1; }

// Once per step, after the last reallocation the step can cause: every rank publishes the addresses and IPC handles of
// the three arrays its neighbours write into; a neighbour whose array moved is mapped again.  Also swaps the ghost
// slot numbers: my k-th send-list entry lands in the neighbour's recv_idx[k].
static int p2p_refresh(asph_sim* sim) {
  DistState* D = sim->dist;
  if (!D->p2p) return ASPH_OK;
  const int R = D->nranks, r = D->rank;
  cudaStream_t st = sim->stream;
  P2PRecord mine;
  memset(&mine, 0, sizeof mine);
  if (D->mbox_cap != 4u * D->cap_h || !D->mbox.p) {  // cap_h is the same on every rank (ensure_halo_capacity decisions are global)
    D->mbox.release();                               // a round of the partner search mails at most four words per border particle
    CUDA_TRY(D->mbox.ensure(size_t(4) * 4u * D->cap_h));
    D->mbox_cap = 4u * D->cap_h;
  }
  mine.ptr[0] = sim->packA.p; mine.ptr[1] = sim->packP[0].p; mine.ptr[2] = sim->packP[1].p; mine.ptr[3] = D->mbox.p; mine.ctl = D->my_ctl;
  for (int k = 0; k < kPeerFields; k++) CUDA_TRY(cudaIpcGetMemHandle(&mine.h[k], mine.ptr[k]));
  CUDA_TRY(cudaIpcGetMemHandle(&mine.hc, D->my_ctl));
  CUDA_TRY(cudaMemcpyAsync(D->rec_dev.p, &mine, sizeof mine, cudaMemcpyHostToDevice, st));
  NCCL_TRY(nccl().AllGather(D->rec_dev.p, D->rec_dev.p + sizeof(P2PRecord), sizeof(P2PRecord), ncclUint8, D->comm, st));
  CUDA_TRY(cudaMemcpyAsync(D->rec_host, D->rec_dev.p + sizeof(P2PRecord), size_t(R) * sizeof(P2PRecord), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  const P2PRecord* all = reinterpret_cast<const P2PRecord*>(D->rec_host);
  bool ok = true;
  for (int q = 0; q < R; q++) {  // every rank's PeerCtl, once
    if (q == r) { D->peer_ctl[q] = D->my_ctl; continue; }
    if (!D->peer_ctl[q]) {
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[q].hc, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; continue; }
      D->peer_ctl[q] = static_cast<PeerCtl*>(p);
    }
  }
  for (int side = 0; side < 2; side++) {
    const int q = side ? r + 1 : r - 1;
    if (q < 0 || q >= R) continue;
    for (int k = 0; k < kPeerFields; k++) {
      if (D->mapped_src[side][k] == all[q].ptr[k] && D->peer_field[side][k]) continue;
      if (D->peer_field[side][k]) { cudaIpcCloseMemHandle(D->peer_field[side][k]); D->peer_field[side][k] = nullptr; }
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[q].h[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; continue; }
      D->peer_field[side][k] = static_cast<float4*>(p);
      D->mapped_src[side][k] = all[q].ptr[k];
    }
  }
  if (ok && !D->peer_ctl_ready) {
    CUDA_TRY(cudaMemcpyAsync(D->peer_ctl_dev.p, D->peer_ctl, size_t(R) * sizeof(PeerCtl*), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    D->peer_ctl_ready = true;
  }
  if (!ok) {  // this rank cannot reach a peer: every rank must take the same path, so agree through the error-free NCCL route
    sim->last_error = "cudaIpcOpenMemHandle failed (no peer access between the GPUs of this job?)";
    return ASPH_ERR_CUDA;
  }
  // ghost slot numbers
  const uint32_t ns = D->n_send[0] + D->n_send[1];
  CUDA_TRY(D->remote_slot.ensure(std::max<size_t>(size_t(D->cap_h) * 2, ns)));
  if (ns | (D->n_recv[0] + D->n_recv[1])) {
    const Nccl& N = nccl();
    NCCL_TRY(N.GroupStart());
    if (D->n_recv[0]) NCCL_TRY(N.Send(D->recv_idx.p, D->n_recv[0], ncclUint32, r - 1, D->comm, st));
    if (D->n_recv[1]) NCCL_TRY(N.Send(D->recv_idx.p + D->n_recv[0], D->n_recv[1], ncclUint32, r + 1, D->comm, st));
    if (D->n_send[0]) NCCL_TRY(N.Recv(D->remote_slot.p, D->n_send[0], ncclUint32, r - 1, D->comm, st));
    if (D->n_send[1]) NCCL_TRY(N.Recv(D->remote_slot.p + D->n_send[0], D->n_send[1], ncclUint32, r + 1, D->comm, st));
    NCCL_TRY(N.GroupEnd());
  }
  // ... and the other way round: the k-th ghost I received is particle send_idx[k] on its owner
  const uint32_t nr = D->n_recv[0] + D->n_recv[1];
  CUDA_TRY(D->owner_tmp.ensure(std::max<size_t>(size_t(D->cap_h) * 2, nr)));
  CUDA_TRY(D->owner_slot.ensure(sim->cap));
  if (ns | nr) {
    const Nccl& N = nccl();
    NCCL_TRY(N.GroupStart());
    if (D->n_send[0]) NCCL_TRY(N.Send(D->send_idx.p, D->n_send[0], ncclUint32, r - 1, D->comm, st));
    if (D->n_send[1]) NCCL_TRY(N.Send(D->send_idx.p + D->n_send[0], D->n_send[1], ncclUint32, r + 1, D->comm, st));
    if (D->n_recv[0]) NCCL_TRY(N.Recv(D->owner_tmp.p, D->n_recv[0], ncclUint32, r - 1, D->comm, st));
    if (D->n_recv[1]) NCCL_TRY(N.Recv(D->owner_tmp.p + D->n_recv[0], D->n_recv[1], ncclUint32, r + 1, D->comm, st));
    NCCL_TRY(N.GroupEnd());
  }
  CUDA_TRY(cudaMemsetAsync(D->owner_slot.p, 0xFF, size_t(sim->n) * sizeof(uint32_t), st));
  if (nr) {
    k_scatter_owner<<<(nr + kThreads - 1) / kThreads, kThreads, 0, st>>>(nr, D->n_recv[0], D->recv_idx.p, D->owner_tmp.p, D->owner_slot.p);
    LAUNCH_CHECK();
  }
  const size_t tiles = (size_t(sim->cap) + ASPH_PAIR_BLOCK - 1) / ASPH_PAIR_BLOCK + 1;
  CUDA_TRY(D->rslot[0].ensure(sim->cap)); CUDA_TRY(D->rslot[1].ensure(sim->cap)); CUDA_TRY(D->tile_border.ensure(tiles));
  CUDA_TRY(cudaMemsetAsync(D->rslot[0].p, 0xFF, size_t(sim->n) * sizeof(uint32_t), st));
  CUDA_TRY(cudaMemsetAsync(D->rslot[1].p, 0xFF, size_t(sim->n) * sizeof(uint32_t), st));
  CUDA_TRY(cudaMemsetAsync(D->tile_border.p, 0, tiles, st));
  if (ns) {
    k_build_rslot<<<(ns + kThreads - 1) / kThreads, kThreads, 0, st>>>(D->n_send[0], D->n_send[1], D->send_idx.p, D->remote_slot.p, D->rslot[0].p,
                                                                     D->rslot[1].p, D->tile_border.p);
    LAUNCH_CHECK();
  }
  const uint32_t ntiles = (sim->n + ASPH_PAIR_BLOCK - 1) / ASPH_PAIR_BLOCK;
  CUDA_TRY(D->tile_order.ensure(tiles + 1));
  k_tile_order<<<1, 1024, 0, st>>>(ntiles, D->tile_border.p, D->tile_order.p);
  LAUNCH_CHECK();
  return ASPH_OK;
}

PeerArgs dist_peer_args(asph_sim* sim, bool wait_halo, bool wait_stats, int field, bool with_stats) {
  PeerArgs a;
  memset(&a, 0, sizeof a);
  a.nranks = 1;
  if (!dist_p2p(sim)) return a;
  DistState* D = sim->dist;
  a.self = D->my_ctl; a.rank = D->rank; a.nranks = D->nranks;
  a.halo_seq = wait_halo ? D->halo_seq : 0u;
  a.stats_seq = wait_stats ? D->stats_seq : 0u;
  if (field >= 0) {
    for (int side = 0; side < 2; side++) {
      const int q = side ? D->rank + 1 : D->rank - 1;
      const bool has = q >= 0 && q < D->nranks;
      a.dst[side] = has ? D->peer_field[side][field] : nullptr;
      a.rslot[side] = D->rslot[side].p;
      a.nb_ctl[side] = has ? D->peer_ctl[q] : nullptr;
    }
    a.tile_border = D->tile_border.p;
    // Edge tiles first + early sequence numbers (sim.cuh) is implemented and parity-tested, but on 2 x B200 it measured
    // no faster than the natural tile order with the numbers sent by the block that finishes the pass last (6.42 vs
    // 6.16 ms per 2 M-particle step), so it is opt-in until the update pass's remaining ~15 us of peer overhead is understood.
    static const bool edge_first = getenv("ASPH_P2P_EDGE_FIRST") != nullptr;
    a.tile_order = edge_first ? D->tile_order.p : nullptr;
    a.edge_done = D->blocks_done.p + 1;
    a.all_ctl = reinterpret_cast<PeerCtl* const*>(D->peer_ctl_dev.p);
    a.halo_seq_out = ++D->halo_seq;
    a.stats_seq_out = with_stats ? ++D->stats_seq : 0u;
    a.blocks_done = D->blocks_done.p;
    D->halo_calls++;
    D->halo_bytes += uint64_t(D->n_send[0] + D->n_send[1]) * 16;
  }
  return a;
}

CoopPeer dist_coop_peer(asph_sim* sim) {
  CoopPeer c;
  memset(&c, 0, sizeof c);
  c.nranks = 1;
  if (!dist_p2p(sim)) return c;
  DistState* D = sim->dist;
  c.self = D->my_ctl; c.all_ctl = reinterpret_cast<PeerCtl* const*>(D->peer_ctl_dev.p);
  c.rank = D->rank; c.nranks = D->nranks; c.seq0 = D->coop_seq;
  c.mbox = D->mbox.p; c.mbox_cap = D->mbox_cap;
  for (int side = 0; side < 2; side++) {
    const int q = side ? D->rank + 1 : D->rank - 1;
    const bool has = q >= 0 && q < D->nranks;
    c.nb_ctl[side] = has ? D->peer_ctl[q] : nullptr;
    c.nb_mbox[side] = has ? reinterpret_cast<uint2*>(D->peer_field[side][3]) : nullptr;
    c.rslot[side] = D->rslot[side].p;
  }
  c.owner_slot = D->owner_slot.p;
  c.verdict = D->verdict.p;
  return c;
}
void dist_coop_advance(asph_sim* sim, unsigned int barriers) { if (sim->dist) sim->dist->coop_seq += barriers; }
int dist_halo_words(asph_sim* sim, void* field) { return halo_impl(sim, field, field, nullptr, 4); }
const uint32_t* dist_ghost_index(asph_sim* sim, uint32_t* count) {
  DistState* D = sim->dist;
  *count = D ? D->n_recv[0] + D->n_recv[1] : 0u;
  return D ? D->recv_idx.p : nullptr;
}

int dist_ranks(asph_sim* sim) { return sim->dist ? sim->dist->nranks : 1; }
uint64_t dist_n_global(asph_sim* sim) { return sim->dist ? sim->dist->n_global : sim->n; }
void dist_set_n_global(asph_sim* sim, uint64_t n) { if (sim->dist) sim->dist->n_global = n; }

int dist_allreduce_host(asph_sim* sim, void* host_values, int count, int is_double) {
  DistState* D = sim->dist;
  if (!D || D->nranks == 1) return ASPH_OK;
  CUDA_TRY(D->red.ensure(size_t(std::max(count, 8))));
  cudaStream_t st = sim->stream;
  CUDA_TRY(cudaMemcpyAsync(D->red.p, host_values, size_t(count) * 8, cudaMemcpyHostToDevice, st));
  NCCL_TRY(nccl().AllReduce(D->red.p, D->red.p, size_t(count), is_double ? ncclDouble : ncclUint64, ncclSum, D->comm, st));
  CUDA_TRY(cudaMemcpyAsync(host_values, D->red.p, size_t(count) * 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return ASPH_OK;
}

int dist_allgather_list(asph_sim* sim, const uint32_t* list, uint32_t n_entries, int words, const uint32_t** gathered, const uint32_t** counts,
                        uint32_t* stride, unsigned long long* total) {
  DistState* D = sim->dist;
  const int R = D->nranks;
  cudaStream_t st = sim->stream;
  CUDA_TRY(D->ag_counts.ensure(size_t(R) + 1));
  CUDA_TRY(cudaMemcpyAsync(D->ag_counts.p + R, &n_entries, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  NCCL_TRY(nccl().AllGather(D->ag_counts.p + R, D->ag_counts.p, 1, ncclUint32, D->comm, st));
  std::vector<uint32_t> cnt_h(static_cast<size_t>(R));
  CUDA_TRY(cudaMemcpyAsync(cnt_h.data(), D->ag_counts.p, size_t(R) * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  uint32_t mx = 0;
  unsigned long long tot = 0;
  for (int q = 0; q < R; q++) { mx = std::max(mx, cnt_h[static_cast<size_t>(q)]); tot += cnt_h[static_cast<size_t>(q)]; }
  *total = tot; *counts = D->ag_counts.p; *stride = mx * uint32_t(words); *gathered = nullptr;
  if (tot == 0) return ASPH_OK;
  const size_t per = size_t(mx) * size_t(words);
  CUDA_TRY(D->ag_send.ensure(per)); CUDA_TRY(D->ag_recv.ensure(per * size_t(R)));
  if (n_entries) CUDA_TRY(cudaMemcpyAsync(D->ag_send.p, list, size_t(n_entries) * words * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  NCCL_TRY(nccl().AllGather(D->ag_send.p, D->ag_recv.p, per, ncclUint32, D->comm, st));
  *gathered = D->ag_recv.p;
  return ASPH_OK;
}

int dist_ref_buffers(asph_sim* sim, size_t n, uint32_t** a, uint32_t** b) {
  DistState* D = sim->dist;
  CUDA_TRY(D->ref_a.ensure(n + n / 8 + 1024)); CUDA_TRY(D->ref_b.ensure(n + n / 8 + 1024));
  *a = D->ref_a.p; *b = D->ref_b.p;
  return ASPH_OK;
}

int dist_local_map(asph_sim* sim) {
  DistState* D = sim->dist;
  if (D->map_valid) return ASPH_OK;
  const uint32_t n = sim->n;
  if (n) {
    cudaStream_t st = sim->stream;
    const uint32_t blocks = (n + kThreads - 1) / kThreads;
    const uint32_t* refid = sim->refid[sim->cur].p;
    k_owned_flag<<<blocks, kThreads, 0, st>>>(n, refid, sim->scratch_u[2].p);
    LAUNCH_CHECK();
    CUDA_TRY(cudaMemcpyAsync(D->words.p, &n, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    TRY(launch_exclusive_scan(sim, sim->scratch_u[2].p, sim->scratch_u[3].p, D->words.p, 0, n));
    k_mask_ghost_slots<<<blocks, kThreads, 0, st>>>(n, refid, sim->scratch_u[3].p);
    LAUNCH_CHECK();
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  D->map_valid = true;
  return ASPH_OK;
}

void dist_destroy(asph_sim* sim) {
  DistState* D = sim->dist;
  if (!D) return;
  if (D->comm && nccl().CommDestroy) nccl().CommDestroy(D->comm);
  D->send_idx.release(); D->recv_idx.release(); D->slot[0].release(); D->slot[1].release();
  D->sendbuf.release(); D->recvbuf.release(); D->words.release(); D->gather.release(); D->hist.release(); D->flag_bits.release();
  if (D->gather_host) cudaFreeHost(D->gather_host);
  for (int side = 0; side < 2; side++)
    for (int k = 0; k < 4; k++) if (D->peer_field[side][k]) cudaIpcCloseMemHandle(D->peer_field[side][k]);
  for (int q = 0; q < D->nranks && q < ASPH_MAX_RANKS; q++) if (q != D->rank && D->peer_ctl[q]) cudaIpcCloseMemHandle(D->peer_ctl[q]);
  if (D->my_ctl) cudaFree(D->my_ctl);
  if (D->rec_host) cudaFreeHost(D->rec_host);
  D->ag_send.release(); D->ag_recv.release(); D->ag_counts.release(); D->ref_a.release(); D->ref_b.release(); D->red.release(); D->mbox.release(); D->owner_slot.release(); D->owner_tmp.release(); D->verdict.release(); D->remote_slot.release(); D->rec_dev.release(); D->peer_ctl_dev.release(); D->rslot[0].release(); D->rslot[1].release(); D->blocks_done.release(); D->tile_border.release();
  delete D;
  sim->dist = nullptr;
}

extern "C" {

int asph_comm_unique_id(uint8_t id_out[128]) {
  if (!id_out) return ASPH_ERR_INVALID;
  if (!nccl().ok) return ASPH_ERR_NCCL;
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  if (nccl().GetUniqueId(&id) != ncclSuccess) return ASPH_ERR_NCCL;
  memcpy(id_out, &id, 128);
  return ASPH_OK;
}

int asph_create_distributed(const asph_params* params, const float* pos, const float* vel, const float* mass, const uint32_t* global_index,
                            uint64_t n_local, uint64_t n_global, const asph_boundary* boundary, const asph_split_patterns* split,
                            int counters_enabled, uint64_t capacity, const uint8_t nccl_id[128], int rank, int n_ranks, int device,
                            asph_sim** out) {
  if (!out) return ASPH_ERR_INVALID;
  *out = nullptr;
  if (!nccl_id || rank < 0 || n_ranks < 1 || rank >= n_ranks || n_global >= 0x7FFFFFF0ull || (n_local && !global_index)) return ASPH_ERR_INVALID;
  if (!nccl().ok) return ASPH_ERR_NCCL;
  if (device >= 0) {
    char buf[32];
    snprintf(buf, sizeof buf, "%d", device);
    setenv("ASPH_DEVICE", buf, 1);
  }
  uint64_t cap = capacity ? capacity : (n_local + n_local / 2 + 65536);
  asph_sim* sim = nullptr;
  int rc = asph_create(params, pos, vel, mass, n_local, boundary, split, counters_enabled, cap, &sim);
  if (rc != ASPH_OK) return rc;
  DistState* D = new DistState();
  sim->dist = D;
  D->rank = rank; D->nranks = n_ranks; D->n_global = n_global;
  auto fail = [&](int code) { asph_destroy(sim); return code; };
  if (n_local) {  // refid = global particle index
    if (cudaMemcpy(sim->refid[sim->cur].p, global_index, n_local * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess) return fail(ASPH_ERR_CUDA);
  }
  ncclUniqueId id;
  memcpy(&id, nccl_id, 128);
  if (nccl().CommInitRank(&D->comm, n_ranks, id, rank) != ncclSuccess) { D->comm = nullptr; return fail(ASPH_ERR_NCCL); }
  if (D->words.ensure(W_COUNT) != cudaSuccess || D->flag_bits.ensure(32) != cudaSuccess || D->gather.ensure(size_t(n_ranks) * W_COUNT) != cudaSuccess ||
      D->hist.ensure(kHistBins) != cudaSuccess ||
      cudaMallocHost((void**)&D->gather_host, std::max<size_t>(size_t(n_ranks) * W_COUNT, 16) * sizeof(uint32_t)) != cudaSuccess)
    return fail(ASPH_ERR_CUDA);
  memset(D->gather_host, 0, std::max<size_t>(size_t(n_ranks) * W_COUNT, 16) * sizeof(uint32_t));
  if (ensure_halo_capacity(sim, uint32_t(std::max<uint64_t>(1u << 16, cap / 8))) != ASPH_OK) return fail(ASPH_ERR_CUDA);
  {  // peer-memory path of the sweeps (ASPH_DIST_P2P=0 keeps them on NCCL send / recv + all-reduce)
    const char* e = getenv("ASPH_DIST_P2P");
    D->p2p = n_ranks > 1 && n_ranks <= ASPH_MAX_RANKS && !(e && e[0] == '0');
    if (D->p2p) {
      if (cudaMalloc((void**)&D->my_ctl, sizeof(PeerCtl)) != cudaSuccess || cudaMemset(D->my_ctl, 0, sizeof(PeerCtl)) != cudaSuccess ||
          D->rec_dev.ensure(size_t(n_ranks + 1) * sizeof(P2PRecord)) != cudaSuccess || D->peer_ctl_dev.ensure(ASPH_MAX_RANKS) != cudaSuccess ||
          D->blocks_done.ensure(4) != cudaSuccess || cudaMemset(D->blocks_done.p, 0, 16) != cudaSuccess || D->verdict.ensure(4) != cudaSuccess ||
          cudaMallocHost((void**)&D->rec_host, size_t(n_ranks) * sizeof(P2PRecord)) != cudaSuccess)
        return fail(ASPH_ERR_CUDA);
    }
  }
  *out = sim;
  return ASPH_OK;
}

int asph_get_global_index(asph_sim* sim, uint32_t* dst, uint64_t cap) {
  if (!sim || !dst) return ASPH_ERR_INVALID;
  const uint32_t no = sim->dist ? sim->n_owned : sim->n;
  if (cap < no) return ASPH_ERR_INVALID;
  if (no == 0) return ASPH_OK;
  CUDA_TRY(cudaSetDevice(sim->device));
  const uint32_t n = sim->n;
  uint32_t* tmp = reinterpret_cast<uint32_t*>(sim->scratch_f.p);
  if (sim->dist) {
    TRY(dist_local_map(sim));
    k_global_index<<<(n + kThreads - 1) / kThreads, kThreads, 0, sim->stream>>>(n, sim->refid[sim->cur].p, sim->scratch_u[3].p, tmp);
    LAUNCH_CHECK();
    CUDA_TRY(cudaMemcpyAsync(dst, tmp, size_t(no) * sizeof(uint32_t), cudaMemcpyDeviceToHost, sim->stream));
    CUDA_TRY(cudaStreamSynchronize(sim->stream));
  } else {
    for (uint32_t i = 0; i < no; i++) dst[i] = i;  // read-backs of a single-GPU handle are in reference order already
  }
  return ASPH_OK;
}

}  // extern "C"
