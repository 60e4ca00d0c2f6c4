// sim.cuh — internal declarations of libasph_b200.so (hand-written sm_100a CUDA; no tensor cores: the
// path has no dense contraction).  Device state layout, device math, launcher prototypes.
//
// Data layout in HBM (all SoA, fp32; particles are physically re-sorted by (size level, grid cell) every step):
//   persistent  pos float2 | vel float2 | mass f32 | refid u32 (index in the reference's ParticleVec) | level f32
//   per step    xyhm float4 {x, y, h, m}   pre-advection snapshot, one 16 B gather per neighbour candidate
//               neighbour lists: sliced ELL of 16-bit relative indices (lists.cuh); no stored pair coefficients
//   solver      packP float4 {x, y, p/rho^2, p} | packA float4 {x, y, a^p_x, a^p_y} | pconst float4 {Gs.x, Gs.y, a_ii, s}
//               xv float4 {x, y, vx, vy} | hm float2 {h, m} (adaptive h only)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/asph.h"

#define ASPH_MAX_LEVELS 12
#define ASPH_POLY_DEV ASPH_MAX_POLY_VERTS  // polygon vertices the device code takes = what the header allows (a box has 4)
#define ASPH_RETRY_LISTS 1000  // internal: the neighbour pool was too small; grow it and redo the step from the neighbour pass
#define ASPH_SLACK 1.00390625f  // 1 + 1/256: cell / search-radius safety factor against fp32 binning error

// ------------------------------------------------------------------------------------------------
// device-visible control block: everything data-dependent that decides control flow lives here, so the host
// synchronises only a handful of times per step (after the neighbour build, once per solver batch).
// ------------------------------------------------------------------------------------------------
enum {
  ERRF_NONFINITE = 1, ERRF_NEG_AII = 2, ERRF_DENSITY = 4, ERRF_NEIGHBOR_OVERFLOW = 8, ERRF_LIST_CAPACITY = 16,
  ERRF_PARTICLE_CAPACITY = 32, ERRF_SPLIT_PATTERN = 64, ERRF_LEVEL_WEIGHT = 128, ERRF_CELL_BUDGET = 256,
  ERRF_SPLIT_CHILDREN = 512, ERRF_SOLVER_NONFINITE = 1024, ERRF_PARTNER_VALIDATION = 2048, ERRF_PEER_TIMEOUT = 4096,
  ERRF_CONSTRAIN = 8192
};

// Cells of a level are numbered STRIP-MAJOR: the grid is cut into vertical strips of 2^strip_log2 columns; inside a
// strip cells run row by row.  Particles are sorted by cell number, so the particles of the 3 x 3 cell block around a
// cell lie within about one strip row of each other in memory, apart from the cells across a strip edge.  The pair
// passes stage that window in shared memory (solver.cu); k_make_levels picks the strip width so that one strip row
// (plus the diagonal cell) fits the window halo at rest density.
struct GridLevel {
  float cell, inv_cell, hmax;
  int nx, ny;
  int strip_log2;
  uint32_t base;  // first cell of this level in the concatenated cell array
};
__host__ __device__ __forceinline__ uint32_t cell_index(const GridLevel& g, int cx, int cy) {
  const uint32_t st = uint32_t(cx) >> g.strip_log2;
  return g.base + ((st * uint32_t(g.ny) + uint32_t(cy)) << g.strip_log2) + (uint32_t(cx) & ((1u << g.strip_log2) - 1u));
}
__host__ __device__ __forceinline__ unsigned long long level_cells(int nx, int ny, int strip_log2) {
  return (unsigned long long)((nx + (1 << strip_log2) - 1) >> strip_log2) * (unsigned long long)ny << strip_log2;
}

// Relaxed-Jacobi solver state.  Sweep number s (0, 1, ...) accumulates its PressureSolverStatistics
// (simulation.rs:397-469) into slot s % 3 with integer atomics only — counts, and the error sum in 2^-32 fixed point —
// so the totals do not depend on the order in which blocks finish, and (multi-GPU) a plain integer all-reduce of the slot
// gives every rank the same numbers.  The loop control (simulation.rs:1453-1477) is evaluated from the slot by the next
// kernel in the stream: every block of the next pressure-acceleration pass derives the same stop decision in its
// prologue (no ticket, no fence, no extra launch); block 0 records it here for the host.
#define ASPH_ACC_COPIES 8   // each slot is spread over 8 copies (by block index) to keep same-address atomics apart
#define ASPH_ACC_WORDS (2 * ASPH_ACC_COPIES + 1)
struct SolverCtl {
  int k;       // index of the sweep being executed (num_pressure_iters, simulation.rs:1388)
  int done;    // set once a sweep satisfies the stop rule
  int sweeps;  // sweeps executed and evaluated
  unsigned long long normal, singular, negative;  // statistics of the last evaluated sweep
  float err_sum, max_err, avg;
  unsigned int maxerr_enc[3];
  // [slot][2 * copy + 0] = normal | negative << 32; [2 * copy + 1] = sum of err * 2^32 (two's complement); [16] = singular
  unsigned long long acc[3][ASPH_ACC_WORDS];
};

struct StepCtl {
  // order-preserving encodings (see enc_f / dec_f) for atomicMin / atomicMax on floats
  unsigned int hmin_enc, hmax_enc, minx_enc, miny_enc, maxx_enc, maxy_enc, cfl_enc;
  int nlevels;
  float hmin, hmax, origin_x, origin_y;
  GridLevel lv[ASPH_MAX_LEVELS];
  uint32_t total_cells;
  float dt;
  unsigned int error_flags;
  unsigned int list_used;  // sliced-ELL pool units (64 uint16) handed out this step
  uint32_t max_count;                 // largest neighbour count
  SolverCtl solver;
  // level set
  uint32_t front_n[2], cand_n[2];
  int level_live[2];
  int level_sweep, level_done;
  // adaptivity
  uint32_t n_new;  // particle count after merge / split
  uint32_t n_shared, n_merged, n_split_parents;
  uint32_t work_n[2], rounds, n_claims, ready_n, greedy_done;
  uint32_t mail_n[2];                   // several GPUs: messages written into the mailbox of neighbour rank - 1 / rank + 1 since the last barrier
  uint32_t mail_sent, greedy_barriers;  // several GPUs: number of the last barrier mail was posted for; barriers a search ran
  uint32_t greedy_duplicates;             // ready-list entries whose donor had been decided already (diagnostics; expected 0)
  uint32_t validate_why, validate_at[4];  // first broken partner invariant: which, particle, counter, partner, what was found
  uint32_t list_n, local_extra;         // several GPUs: entries in this rank's delete / split list; children this rank appends
  double mass_before, mass_after;
};

struct PackedParams {  // SimulationParams rounded once to fp32 (what serde does for the f32 build)
  float rest_density, cfl_factor, max_dt, viscosity, gravity, jacobi_omega, sdf_gradient_eps;
  float particle_radius_fine, particle_radius_base, maximum_surface_distance, mass_fine, mass_base;
  float max_mass_transfer_sharing, max_share_distance, max_merge_distance;
  float max_avg_density_error_iisph, hybrid_factor, max_avg_density_error, max_avg_divergence_error;
  float f_ext, f_near;  // range factors: level_estimation_range / ETA and 2
  float pull_x, pull_y;
  int has_pull, viscosity_type, level_method, solver, density_source, np_before_div, penalty, sizing, opdisc;
  int boundary_is_fluid_surface, max_iters, constrain;
  int min_share_partners, min_merge_partners, allow_merge_optimal, allow_share_optimal, allow_share_too_small,
      allow_merge_size_diff, fail_on_missing_split_pattern;
  int n_planes;
  float planes[ASPH_MAX_PLANES][3];
  // AnalyticUnderestimate: one polygon SDF (Sdf2D with one connected component, sdf/sdf2d.rs); 0 vertices = planes
  int n_poly;
  float poly_pt[ASPH_POLY_DEV][2], poly_dir[ASPH_POLY_DEV][2], poly_pn[ASPH_POLY_DEV][2];
  // support_length_estimation (ASPH_H_*); level_cut = maximum_range for FromDistribution / FromDistribution2 (the surface
  // detector ignores neighbours beyond particle_radius * maximum_range, simulation.rs:698-723), 0 = no cut
  int h_mode;
  float level_cut;
  // the neighbour pass writes a particle's own W row LAST, so that the sweep kernels can stop one row early and in steps
  // of 4 rows instead of 8 (solver.cu, R4); always 1 (kept as a parameter of the list layout, tests/test_list_layout.py)
  int self_last;
};

template <class T> struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc((void**)&p, n * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct DistState;  // dist.cu

// ---- multi-GPU, inside the sweeps: direct peer-memory exchange over NVLink (dist.cu) -------------------------------
// Every rank owns one PeerCtl in device memory that all other ranks of the node have mapped (CUDA IPC).  A producer
// (k_push / k_stats_push, launched right after the pass that computed the values) stores border values straight into
// the neighbour's particle arrays, or its sweep statistics into every rank's stats_in[], fences, and then publishes a
// sequence number in the consumer's PeerCtl; the consumer's next sweep kernel spins on that number in its prologue.
// Sequence numbers only grow; all ranks launch the same sequence of passes, so rank r waits for the value its own
// launch counter has.
#define ASPH_MAX_RANKS 16
struct PeerCtl {
  unsigned int halo_flag[2];                 // [0] written by rank - 1, [1] by rank + 1
  unsigned int stats_flag[ASPH_MAX_RANKS];   // [r] written by rank r
  unsigned long long stats_in[ASPH_MAX_RANKS][3][ASPH_ACC_WORDS];  // rank r's SolverCtl::acc[slot] of its own particles
  // the persistent cooperative kernels (level.cu, adapt.cu): barriers across all GPUs with a 4-bit payload per rank, and
  // mailboxes the two neighbour ranks append to; both double buffered by the parity of the barrier number
  unsigned int coop_flag[2][ASPH_MAX_RANKS];  // [parity][from rank] = barrier number << 4 | payload
  unsigned int mbox_n[2][2];                  // [parity][from side (0 = rank - 1)] entries waiting in this rank's mailbox
};
// What a persistent cooperative kernel needs to talk to the other GPUs (by value; self == nullptr: single GPU).
// A message is a uint2 {ghost slot on the receiving rank, payload}: a border particle's new value travels to its ghost
// copy as one posted 8-byte store into the neighbour's mailbox over NVLink.  The sender counts its messages in its own
// memory (StepCtl::mail_n) and hands the count over with the barrier (coop_publish_mail): nothing waits for a round trip.
struct CoopPeer {
  PeerCtl* self;
  PeerCtl* const* all_ctl;   // every rank's PeerCtl (device array)
  PeerCtl* nb_ctl[2];
  int rank, nranks;
  unsigned int seq0;         // barriers completed before this launch (same on every rank)
  uint2* mbox;               // this rank's mailboxes: entry k of [parity][side] at ((parity * 2 + side) * mbox_cap + k)
  uint2* nb_mbox[2];         // the neighbours' mailboxes (peer mappings)
  uint32_t mbox_cap;
  const uint32_t* rslot[2];  // per local particle: its ghost slot on that neighbour, or ~0
  const uint32_t* owner_slot;  // per local particle: for a ghost, its index on the rank that owns it (else ~0)
  unsigned int* verdict;     // device word (this rank) in which the thread that ran the barrier leaves the combined payload
};
struct PeerArgs {          // by value into the sweep kernels; self == nullptr: single GPU, or the NCCL path
  PeerCtl* self;
  int rank, nranks;
  unsigned int halo_seq;   // halo_flag value that says the ghosts this pass reads have arrived (from both neighbours)
  unsigned int stats_seq;  // stats_flag value that says every rank's totals of the previous sweep have arrived
  // what this pass publishes itself: its border particles' results go straight into the neighbours' arrays as they are
  // computed; the block that finishes last fences and writes the sequence numbers (and, after the update pass, this
  // rank's statistics into every rank's PeerCtl)
  float4* dst[2];               // the pass's output array on rank - 1 / rank + 1 (nullptr: no such neighbour)
  const uint32_t* rslot[2];     // per local particle: its ghost slot on that neighbour, or ~0
  const unsigned char* tile_border;  // per tile: does it hold border particles at all?
  // Tiles with border particles ("edge tiles": the only ones that read ghost values, too) are processed first, and the
  // sequence number goes to the neighbours as soon as the last of them is done — a whole pass ahead of when the
  // neighbour needs it, so the handshake latency hides behind the interior tiles.
  const uint32_t* tile_order;   // permutation of the tiles, edge tiles first; [ntiles] = number of edge tiles
  unsigned int* edge_done;      // counter, returns to 0 in every pass
  PeerCtl* nb_ctl[2];
  PeerCtl* const* all_ctl;      // every rank's PeerCtl (device array)
  unsigned int halo_seq_out, stats_seq_out;  // stats_seq_out == 0: no statistics to publish (acceleration pass)
  unsigned int* blocks_done;    // counter in this rank's memory, returns to 0 at the end of every pass
};
#define ASPH_GHOST_BIT 0x80000000u  // refid of a ghost particle (owned by a neighbouring slab) carries this bit

struct asph_sim {
  int device = 0;
  cudaStream_t stream = nullptr;
  uint32_t n = 0;        // particles on this device (owned + ghosts)
  uint32_t cap = 0;      // particle capacity
  uint32_t cells_budget = 0;
  // persistent, double buffered for the per-step reorder
  DevBuf<float2> pos[2], vel[2];
  DevBuf<float> mass[2], level[2];
  DevBuf<uint32_t> refid[2];
  int cur = 0;
  // per step
  DevBuf<float4> xyhm, xv[2], packP[2], packA, pconst;
  int xv_cur = 0, p_cur = 0;
  DevBuf<float> h_tmp, rho, lam_sum;
  DevBuf<float2> nrm, gB, lam_grad;
  DevBuf<uint32_t> key, cellcount, cellstart, order, scan_sums;
  DevBuf<uint32_t> cnt, cnt_ext, slice_base, far_idx, far_cnt;
  DevBuf<uint16_t> nbpool;  // sliced-ELL neighbour lists (lists.cuh)
  DevBuf<float2> hm;        // {h, m} per particle: second gather of the adaptive-h pair passes
  DevBuf<float2> hv;        // {h, m / rho}: what the divergence passes gather instead under the Winchenbach2020 operator
  // support_length_estimation != FromMass: the two per-particle values that survive a step besides
  // x, v, m — h2_next (the smoothing length of the NEXT step: this step's estimate, or h from the new mass where
  // resampling changed a particle) and the boundary handler's lambda sum of this step (read by the next step's estimate).
  // Double buffered like `mass`; carried through the reorder, the merge compaction and the split.
  DevBuf<float> hnext[2], lamprev[2];
  bool hdist_valid = false;
  // IISPH2 (simulation.rs:2262-2387): the correction factors of the step, and the size classes the LAST resampling phase
  // assigned (particle_size_class outlives the step in the reference; the omega pass reads it), carried like hnext
  DevBuf<float> omega;
  DevBuf<uint8_t> cls[2];
  bool cls_valid = false;
  int predicted_sweeps[2] = {1, 1};  // sweeps of the divergence / density solve in the previous step
  DevBuf<uint8_t> size_class, flags;  // flags: bit0 surface, bit1 insufficient neighbours
  DevBuf<uint32_t> merge_partner, front[2], cand, work[2], scratch_u[4];
  DevBuf<uint32_t> merge_counter;
  DevBuf<int> stamp;
  // partner search (adapt.cu, k_greedy): donor state | |E'| << 8, offered mass, wait lists (head / next), scan resume point
  DevBuf<uint32_t> g_info, g_head, g_next, g_resume;
  DevBuf<float> g_drop;
  uint64_t adapt_rounds = 0;
  uint64_t greedy_duplicates = 0;  // over the lifetime of the handle (diagnostics)
  DevBuf<float> scratch_f;
  DevBuf<float> lut;         // 2 * 10001 floats: λ then λ′
  DevBuf<float> split_pos;   // flattened patterns
  DevBuf<int> split_off;
  DevBuf<float> blockstats;  // per-block partials of the Jacobi reduction
  StepCtl* ctl = nullptr;       // device
  StepCtl* ctl_host = nullptr;  // pinned mirror
  PackedParams* pp_dev = nullptr;
  PackedParams pp;
  int max_children = 0;
  asph_boundary boundary;
  bool lists_valid = false;
  bool level_valid = false;
  bool step_fields_valid = false;
  bool level_cutoff = true;  // stop the level-set propagation once every new value is below -maximum_surface_distance
  bool share_enabled = false, merge_enabled = false, split_enabled = false;
  float last_dt = 0;
  // bookkeeping
  double time = 0;
  float time_f = 0;
  uint64_t step_number = 0;
  asph_step_info info;
  bool counters = false;
  double pc_ms[ASPH_PC_COUNT] = {0};
  uint64_t pc_calls[ASPH_PC_COUNT] = {0};
  void* pc = nullptr;  // PerformanceCounters state (capi.cu: open intervals + event pool)
  int sm_count = 148;
  float cell_scale = 0.f;  // cell edge of a grid level in units of the level's largest search radius; 0 = chosen per step (grid.cu); ASPH_CELL_SCALE overrides
  bool bulk = true;    // bulk-copy (cp.async.bulk + mbarrier) stage fill of the single-GPU sweep kernels; ASPH_BULK=0 at asph_create: per-thread copies
  bool sweep_attr_done = false;  // dynamic shared-memory limit of the sweep kernels raised on this handle's device
  int prop_grid = 0, greedy_grid = 0, greedy_grid_peer = 0;  // co-resident blocks of the persistent cooperative kernels (level.cu, adapt.cu) on this device
  bool ctl_seen = false;  // ctl_host holds a control block read back from the device (possibly of the previous step)
  std::string last_error;
  uint64_t kernel_launches = 0;
  // kernel timing (asph_set_kernel_timing)
  int kt_every = 0;
  double kt_ms[ASPH_KT_COUNT] = {0};
  uint64_t kt_samples[ASPH_KT_COUNT] = {0};
  std::vector<cudaEvent_t> kt_pool;
  // multi-GPU
  DistState* dist = nullptr;
  uint32_t n_owned = 0;  // == n when single GPU
};

#define CUDA_TRY(x)                                                                              \
  do {                                                                                           \
    cudaError_t e__ = (x);                                                                       \
    if (e__ != cudaSuccess) {                                                                    \
      sim->last_error = std::string(#x) + ": " + cudaGetErrorString(e__);                        \
      return ASPH_ERR_CUDA;                                                                      \
    }                                                                                            \
  } while (0)
#define TRY(x)                       \
  do {                               \
    int rc__ = (x);                  \
    if (rc__ != ASPH_OK) return rc__; \
  } while (0)
#define LAUNCH_CHECK()                                                       \
  do {                                                                       \
    sim->kernel_launches++;                                                  \
    cudaError_t e__ = cudaGetLastError();                                    \
    if (e__ != cudaSuccess) {                                                \
      sim->last_error = std::string("kernel launch: ") + cudaGetErrorString(e__); \
      return ASPH_ERR_CUDA;                                                  \
    }                                                                        \
  } while (0)

// ------------------------------------------------------------------------------------------------
// host-side λ tables (plane_lambda.cpp)
// ------------------------------------------------------------------------------------------------
double asph_host_lambda(double d);
double asph_host_dlambda(double d);
void asph_host_build_luts(std::vector<float>& lam, std::vector<float>& dlam);
float asph_host_lut_get(const std::vector<float>& data, float x);

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
// grid.cu
int ensure_capacity(asph_sim* sim, uint32_t n_particles);
int sync_ctl(asph_sim* sim);  // copies the control block to the pinned mirror and synchronises the stream
int launch_sort_and_grid(asph_sim* sim, float f_search);
inline bool h_from_distribution(const asph_sim* sim) { return sim->pp.h_mode != ASPH_H_FROM_MASS; }
int launch_exclusive_scan(asph_sim* sim, const uint32_t* in, uint32_t* out, const uint32_t* n_dev, uint32_t n_add,
                          uint32_t n_max);  // out[i] = sum(in[0..i)) for i < *n_dev + n_add
// neighbors.cu
int launch_neighbors(asph_sim* sim, float f_ext, float f_near, bool lists_only = false);
int launch_constrain_neighborhood(asph_sim* sim);       // constrain_neighborhood_count (simulation.rs:2145-2177): shrinks h, then the lists are rebuilt
inline bool op_w2020(const asph_sim* sim) { return sim->pp.opdisc == ASPH_OP_WINCHENBACH2020; }
int neighbors_grow(asph_sim* sim);
// solver.cu
int launch_viscosity(asph_sim* sim);
int launch_source(asph_sim* sim, int kind);  // 0 divergence, 1 only density, 2 full, 3 full with the IISPH2 omega factors
int launch_omega(asph_sim* sim);             // IISPH2: omega_i (simulation.rs:2263-2311)
int launch_scale_pressure(asph_sim* sim);    // IISPH2: p /= sqrt(omega) after the solve (simulation.rs:2358-2360)
inline bool solver_iisph2(const asph_sim* sim) { return sim->pp.solver == ASPH_SOLVER_IISPH2; }
int launch_solver(asph_sim* sim, bool density_mode, float max_avg_error, int* iters_out, int* sweeps_out, double* avg_out);
int launch_final_accel(asph_sim* sim, int mode);  // 1: v += dt a (xv pack); 2: hybrid integrate; 3: IISPH integrate
// level.cu
int launch_level_estimation(asph_sim* sim);
int launch_level_smoothing(asph_sim* sim);
// adapt.cu
int launch_adaptivity(asph_sim* sim, float dt);
// dist.cu — multi-GPU slab decomposition (all are no-ops / never called when sim->dist == nullptr)
int dist_begin_step(asph_sim* sim, float f_search);        // migrate, exchange ghosts; afterwards sim->n = owned + ghosts
int dist_allreduce_cfl(asph_sim* sim);                     // min over ranks of the CFL term, between k_prepare and k_make_levels
int dist_after_sort(asph_sim* sim);                        // halo index maps in sorted order
int dist_halo(asph_sim* sim, void* field, int elem_bytes); // owner -> ghost copies of one per-particle field
int dist_solver_reduce(asph_sim* sim, int slot);           // sum SolverCtl::acc[slot] over ranks
bool dist_p2p(asph_sim* sim);                              // peer-memory path available (all neighbours mapped)?
// arguments of the next sweep pass: what it waits for, and (field >= 0) where it publishes: packA (0) / packP[0] (1) / packP[1] (2)
PeerArgs dist_peer_args(asph_sim* sim, bool wait_halo, bool wait_stats, int field, bool with_stats);
int dist_reduce_flags(asph_sim* sim, bool with_lists);     // make error flags (and "lists too small") agree on all ranks
CoopPeer dist_coop_peer(asph_sim* sim);                    // arguments of a persistent cooperative kernel (self == nullptr when single GPU)
void dist_coop_advance(asph_sim* sim, unsigned int barriers);  // the launch ran that many cross-GPU barriers
int dist_halo_words(asph_sim* sim, void* field);           // dist_halo of a 4-byte field whatever its type
const uint32_t* dist_ghost_index(asph_sim* sim, uint32_t* count);  // sorted indices of the ghost particles of this step
int dist_ranks(asph_sim* sim);
uint64_t dist_n_global(asph_sim* sim);                     // particles of the whole fluid (the reference index space)
void dist_set_n_global(asph_sim* sim, uint64_t n);
int dist_allreduce_host(asph_sim* sim, void* host_values, int count, int is_double);  // sum over ranks of host-side uint64 / double values
// every rank's list (entries of `words` 32-bit words) on every rank: gathered[q * stride + ...], counts[q] entries of rank q (device
// pointers owned by the distributed state, valid until the next call); *total = entries of all ranks
int dist_allgather_list(asph_sim* sim, const uint32_t* list, uint32_t n_entries, int words, const uint32_t** gathered, const uint32_t** counts,
                        uint32_t* stride, unsigned long long* total);
int dist_ref_buffers(asph_sim* sim, size_t n, uint32_t** a, uint32_t** b);  // two arrays over the reference index space
int dist_local_map(asph_sim* sim);                         // scratch_u[3][i] = slot of owned particle i in read-backs, ~0u for ghosts
void dist_destroy(asph_sim* sim);
// capi.cu
int check_error_flags(asph_sim* sim);
cudaEvent_t kt_event(asph_sim* sim);           // event from the handle's pool
void kt_release(asph_sim* sim, cudaEvent_t e);

// ------------------------------------------------------------------------------------------------
// device math
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
#define ASPH_PI_F 3.14159265358979323846f
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Barrier number `seq` across the persistent kernels of all ranks, called by ONE thread of each after its own grid has
// synchronised and fenced (system scope): every rank leaves seq << 4 | payload with every other rank and waits for
// theirs; returns the OR of all payloads, or 0x80000000 if a rank did not show up within about 2 s (the step then fails
// on the host instead of hanging).  A rank can be at most one barrier ahead of another, hence two flag sets.
// position of the next message to the neighbour on `side` (0 = rank - 1) before barrier number `seq`; ~0 = mailbox full
__device__ __forceinline__ void coop_mail(const CoopPeer& P, StepCtl* ctl, unsigned int seq, int side, uint32_t slot_and_kind, uint32_t payload) {
  const uint32_t par = seq & 1u;
  const uint32_t k = atomicAdd(&ctl->mail_n[side], 1u);
  if (k < P.mbox_cap) P.nb_mbox[side][size_t(par * 2u + uint32_t(1 - side)) * P.mbox_cap + k] = make_uint2(slot_and_kind, payload);  // I am that neighbour's other side
  else atomicOr(&ctl->error_flags, ERRF_PEER_TIMEOUT);
}
// ONE thread, after the grid has synchronised behind the last coop_mail and before coop_barrier(seq): the message counts
// go to the neighbours (ordered before the barrier flag by its release); returns whether anything was mailed
__device__ __forceinline__ bool coop_publish_mail(const CoopPeer& P, StepCtl* ctl, unsigned int seq) {
  const uint32_t par = seq & 1u;
  bool any = false;
  for (int side = 0; side < 2; side++) {
    const uint32_t c = min(*reinterpret_cast<volatile uint32_t*>(&ctl->mail_n[side]), P.mbox_cap);
    if (P.nb_ctl[side]) *reinterpret_cast<volatile unsigned int*>(&P.nb_ctl[side]->mbox_n[par][1 - side]) = c;
    any = any || c != 0u;
    ctl->mail_n[side] = 0u;
  }
  return any;
}
__device__ __forceinline__ unsigned int coop_barrier(const CoopPeer& P, unsigned int seq, unsigned int payload, StepCtl* ctl) {
  const unsigned int par = seq & 1u, word = (seq << 4) | (payload & 15u);
  __threadfence_system();  // ONE system-scope fence (a release store per peer would pay for one each), then plain posted stores
  for (int q = 0; q < P.nranks; q++)
    if (q != P.rank) *reinterpret_cast<volatile unsigned int*>(&P.all_ctl[q]->coop_flag[par][P.rank]) = word;
  unsigned int acc = payload & 15u;
  unsigned long long t0 = 0;
  for (int q = 0; q < P.nranks; q++) {
    if (q == P.rank) continue;
    const volatile unsigned int* f = &P.self->coop_flag[par][q];
    unsigned int w;
    for (unsigned int spins = 0; ((w = *f) >> 4) != (seq & 0x0fffffffu); spins++) {
      if ((spins & 1023u) == 1023u) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 2000000000ull) { atomicOr(&ctl->error_flags, ERRF_PEER_TIMEOUT); return 0x80000000u; }
      }
    }
    acc |= w & 15u;
  }
  (void)ld_acquire_sys(&P.self->coop_flag[par][P.rank == 0 ? 1 : 0]);  // orders the reads of what the peers stored before their flag
  return acc;
}
#define ASPH_FRAC_1_PI_F 0.318309886183790671538f

__device__ __forceinline__ unsigned int enc_f(float f) {
  unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_f(unsigned int e) {
  return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e);
}

// h = ETA * sqrt((m / rho0) / pi), simulation.rs:372-380 — IEEE ops in the reference's order (bit exact: the
// neighbour predicate depends on it)
__device__ __forceinline__ float h_from_mass(float m, float rho0) {
  return __fmul_rn(1.9f, __fsqrt_rn(__fmul_rn(__fdiv_rn(m, rho0), ASPH_FRAC_1_PI_F)));
}
// strict predicate |x_ij|^2 < ((h_i + h_j) * 0.5 * f)^2 without FMA contraction
// (neighborhood_search.rs:141-146, 63-67)
// Boundary SDFs of the semi-analytic handler.  Planes: SdfPlane::probe (sdf/sdf_plane.rs:36-38).  Polygon: Sdf2D::probe
// (sdf/sdf2d.rs:73-141, 179-210) — nearest edge or vertex, positive on the air side; same operation order as the
// reference and the oracle, no contraction.
__device__ __forceinline__ int sdf_count(const PackedParams& P) { return P.n_poly > 0 ? 1 : P.n_planes; }
__device__ __forceinline__ float sdf_probe(const PackedParams& P, int s, float x, float y) {
  if (P.n_poly == 0) {
    const float* pl = P.planes[s];
    return __fadd_rn(__fadd_rn(__fmul_rn(pl[0], x), __fmul_rn(pl[1], y)), pl[2]);
  }
  const int np = P.n_poly;
  float min_dist_sq = __int_as_float(0x7f800000);
  bool is_line = false;
  float line_dist = 0.f, pdx_c = 0.f, pdy_c = 0.f, pdist_sq = 0.f;
  int pidx = 0;
  for (int k = 0; k < np; k++) {
    const int k1 = k + 1 == np ? 0 : k + 1;
    const float lx = __fsub_rn(P.poly_pt[k1][0], P.poly_pt[k][0]), ly = __fsub_rn(P.poly_pt[k1][1], P.poly_pt[k][1]);
    const float len_sq = __fadd_rn(__fmul_rn(lx, lx), __fmul_rn(ly, ly));
    const float dx = P.poly_dir[k][0], dy = P.poly_dir[k][1];
    const float px = __fsub_rn(x, P.poly_pt[k][0]), py = __fsub_rn(y, P.poly_pt[k][1]);
    const float proj = __fadd_rn(__fmul_rn(px, dx), __fmul_rn(py, dy));
    if (proj > 0.f && __fmul_rn(proj, proj) < len_sq) {
      const float dl = __fadd_rn(__fmul_rn(px, -dy), __fmul_rn(py, dx));  // dot with the left normal (-dy, dx)
      const float dl2 = __fmul_rn(dl, dl);
      if (dl2 < min_dist_sq) { is_line = true; line_dist = dl; min_dist_sq = dl2; }
    }
    const float c = __fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py));
    if (c < min_dist_sq) { is_line = false; pidx = k; pdx_c = px; pdy_c = py; pdist_sq = c; min_dist_sq = c; }
  }
  if (is_line) return line_dist;
  const float sg = __fadd_rn(__fmul_rn(P.poly_pn[pidx][0], pdx_c), __fmul_rn(P.poly_pn[pidx][1], pdy_c)) >= 0.f ? 1.f : -1.f;
  return __fmul_rn(__fsqrt_rn(pdist_sq), sg);
}

__device__ __forceinline__ float dist_sq_exact(float dx, float dy) {
  return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}
__device__ __forceinline__ float support_sq_exact(float hi, float hj, float f) {
  float s = __fmul_rn(__fmul_rn(__fadd_rn(hi, hj), 0.5f), f);
  return __fmul_rn(s, s);
}

// cubic spline, sph_kernels.rs:23-71; h = smoothing length (support 2h)
__device__ __forceinline__ float cubic_w(float q) {
  if (q < 0.5f) return 6.f * (q * q * q - q * q) + 1.f;
  if (q < 1.f) { float v = 1.f - q; return 2.f * (v * v * v); }
  return 0.f;
}
__device__ __forceinline__ float cubic_dw(float q) {
  if (q < 0.5f) return 18.f * q * q - 12.f * q;
  if (q < 1.f) { float v = 1.f - q; return -6.f * v * v; }
  return 0.f;
}
__device__ __forceinline__ float kernel_norm(float h) { return 10.f / (7.f * ASPH_PI_F * (h * h)); }
__device__ __forceinline__ float kernel_w(float r, float h) { return kernel_norm(h) * cubic_w(r / (2.f * h)); }
// dW/dr (zero when q <= 1e-5, sph_kernels.rs:64-66); gradW = dwdr * x_ij / r
__device__ __forceinline__ float kernel_dwdr(float r, float h) {
  float q = r / (2.f * h);
  if (q <= 1.0e-5f) return 0.f;
  return kernel_norm(h) * cubic_dw(q) / (2.f * h);
}
__device__ __forceinline__ float radius_to_volume(float r) { return ASPH_PI_F * r * r; }
__device__ __forceinline__ float volume_to_radius(float a) { return sqrtf(a * ASPH_FRAC_1_PI_F); }

// LevelEstimationState::target_mass simulation.rs:213-237
__device__ __forceinline__ float target_mass(float level, const PackedParams& P) {
  float lv = fmaxf(level, -P.maximum_surface_distance);
  float t = lv / -P.maximum_surface_distance;
  if (P.sizing == ASPH_SIZING_MASS) return P.mass_fine * (1.f - t) + P.mass_base * t;
  if (P.sizing == ASPH_SIZING_RADIUS) {
    float tr = P.particle_radius_fine * (1.f - t) + P.particle_radius_base * t;
    return radius_to_volume(tr) * P.rest_density;
  }
  // t.powf(0.5) in the reference: libm's powf is correctly rounded here (CUDA's powf is not), and the correctly rounded
  // square root is what it returns
  float st = __fsqrt_rn(t);
  float tr = P.particle_radius_fine * (1.f - st) + P.particle_radius_base * st;
  return radius_to_volume(tr) * P.rest_density;
}
// classify_particle adaptivity/mod.rs:32-48
__device__ __forceinline__ uint8_t classify_particle(float level, float mass, const PackedParams& P) {
  float mrel = mass / target_mass(level, P);
  if (mrel <= 0.5f) return ASPH_CLASS_TOO_SMALL;
  if (mrel <= 1.f / 1.1f) return ASPH_CLASS_SMALL;
  if (mrel < 1.1f) return ASPH_CLASS_OPTIMAL;
  if (mrel < 2.0f) return ASPH_CLASS_LARGE;
  return ASPH_CLASS_TOO_LARGE;
}
#endif
