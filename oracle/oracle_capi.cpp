// oracle_capi.cpp — exports include/asph.h on top of sph_oracle.hpp (CPU oracle; TEST INFRASTRUCTURE ONLY).
// Built twice: -DORACLE_FT=float -> liboracle_f32.so, -DORACLE_FT=double -> liboracle_f64.so.
// Nothing under adaptive-sph_b200/ links or loads this.
#include "sph_oracle.hpp"

#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef ORACLE_FT
#define ORACLE_FT float
#endif
typedef ORACLE_FT FT;
using namespace oracle;

struct asph_sim {
  Sim<FT> s;
};

static int fail(asph_sim* sim, const StepError& e) {
  sim->s.last_error = e.msg;
  return e.code;
}

extern "C" {

const char* asph_backend_name(void) { return sizeof(FT) == 4 ? "oracle-f32" : "oracle-f64"; }

int asph_set_state(asph_sim* sim, const float* pos, const float* vel, const float* mass, uint64_t n) {
  Sim<FT>& s = sim->s;
  s.resize_all(n);
  for (uint64_t i = 0; i < n; i++) {
    s.position[i] = V2<FT>(FT(pos[2 * i]), FT(pos[2 * i + 1]));
    s.velocity[i] = V2<FT>(FT(vel[2 * i]), FT(vel[2 * i + 1]));
    s.mass[i] = FT(mass[i]);
    s.level_estimation[i] = Level<FT>();
    s.neighs[i].clear();
    s.lambda[i].clear();
    // FluidSimulation::new (sim.rs:505-520): h2 = 0, h2_next = h from mass at INIT_REST_DENSITY = 1
    s.h2[i] = 0;
    s.h2_next[i] = h_from_mass<FT>(s.mass[i], FT(1));
    s.size_class[i] = ASPH_CLASS_OPTIMAL;
  }
  return ASPH_OK;
}

int asph_create(const asph_params* params, const float* pos, const float* vel, const float* mass, uint64_t n,
                const asph_boundary* boundary, const asph_split_patterns* split, int counters_enabled, uint64_t,
                asph_sim** out) {
  if (!params || !out || (n && (!pos || !vel || !mass))) return ASPH_ERR_INVALID;
  asph_sim* sim = new asph_sim();
  Sim<FT>& s = sim->s;
  s.pc.enabled = counters_enabled != 0;
  if (boundary) s.boundary = *boundary; else { memset(&s.boundary, 0, sizeof(s.boundary)); }
  if (s.boundary.kind == ASPH_BND_POLYGON) s.init_polygon();
  s.init_luts();
  if (split && split->max_children >= 2) {
    s.max_children = split->max_children;
    s.split_offset.assign(split->offset, split->offset + (split->max_children - 1));
    int total = split->offset[split->max_children - 2] + split->max_children;
    s.split_pos.assign(split->pos_xy, split->pos_xy + 2 * size_t(total));
  }
  asph_set_state(sim, pos, vel, mass, n);
  memset(&s.info, 0, sizeof(s.info));
  *out = sim;
  return ASPH_OK;
}

void asph_destroy(asph_sim* sim) { delete sim; }

int asph_step_physics(asph_sim* sim, const asph_params* params, float* dt_out) {
  Params<FT> P(*params);
  StepError e{0, ""};
  FT dt = 0;
  if (!sim->s.step_physics(P, dt, e)) return fail(sim, e);
  if (dt_out) *dt_out = float(dt);
  sim->s.info.dt = float(dt);
  return ASPH_OK;
}
static FT g_last_dt = 0;
int asph_step_adaptivity(asph_sim* sim, const asph_params* params, float dt) {
  Params<FT> P(*params);
  StepError e{0, ""};
  if (!sim->s.step_adaptivity(P, FT(dt), e)) return fail(sim, e);
  return ASPH_OK;
}
int asph_step(asph_sim* sim, const asph_params* params, float* dt_out) {  // sim.rs:1973-1978
  Params<FT> P(*params);
  StepError e{0, ""};
  FT dt = 0;
  if (!sim->s.step_physics(P, dt, e)) return fail(sim, e);
  if (dt_out) *dt_out = float(dt);
  (void)g_last_dt;
  if (!sim->s.step_adaptivity(P, dt, e)) return fail(sim, e);  // dt stays in FT (no float round trip)
  return ASPH_OK;
}

uint64_t asph_num_particles(const asph_sim* sim) { return sim->s.n(); }
double asph_time(const asph_sim* sim) { return sim->s.time; }
uint64_t asph_step_number(const asph_sim* sim) { return sim->s.step_number; }

}  // extern "C"
template <class T> static int put(void* dst, uint64_t bytes, uint64_t count, T gen) {
  typedef decltype(gen(0)) E;
  if (bytes < count * sizeof(E)) return ASPH_ERR_INVALID;
  E* d = (E*)dst;
  for (uint64_t i = 0; i < count; i++) d[i] = gen(i);
  return ASPH_OK;
}

// element type: float for every real field (double when built as f64 AND the _f64 getter is used)
template <class OUT> static int get_field_t(asph_sim* sim, int field, void* dst, uint64_t bytes) {
  Sim<FT>& s = sim->s;
  const uint64_t n = s.n();
  switch (field) {
    case ASPH_F_POSITION: return put(dst, bytes, 2 * n, [&](uint64_t k) { return OUT(k & 1 ? s.position[k / 2].y : s.position[k / 2].x); });
    case ASPH_F_VELOCITY: return put(dst, bytes, 2 * n, [&](uint64_t k) { return OUT(k & 1 ? s.velocity[k / 2].y : s.velocity[k / 2].x); });
    case ASPH_F_PRESSURE_ACCEL: return put(dst, bytes, 2 * n, [&](uint64_t k) { return OUT(k & 1 ? s.pressure_accel[k / 2].y : s.pressure_accel[k / 2].x); });
    case ASPH_F_MASS: return put(dst, bytes, n, [&](uint64_t i) { return OUT(s.mass[i]); });
    case ASPH_F_H: return put(dst, bytes, n, [&](uint64_t i) { return OUT(s.h2[i]); });
    case ASPH_F_DENSITY: return put(dst, bytes, n, [&](uint64_t i) { return OUT(s.density[i]); });
    case ASPH_F_PRESSURE: return put(dst, bytes, n, [&](uint64_t i) { return OUT(s.pressure[i]); });
    case ASPH_F_AII: return put(dst, bytes, n, [&](uint64_t i) { return OUT(s.aii[i]); });
    case ASPH_F_SOURCE_TERM: return put(dst, bytes, n, [&](uint64_t i) { return OUT(s.ppe_source_term[i]); });
    case ASPH_F_DENSITY_ERROR: return put(dst, bytes, n, [&](uint64_t i) { return OUT(s.density_error[i]); });
    case ASPH_F_CONSTANT_FIELD: return put(dst, bytes, n, [&](uint64_t i) { return OUT(s.constant_field[i]); });
    case ASPH_F_LEVEL: return put(dst, bytes, n, [&](uint64_t i) { return s.level_estimation[i].surface ? OUT(s.level_estimation[i].v) : OUT(ASPH_LEVEL_INTERIOR); });
    case ASPH_F_LAMBDA_SUM: return put(dst, bytes, n, [&](uint64_t i) { return OUT(s.lambda_sum(i)); });
    case ASPH_F_LAMBDA_GRAD: return put(dst, bytes, 2 * n, [&](uint64_t k) {
      FT g = 0;
      for (auto& e : s.lambda[k / 2]) g += (k & 1) ? e.grad.y : e.grad.x;
      return OUT(g);
    });
    case ASPH_F_SIZE_CLASS: return put(dst, bytes, n, [&](uint64_t i) { return uint8_t(s.size_class[i]); });
    case ASPH_F_NEIGHBOR_COUNT: return put(dst, bytes, n, [&](uint64_t i) { return uint32_t(s.neighs[i].size()); });
    case ASPH_F_FLAG_SURFACE: return put(dst, bytes, n, [&](uint64_t i) { return uint8_t(s.flag_is_fluid_surface[i]); });
    case ASPH_F_FLAG_INSUFFICIENT: return put(dst, bytes, n, [&](uint64_t i) { return uint8_t(s.flag_insufficient_neighs[i]); });
    case ASPH_F_MERGE_PARTNER: return put(dst, bytes, n, [&](uint64_t i) { return uint32_t(s.merge_partner[i]); });
    case ASPH_F_MERGE_COUNTER: return put(dst, bytes, n, [&](uint64_t i) { return uint16_t(s.merge_counter[i]); });
  }
  return ASPH_ERR_INVALID;
}
extern "C" {
int asph_get_field(asph_sim* sim, int field, void* dst, uint64_t bytes) { return get_field_t<float>(sim, field, dst, bytes); }
// oracle-only extension: real fields in double (used to measure the fp32 noise floor against the f64 build)
int oracle_get_field_f64(asph_sim* sim, int field, void* dst, uint64_t bytes) { return get_field_t<double>(sim, field, dst, bytes); }

int asph_get_neighbors_csr(asph_sim* sim, uint64_t* offsets, uint32_t* idx, uint64_t cap, uint64_t* nnz_out) {
  Sim<FT>& s = sim->s;
  uint64_t nnz = 0;
  for (auto& l : s.neighs) nnz += l.size();
  if (nnz_out) *nnz_out = nnz;
  if (!idx) return ASPH_OK;
  if (cap < nnz) return ASPH_ERR_INVALID;
  uint64_t o = 0;
  for (uint64_t i = 0; i < s.n(); i++) {
    offsets[i] = o;
    std::vector<uint32_t> row = s.neighs[i];
    std::sort(row.begin(), row.end());
    for (uint32_t j : row) idx[o++] = j;
  }
  offsets[s.n()] = o;
  return ASPH_OK;
}

int asph_build_neighbors(asph_sim* sim, const asph_params* params, float f) {
  Sim<FT>& s = sim->s;
  Params<FT> P(*params);
  for (uint64_t i = 0; i < s.n(); i++) s.h2[i] = h_from_mass<FT>(s.mass[i], P.rest_density);
  StepError e{0, ""};
  if (!s.build_neighbors(FT(f), e)) return fail(sim, e);
  return ASPH_OK;
}
// oracle-only: O(N^2) statement of the neighbour predicate (sim.rs:1810-1863)
int oracle_build_neighbors_bruteforce(asph_sim* sim, const asph_params* params, float f) {
  Sim<FT>& s = sim->s;
  Params<FT> P(*params);
  for (uint64_t i = 0; i < s.n(); i++) s.h2[i] = h_from_mass<FT>(s.mass[i], P.rest_density);
  s.build_neighbors_bruteforce(FT(f));
  return ASPH_OK;
}
// oracle-only: operator diagonal through the full operator (check_aii sim.rs:1347-1375)
double oracle_aii_inefficient(asph_sim* sim, const asph_params* params, uint64_t i) {
  Params<FT> P(*params);
  return double(sim->s.aii_inefficient(i, P));
}
// oracle-only: the IISPH2 correction factors of the last step (sim.rs:2263-2311), for the numpy cross-check in tests/
int oracle_get_omega(asph_sim* sim, double* out, uint64_t n) {
  if (n != sim->s.n()) return ASPH_ERR_INVALID;
  for (uint64_t i = 0; i < n; i++) out[i] = double(sim->s.omega[i]);
  return ASPH_OK;
}
// oracle-only: run phases of the adaptivity separately for per-phase parity
int oracle_set_level(asph_sim* sim, const float* level, uint64_t n) {
  Sim<FT>& s = sim->s;
  if (n != s.n()) return ASPH_ERR_INVALID;
  for (uint64_t i = 0; i < n; i++) {
    s.level_estimation[i].surface = !(level[i] > 0.0f);
    s.level_estimation[i].v = s.level_estimation[i].surface ? FT(level[i]) : FT(0);
  }
  return ASPH_OK;
}
void oracle_set_step_number(asph_sim* sim, uint64_t k) { sim->s.step_number = k; }
// the same hooks under the names include/asph.h declares
int asph_set_level(asph_sim* sim, const float* level, uint64_t n) { return oracle_set_level(sim, level, n); }
void asph_set_step_number(asph_sim* sim, uint64_t k) { oracle_set_step_number(sim, k); }
uint64_t asph_adapt_rounds(const asph_sim*) { return 0; }
uint64_t asph_debug_greedy_duplicates(const asph_sim*) { return 0; }
void oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int asph_get_step_info(const asph_sim* sim, asph_step_info* out) { *out = sim->s.info; return ASPH_OK; }
int asph_get_counters(const asph_sim* sim, double ms[ASPH_PC_COUNT], uint64_t calls[ASPH_PC_COUNT]) {
  for (int i = 0; i < ASPH_PC_COUNT; i++) { ms[i] = sim->s.pc.ms[i]; calls[i] = sim->s.pc.calls[i]; }
  return ASPH_OK;
}
const char* asph_last_error(const asph_sim* sim) { return sim->s.last_error.c_str(); }
uint64_t asph_kernel_launches(const asph_sim*) { return 0; }
int asph_set_kernel_timing(asph_sim*, int) { return ASPH_OK; }
int asph_get_kernel_timing(asph_sim*, double ms[ASPH_KT_COUNT], uint64_t n[ASPH_KT_COUNT]) {
  for (int k = 0; k < ASPH_KT_COUNT; k++) { ms[k] = 0; n[k] = 0; }
  return ASPH_OK;
}

float asph_kernel_w(float r, float h) { return float(kernel_w<FT>(FT(r), FT(h))); }
void asph_kernel_grad(float dx, float dy, float h, float* gx, float* gy) {
  V2<FT> g = kernel_derivh<FT>(V2<FT>(FT(dx), FT(dy)), FT(h));
  *gx = float(g.x); *gy = float(g.y);
}
double asph_lambda(double d) { return lambda2(d); }
double asph_dlambda(double d) { return dlambda2(d); }
static Lut<FT>& lut(int which) {
  static Lut<FT> l, dl;
  static bool init = false;
  if (!init) {
    l.init(FT(-1), FT(1), 10000, [](double x) { return lambda2(x); });
    dl.init(FT(-1), FT(1), 10000, [](double x) { return dlambda2(x); });
    init = true;
  }
  return which ? dl : l;
}
float asph_lambda_lut(float d) { return float(lut(0).get(FT(d))); }
float asph_dlambda_lut(float d) { return float(lut(1).get(FT(d))); }
// oracle-only: double-precision kernel helpers for the integral / derivative known-answer tests
double oracle_kernel_w_ft(double r, double h) { return double(kernel_w<FT>(FT(r), FT(h))); }
double oracle_volume_to_radius(double a) { return double(volume_to_radius<FT>(FT(a))); }
double oracle_radius_to_volume(double r) { return double(radius_to_volume<FT>(FT(r))); }
double oracle_target_mass(const asph_params* params, float level) {
  Params<FT> P(*params);
  Level<FT> l; l.surface = true; l.v = FT(level);
  return double(target_mass<FT>(l, P));
}
int oracle_classify(const asph_params* params, float level, float mass) {
  Params<FT> P(*params);
  Level<FT> l; l.surface = true; l.v = FT(level);
  return classify<FT>(l, FT(mass), P);
}

// multi-GPU entry points do not exist on the CPU oracle
int asph_comm_unique_id(uint8_t*) { return ASPH_ERR_UNSUPPORTED; }
int asph_create_distributed(const asph_params*, const float*, const float*, const float*, const uint32_t*, uint64_t,
                            uint64_t, const asph_boundary*, const asph_split_patterns*, int, uint64_t, const uint8_t*,
                            int, int, int, asph_sim**) { return ASPH_ERR_UNSUPPORTED; }
int asph_get_global_index(asph_sim* sim, uint32_t* dst, uint64_t cap) {
  if (cap < sim->s.n()) return ASPH_ERR_INVALID;
  for (uint64_t i = 0; i < sim->s.n(); i++) dst[i] = uint32_t(i);
  return ASPH_OK;
}
}
