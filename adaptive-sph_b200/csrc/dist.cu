// dist.cu — multi-GPU slab decomposition (placeholder until the halo exchange lands)
#include "sim.cuh"
int dist_step_physics(asph_sim* sim) {
  sim->last_error = "multi-GPU step not built yet";
  return ASPH_ERR_UNSUPPORTED;
}
extern "C" {
int asph_comm_unique_id(uint8_t*) { return ASPH_ERR_UNSUPPORTED; }
int asph_create_distributed(const asph_params*, const float*, const float*, const float*, const uint32_t*, uint64_t, uint64_t,
                            const asph_boundary*, const asph_split_patterns*, int, uint64_t, const uint8_t*, int, int, int,
                            asph_sim**) { return ASPH_ERR_UNSUPPORTED; }
int asph_get_global_index(asph_sim* sim, uint32_t* dst, uint64_t cap) {
  if (!sim || cap < sim->n) return ASPH_ERR_INVALID;
  for (uint32_t i = 0; i < sim->n; i++) dst[i] = i;
  return ASPH_OK;
}
}
