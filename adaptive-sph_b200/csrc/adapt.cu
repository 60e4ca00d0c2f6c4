// adapt.cu — share / merge / split resampling (placeholder until the kernels land)
#include "sim.cuh"
int launch_adaptivity(asph_sim* sim, float) {
  sim->last_error = "resampling kernels not built yet";
  return ASPH_ERR_UNSUPPORTED;
}
