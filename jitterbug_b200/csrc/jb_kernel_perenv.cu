// Per-env-model (augmented Jitterbug) instantiation of the fused step kernel.
#define JB_PER_ENV 1
#include "jb_sim_kernel.cuh"
