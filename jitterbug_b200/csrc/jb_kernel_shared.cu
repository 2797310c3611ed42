// Shared-model instantiation of the fused step kernel (model constants ride in the kernel parameter bank).
#define JB_PER_ENV 0
#include "jb_sim_kernel.cuh"
