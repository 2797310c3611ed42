// TEST-ONLY host build of the on-device model compiler (jitterbug_b200/csrc/jb_augment.cuh): lets the CPU suite check
// the kernel's algorithm against the host XML compile without a GPU.  Never loaded by the product package.
#include "../../jitterbug_b200/csrc/jb_augment.cuh"
#include "../../jitterbug_b200/csrc/jb_packed.h"

extern "C" {
int ha_sizeof_base(void) { return (int)sizeof(JbAugBase); }
void ha_augment(const JbAugBase* base, const double* deltas, int n, float* packed, JbModel* models, JbModelD* models64) {
  for (int e = 0; e < n; e++) {
    JbModelD M;
    jba::augment_model(*base, deltas + (size_t)e * JB_AUG_NDELTA, M);
    jb_pack_model<float, JbModelD>(M, packed + (size_t)e * JBP_SIZE);
    if (models) jba::to_float(M, models[e]);
    if (models64) models64[e] = M;
  }
}
}
