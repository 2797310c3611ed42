// Batched on-device model compilation for augmented Jitterbugs (BASELINE config 4).
//
// The reference perturbs the MJCF text and lets MuJoCo recompile one model at a time
// (jitterbug_dmc/augmented_jitterbug.py:95-267); at 1 Mi envs that is 1 Mi host compiles.  Here one thread per env
// (fp64, run once) applies the same perturbation to the parametric part of the model -- leg end points, rotor offset,
// densities, gear -- and redoes the part of MuJoCo's compiler that depends on it: geom masses and inertias, body
// centres of mass and inertias, hinge axes, then mj_setConst (mass matrix at qpos0, its inverse, body_invweight0,
// meaninertia).  The result is written straight into the lane-interleaved packed layout the step kernel stages.
// Host-side twin (parity oracle for the tests): augment_Jitterbug -> mjcf.compile_xml -> model.flatten_model.
#include <stdint.h>
#include <string>

#include "../../include/jbsim.h"
#include "jb_packed.h"
#include "jb_task.cuh"

int jb_set_error(const std::string& msg);   // jb_api.cu

#include "jb_augment.cuh"

namespace {

struct AugArgs {
  const JbAugBase* base;     // device copy
  const double* deltas;      // [n][JB_AUG_NDELTA]
  float* packed;             // [n][JBP_SIZE]
  JbModel* models;           // optional [n] float structs (inspection / parity)
  int n;
};

__global__ void __launch_bounds__(64) jb_augment_compile_kernel(const AugArgs a) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.n) return;
  JbModelD M;
  jba::augment_model(*a.base, a.deltas + (size_t)e * JB_AUG_NDELTA, M);
  jb_pack_model<float, JbModelD>(M, a.packed + (size_t)e * JBP_SIZE);
  if (a.models) jba::to_float(M, a.models[e]);
}

// Perturbations drawn like augment_Jitterbug draws them (one Gaussian per parameter, the rotor's y and z offsets
// clipped to [-0.001, 1]; :218-222), from Philox streams keyed by (seed, global env id) so a shard of the batch gets
// the models it would get in the full batch.
struct SampleArgs { JbAugSpec spec; double* deltas; int n; uint32_t env_id_offset; uint64_t seed; };

__device__ __forceinline__ double normal_from(uint32_t a, uint32_t b) {
  float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f), u2 = jb_u01(b);
  return (double)(sqrtf(-2.0f * logf(u1)) * cosf(6.28318530717958647692f * u2));
}

__global__ void jb_augment_sample_kernel(const SampleArgs a) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.n) return;
  double* d = a.deltas + (size_t)e * JB_AUG_NDELTA;
  double z[JB_AUG_NDELTA + 1];
  for (int i = 0; i < JB_AUG_NDELTA + 1; i += 2) {
    JbPhilox r = jb_philox((uint32_t)e + a.env_id_offset, (uint32_t)(i / 2), 0u, 0x415547u, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    z[i] = normal_from(r.v[0], r.v[1]);
    z[i + 1] = normal_from(r.v[2], r.v[3]);
  }
  const JbAugSpec& s = a.spec;
  d[0] = s.modify_global_density ? s.sd_global_density * z[0] : 0.0;
  d[1] = s.modify_core1 ? s.sd_core1_density * z[1] : 0.0;
  d[2] = s.modify_core2 ? s.sd_core2_density * z[2] : 0.0;
  for (int i = 0; i < 24; i++) d[3 + i] = s.modify_legs ? s.sd_legs[i % 3] * z[3 + i] : 0.0;
  for (int i = 0; i < 3; i++) {
    double v = s.modify_mass ? s.sd_mass_pos[i] * z[27 + i] : 0.0;
    if (i > 0) v = fmin(fmax(v, -0.001), 1.0);
    d[27 + i] = v;
  }
  d[30] = s.modify_gear ? s.sd_gear * z[30] : 0.0;
}

}  // namespace

extern "C" int jb_packed_model_size(void) { return JBP_SIZE; }

extern "C" int jb_augment_sample(const JbAugSpec* spec, uint64_t seed, uint32_t env_id_offset, int n, double* deltas_dev,
                                 void* stream) {
  if (!spec || !deltas_dev) return jb_set_error("jb_augment_sample: null argument");
  if (n <= 0) return 0;
  SampleArgs a{*spec, deltas_dev, n, env_id_offset, seed};
  jb_augment_sample_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return jb_set_error(std::string("jb_augment_sample: ") + cudaGetErrorString(e));
  return 0;
}

extern "C" int jb_augment_compile(const JbAugBase* base_host, const double* deltas_dev, int n, float* packed_dev,
                                  JbModel* models_dev, void* stream) {
  if (!base_host || !deltas_dev || !packed_dev) return jb_set_error("jb_augment_compile: null argument");
  if (n <= 0) return 0;
  JbAugBase* bdev = nullptr;
  cudaError_t e = cudaMalloc(&bdev, sizeof(JbAugBase));
  if (e == cudaSuccess) e = cudaMemcpyAsync(bdev, base_host, sizeof(JbAugBase), cudaMemcpyHostToDevice, (cudaStream_t)stream);
  if (e == cudaSuccess) {
    AugArgs a{bdev, deltas_dev, packed_dev, models_dev, n};
    jb_augment_compile_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(a);
    e = cudaGetLastError();
  }
  // the base is pageable host memory: the copy above has been staged by the time it returns; free after the kernel
  if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
  if (bdev) cudaFree(bdev);
  if (e != cudaSuccess) return jb_set_error(std::string("jb_augment_compile: ") + cudaGetErrorString(e));
  return 0;
}
