// TEST-ONLY host instantiation of the GPU sub-step algorithm (jitterbug_b200/csrc/jb_dynamics.cuh) in fp64 and fp32.
// It lets the CPU test-suite compare the *algorithm* the CUDA kernel runs against the oracle without a GPU and
// separates algorithmic parity (fp64, ~1e-10) from fp32 rounding.  Never loaded by the product package.
#include "ref_thread_per_env.cuh"
#include <string.h>

struct JbGeomD { JB_GEOM_FIELDS(double) };
struct JbModelD { JB_MODEL_FIELDS(double, JbGeomD) };

template <typename T, typename M>
static int run(const M* m, T* qpos, T* qvel, T* warm, T ctrl, int nsub, SubstepDiag<T>* diag) {
  JbState<T> s;
  for (int i = 0; i < 3; i++) s.pos[i] = qpos[i];
  for (int i = 0; i < 4; i++) s.quat[i] = qpos[3 + i];
  for (int k = 0; k < 4; k++) { s.ql[k][0] = qpos[7 + 2 * k]; s.ql[k][1] = qpos[8 + 2 * k]; }
  s.qm = qpos[15]; s.rev = 0;
  while (s.qm > 3.14159265358979323846) { s.qm -= 6.28318530717958647692; s.rev++; }
  while (s.qm <= -3.14159265358979323846) { s.qm += 6.28318530717958647692; s.rev--; }
  for (int i = 0; i < 3; i++) { s.vlin[i] = qvel[i]; s.w[i] = qvel[3 + i]; }
  for (int k = 0; k < 4; k++) { s.vl[k][0] = qvel[6 + 2 * k]; s.vl[k][1] = qvel[7 + 2 * k]; }
  s.vm = qvel[14];
  for (int i = 0; i < 15; i++) s.warm[i] = warm[i];
  int iters = 0;
  for (int k = 0; k < nsub; k++) {
    if (diag) jb_substep<T, M, true>(*m, s, ctrl, diag, &iters);
    else jb_substep<T, M, false>(*m, s, ctrl, nullptr, &iters);
  }
  for (int i = 0; i < 3; i++) qpos[i] = s.pos[i];
  for (int i = 0; i < 4; i++) qpos[3 + i] = s.quat[i];
  for (int k = 0; k < 4; k++) { qpos[7 + 2 * k] = s.ql[k][0]; qpos[8 + 2 * k] = s.ql[k][1]; }
  qpos[15] = (T)((double)s.qm + 6.28318530717958647692 * s.rev);
  for (int i = 0; i < 3; i++) { qvel[i] = s.vlin[i]; qvel[3 + i] = s.w[i]; }
  for (int k = 0; k < 4; k++) { qvel[6 + 2 * k] = s.vl[k][0]; qvel[7 + 2 * k] = s.vl[k][1]; }
  qvel[14] = s.vm;
  for (int i = 0; i < 15; i++) warm[i] = s.warm[i];
  return iters;
}

extern "C" {
int hs_substeps_f64(const JbModelD* m, double* qpos, double* qvel, double* warm, double ctrl, int nsub, SubstepDiag<double>* diag) {
  return run<double, JbModelD>(m, qpos, qvel, warm, ctrl, nsub, diag);
}
int hs_substeps_f32(const JbModel* m, float* qpos, float* qvel, float* warm, float ctrl, int nsub, SubstepDiag<float>* diag) {
  return run<float, JbModel>(m, qpos, qvel, warm, ctrl, nsub, diag);
}
int hs_sizeof_model_f32(void) { return (int)sizeof(JbModel); }
int hs_sizeof_model_f64(void) { return (int)sizeof(JbModelD); }
int hs_sizeof_diag_f64(void) { return (int)sizeof(SubstepDiag<double>); }
int hs_sizeof_diag_f32(void) { return (int)sizeof(SubstepDiag<float>); }
}
