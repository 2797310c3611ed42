// TEST-ONLY host emulation of the quad (4 lanes per env) sub-step of jitterbug_b200/csrc/jb_quad.cuh.
// The SPMD lane program runs unchanged in four ucontext fibers; Comm::sum* is a rendezvous that reproduces the GPU's
// xor-butterfly summation order ((x_l + x_{l^1}) + (x_{l^2} + x_{l^3})), so the host result is what the kernel computes
// (in fp64 for algorithmic parity, in fp32 for rounding behaviour).  Never loaded by the product package.
#define JBQ_DEBUG_COUNT 1
#ifndef JBQ_HOST_PHASE_HOOK
#define JBQ_HOST_PHASE_HOOK(lane, p) ((void)0)      /* quadflops.cpp: attribute counted flops to (lane, phase) */
#define JBQ_HOST_LANE_HOOK(lane) ((void)0)
#endif
#include "../../jitterbug_b200/csrc/jb_quad.cuh"
#include "../../jitterbug_b200/csrc/jb_task.cuh"
#include <string.h>
#include <ucontext.h>
#include <vector>

struct JbGeomD { JB_GEOM_FIELDS(double) };
struct JbModelD { JB_MODEL_FIELDS(double, JbGeomD) };

namespace {
struct Sched {
  ucontext_t main_ctx, fib[4];
  double slot[2][4][32];
  bool done[4];
  std::vector<char> stacks[4];
};
Sched* g_s = nullptr;

template <typename T> struct FiberComm {
  int lane;
  int phase = 0;
  void exchange(T* v, int n) {
    Sched& S = *g_s;
    int ph = phase & 1;
    for (int i = 0; i < n; i++) S.slot[ph][lane][i] = (double)v[i];
    swapcontext(&S.fib[lane], &S.main_ctx);          // everybody has posted when we are resumed
    for (int i = 0; i < n; i++) {
      T x0 = (T)S.slot[ph][lane][i], x1 = (T)S.slot[ph][lane ^ 1][i];
      T x2 = (T)S.slot[ph][lane ^ 2][i], x3 = (T)S.slot[ph][lane ^ 3][i];
      T a = x0 + x1, b = x2 + x3;                     // step 1: xor 1 ; step 2: xor 2
      v[i] = a + b;
    }
    phase++;
  }
  T get(T v, int src) {                               // value held by lane `src` of the quad
    Sched& S = *g_s;
    int ph = phase & 1;
    S.slot[ph][lane][0] = (double)v;
    swapcontext(&S.fib[lane], &S.main_ctx);
    T out = (T)S.slot[ph][src][0];
    phase++;
    return out;
  }
  int iget(int v, int src) { return (int)get((T)v, src); }
  T sum(T v) { exchange(&v, 1); return v; }
  void sum2(T& a, T& b) { T v[2] = {a, b}; exchange(v, 2); a = v[0]; b = v[1]; }
  void sumN(T* v, int n) { exchange(v, n); }
  int isum(int v) { T x = (T)v; exchange(&x, 1); return (int)(x + (T)0.5); }
  bool warp_any(bool p) { return isum(p ? 1 : 0) > 0; }   // the emulated "warp" is one quad
  bool loop_any(bool p) { return warp_any(p); }
  void phase_sync(int p) { JBQ_HOST_PHASE_HOOK(lane, p); }
};

template <typename T> struct Job {
  const T* P; QState<T> st[4]; T ctrl; int nsub; bool fwd; int iters[4], ncon[4], ovf[4];
  int rgeom[4 * JBQ_CAP]; T rdist[4 * JBQ_CAP]; bool want_report;
};
void* g_job = nullptr;

template <typename T> void fiber_main(int lane) {
  Job<T>* J = (Job<T>*)g_job;
  FiberComm<T> cm;
  cm.lane = lane;
  QReport<T> rep = {J->rgeom, J->rdist, 1};
  for (int k = 0; k < J->nsub; k++) {
    if (J->want_report) for (int i = 0; i < JBQ_CAP; i++) J->rgeom[lane * JBQ_CAP + i] = -1;
    jbq_substep<T, FiberComm<T>>(cm, J->P, lane, J->st[lane], J->ctrl, J->fwd, &J->iters[lane], &J->ncon[lane], &J->ovf[lane],
                                 J->want_report ? &rep : nullptr);
  }
  g_s->done[lane] = true;
  swapcontext(&g_s->fib[lane], &g_s->main_ctx);
}
template <typename T> void fiber_entry(int lane) { fiber_main<T>(lane); }

template <typename T> void run_quad(Job<T>& J) {
  Sched S;
  g_s = &S;
  g_job = &J;
  for (int l = 0; l < 4; l++) {
    S.done[l] = false;
    S.stacks[l].resize(1 << 20);
    getcontext(&S.fib[l]);
    S.fib[l].uc_stack.ss_sp = S.stacks[l].data();
    S.fib[l].uc_stack.ss_size = S.stacks[l].size();
    S.fib[l].uc_link = &S.main_ctx;
    makecontext(&S.fib[l], (void (*)())fiber_entry<T>, 1, l);
  }
  for (;;) {
    bool all = true;
    for (int l = 0; l < 4; l++) {
      if (S.done[l]) continue;
      all = false;
      JBQ_HOST_LANE_HOOK(l);
      swapcontext(&S.main_ctx, &S.fib[l]);
    }
    if (all) break;
  }
  g_s = nullptr;
}

template <typename T>
int run(const T* P, T* qpos, T* qvel, T* warm, T ctrl, int nsub, int fwd, int* ncon_out, int* geom_out, T* dist_out, int* ovf_out) {
  Job<T> J;
  J.P = P; J.ctrl = ctrl; J.nsub = fwd ? 1 : nsub; J.fwd = fwd != 0; J.want_report = geom_out != nullptr;
  int rev = 0;
  double qm = (double)qpos[15];
  while (qm > 3.14159265358979323846) { qm -= 6.28318530717958647692; rev++; }
  while (qm <= -3.14159265358979323846) { qm += 6.28318530717958647692; rev--; }
  for (int l = 0; l < 4; l++) {
    QState<T>& s = J.st[l];
    for (int i = 0; i < 3; i++) { s.pos[i] = qpos[i]; s.vlin[i] = qvel[i]; s.w[i] = qvel[3 + i]; }
    for (int i = 0; i < 4; i++) s.quat[i] = qpos[3 + i];
    s.qa = qpos[7 + 2 * l]; s.qb = qpos[8 + 2 * l]; s.va = qvel[6 + 2 * l]; s.vb = qvel[7 + 2 * l];
    s.qm = (T)qm; s.rev = rev; s.vm = qvel[14];
    for (int i = 0; i < 6; i++) s.wr[i] = warm[i];
    s.wla = warm[6 + 2 * l]; s.wlb = warm[7 + 2 * l]; s.wmo = warm[14];
    J.iters[l] = J.ncon[l] = J.ovf[l] = 0;
  }
  run_quad<T>(J);
  const QState<T>& s0 = J.st[0];
  for (int i = 0; i < 3; i++) { qpos[i] = s0.pos[i]; qvel[i] = s0.vlin[i]; qvel[3 + i] = s0.w[i]; }
  for (int i = 0; i < 4; i++) qpos[3 + i] = s0.quat[i];
  qpos[15] = (T)((double)s0.qm + 6.28318530717958647692 * s0.rev); qvel[14] = s0.vm;
  for (int i = 0; i < 6; i++) warm[i] = s0.wr[i];
  warm[14] = s0.wmo;
  for (int l = 0; l < 4; l++) {
    const QState<T>& s = J.st[l];
    qpos[7 + 2 * l] = s.qa; qpos[8 + 2 * l] = s.qb; qvel[6 + 2 * l] = s.va; qvel[7 + 2 * l] = s.vb;
    warm[6 + 2 * l] = s.wla; warm[7 + 2 * l] = s.wlb;
    // replicated state must be bitwise identical in the four lanes
    if (memcmp(s.pos, s0.pos, sizeof(s.pos)) || memcmp(s.quat, s0.quat, sizeof(s.quat)) || memcmp(s.wr, s0.wr, sizeof(s.wr)) ||
        memcmp(s.vlin, s0.vlin, sizeof(s.vlin)) || memcmp(s.w, s0.w, sizeof(s.w)) || s.qm != s0.qm || s.vm != s0.vm)
      return -1;
  }
  if (ncon_out) *ncon_out = J.ncon[0];
  if (ovf_out) *ovf_out = J.ovf[0] + J.ovf[1] + J.ovf[2] + J.ovf[3];
  if (geom_out) {
    int n = 0;
    for (int i = 0; i < 4 * JBQ_CAP; i++) if (J.rgeom[i] >= 0) { geom_out[n] = J.rgeom[i]; dist_out[n] = J.rdist[i]; n++; }
    for (int i = n; i < 4 * JBQ_CAP; i++) geom_out[i] = -1;
  }
  return J.iters[0];
}
}  // namespace

extern "C" {
// host instantiation of the device reset code: out[i] = quat(4), target x, y, qw, qz
void hq_reset_states(const float* P, unsigned long long seed, const unsigned* env_ids, int n, unsigned episode, int task,
                     int random_pose, float* out) {
  for (int i = 0; i < n; i++) {
    QState<float> s;
    JbTarget<float> tg;
    jb_reset_env<float>(P, s, tg, task, random_pose, seed, env_ids[i], episode);
    for (int k = 0; k < 4; k++) out[8 * i + k] = s.quat[k];
    out[8 * i + 4] = tg.x; out[8 * i + 5] = tg.y; out[8 * i + 6] = tg.qw; out[8 * i + 7] = tg.qz;
  }
}
// the reset mapping alone, from given uniforms u[n][6] (double and float instantiations)
void hq_reset_from_uniforms_f32(const float* P, const float* u, int n, int task, int random_pose, float* out) {
  for (int i = 0; i < n; i++) {
    QState<float> s;
    JbTarget<float> tg;
    jb_reset_from_uniforms<float>(P, s, tg, task, random_pose, u + 6 * i);
    for (int k = 0; k < 4; k++) out[8 * i + k] = s.quat[k];
    out[8 * i + 4] = tg.x; out[8 * i + 5] = tg.y; out[8 * i + 6] = tg.qw; out[8 * i + 7] = tg.qz;
  }
}
void hq_reset_from_uniforms_f64(const double* P, const double* u, int n, int task, int random_pose, double* out) {
  for (int i = 0; i < n; i++) {
    QState<double> s;
    JbTarget<double> tg;
    jb_reset_from_uniforms<double>(P, s, tg, task, random_pose, u + 6 * i);
    for (int k = 0; k < 4; k++) out[8 * i + k] = s.quat[k];
    out[8 * i + 4] = tg.x; out[8 * i + 5] = tg.y; out[8 * i + 6] = tg.qw; out[8 * i + 7] = tg.qz;
  }
}
// host instantiation of the device obs/reward and heuristic code for one env
int hq_obs_reward(const float* P, const float* qpos, const float* qvel, int task, const float* tgt, float* obs, float* rew) {
  QState<float> s;
  for (int i = 0; i < 3; i++) { s.pos[i] = qpos[i]; s.vlin[i] = qvel[i]; s.w[i] = qvel[3 + i]; }
  for (int i = 0; i < 4; i++) s.quat[i] = qpos[3 + i];
  s.qm = qpos[15]; s.vm = qvel[14];
  JbTarget<float> tg = {tgt[0], tgt[1], tgt[2], tgt[3]};
  return jb_obs_reward<float>(P, s, task, tg, obs, rew);
}
float hq_heuristic(int task, const float* obs) { return jb_heuristic<float>(task, obs); }
void hq_debug_counts(long long* out) { out[0] = g_jbq_ls_iters; out[1] = g_jbq_newton_trips; out[2] = g_jbq_spread_moves; }
int hq_packed_size(void) { return JBP_SIZE; }
void hq_pack_f64(const JbModelD* m, double* out) { jb_pack_model<double, JbModelD>(*m, out); }
void hq_pack_f32(const JbModel* m, float* out) { jb_pack_model<float, JbModel>(*m, out); }
// returns Newton iterations (>= 0) or -1 if the replicated state diverged between lanes
int hq_substeps_f64(const double* P, double* qpos, double* qvel, double* warm, double ctrl, int nsub, int fwd, int* ncon,
                    int* geom, double* dist, int* ovf) {
  return run<double>(P, qpos, qvel, warm, ctrl, nsub, fwd, ncon, geom, dist, ovf);
}
int hq_substeps_f32(const float* P, float* qpos, float* qvel, float* warm, float ctrl, int nsub, int fwd, int* ncon,
                    int* geom, float* dist, int* ovf) {
  return run<float>(P, qpos, qvel, warm, ctrl, nsub, fwd, ncon, geom, dist, ovf);
}
}
