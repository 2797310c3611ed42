// TEST/TOOL-ONLY: instantiates the sub-step algorithm with a flop-counting scalar to produce the algorithmic FLOP
// model (SURVEY.md section 8d asks for instrumented counts instead of estimates).  Counting rules: + - * / compare = 1
// flop each (an FMA therefore counts 2), sqrt = 1, sincos = 2, abs/min/max = 1, casts and moves = 0.
#include <math.h>
#include <stdint.h>
static long long g_flops = 0;
struct Cnt {
  double v;
  Cnt() : v(0) {}
  Cnt(double x) : v(x) {}
  Cnt(float x) : v(x) {}
  Cnt(int x) : v(x) {}
  explicit operator float() const { return (float)v; }
  explicit operator double() const { return v; }
};
#define BIN(op) \
  inline Cnt operator op(Cnt a, Cnt b) { g_flops++; return Cnt(a.v op b.v); }
BIN(+) BIN(-) BIN(*) BIN(/)
#define CMP(op) inline bool operator op(Cnt a, Cnt b) { g_flops++; return a.v op b.v; }
CMP(<) CMP(>) CMP(<=) CMP(>=) CMP(==) CMP(!=)
inline Cnt operator-(Cnt a) { return Cnt(-a.v); }
inline Cnt& operator+=(Cnt& a, Cnt b) { g_flops++; a.v += b.v; return a; }
inline Cnt& operator-=(Cnt& a, Cnt b) { g_flops++; a.v -= b.v; return a; }
inline Cnt& operator*=(Cnt& a, Cnt b) { g_flops++; a.v *= b.v; return a; }
#define JB_HD inline
#include "../../jitterbug_b200/csrc/jb_model.h"
inline Cnt jb_sqrt(Cnt x) { g_flops++; return Cnt(sqrt(x.v)); }
inline Cnt jb_abs(Cnt x) { g_flops++; return Cnt(fabs(x.v)); }
inline void jb_sincos(Cnt x, Cnt* s, Cnt* c) { g_flops += 2; *s = Cnt(sin(x.v)); *c = Cnt(cos(x.v)); }
#include "ref_thread_per_env.cuh"
template <> struct JbEps<Cnt> {
  static Cnt newton_tol() { return Cnt(2e-6); }
  static Cnt ls_tol() { return Cnt(1e-5); }
  static int newton_iters() { return 16; }
};
struct JbGeomD { JB_GEOM_FIELDS(double) };
struct JbModelD { JB_MODEL_FIELDS(double, JbGeomD) };

extern "C" long long hs_count_flops(const JbModelD* m, double* qpos, double* qvel, double* warm, double ctrl, int nsub,
                                    int* iters_out, int* ncon_out) {
  JbState<Cnt> s;
  for (int i = 0; i < 3; i++) s.pos[i] = qpos[i];
  for (int i = 0; i < 4; i++) s.quat[i] = qpos[3 + i];
  for (int k = 0; k < 4; k++) { s.ql[k][0] = qpos[7 + 2 * k]; s.ql[k][1] = qpos[8 + 2 * k]; }
  s.qm = qpos[15]; s.rev = 0;
  for (int i = 0; i < 3; i++) { s.vlin[i] = qvel[i]; s.w[i] = qvel[3 + i]; }
  for (int k = 0; k < 4; k++) { s.vl[k][0] = qvel[6 + 2 * k]; s.vl[k][1] = qvel[7 + 2 * k]; }
  s.vm = qvel[14];
  for (int i = 0; i < 15; i++) s.warm[i] = warm[i];
  g_flops = 0;
  int iters = 0, ncon = 0, ovf = 0;
  for (int k = 0; k < nsub; k++) jb_substep<Cnt, JbModelD, false>(*m, s, Cnt(ctrl), nullptr, &iters, false, &ncon, &ovf);
  long long f = g_flops;
  for (int i = 0; i < 3; i++) qpos[i] = s.pos[i].v;
  for (int i = 0; i < 4; i++) qpos[3 + i] = s.quat[i].v;
  for (int k = 0; k < 4; k++) { qpos[7 + 2 * k] = s.ql[k][0].v; qpos[8 + 2 * k] = s.ql[k][1].v; }
  qpos[15] = s.qm.v + 6.28318530717958647692 * s.rev;
  for (int i = 0; i < 3; i++) { qvel[i] = s.vlin[i].v; qvel[3 + i] = s.w[i].v; }
  for (int k = 0; k < 4; k++) { qvel[6 + 2 * k] = s.vl[k][0].v; qvel[7 + 2 * k] = s.vl[k][1].v; }
  qvel[14] = s.vm.v;
  for (int i = 0; i < 15; i++) warm[i] = s.warm[i].v;
  *iters_out = iters; *ncon_out = ncon;
  return f;
}
