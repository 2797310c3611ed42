// TEST/TOOL-ONLY: algorithmic FLOP count of the QUAD sub-step (jitterbug_b200/csrc/jb_quad.cuh, the algorithm the step
// kernel runs) by instantiating it with a counting scalar in the fiber emulation of quadsim.cpp.
// Counting rules: + - * / compare = 1 flop each (an FMA counts 2), sqrt = 1, sincos = 2, abs / min / max = 1, casts and
// moves = 0.  The quad program REPLICATES root / rotor / solver-control arithmetic in its four lanes; the algorithmic
// figure counts replicated work once: per phase of the sub-step (Comm::phase_sync marks) the flops of lane 0 are taken
// for replicated phases and the sum over the four lanes for lane-owned phases (leg kinematics + collision, extra-geom
// collision, the per-contact passes of the Newton trip).  `executed` is the plain sum over the four lanes.
#include <math.h>
#include <stdint.h>
static long long g_flops_lane_phase[4][16];
static int g_cur_lane = 0, g_cur_phase[4] = {0, 0, 0, 0};
#define FL() (g_flops_lane_phase[g_cur_lane][g_cur_phase[g_cur_lane]]++)
struct Cnt {
  double v;
  Cnt() : v(0) {}
  Cnt(double x) : v(x) {}
  Cnt(float x) : v(x) {}
  Cnt(int x) : v(x) {}
  explicit operator float() const { return (float)v; }
  explicit operator double() const { return v; }
  explicit operator int() const { return (int)v; }
};
#define BIN(op) inline Cnt operator op(Cnt a, Cnt b) { FL(); return Cnt(a.v op b.v); }
BIN(+) BIN(-) BIN(*) BIN(/)
#define CMP(op) inline bool operator op(Cnt a, Cnt b) { FL(); return a.v op b.v; }
CMP(<) CMP(>) CMP(<=) CMP(>=) CMP(==) CMP(!=)
inline Cnt operator-(Cnt a) { return Cnt(-a.v); }
inline Cnt& operator+=(Cnt& a, Cnt b) { FL(); a.v += b.v; return a; }
inline Cnt& operator-=(Cnt& a, Cnt b) { FL(); a.v -= b.v; return a; }
inline Cnt& operator*=(Cnt& a, Cnt b) { FL(); a.v *= b.v; return a; }
inline Cnt jb_sqrt(Cnt x) { FL(); return Cnt(sqrt(x.v)); }
inline Cnt jb_abs(Cnt x) { FL(); return Cnt(fabs(x.v)); }
inline void jb_sincos(Cnt x, Cnt* s, Cnt* c) { FL(); FL(); *s = Cnt(sin(x.v)); *c = Cnt(cos(x.v)); }
inline void jb_sincos_fast(Cnt x, Cnt* s, Cnt* c) { jb_sincos(x, s, c); }
#define JBQ_HOST_PHASE_HOOK(lane, p) (g_cur_phase[lane] = (p))
#define JBQ_HOST_LANE_HOOK(l) (g_cur_lane = (l))
#include "../../jitterbug_b200/csrc/jb_math.cuh"
template <> struct JbEps<Cnt> {
  static Cnt newton_tol() { return Cnt(2e-6); }
  static Cnt ls_tol() { return Cnt(1e-5); }
  static int newton_iters() { return 16; }
};
#define JBQ_FLOPCOUNT 1
#include "quadsim.cpp"

// phases (value of the last phase_sync mark passed): 0 leg kinematics + collision (owned), 1 quad sums + root / rotor
// (replicated), 2 extra geoms (owned), 3 smooth force (replicated), 4 Newton entry: warm start, first J a pass, then each
// trip's forces / gradient / Hessian until mark 8 (owned per-contact passes dominate), 8 factor, 9 solve + M s + J s
// (replicated except the J s pass), 10 line search (owned ls passes), 11 move (replicated), 6 Euler (replicated), 7 end
static const int kOwned[16] = {1, 0, 1, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0};

extern "C" void hq_count_flops(const double* P, double* qpos, double* qvel, double* warm, double ctrl, int nsub,
                               long long* algorithmic, long long* executed, int* iters_out, int* ncon_out) {
  const int n = JBP_SIZE;
  std::vector<Cnt> Pc(n), q(16), v(15), w(15);
  for (int i = 0; i < n; i++) Pc[i] = Cnt(P[i]);
  for (int i = 0; i < 16; i++) q[i] = Cnt(qpos[i]);
  for (int i = 0; i < 15; i++) { v[i] = Cnt(qvel[i]); w[i] = Cnt(warm[i]); }
  for (int l = 0; l < 4; l++) for (int p = 0; p < 16; p++) g_flops_lane_phase[l][p] = 0;
  int ncon = 0, ovf = 0;
  int it = run<Cnt>(Pc.data(), q.data(), v.data(), w.data(), Cnt(ctrl), nsub, 0, &ncon, nullptr, nullptr, &ovf);
  long long alg = 0, exe = 0;
  for (int p = 0; p < 16; p++) {
    long long s = 0;
    for (int l = 0; l < 4; l++) s += g_flops_lane_phase[l][p];
    exe += s;
    alg += kOwned[p] ? s : g_flops_lane_phase[0][p];
  }
  for (int i = 0; i < 16; i++) qpos[i] = q[i].v;
  for (int i = 0; i < 15; i++) { qvel[i] = v[i].v; warm[i] = w[i].v; }
  *algorithmic = alg; *executed = exe; *iters_out = it; *ncon_out = ncon;
}

// per-(lane, phase) counts of the last hq_count_flops call (diagnostics: which phases replicate how much)
extern "C" void hq_flops_lane_phase(long long* out64) {
  for (int l = 0; l < 4; l++) for (int p = 0; p < 16; p++) out64[16 * l + p] = g_flops_lane_phase[l][p];
}
