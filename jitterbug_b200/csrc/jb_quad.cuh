// The physics sub-step for ONE Jitterbug executed by a QUAD of 4 cooperating lanes (lane l owns leg l).
//
// Same mathematics as documented in DESIGN.md section 3 (MuJoCo's mj_step2 ; mj_step1 restated in the root body's
// frame, product-of-exponentials kinematics, arrow-structured L'DL, primal Newton with exact line search), but
// written SPMD: every lane runs this code; data that belongs to one leg (joint state, the 2x2 inertia block and its
// coupling to the root, that leg's contacts) lives only in that leg's lane, data of the root / rotor is replicated in
// all four lanes, and the few sums over legs go through `Comm::sum*` (two __shfl_xor_sync steps on the GPU, a fiber
// rendezvous in the test-only host build).  Butterfly sums are bitwise identical in the 4 lanes, so replicated
// control flow (Newton iterations, line search) never diverges inside a quad.
//
// Why a quad and not a thread per env: the thread-per-env version (profiles/r01_v1_*) needed ~MBs of unrolled SASS,
// 255 registers + 4 KB of stack and diverged 32 ways per warp; it stalled ~90% on instruction fetch.  With one leg per
// lane the per-lane working set fits in registers, the code is 4x smaller, and a warp holds 8 envs.
#ifndef JB_QUAD_CUH
#define JB_QUAD_CUH

#include "jb_math.cuh"
#include "jb_packed.h"
#include <stdint.h>

// cold paths (contact creation, cylinder / box / ellipsoid narrow phase) are real function calls on the device: they
// run rarely and inlining them at every call site made the kernel several times larger than the instruction cache
#if defined(__CUDACC__) && defined(JBQ_INLINE_ALL)      /* tuning: every narrow-phase routine in line */
#define JB_COLD __host__ __device__ __forceinline__
#elif defined(__CUDACC__)
#define JB_COLD __host__ __device__ __noinline__
#else
#define JB_COLD inline
#endif

// optional in-kernel phase timers (profiling builds only): Comm::tick(phase) closes the interval that ends here
#ifdef JBQ_PHASE_TIMERS
#define JBQ_TICK(p) do { cm.phase_sync(p); cm.tick(p); } while (0)
#else
#define JBQ_TICK(p) cm.phase_sync(p)
#endif
#ifdef JBQ_DEBUG_COUNT
static long long g_jbq_ls_iters = 0, g_jbq_newton_trips = 0, g_jbq_spread_moves = 0;
#endif
#if JBQ_CAP > 16
#error "the active-row masks hold 16 slots"
#endif
#if JBQ_CAP > 8
typedef uint64_t jbq_mask_t;
#else
typedef uint32_t jbq_mask_t;
#endif
// JBQ_SPREAD: balance a heavy lane's leg contacts over its quad (see the leg phase of jbq_substep).  It shortens the
// contact loops of the fullest lane from 4 to 2 trips, but costs ~800 instructions and 15 live accumulators.  It paid in
// round 1's per-env-model kernel (+10 %); since the load sort packs heavy envs into CTAs of their own (jb_sort.cu) both
// product builds compile it out (-16 % in the shared-model kernel, -9 % in the per-env one).  The host emulation keeps
// it on, so the path stays tested.
#ifndef JBQ_SPREAD
#define JBQ_SPREAD 1
#endif
#define JBQ_TRI(i, j) ((i) * ((i) + 1) / 2 + (j))
// a geom close enough to the floor for its narrow phase to run is the exception: keeps those blocks out of the hot
// instruction stream (tuning: -DJBQ_NO_UNLIKELY)
#if defined(__CUDACC__) && !defined(JBQ_NO_UNLIKELY)
#define JBQ_UNLIKELY(x) __builtin_expect(!!(x), 0)
#else
#define JBQ_UNLIKELY(x) (x)
#endif

template <typename T> struct QState {
  T pos[3], quat[4], vlin[3], w[3];     // root: replicated in the 4 lanes (MuJoCo coordinates: world linear, body angular)
  T qa, qb, va, vb;                     // this lane's leg: hinge 1 / hinge 2 angle and velocity
  T qm, vm; int rev;                    // motor: replicated; qpos[15] = qm + 2*pi*rev
  T wr[6], wla, wlb, wmo;               // qacc_warmstart, same split
};
template <typename T> struct QVec { T r[6]; T la, lb, mo; };
template <typename T> struct QArrow {   // R, cm, mm replicated; aa..cb this lane's leg
  T R[21]; T cm[6], mm; T aa, ab, bb, ca[6], cb[6];
};
enum { QK_LEG = 0, QK_ROOT = 1, QK_ROTOR = 2, QK_LEG_F = 3 };   // LEG_F: a contact of ANOTHER lane's leg (see jbq_spread)
// raw candidates parked in a slot's `kind` until they are built: geom id | JBQ_TAG_LOWER | JBQ_TAG_FOREIGN
enum { JBQ_TAG_LOWER = 256, JBQ_TAG_FOREIGN = 512 };
template <typename T> struct QCon {
  // world-frame rows of the 3 x 8 contact Jacobian over (root lin 3, root ang 3, jA, jB): u_world = J (a_root, qA, qB)
  // with (qA, qB) = (la, lb) of this lane's leg, or (motor, -) for a rotor contact.  Row k = [r_k, x cross r_k, r_k.ja,
  // r_k.jb], r_k = world axis k in root coordinates = row k of the sub-step's Rw (QFrame), which every contact shares:
  // only the five contact-specific entries [x cross r_k (3), r_k.ja, r_k.jb] are stored.  Built once per sub-step.
  T Jx[5], Jy[5], Jz[5];
  T D, ar0, uref[3];                    // 1/Rpy, -K*imp*dist, -B*v_point (WORLD comps)
  T ww[3], jv[3];                       // solver scratch (WORLD comps): J a - uref ; J search
  int kind;
};
template <typename T> struct QConSet { QCon<T> c[JBQ_CAP]; int n, overflow; };
template <typename T> struct QReport { int* geom; T* dist; int stride; };     // parity hook: [lane*JBQ_CAP + i]

template <typename T> JB_HD void qv_zero(QVec<T>& v) {
#pragma unroll
  for (int i = 0; i < 6; i++) v.r[i] = 0;
  v.la = 0; v.lb = 0; v.mo = 0;
}
template <typename T> JB_HD void qv_axpy(QVec<T>& y, T a, const QVec<T>& x) {
#pragma unroll
  for (int i = 0; i < 6; i++) y.r[i] += a * x.r[i];
  y.la += a * x.la; y.lb += a * x.lb; y.mo += a * x.mo;
}
// replicated part of a dot product (root + motor); the leg part must be quad-summed by the caller
template <typename T> JB_HD T qv_dot_rep(const QVec<T>& a, const QVec<T>& b) {
  T s = a.mo * b.mo;
#pragma unroll
  for (int i = 0; i < 6; i++) s += a.r[i] * b.r[i];
  return s;
}
template <typename T> JB_HD T qv_dot_leg(const QVec<T>& a, const QVec<T>& b) { return a.la * b.la + a.lb * b.lb; }

// y = A x
template <typename T, typename Comm> JB_HD void qarrow_mul(Comm& cm, QVec<T>& y, const QArrow<T>& A, const QVec<T>& x) {
  T p[6];
#pragma unroll
  for (int i = 0; i < 6; i++) p[i] = A.ca[i] * x.la + A.cb[i] * x.lb;
  cm.sumN(p, 6);
  T sa = A.aa * x.la + A.ab * x.lb, sb = A.ab * x.la + A.bb * x.lb, sm = A.mm * x.mo;
#pragma unroll
  for (int i = 0; i < 6; i++) {
    T s = p[i] + A.cm[i] * x.mo;
#pragma unroll
    for (int j = 0; j < 6; j++) s += A.R[i >= j ? JBQ_TRI(i, j) : JBQ_TRI(j, i)] * x.r[j];
    y.r[i] = s;
    sa += A.ca[i] * x.r[i]; sb += A.cb[i] * x.r[i]; sm += A.cm[i] * x.r[i];
  }
  y.la = sa; y.lb = sb; y.mo = sm;
}

// Leaf-first L'DL (MuJoCo's elimination order).  The replicated block (R, cm, mm) comes in one of two forms:
//   shares == false : every lane holds the full block (M, M + hB);
//   shares == true  : every lane holds its additive SHARE of the block (the Newton Hessian: lane 0 carries M's block
//                     plus its own contact terms, lanes 1-3 only their contact terms) and the quad sum forms the block.
// Either way the per-leg Schur complements are folded into the same quad reduction.
template <typename T, typename Comm> JB_HD void qarrow_factor(Comm& cm, QArrow<T>& A, bool shares) {
  T red[28];
  T ibb = (T)1 / A.bb;
  T lab = A.ab * ibb;
  T Lb[6], La[6];
#pragma unroll
  for (int i = 0; i < 6; i++) Lb[i] = A.cb[i] * ibb;
  A.aa -= lab * A.ab;
#pragma unroll
  for (int i = 0; i < 6; i++) A.ca[i] -= lab * A.cb[i];
  T iaa = (T)1 / A.aa;
#pragma unroll
  for (int i = 0; i < 6; i++) La[i] = A.ca[i] * iaa;
  if (shares) {
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
      for (int j = 0; j <= i; j++) red[JBQ_TRI(i, j)] = A.R[JBQ_TRI(i, j)] - Lb[i] * A.cb[j] - La[i] * A.ca[j];
#pragma unroll
    for (int i = 0; i < 6; i++) red[21 + i] = A.cm[i];
    red[27] = A.mm;
    cm.sumN(red, 28);
#pragma unroll
    for (int i = 0; i < 21; i++) A.R[i] = red[i];
#pragma unroll
    for (int i = 0; i < 6; i++) A.cm[i] = red[21 + i];
    A.mm = red[27];
  } else {
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
      for (int j = 0; j <= i; j++) red[JBQ_TRI(i, j)] = -Lb[i] * A.cb[j] - La[i] * A.ca[j];
    cm.sumN(red, 21);
#pragma unroll
    for (int i = 0; i < 21; i++) A.R[i] += red[i];
  }
  A.bb = ibb; A.ab = lab; A.aa = iaa;
#pragma unroll
  for (int i = 0; i < 6; i++) { A.cb[i] = Lb[i]; A.ca[i] = La[i]; }
  // motor dof (replicated)
  T imm = (T)1 / A.mm;
  T Lm[6];
#pragma unroll
  for (int i = 0; i < 6; i++) Lm[i] = A.cm[i] * imm;
#pragma unroll
  for (int i = 0; i < 6; i++)
#pragma unroll
    for (int j = 0; j <= i; j++) A.R[JBQ_TRI(i, j)] -= Lm[i] * A.cm[j];
  A.mm = imm;
#pragma unroll
  for (int i = 0; i < 6; i++) A.cm[i] = Lm[i];
  // dense LDL' of the 6x6 Schur complement (unit L below the diagonal, reciprocal pivots on it), replicated
  T dv[6];
#pragma unroll
  for (int j = 0; j < 6; j++) {
    T d = A.R[JBQ_TRI(j, j)];
    T ld[6];
#pragma unroll
    for (int k = 0; k < j; k++) { ld[k] = A.R[JBQ_TRI(j, k)] * dv[k]; d -= A.R[JBQ_TRI(j, k)] * ld[k]; }
    dv[j] = d;
    T id = (T)1 / d;
#pragma unroll
    for (int i = j + 1; i < 6; i++) {
      T s = A.R[JBQ_TRI(i, j)];
#pragma unroll
      for (int k = 0; k < j; k++) s -= A.R[JBQ_TRI(i, k)] * ld[k];
      A.R[JBQ_TRI(i, j)] = s * id;
    }
    A.R[JBQ_TRI(j, j)] = id;
  }
}

template <typename T, typename Comm> JB_HD void qarrow_solve(Comm& cm, QVec<T>& x, const QArrow<T>& F, const QVec<T>& r) {
  T ra = r.la - F.ab * r.lb;
  T rr[6];
#pragma unroll
  for (int i = 0; i < 6; i++) rr[i] = -(F.cb[i] * r.lb + F.ca[i] * ra);
  cm.sumN(rr, 6);
#pragma unroll
  for (int i = 0; i < 6; i++) rr[i] += r.r[i] - F.cm[i] * r.mo;
#pragma unroll
  for (int i = 0; i < 6; i++) {
#pragma unroll
    for (int k = 0; k < i; k++) rr[i] -= F.R[JBQ_TRI(i, k)] * rr[k];
  }
#pragma unroll
  for (int i = 0; i < 6; i++) rr[i] *= F.R[JBQ_TRI(i, i)];
#pragma unroll
  for (int i = 5; i >= 0; i--) {
#pragma unroll
    for (int k = i + 1; k < 6; k++) rr[i] -= F.R[JBQ_TRI(k, i)] * rr[k];
  }
  T xa = ra * F.aa, xb = r.lb * F.bb, xm = r.mo * F.mm;
#pragma unroll
  for (int i = 0; i < 6; i++) { x.r[i] = rr[i]; xa -= F.ca[i] * rr[i]; xb -= F.cb[i] * rr[i]; xm -= F.cm[i] * rr[i]; }
  xb -= F.ab * xa;
  x.la = xa; x.lb = xb; x.mo = xm;
}

// ------------------------------------------------------------------------------------------------ contacts
// MuJoCo's solimp sigmoid (getimpedance), impedance only; si = packed solimp[5].  The general power law is a call: the
// Jitterbug models use power = 2, and this function is inlined into every contact-creation site.
template <typename T> JB_COLD T jbq_impedance_pow(T x, T mid, T power) {
#if defined(__CUDA_ARCH__)
  return x <= mid ? (T)(__powf((float)x, (float)power) / __powf((float)mid, (float)power - 1.f))
                  : (T)(1.f - __powf(1.f - (float)x, (float)power) / __powf(1.f - (float)mid, (float)power - 1.f));
#else
  return x <= mid ? (T)(pow((double)x, (double)power) / pow((double)mid, (double)power - 1))
                  : (T)(1 - pow(1 - (double)x, (double)power) / pow(1 - (double)mid, (double)power - 1));
#endif
}
template <typename T> JB_HD T jbq_impedance(const T* si, T dist) {
  T d0 = jb_min((T)0.9999, jb_max((T)0.0001, si[0])), d1 = jb_min((T)0.9999, jb_max((T)0.0001, si[1]));
  T width = si[2], mid = jb_min((T)0.9999, jb_max((T)0.0001, si[3])), power = si[4];
  if (d0 == d1 || width <= (T)1e-15) return (T)0.5 * (d0 + d1);
  T x = jb_abs(dist) / width;
  if (x >= (T)1) return d1;
  if (x <= (T)0) return d0;
  T y;
  if (power == (T)2) y = x <= mid ? x * x / mid : (T)1 - ((T)1 - x) * ((T)1 - x) / ((T)1 - mid);
  else if (power <= (T)1) y = x;
  else y = jbq_impedance_pow(x, mid, power);
  return d0 + y * (d1 - d0);
}

template <typename T> struct QFrame {    // per-substep root-frame quantities (replicated)
  T Rw[9], n[3], pz, vr[3], w[3];
};
template <typename T> struct QBody {     // how a geom's body moves: joint columns for a point x
  int kind;                              // QK_*
  const T* a1; const T* p1; T qd1;       // first joint above the body (leg hinge 1, or the motor)
  const T* a2; const T* p2; T qd2; bool lower;
  T invw;
  bool defer;                            // park the raw contact (point, dist, geom) in its slot; jbq_build comes later
};

// Fill slot `c` for a contact at root-frame point x (on a body described by b) with penetration `dist`: Jacobian rows,
// regulariser, reference acceleration.
#if JBQ_SPREAD          /* the out-of-line unit is the builder (called from jbq_emit and from the spread step) ... */
#define JBQ_BUILD_ATTR JB_COLD
#define JBQ_EMIT_ATTR JB_HD
#else                   /* ... or, without spreading, the whole emit as one call per contact site */
#define JBQ_BUILD_ATTR JB_HD
#define JBQ_EMIT_ATTR JB_COLD
#endif
template <typename T>
JBQ_BUILD_ATTR void jbq_build(const T* g, QCon<T>& c, const QFrame<T>& fr, const QBody<T>& b, int geom, const T* x, T dist,
                     const QReport<T>* rep, int lane, int i) {
  c.kind = b.kind;
  T ja[3] = {0, 0, 0}, jb[3] = {0, 0, 0}, r[3];
  if (b.kind != QK_ROOT) {
    sub3(r, x, b.p1); cross3(ja, b.a1, r);
    if (b.lower) { sub3(r, x, b.p2); cross3(jb, b.a2, r); }
  }
  T* const rows[3] = {c.Jx, c.Jy, c.Jz};
  // point velocity in WORLD comps from the same rows: v_w[k] = J_k . (vr, w, qdA, qdB)
  T vw[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    T* J = rows[k];
    const T* rk = fr.Rw + 3 * k;
    cross3(J, x, rk);
    J[3] = dot3(rk, ja); J[4] = dot3(rk, jb);
    vw[k] = dot3(rk, fr.vr) + dot3(J, fr.w) + J[3] * b.qd1 + J[4] * b.qd2;
  }
  T imp = jbq_impedance<T>(g + G_SOLIMP, dist);
  T mu = g[G_FRIC];
  // diagApprox = tran*(1+mu^2); R0 = (1-imp)/imp*diagApprox; Rpy = 2*mu_reg^2*R0, mu_reg = mu/sqrt(impratio)
  T Rpy = (T)2 * mu * mu / g[G_IMPRATIO] * ((T)1 - imp) / imp * b.invw * ((T)1 + mu * mu);
  c.D = (T)1 / jb_max(Rpy, (T)1e-15);
  c.ar0 = -g[G_SOLK] * imp * dist;
  c.uref[0] = -g[G_SOLB] * vw[0]; c.uref[1] = -g[G_SOLB] * vw[1]; c.uref[2] = -g[G_SOLB] * vw[2];
  if (rep) { rep->geom[(lane * JBQ_CAP + i) * rep->stride] = geom; rep->dist[(lane * JBQ_CAP + i) * rep->stride] = dist; }
}

template <typename T>
JBQ_EMIT_ATTR void jbq_emit(const T* g, QConSet<T>& cs, const QFrame<T>& fr, const QBody<T>& b, int geom, const T* x, T dist,
                    const QReport<T>* rep, int lane) {
  if (!(dist < (T)0)) return;            // dist >= includemargin (= 0): MuJoCo excludes it from the solver
  int i = cs.n;
  if (i >= JBQ_CAP) { cs.overflow = 1; return; }
  cs.n = i + 1;
  QCon<T>& c = cs.c[i];
  if (b.defer) {
    c.ww[0] = x[0]; c.ww[1] = x[1]; c.ww[2] = x[2]; c.ar0 = dist; c.kind = geom | (b.lower ? JBQ_TAG_LOWER : 0);
    return;
  }
  jbq_build<T>(g, c, fr, b, geom, x, dist, rep, lane, i);
}

#define JBQ_EMIT(geom, x, dist) jbq_emit<T>(g, cs, fr, b, (geom), (x), (dist), rep, lane)

// the same, always inlined: for the one hot call site (a foot on the floor is THE common contact), see jbq_col_sphere
template <typename T>
JB_HD void jbq_emit_hot(const T* g, QConSet<T>& cs, const QFrame<T>& fr, const QBody<T>& b, int geom, const T* x, T dist,
                        const QReport<T>* rep, int lane) {
  if (!(dist < (T)0)) return;
  int i = cs.n;
  if (i >= JBQ_CAP) { cs.overflow = 1; return; }
  cs.n = i + 1;
  QCon<T>& c = cs.c[i];
  if (b.defer) {
    c.ww[0] = x[0]; c.ww[1] = x[1]; c.ww[2] = x[2]; c.ar0 = dist; c.kind = geom | (b.lower ? JBQ_TAG_LOWER : 0);
    return;
  }
  c.kind = b.kind;
  T ja[3] = {0, 0, 0}, jb[3] = {0, 0, 0}, r[3];
  if (b.kind != QK_ROOT) {
    sub3(r, x, b.p1); cross3(ja, b.a1, r);
    if (b.lower) { sub3(r, x, b.p2); cross3(jb, b.a2, r); }
  }
  T* const rows[3] = {c.Jx, c.Jy, c.Jz};
  T vw[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    T* J = rows[k];
    const T* rk = fr.Rw + 3 * k;
    cross3(J, x, rk);
    J[3] = dot3(rk, ja); J[4] = dot3(rk, jb);
    vw[k] = dot3(rk, fr.vr) + dot3(J, fr.w) + J[3] * b.qd1 + J[4] * b.qd2;
  }
  T imp = jbq_impedance<T>(g + G_SOLIMP, dist);
  T mu = g[G_FRIC];
  T Rpy = (T)2 * mu * mu / g[G_IMPRATIO] * ((T)1 - imp) / imp * b.invw * ((T)1 + mu * mu);
  c.D = (T)1 / jb_max(Rpy, (T)1e-15);
  c.ar0 = -g[G_SOLK] * imp * dist;
  c.uref[0] = -g[G_SOLB] * vw[0]; c.uref[1] = -g[G_SOLB] * vw[1]; c.uref[2] = -g[G_SOLB] * vw[2];
  if (rep) { rep->geom[(lane * JBQ_CAP + i) * rep->stride] = geom; rep->dist[(lane * JBQ_CAP + i) * rep->stride] = dist; }
}

template <typename T>
JB_HD void jbq_col_sphere(const T* g, QConSet<T>& cs, const QFrame<T>& fr, const QBody<T>& b, int geom, const T* c, T r,
                          const QReport<T>* rep, int lane, bool hot = false) {
  T dist = fr.pz + dot3(fr.n, c) - r;
  if (dist > (T)0) return;
  T s = r + (T)0.5 * dist;
  T x[3] = {c[0] - fr.n[0] * s, c[1] - fr.n[1] * s, c[2] - fr.n[2] * s};
  // a foot on the floor is THE common contact: building it in line instead of through the out-of-line jbq_emit saves the
  // call and the round trip of its by-reference arguments through local memory (+6 % on the whole step)
  if (hot) { jbq_emit_hot<T>(g, cs, fr, b, geom, x, dist, rep, lane); return; }
  JBQ_EMIT(geom, x, dist);
}

// mjc_PlaneCylinder; xax is only needed when the axis is exactly parallel to the floor normal.  HOT: everything in line
// (the lower-leg cylinder of a leg lying on the floor, see the leg phase); otherwise the emits are calls
#define JBQ_EMIT_SEL(geom, x, dist) do { if (HOT) jbq_emit_hot<T>(g, cs, fr, b, (geom), (x), (dist), rep, lane); else JBQ_EMIT(geom, x, dist); } while (0)
template <typename T, bool HOT>
JB_HD void jbq_col_cylinder_impl(const T* g, QConSet<T>& cs, const QFrame<T>& fr, const QBody<T>& b, int geom, const T* c,
                            const T* ax_in, const T* xax, T r, T hl, const QReport<T>* rep, int lane) {
  T dist0 = fr.pz + dot3(fr.n, c);
  T ax[3] = {ax_in[0], ax_in[1], ax_in[2]};
  T prjaxis = dot3(fr.n, ax);
  if (prjaxis > (T)0) { ax[0] = -ax[0]; ax[1] = -ax[1]; ax[2] = -ax[2]; prjaxis = -prjaxis; }
  T vec[3] = {ax[0] * prjaxis - fr.n[0], ax[1] * prjaxis - fr.n[1], ax[2] * prjaxis - fr.n[2]};
  T len_sqr = dot3(vec, vec);
  if (len_sqr >= (T)1e-30) {
    T s = r / jb_sqrt(len_sqr);
    vec[0] *= s; vec[1] *= s; vec[2] *= s;
  } else {
    vec[0] = xax[0] * r; vec[1] = xax[1] * r; vec[2] = xax[2] * r;
  }
  T prjvec = dot3(vec, fr.n);
  ax[0] *= hl; ax[1] *= hl; ax[2] *= hl;
  prjaxis *= hl;
  T d1 = dist0 + prjaxis + prjvec;
  if (d1 > (T)0) return;
  const T d2 = dist0 - prjaxis + prjvec, d3 = dist0 + prjaxis - prjvec * (T)0.5;
  // mjc_PlaneCylinder's up to four points -- the deepest rim point, its counterpart on the other cap, and two more on the
  // deepest cap's rim -- are x_k = c + al_k vec + be_k ax + ga_k v1 - n d_k / 2; one loop body (one copy of the contact
  // builder) instead of four
  T v1[3] = {0, 0, 0};
  if (d3 <= (T)0) {
    cross3(v1, vec, ax);
    T nn = jb_sqrt(dot3(v1, v1));
    if (nn < (T)1e-15) { v1[0] = 1; v1[1] = 0; v1[2] = 0; nn = 1; }
    T s = r * (T)0.8660254037844386 / nn;
    v1[0] *= s; v1[1] *= s; v1[2] *= s;
  }
#if defined(__CUDACC__)
#pragma unroll 1
#endif
  for (int k = 0; k < 4; k++) {
    const T dk = k == 0 ? d1 : (k == 1 ? d2 : d3);
    if (!(dk <= (T)0)) continue;
    const T al = k < 2 ? (T)1 : (T)-0.5, be = k == 1 ? (T)-1 : (T)1, ga = k == 2 ? (T)1 : (k == 3 ? (T)-1 : (T)0);
    T x[3];
#pragma unroll
    for (int i = 0; i < 3; i++) x[i] = c[i] + al * vec[i] + be * ax[i] + ga * v1[i] - fr.n[i] * dk * (T)0.5;
    JBQ_EMIT_SEL(geom, x, dk);
  }
}

#undef JBQ_EMIT_SEL
template <typename T>
JB_COLD void jbq_col_cylinder(const T* g, QConSet<T>& cs, const QFrame<T>& fr, const QBody<T>& b, int geom, const T* c,
                            const T* ax_in, const T* xax, T r, T hl, const QReport<T>* rep, int lane) {
  jbq_col_cylinder_impl<T, false>(g, cs, fr, b, geom, c, ax_in, xax, r, hl, rep, lane);
}

template <typename T>
JB_COLD void jbq_col_box(const T* g, QConSet<T>& cs, const QFrame<T>& fr, const QBody<T>& b, int geom, const T* c, const T* Rg,
                       const T* half, const QReport<T>* rep, int lane) {
  T dist = fr.pz + dot3(fr.n, c);
  int cnt = 0;
  for (int i = 0; i < 8; i++) {
    T v[3] = {(i & 1) ? half[0] : -half[0], (i & 2) ? half[1] : -half[1], (i & 4) ? half[2] : -half[2]};
    T corner[3];
    mulMV(corner, Rg, v);
    T ldist = dot3(fr.n, corner);
    if (dist + ldist > (T)0 || ldist > (T)0) continue;
    T cd = dist + ldist;
    T x[3] = {corner[0] + c[0] - fr.n[0] * cd * (T)0.5, corner[1] + c[1] - fr.n[1] * cd * (T)0.5,
              corner[2] + c[2] - fr.n[2] * cd * (T)0.5};
    JBQ_EMIT(geom, x, cd);
    if (++cnt >= 4) return;
  }
}

template <typename T>
JB_COLD void jbq_col_ellipsoid(const T* g, QConSet<T>& cs, const QFrame<T>& fr, const QBody<T>& b, int geom, const T* c,
                             const T* Rg, const T* size, const QReport<T>* rep, int lane) {
  T dir[3] = {-fr.n[0], -fr.n[1], -fr.n[2]}, loc[3];
  mulMTV(loc, Rg, dir);
  loc[0] *= size[0]; loc[1] *= size[1]; loc[2] *= size[2];
  T nrm = jb_sqrt(dot3(loc, loc));
  if (nrm < (T)1e-15) nrm = (T)1e-15;
  loc[0] = loc[0] / nrm * size[0]; loc[1] = loc[1] / nrm * size[1]; loc[2] = loc[2] / nrm * size[2];
  T wv[3], sup[3];
  mulMV(wv, Rg, loc);
  add3(sup, wv, c);
  T dist = fr.pz + dot3(fr.n, sup);
  if (dist > (T)0) return;
  T x[3] = {sup[0] - fr.n[0] * (T)0.5 * dist, sup[1] - fr.n[1] * (T)0.5 * dist, sup[2] - fr.n[2] * (T)0.5 * dist};
  JBQ_EMIT(geom, x, dist);
}

// the four pyramid rows of a contact, from w = (J a - uref) in WORLD comps: n +- mu*t1, n +- mu*t2 (t1 = y, t2 = -x)
template <typename T> JB_HD void jbq_rows(T* jar, const T* ww, T mu, T ar0) {
  jar[0] = ww[2] + mu * ww[1] - ar0;
  jar[1] = ww[2] - mu * ww[1] - ar0;
  jar[2] = ww[2] - mu * ww[0] - ar0;
  jar[3] = ww[2] + mu * ww[0] - ar0;
}

// out = J_c a in WORLD comps; lin = Rw a_lin (the part every contact shares)
template <typename T> JB_HD void jbq_J_apply(T* out, const QCon<T>& c, const T* lin, const QVec<T>& a, T laf, T lbf) {
  const T qa = c.kind == QK_ROTOR ? a.mo : ((JBQ_SPREAD && c.kind == QK_LEG_F) ? laf : a.la);
  const T qb = (JBQ_SPREAD && c.kind == QK_LEG_F) ? lbf : a.lb;
  const T* const rows[3] = {c.Jx, c.Jy, c.Jz};
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const T* J = rows[k];
    out[k] = lin[k] + J[0] * a.r[3] + J[1] * a.r[4] + J[2] * a.r[5] + J[3] * qa + J[4] * qb;
  }
}

// ww = J_c a - uref (dst = 0) or jv = J_c a (dst = 1) for every contact of this lane; (laf, lbf): the leg entries of `a`
// in the lane that owns this lane's foreign contacts
template <typename T> JB_HD void jbq_apply_all(QConSet<T>& cs, const QFrame<T>& fr, const QVec<T>& a, int dst, T laf, T lbf) {
  T lin[3];
  mulMV(lin, fr.Rw, a.r);
  for (int i = 0; i < cs.n; i++) {
    QCon<T>& c = cs.c[i];
    T u[3];
    jbq_J_apply(u, c, lin, a, laf, lbf);
    if (dst == 0) { c.ww[0] = u[0] - c.uref[0]; c.ww[1] = u[1] - c.uref[1]; c.ww[2] = u[2] - c.uref[2]; }
    else { c.jv[0] = u[0]; c.jv[1] = u[1]; c.jv[2] = u[2]; }
  }
}

// This lane's constraint forces at the current ww: partial generalized force (root 6 + motor partial in out7[6], leg in
// *fla / *flb) and the active-row mask (4 bits per contact).
template <typename T>
JB_HD jbq_mask_t jbq_forces(const QConSet<T>& cs, const QFrame<T>& fr, T mu, T* out7, T* fla, T* flb, T* ffa, T* ffb) {
  jbq_mask_t mask = 0;
#pragma unroll
  for (int i = 0; i < 7; i++) out7[i] = 0;
  *fla = 0; *flb = 0; *ffa = 0; *ffb = 0;
  T fsum[3] = {0, 0, 0};            // sum of the contact forces in WORLD comps: the linear root force is Rw' fsum
  for (int i = 0; i < cs.n; i++) {
    const QCon<T>& c = cs.c[i];
    T jar[4];
    jbq_rows(jar, c.ww, mu, c.ar0);
    mask |= (jbq_mask_t)((jar[0] < (T)0 ? 1u : 0u) | (jar[1] < (T)0 ? 2u : 0u) | (jar[2] < (T)0 ? 4u : 0u) | (jar[3] < (T)0 ? 8u : 0u)) << (4 * i);
    const T f0 = jar[0] < (T)0 ? -c.D * jar[0] : (T)0, f1 = jar[1] < (T)0 ? -c.D * jar[1] : (T)0;
    const T f2 = jar[2] < (T)0 ? -c.D * jar[2] : (T)0, f3 = jar[3] < (T)0 ? -c.D * jar[3] : (T)0;
    const T fw0 = mu * (f3 - f2), fw1 = mu * (f0 - f1), fw2 = f0 + f1 + f2 + f3;
    fsum[0] += fw0; fsum[1] += fw1; fsum[2] += fw2;
#pragma unroll
    for (int p = 0; p < 3; p++) out7[3 + p] += c.Jx[p] * fw0 + c.Jy[p] * fw1 + c.Jz[p] * fw2;
    const T gA = c.Jx[3] * fw0 + c.Jy[3] * fw1 + c.Jz[3] * fw2, gB = c.Jx[4] * fw0 + c.Jy[4] * fw1 + c.Jz[4] * fw2;
    if (c.kind == QK_ROTOR) out7[6] += gA;
    else if ((JBQ_SPREAD && c.kind == QK_LEG_F)) { *ffa += gA; *ffb += gB; }
    else { *fla += gA; *flb += gB; }
  }
  mulMTV(out7, fr.Rw, fsum);
  return mask;
}

// Hessian terms J' W J of this lane's contacts for the active set `mask` (from jbq_forces at the same point), added to
// H: the leg block, and this lane's SHARE of the replicated block (see qarrow_factor).  W = sum over active rows of
// D e e' has 5 distinct entries for the pyramid {z +- mu y, z -+ mu x}.
template <typename T>
JB_HD void jbq_hessian(const QConSet<T>& cs, const QFrame<T>& fr, T mu, jbq_mask_t mask, QArrow<T>& H, T* F15) {
  // F15: leg-block terms of this lane's FOREIGN contacts (ca[6], aa, cb[6], ab, bb of the owning lane)
  if (JBQ_SPREAD) {
#pragma unroll
    for (int q = 0; q < 15; q++) F15[q] = 0;
  }
  for (int i = 0; i < cs.n; i++) {
    const QCon<T>& c = cs.c[i];
    const uint32_t m4 = (uint32_t)(mask >> (4 * i));
    const T a0 = (m4 & 1u) ? (T)1 : (T)0, a1 = (m4 & 2u) ? (T)1 : (T)0;
    const T a2 = (m4 & 4u) ? (T)1 : (T)0, a3 = (m4 & 8u) ? (T)1 : (T)0;
    const T dm = c.D * mu;
    const T Wzz = c.D * (a0 + a1 + a2 + a3), Wyz = dm * (a0 - a1), Wxz = dm * (a3 - a2);
    const T Wyy = dm * mu * (a0 + a1), Wxx = dm * mu * (a2 + a3);
    // full rows: the three linear entries are the shared Rw rows, the other five come from the slot
    T jx[8], jy[8], jz[8];
#pragma unroll
    for (int q = 0; q < 3; q++) { jx[q] = fr.Rw[q]; jy[q] = fr.Rw[3 + q]; jz[q] = fr.Rw[6 + q]; }
#pragma unroll
    for (int q = 0; q < 5; q++) { jx[3 + q] = c.Jx[q]; jy[3 + q] = c.Jy[q]; jz[3 + q] = c.Jz[q]; }
    T Gx[8], Gy[8], Gz[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
      Gx[q] = Wxx * jx[q] + Wxz * jz[q];
      Gy[q] = Wyy * jy[q] + Wyz * jz[q];
      Gz[q] = Wxz * jx[q] + Wyz * jy[q] + Wzz * jz[q];
    }
#pragma unroll
    for (int p = 0; p < 6; p++)
#pragma unroll
      for (int q = 0; q <= p; q++) H.R[JBQ_TRI(p, q)] += jx[p] * Gx[q] + jy[p] * Gy[q] + jz[p] * Gz[q];
    T hA[7], hB[8];      // H[A][0..5], H[A][A]; H[B][0..5], H[B][A], H[B][B]
#pragma unroll
    for (int q = 0; q < 7; q++) hA[q] = jx[6] * Gx[q] + jy[6] * Gy[q] + jz[6] * Gz[q];
#pragma unroll
    for (int q = 0; q < 8; q++) hB[q] = jx[7] * Gx[q] + jy[7] * Gy[q] + jz[7] * Gz[q];
    if (c.kind == QK_ROTOR) {
#pragma unroll
      for (int q = 0; q < 6; q++) H.cm[q] += hA[q];
      H.mm += hA[6];
    } else if ((JBQ_SPREAD && c.kind == QK_LEG_F)) {
#pragma unroll
      for (int q = 0; q < 7; q++) F15[q] += hA[q];
#pragma unroll
      for (int q = 0; q < 8; q++) F15[7 + q] += hB[q];
    } else {
#pragma unroll
      for (int q = 0; q < 6; q++) { H.ca[q] += hA[q]; H.cb[q] += hB[q]; }
      H.aa += hA[6]; H.ab += hB[6]; H.bb += hB[7];
    }
  }
}

// this lane's part of phi'(alpha) - (g1 + alpha g2) and of phi''(alpha) - g2; returns the active mask at alpha
template <typename T> JB_HD jbq_mask_t jbq_ls_local(const QConSet<T>& cs, T mu, T alpha, T* d_out, T* h_out) {
  T d = 0, h = 0;
  jbq_mask_t mask = 0;
  for (int i = 0; i < cs.n; i++) {
    const QCon<T>& c = cs.c[i];
    T jar[4], jv[4];
    jbq_rows(jar, c.ww, mu, c.ar0);
    jbq_rows(jv, c.jv, mu, (T)0);
#pragma unroll
    for (int r = 0; r < 4; r++) {
      T x = jar[r] + alpha * jv[r];
      if (x < (T)0) { d += c.D * x * jv[r]; h += c.D * jv[r] * jv[r]; mask |= (jbq_mask_t)1 << (4 * i + r); }
    }
  }
  *d_out = d; *h_out = h;
  return mask;
}

// ------------------------------------------------------------------------------------------------ the sub-step (SPMD)
// P: packed model (jb_packed.h); lane in [0,4).  forward_only: stop after the constraint solve (mj_forward).
template <typename T, typename Comm>
JB_HD void jbq_substep(Comm& cm, const T* P, int lane, QState<T>& s, T ctrl, bool forward_only, int* iter_accum,
                       int* ncon_accum, int* overflow_accum, const QReport<T>* rep, int* lane_load = nullptr) {
  // forward_only (mj_forward: everything up to qacc / qacc_warmstart) is a sub-step of length zero, so that every
  // thread of a CTA follows the same control flow whatever its env is doing (the phase barriers below rely on it)
  const T* g = P;
  const T* X = P + JBP_X + lane;
  const T* L = P + JBP_L + lane;
#define LX(i) X[4 * (i)]
#define LL(i) L[4 * (i)]
  const T h = forward_only ? (T)0 : g[G_H];
  JBQ_TICK(0);
  // ---------------- root frame (replicated)
  QFrame<T> fr;
  {
    T qn = (T)1 / jb_sqrt(s.quat[0] * s.quat[0] + s.quat[1] * s.quat[1] + s.quat[2] * s.quat[2] + s.quat[3] * s.quat[3]);
    s.quat[0] *= qn; s.quat[1] *= qn; s.quat[2] *= qn; s.quat[3] *= qn;
    quat2mat(fr.Rw, s.quat);
  }
  fr.pz = s.pos[2];
  fr.n[0] = fr.Rw[6]; fr.n[1] = fr.Rw[7]; fr.n[2] = fr.Rw[8];
  mulMTV(fr.vr, fr.Rw, s.vlin);
  fr.w[0] = s.w[0]; fr.w[1] = s.w[1]; fr.w[2] = s.w[2];
  const T* w = fr.w;
  const T aO[3] = {g[G_GRAV] * fr.n[0], g[G_GRAV] * fr.n[1], g[G_GRAV] * fr.n[2]};   // fictitious -g at the root origin

  QConSet<T> cs;
  cs.n = 0; cs.overflow = 0;
  int hv = 0;            // the lane of this quad whose contacts were spread over the others (jbq_spread step below)
  bool wf = false;       // warp-uniform: some quad of the warp carries foreign contacts

  // first moment / inertia about the origin and bias wrench: per-lane partials (this leg), summed over the quad below
  T ht[3] = {0, 0, 0}, It[6] = {0, 0, 0, 0, 0, 0};
  T bF[3] = {0, 0, 0}, bT[3] = {0, 0, 0};
  auto accum_body = [&](T mass, const T* c, const T* Ic) {
    addscl3(ht, c, mass);
    T cc = dot3(c, c);
    It[0] += Ic[0] + mass * (cc - c[0] * c[0]);
    It[1] += Ic[1] + mass * (cc - c[1] * c[1]);
    It[2] += Ic[2] + mass * (cc - c[2] * c[2]);
    It[3] += Ic[3] - mass * c[0] * c[1];
    It[4] += Ic[4] - mass * c[0] * c[2];
    It[5] += Ic[5] - mass * c[1] * c[2];
  };
  auto acc_root_point = [&](T* a, const T* p) {
    T t[3], u[3];
    cross3(t, w, p); cross3(u, w, t);
    a[0] = aO[0] + u[0]; a[1] = aO[1] + u[1]; a[2] = aO[2] + u[2];
  };
  auto acc_child = [&](T* out, const T* abase, const T* al, const T* om, const T* rr) {
    T x1[3], x2[3], x3[3];
    cross3(x1, al, rr); cross3(x2, om, rr); cross3(x3, om, x2);
    out[0] = abase[0] + x1[0] + x3[0]; out[1] = abase[1] + x1[1] + x3[1]; out[2] = abase[2] + x1[2] + x3[2];
  };
  auto body_wrench = [&](T* F, T* N, T mass, const T* Ic, const T* acom, const T* om, const T* al) {
    scl3(F, acom, mass);
    T Iw[3], Ia[3], t[3];
    mulSV(Iw, Ic, om); mulSV(Ia, Ic, al); cross3(t, om, Iw);
    add3(N, Ia, t);
  };

  QArrow<T> Mx;
  QVec<T> bias, fpass;
  qv_zero(bias); qv_zero(fpass);

  // ---------------- this lane's leg
  T a1[3] = {LL(L_A1), LL(L_A1 + 1), LL(L_A1 + 2)}, p1[3] = {LL(L_P1), LL(L_P1 + 1), LL(L_P1 + 2)};
  T a2[3], p2[3];
  {
    T a2h[3] = {LL(L_A2), LL(L_A2 + 1), LL(L_A2 + 2)}, p2h[3] = {LL(L_P2), LL(L_P2 + 1), LL(L_P2 + 2)};
    T sa, ca, sb, cb;
    jb_sincos_fast(s.qa, &sa, &ca);
    jb_sincos_fast(s.qb, &sb, &cb);
    T R1[9], R2[9], R12[9];
    rodrigues(R1, a1, sa, ca);
    rodrigues(R2, a2h, sb, cb);
    mulMM(R12, R1, R2);
    T t[3], r[3], tmp[3], tmp2[3];
    mulMV(a2, R1, a2h);
    sub3(r, p2h, p1); mulMV(t, R1, r); add3(p2, p1, t);
    T cu[3], cl[3], Iu[6], Il[6], ch[3], Ih[6];
    ch[0] = LL(L_CU); ch[1] = LL(L_CU + 1); ch[2] = LL(L_CU + 2);
    sub3(r, ch, p1); mulMV(t, R1, r); add3(cu, p1, t);
    ch[0] = LL(L_CL); ch[1] = LL(L_CL + 1); ch[2] = LL(L_CL + 2);
    sub3(r, ch, p2h); mulMV(t, R12, r); add3(cl, p2, t);
#pragma unroll
    for (int i = 0; i < 6; i++) Ih[i] = LL(L_IU + i);
    rotSym(Iu, R1, Ih);
#pragma unroll
    for (int i = 0; i < 6; i++) Ih[i] = LL(L_IL + i);
    rotSym(Il, R12, Ih);
    const T mu_ = LL(L_MU), ml_ = LL(L_ML);
    accum_body(mu_, cu, Iu);
    accum_body(ml_, cl, Il);
    // inertia columns: P = sum m a x (c - p), L = sum (I a + c x P_b)
    T Pb[3], Lb[3], Pa[3], La[3], Pal[3], Pau[3];
    sub3(r, cl, p2); cross3(tmp, a2, r); scl3(Pb, tmp, ml_);
    mulSV(tmp, Il, a2); cross3(tmp2, cl, Pb); add3(Lb, tmp, tmp2);
    sub3(r, cl, p1); cross3(tmp, a1, r); scl3(Pal, tmp, ml_);
    sub3(r, cu, p1); cross3(tmp, a1, r); scl3(Pau, tmp, mu_);
    add3(Pa, Pal, Pau);
    mulSV(La, Il, a1); mulSV(tmp, Iu, a1); addto3(La, tmp);
    cross3(tmp, cl, Pal); addto3(La, tmp);
    cross3(tmp, cu, Pau); addto3(La, tmp);
#pragma unroll
    for (int i = 0; i < 3; i++) { Mx.cb[i] = Pb[i]; Mx.cb[3 + i] = Lb[i]; Mx.ca[i] = Pa[i]; Mx.ca[3 + i] = La[i]; }
    cross3(tmp, p2, Pb); sub3(tmp2, Lb, tmp); Mx.bb = dot3(a2, tmp2);
    cross3(tmp, p1, Pb); sub3(tmp2, Lb, tmp); Mx.ab = dot3(a1, tmp2);
    cross3(tmp, p1, Pa); sub3(tmp2, La, tmp); Mx.aa = dot3(a1, tmp2);
    // Newton-Euler bias
    T om1[3], al1[3], om2[3], al2[3], ap1[3], ap2[3], acu[3], acl[3];
    om1[0] = w[0] + a1[0] * s.va; om1[1] = w[1] + a1[1] * s.va; om1[2] = w[2] + a1[2] * s.va;
    cross3(al1, w, a1); scl3(al1, al1, s.va);
    om2[0] = om1[0] + a2[0] * s.vb; om2[1] = om1[1] + a2[1] * s.vb; om2[2] = om1[2] + a2[2] * s.vb;
    cross3(tmp, om1, a2);
    al2[0] = al1[0] + tmp[0] * s.vb; al2[1] = al1[1] + tmp[1] * s.vb; al2[2] = al1[2] + tmp[2] * s.vb;
    acc_root_point(ap1, p1);
    T ru[3], r12[3], rl[3];
    sub3(ru, cu, p1); acc_child(acu, ap1, al1, om1, ru);
    sub3(r12, p2, p1); acc_child(ap2, ap1, al1, om1, r12);
    sub3(rl, cl, p2); acc_child(acl, ap2, al2, om2, rl);
    T Fu[3], Nu[3], Fl[3], Nl[3];
    body_wrench(Fu, Nu, mu_, Iu, acu, om1, al1);
    body_wrench(Fl, Nl, ml_, Il, acl, om2, al2);
    cross3(tmp, rl, Fl); addto3(tmp, Nl); bias.lb = dot3(a2, tmp);
    cross3(tmp, ru, Fu); addto3(tmp, Nu);
    sub3(r, cl, p1); cross3(tmp2, r, Fl); addto3(tmp, tmp2); addto3(tmp, Nl);
    bias.la = dot3(a1, tmp);
    bF[0] = Fu[0] + Fl[0]; bF[1] = Fu[1] + Fl[1]; bF[2] = Fu[2] + Fl[2];
    cross3(tmp, cu, Fu); cross3(tmp2, cl, Fl);
    bT[0] = tmp[0] + tmp2[0] + Nu[0] + Nl[0]; bT[1] = tmp[1] + tmp2[1] + Nu[1] + Nl[1]; bT[2] = tmp[2] + tmp2[2] + Nu[2] + Nl[2];
    fpass.la = -LL(L_K1) * s.qa - LL(L_D1) * s.va;
    fpass.lb = -LL(L_K2) * s.qb - LL(L_D2) * s.vb;
    // collision: this leg's geoms.  Cheap height test on the sphere centres / cylinder bounding spheres first.
    {
      // the leg's contacts are parked raw (defer) and built after jbq_spread has balanced them over the quad
      // The out-of-line collision routines take the frame and the body description by reference, i.e. through local
      // memory.  They run rarely, so each call site makes its own copy (JBQ_COLD_ARGS) inside its branch: `fr` and the
      // description used by the in-line foot path then never leave the registers on the common path.
#define JBQ_COLD_ARGS(body) const QFrame<T> frc = fr; const QBody<T> bc = body
#define JBQ_BODY_U {QK_LEG, a1, p1, s.va, a2, p2, (T)0, false, LL(L_INVWU), JBQ_SPREAD != 0}
#define JBQ_BODY_L {QK_LEG, a1, p1, s.va, a2, p2, s.vb, true, LL(L_INVWL), JBQ_SPREAD != 0}
      const QBody<T> bl = JBQ_BODY_L;
      T c[3], gh[3], gv[3], ax[3], xa[3];
      // foot sphere (lower body)
      gh[0] = LL(L_FT_C); gh[1] = LL(L_FT_C + 1); gh[2] = LL(L_FT_C + 2);
      sub3(r, gh, p2h); mulMV(t, R12, r); add3(c, p2, t);
      jbq_col_sphere<T>(g, cs, fr, bl, (int)LL(L_FT_ID), c, LL(L_FT_R), rep, lane, true);
      // lower cylinder
      gh[0] = LL(L_LC_C); gh[1] = LL(L_LC_C + 1); gh[2] = LL(L_LC_C + 2);
      sub3(r, gh, p2h); mulMV(t, R12, r); add3(c, p2, t);
      {
        T rr_ = LL(L_LC_R), hl_ = LL(L_LC_HL);
        if (JBQ_UNLIKELY(fr.pz + dot3(fr.n, c) - jb_sqrt(rr_ * rr_ + hl_ * hl_) <= (T)0)) {
          gv[0] = LL(L_LC_AX); gv[1] = LL(L_LC_AX + 1); gv[2] = LL(L_LC_AX + 2); mulMV(ax, R12, gv);
          gv[0] = LL(L_LC_XA); gv[1] = LL(L_LC_XA + 1); gv[2] = LL(L_LC_XA + 2); mulMV(xa, R12, gv);
          // in line as well: a leg lying on the floor (the heavy envs) comes through here every sub-step; measured
          // +4 % on the whole step.  Knee / upper cylinder / extra geoms in line too: no further gain (JBQ_UPPER_INLINE,
          // JBQ_INLINE_ALL), so those stay calls.
#ifndef JBQ_LCYL_CALL
          jbq_col_cylinder_impl<T, true>(g, cs, fr, bl, (int)LL(L_LC_ID), c, ax, xa, rr_, hl_, rep, lane);
#else
          JBQ_COLD_ARGS(JBQ_BODY_L);
          const T cc[3] = {c[0], c[1], c[2]};
          jbq_col_cylinder<T>(g, cs, frc, bc, (int)LL(L_LC_ID), cc, ax, xa, rr_, hl_, rep, lane);
#endif
        }
      }
        // knee sphere (upper body)
      gh[0] = LL(L_KN_C); gh[1] = LL(L_KN_C + 1); gh[2] = LL(L_KN_C + 2);
      sub3(r, gh, p1); mulMV(t, R1, r); add3(c, p1, t);
#ifdef JBQ_UPPER_INLINE
      const QBody<T> bu = JBQ_BODY_U;
      jbq_col_sphere<T>(g, cs, fr, bu, (int)LL(L_KN_ID), c, LL(L_KN_R), rep, lane, true);
#else
      if (JBQ_UNLIKELY(fr.pz + dot3(fr.n, c) - LL(L_KN_R) <= (T)0)) {          // (the same test jbq_col_sphere starts with)
        JBQ_COLD_ARGS(JBQ_BODY_U);
        const T cc[3] = {c[0], c[1], c[2]};
        jbq_col_sphere<T>(g, cs, frc, bc, (int)LL(L_KN_ID), cc, LL(L_KN_R), rep, lane);
      }
#endif
      // upper cylinder
      gh[0] = LL(L_UC_C); gh[1] = LL(L_UC_C + 1); gh[2] = LL(L_UC_C + 2);
      sub3(r, gh, p1); mulMV(t, R1, r); add3(c, p1, t);
      {
        T rr_ = LL(L_UC_R), hl_ = LL(L_UC_HL);
        if (JBQ_UNLIKELY(fr.pz + dot3(fr.n, c) - jb_sqrt(rr_ * rr_ + hl_ * hl_) <= (T)0)) {
          gv[0] = LL(L_UC_AX); gv[1] = LL(L_UC_AX + 1); gv[2] = LL(L_UC_AX + 2); mulMV(ax, R1, gv);
          gv[0] = LL(L_UC_XA); gv[1] = LL(L_UC_XA + 1); gv[2] = LL(L_UC_XA + 2); mulMV(xa, R1, gv);
#ifdef JBQ_UPPER_INLINE
          jbq_col_cylinder_impl<T, true>(g, cs, fr, bu, (int)LL(L_UC_ID), c, ax, xa, rr_, hl_, rep, lane);
#else
          JBQ_COLD_ARGS(JBQ_BODY_U);
          const T cc[3] = {c[0], c[1], c[2]};
          jbq_col_cylinder<T>(g, cs, frc, bc, (int)LL(L_UC_ID), cc, ax, xa, rr_, hl_, rep, lane);
#endif
        }
      }
    }
    // ---------------- spread a heavy lane's contacts over the quad.  The per-contact loops of the solver run as many
    // trips as the fullest lane of the WARP has contacts, and a leg that stands on the rim of its lower cylinder brings
    // four (foot + three cap points) where the other lanes have one: the warp -- and through the phase barriers the
    // CTA -- then waits for that lane.  The fullest lane of the quad (hv) hands its last raw contacts to the emptiest
    // lanes until the counts differ by at most one; a helper builds them with hv's leg frame (QK_LEG_F) and the solver
    // routes their leg-dof terms back to hv (foreign accumulators + quad sums in jbq_substep).
    if (JBQ_SPREAD) {
      const int n0 = cs.n;
      int give = 0, M = 0;
      if (cm.warp_any(n0 >= 3)) {            // the usual sub-step (at most two contacts per lane) skips the planning
      const int c0 = cm.iget(n0, 0), c1 = cm.iget(n0, 1), c2 = cm.iget(n0, 2), c3 = cm.iget(n0, 3);
      int pk = c0 | (c1 << 4) | (c2 << 8) | (c3 << 12), hmax = c0;
      hv = 0;
      if (c1 > hmax) { hv = 1; hmax = c1; }
      if (c2 > hmax) { hv = 2; hmax = c2; }
      if (c3 > hmax) { hv = 3; hmax = c3; }
      if (hmax >= 3) {
#pragma unroll
        for (int it = 0; it < 4; it++) {
          int m = -1, cmin = 99;
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int cj = (pk >> (4 * j)) & 15;
            if (j != hv && cj < cmin) { m = j; cmin = cj; }
          }
          const int ch = (pk >> (4 * hv)) & 15;
          if (ch - cmin >= 2) { pk += (1 << (4 * m)) - (1 << (4 * hv)); give += 1 << (4 * m); M++; }
        }
      }
      }
#ifdef JBQ_DEBUG_COUNT
      if (lane == 0) g_jbq_spread_moves += M;
#endif
      wf = cm.warp_any(M > 0);
      if (wf) {
        // move k (k = 0..M-1) takes hv's slot n_hv - M + k to the lanes in index order: give_0 moves to lane 0, ...
        for (int k = 0; cm.warp_any(k < M); k++) {
          T v[5] = {0, 0, 0, 0, 0};
          if (lane == hv && k < M) {
            const QCon<T>& c = cs.c[n0 - M + k];
            v[0] = c.ww[0]; v[1] = c.ww[1]; v[2] = c.ww[2]; v[3] = c.ar0; v[4] = (T)c.kind;
          }
#pragma unroll
          for (int i = 0; i < 5; i++) v[i] = cm.get(v[i], hv);
          int dst = 0, acc = give & 15;
          if (k >= acc) { dst = 1; acc += (give >> 4) & 15; }
          if (k >= acc) { dst = 2; acc += (give >> 8) & 15; }
          if (k >= acc) dst = 3;
          if (k < M && lane == dst) {
            QCon<T>& c = cs.c[cs.n];
            cs.n++;
            c.ww[0] = v[0]; c.ww[1] = v[1]; c.ww[2] = v[2]; c.ar0 = v[3]; c.kind = (int)v[4] | JBQ_TAG_FOREIGN;
          }
        }
        if (lane == hv) cs.n = n0 - M;
      }
      // build: own contacts with this leg's frame, foreign ones with hv's (hinge 1 is a model constant of leg hv,
      // hinge 2 and the joint velocities come over the quad)
      T fa2[3] = {a2[0], a2[1], a2[2]}, fp2[3] = {p2[0], p2[1], p2[2]}, fva = s.va, fvb = s.vb;
      if (wf) {
#pragma unroll
        for (int i = 0; i < 3; i++) { fa2[i] = cm.get(fa2[i], hv); fp2[i] = cm.get(fp2[i], hv); }
        fva = cm.get(fva, hv); fvb = cm.get(fvb, hv);
      }
      const T* Lh = P + JBP_L + hv;
      const T fa1[3] = {Lh[4 * L_A1], Lh[4 * (L_A1 + 1)], Lh[4 * (L_A1 + 2)]};
      const T fp1[3] = {Lh[4 * L_P1], Lh[4 * (L_P1 + 1)], Lh[4 * (L_P1 + 2)]};
      for (int i = 0; i < cs.n; i++) {
        QCon<T>& c = cs.c[i];
        const int tag = c.kind;
        const bool lower = (tag & JBQ_TAG_LOWER) != 0, foreign = (tag & JBQ_TAG_FOREIGN) != 0;
        const T x[3] = {c.ww[0], c.ww[1], c.ww[2]};
        const T dist = c.ar0;
        if (foreign) {
          QBody<T> b = {QK_LEG_F, fa1, fp1, fva, fa2, fp2, lower ? fvb : (T)0, lower, Lh[4 * (lower ? L_INVWL : L_INVWU)], false};
          jbq_build<T>(g, c, fr, b, tag & 255, x, dist, rep, lane, i);
        } else {
          QBody<T> b = {QK_LEG, a1, p1, s.va, a2, p2, lower ? s.vb : (T)0, lower, LL(lower ? L_INVWL : L_INVWU), false};
          jbq_build<T>(g, c, fr, b, tag & 255, x, dist, rep, lane, i);
        }
      }
    }
  }

  JBQ_TICK(1);
  // ---------------- quad-sum the leg partials: first moment (3), origin inertia (6), bias force/torque (6)
  {
    T red[15] = {ht[0], ht[1], ht[2], It[0], It[1], It[2], It[3], It[4], It[5], bF[0], bF[1], bF[2], bT[0], bT[1], bT[2]};
    cm.sumN(red, 15);
    ht[0] = red[0]; ht[1] = red[1]; ht[2] = red[2];
#pragma unroll
    for (int i = 0; i < 6; i++) It[i] = red[3 + i];
    bias.r[0] = red[9]; bias.r[1] = red[10]; bias.r[2] = red[11]; bias.r[3] = red[12]; bias.r[4] = red[13]; bias.r[5] = red[14];
  }

  // ---------------- root (core) body and rotor: replicated in the 4 lanes
  T pm[3] = {g[G_PM], g[G_PM + 1], g[G_PM + 2]}, am[3] = {g[G_AM], g[G_AM + 1], g[G_AM + 2]};
  T Rm[9];
  {
    T c0[3] = {g[G_C0], g[G_C0 + 1], g[G_C0 + 2]}, I0[6];
#pragma unroll
    for (int i = 0; i < 6; i++) I0[i] = g[G_I0 + i];
    accum_body(g[G_M0], c0, I0);
    T ac[3], F[3], N[3], zero[3] = {0, 0, 0}, t[3];
    acc_root_point(ac, c0);
    body_wrench(F, N, g[G_M0], I0, ac, w, zero);
    bias.r[0] += F[0]; bias.r[1] += F[1]; bias.r[2] += F[2];
    cross3(t, c0, F);
    bias.r[3] += t[0] + N[0]; bias.r[4] += t[1] + N[1]; bias.r[5] += t[2] + N[2];
    // rotor
    T sm, cmn, r[3], cmc[3], Imc[6], ch[3] = {g[G_CM], g[G_CM + 1], g[G_CM + 2]}, Ih[6];
    jb_sincos_fast(s.qm, &sm, &cmn);
    rodrigues(Rm, am, sm, cmn);
    sub3(r, ch, pm); mulMV(t, Rm, r); add3(cmc, pm, t);
#pragma unroll
    for (int i = 0; i < 6; i++) Ih[i] = g[G_IM + i];
    rotSym(Imc, Rm, Ih);
    const T mm_ = g[G_MM];
    accum_body(mm_, cmc, Imc);
    T Pv[3], Lv[3], tmp[3], tmp2[3];
    sub3(r, cmc, pm); cross3(tmp, am, r); scl3(Pv, tmp, mm_);
    mulSV(Lv, Imc, am); cross3(tmp, cmc, Pv); addto3(Lv, tmp);
#pragma unroll
    for (int i = 0; i < 3; i++) { Mx.cm[i] = Pv[i]; Mx.cm[3 + i] = Lv[i]; }
    cross3(tmp, pm, Pv); sub3(tmp2, Lv, tmp); Mx.mm = dot3(am, tmp2);
    T om[3], al[3], ap[3];
    om[0] = w[0] + am[0] * s.vm; om[1] = w[1] + am[1] * s.vm; om[2] = w[2] + am[2] * s.vm;
    cross3(al, w, am); scl3(al, al, s.vm);
    acc_root_point(ap, pm);
    acc_child(ac, ap, al, om, r);
    body_wrench(F, N, mm_, Imc, ac, om, al);
    cross3(tmp, r, F); addto3(tmp, N); bias.mo = dot3(am, tmp);
    bias.r[0] += F[0]; bias.r[1] += F[1]; bias.r[2] += F[2];
    cross3(tmp, cmc, F);
    bias.r[3] += tmp[0] + N[0]; bias.r[4] += tmp[1] + N[1]; bias.r[5] += tmp[2] + N[2];
    T u = jb_min(g[G_CHI], jb_max(g[G_CLO], ctrl));
    T force = g[G_GAIN] * u + g[G_BIASV] * (g[G_GEAR] * s.vm);
    fpass.mo = g[G_GEAR] * force - g[G_DAMPM] * s.vm;
  }
  // root block of M
  {
    const T mt = g[G_MTOT];
#pragma unroll
    for (int i = 0; i < 21; i++) Mx.R[i] = 0;
    Mx.R[JBQ_TRI(0, 0)] = mt; Mx.R[JBQ_TRI(1, 1)] = mt; Mx.R[JBQ_TRI(2, 2)] = mt;
    Mx.R[JBQ_TRI(3, 1)] = -ht[2]; Mx.R[JBQ_TRI(3, 2)] = ht[1];
    Mx.R[JBQ_TRI(4, 0)] = ht[2];  Mx.R[JBQ_TRI(4, 2)] = -ht[0];
    Mx.R[JBQ_TRI(5, 0)] = -ht[1]; Mx.R[JBQ_TRI(5, 1)] = ht[0];
    Mx.R[JBQ_TRI(3, 3)] = It[0]; Mx.R[JBQ_TRI(4, 4)] = It[1]; Mx.R[JBQ_TRI(5, 5)] = It[2];
    Mx.R[JBQ_TRI(4, 3)] = It[3]; Mx.R[JBQ_TRI(5, 3)] = It[4]; Mx.R[JBQ_TRI(5, 4)] = It[5];
  }

  JBQ_TICK(2);
  // ---------------- collision of the extra geoms: lanes 0,1 -> core boxes; lane 2 -> root cylinder + ellipsoid;
  //                  lane 3 -> rotor cylinder + ellipsoid
  if (lane < 2) {
    T c[3] = {LX(X_BOX_C), LX(X_BOX_C + 1), LX(X_BOX_C + 2)}, half[3] = {LX(X_BOX_H), LX(X_BOX_H + 1), LX(X_BOX_H + 2)};
    if (JBQ_UNLIKELY(fr.pz + dot3(fr.n, c) - jb_sqrt(dot3(half, half)) <= (T)0)) {
      T Rg[9];
#pragma unroll
      for (int i = 0; i < 9; i++) Rg[i] = LX(X_BOX_R + i);
      QBody<T> b = {QK_ROOT, nullptr, nullptr, (T)0, nullptr, nullptr, (T)0, false, g[G_INVW0]};
      const QFrame<T> frc = fr;
      const T cc[3] = {c[0], c[1], c[2]}, hh[3] = {half[0], half[1], half[2]};
      jbq_col_box<T>(g, cs, frc, b, (int)LX(X_BOX_ID), cc, Rg, hh, rep, lane);
    }
  } else {
    const bool rotor = lane == 3;
#define JBQ_BODY_X {rotor ? QK_ROTOR : QK_ROOT, am, pm, rotor ? s.vm : (T)0, nullptr, nullptr, (T)0, false, \
                    rotor ? g[G_INVWM] : g[G_INVW0]}
    T ch[3] = {LX(X_CYL_C), LX(X_CYL_C + 1), LX(X_CYL_C + 2)}, c[3], r[3], t[3];
    if (rotor) { sub3(r, ch, pm); mulMV(t, Rm, r); add3(c, pm, t); } else { c[0] = ch[0]; c[1] = ch[1]; c[2] = ch[2]; }
    T rr_ = LX(X_CYL_R), hl_ = LX(X_CYL_HL);
    if (JBQ_UNLIKELY(fr.pz + dot3(fr.n, c) - jb_sqrt(rr_ * rr_ + hl_ * hl_) <= (T)0)) {
      T gv[3] = {LX(X_CYL_AX), LX(X_CYL_AX + 1), LX(X_CYL_AX + 2)}, gx[3] = {LX(X_CYL_XA), LX(X_CYL_XA + 1), LX(X_CYL_XA + 2)};
      T ax[3], xa[3];
      if (rotor) { mulMV(ax, Rm, gv); mulMV(xa, Rm, gx); }
      else { ax[0] = gv[0]; ax[1] = gv[1]; ax[2] = gv[2]; xa[0] = gx[0]; xa[1] = gx[1]; xa[2] = gx[2]; }
      JBQ_COLD_ARGS(JBQ_BODY_X);
      const T cc[3] = {c[0], c[1], c[2]};
      jbq_col_cylinder<T>(g, cs, frc, bc, (int)LX(X_CYL_ID), cc, ax, xa, rr_, hl_, rep, lane);
    }
    ch[0] = LX(X_ELL_C); ch[1] = LX(X_ELL_C + 1); ch[2] = LX(X_ELL_C + 2);
    if (rotor) { sub3(r, ch, pm); mulMV(t, Rm, r); add3(c, pm, t); } else { c[0] = ch[0]; c[1] = ch[1]; c[2] = ch[2]; }
    T sz[3] = {LX(X_ELL_S), LX(X_ELL_S + 1), LX(X_ELL_S + 2)};
    if (JBQ_UNLIKELY(fr.pz + dot3(fr.n, c) - jb_max(sz[0], jb_max(sz[1], sz[2])) <= (T)0)) {
      T Rh[9], Rg[9];
#pragma unroll
      for (int i = 0; i < 9; i++) Rh[i] = LX(X_ELL_R + i);
      if (rotor) mulMM(Rg, Rm, Rh);
      else {
#pragma unroll
        for (int i = 0; i < 9; i++) Rg[i] = Rh[i];
      }
      JBQ_COLD_ARGS(JBQ_BODY_X);
      const T cc[3] = {c[0], c[1], c[2]}, ss[3] = {sz[0], sz[1], sz[2]};
      jbq_col_ellipsoid<T>(g, cs, frc, bc, (int)LX(X_ELL_ID), cc, Rg, ss, rep, lane);
    }
  }

#if defined(JBQ_STOP_AFTER) && JBQ_STOP_AFTER == 1      /* timing experiments only: kinematics + collision */
  s.wla += Mx.aa + Mx.R[3] + bias.r[0] + (T)cs.n;
  return;
#endif
  JBQ_TICK(3);
  // ---------------- smooth force: f_s = passive + actuator - bias.  The unconstrained acceleration a_s = M^-1 f_s is
  // never formed: the primal Newton solver only needs f_s (gradient M a - f_s - J'f), it starts from the warm start,
  // and the contact-free case goes straight to the implicit-damping Euler solve below.
  QVec<T> fs = fpass;
  qv_axpy(fs, (T)-1, bias);
  JBQ_TICK(4);
  const int ncon = cm.isum(cs.n);
  // this lane's share of the env's load record: sum of contacts | sub-steps with >= 3 | sub-steps with >= 2 | latest count
  if (lane_load)                         // bits 0-11 / 12-19 / 20-27 accumulate, bits 28-31: contacts at the latest sub-step
    *lane_load = (int)((((unsigned)*lane_load & 0x0fffffffu) + (unsigned)(cs.n + (cs.n >= 3 ? (1 << 12) : 0) + (cs.n >= 2 ? (1 << 20) : 0))) |
                       ((unsigned)cs.n << 28));

  // ---------------- constraint solve (primal Newton).  Control flow is kept WARP-uniform (Comm::warp_any): every quad
  // of the warp runs the same number of Newton / line-search iterations, quads that are finished take zero-length
  // steps.  That is what SIMT divergence would cost anyway, and it lets every quad shuffle run with the full mask.
  QVec<T> a, fcon;
  qv_zero(a); qv_zero(fcon);
  int iters = 0;
  const T mu = g[G_FRIC];
  const bool has_con = ncon > 0;
  if (cm.loop_any(has_con)) {              // uniform over the same scope as the Newton loop below
    {
      // start from qacc_warmstart (MuJoCo starts from the better of warm start and a_s; the minimiser is the same)
      T lw[3] = {s.wr[0], s.wr[1], s.wr[2]}, lr[3];
      mulMTV(lr, fr.Rw, lw);
      a.r[0] = lr[0]; a.r[1] = lr[1]; a.r[2] = lr[2]; a.r[3] = s.wr[3]; a.r[4] = s.wr[4]; a.r[5] = s.wr[5];
      a.la = s.wla; a.lb = s.wlb; a.mo = s.wmo;
    }
    QVec<T> Ma, grad;
    qarrow_mul(cm, Ma, Mx, a);
    T laf = 0, lbf = 0;                       // leg entries of the lane that owns this lane's foreign contacts
    if (JBQ_SPREAD && wf) { laf = cm.get(a.la, hv); lbf = cm.get(a.lb, hv); }
    jbq_apply_all(cs, fr, a, 0, laf, lbf);
    const T scale = (T)1 / (g[G_MEANINERTIA] * (T)JB_NV);
    const int maxit = JbEps<T>::newton_iters();
    bool done = !has_con, full_prev = false, exact = false;
    jbq_mask_t mask_prev = ~(jbq_mask_t)0;
    // Each trip: (1) forces at the current point, gradient, convergence test; (2) Hessian, Newton direction;
    // (3) exact line search; (4) move.  The final pass through (1) leaves the constraint force of the solution --
    // unless the step just taken was a full Newton step inside one active set (the usual case from a good warm start):
    // the cost is quadratic there, the new gradient is exactly H s + grad = 0, and the quad is finished without
    // another pass (`exact`).
    for (int trip = 0;; trip++) {
      if (trip > 0 && !cm.loop_any(!done)) break;
      T f8[8], fla, flb, ffa, ffb;
      const jbq_mask_t mask = jbq_forces(cs, fr, mu, f8, &fla, &flb, &ffa, &ffb);
      if (JBQ_SPREAD && wf) {                 // leg-dof forces of foreign contacts go home to lane hv
        cm.sum2(ffa, ffb);
        if (lane == hv) { fla += ffa; flb += ffb; }
      }
      T f9[9];
#pragma unroll
      for (int i = 0; i < 7; i++) f9[i] = f8[i];
      f9[7] = mask == mask_prev ? (T)0 : (T)1;
      grad = Ma;
      qv_axpy(grad, (T)-1, fs);
      grad.la -= fla; grad.lb -= flb;                 // leg entries are lane-local: their norm rides in the same reduce
      f9[8] = qv_dot_leg(grad, grad);
      cm.sumN(f9, 9);
#pragma unroll
      for (int i = 0; i < 6; i++) grad.r[i] -= f9[i];
      grad.mo -= f9[6];
      f8[7] = f9[7];
      if (exact) qv_zero(grad);            // keeps the result independent of what the warp's other quads still do
      const T gn = scale * jb_sqrt(f9[8] + qv_dot_rep(grad, grad));
      if (!done && iters > 0 && ((full_prev && f8[7] == (T)0) || gn < JbEps<T>::newton_tol() || iters >= maxit)) done = true;
      if (!cm.loop_any(!done)) break;        // CTA- or warp-uniform trip count (Comm decides)
#ifdef JBQ_DEBUG_COUNT
      if (lane == 0) g_jbq_newton_trips++;
#endif
      QArrow<T> H = Mx;
      if (lane != 0) {                   // lane 0's share carries M's replicated block (see qarrow_factor)
#pragma unroll
        for (int i = 0; i < 21; i++) H.R[i] = 0;
#pragma unroll
        for (int i = 0; i < 6; i++) H.cm[i] = 0;
        H.mm = 0;
      }
      {
        T F15[15];
        jbq_hessian(cs, fr, mu, mask, H, F15);
        if (JBQ_SPREAD && wf) {               // and so do their leg-block Hessian terms
          cm.sumN(F15, 15);
          if (lane == hv) {
#pragma unroll
            for (int q = 0; q < 6; q++) { H.ca[q] += F15[q]; H.cb[q] += F15[7 + q]; }
            H.aa += F15[6]; H.ab += F15[13]; H.bb += F15[14];
          }
        }
      }
      cm.phase_sync(8);
      qarrow_factor<T, Comm>(cm, H, true);
      cm.phase_sync(9);
      QVec<T> search;
      qarrow_solve(cm, search, H, grad);
#pragma unroll
      for (int i = 0; i < 6; i++) search.r[i] = -search.r[i];
      search.la = -search.la; search.lb = -search.lb; search.mo = -search.mo;
      QVec<T> Mv;
      qarrow_mul(cm, Mv, Mx, search);
      if (JBQ_SPREAD && wf) { laf = cm.get(search.la, hv); lbf = cm.get(search.lb, hv); }
      jbq_apply_all(cs, fr, search, 1, laf, lbf);
      cm.phase_sync(10);
      QVec<T> res = Ma; qv_axpy(res, (T)-1, fs);
      // g1 = s.(Ma - fs), g2 = s.Mv, d0 = s.grad (= phi'(0)); at alpha = 1: contact part of phi' and phi''
      T red[6];
      red[0] = qv_dot_leg(search, res); red[1] = qv_dot_leg(search, Mv); red[2] = qv_dot_leg(search, grad);
      jbq_mask_t m1 = jbq_ls_local(cs, mu, (T)1, &red[3], &red[4]);
      red[5] = m1 == mask ? (T)0 : (T)1;
      cm.sumN(red, 6);
      const T g1 = red[0] + qv_dot_rep(search, res), g2 = red[1] + qv_dot_rep(search, Mv);
      const T d0 = red[2] + qv_dot_rep(search, grad);
      const bool descent = d0 < (T)0;                    // otherwise: converged to rounding
      const bool full = red[5] == (T)0;                  // same active set at alpha = 1: the Newton step is exact
      T alpha = 1;
      {
        // exact line search on the piecewise-quadratic 1-D cost: safeguarded Newton on phi'
        T d = g1 + g2 + red[3], dd = g2 + red[4];
        T lo = 0, hi = -1;
        const T tol = JbEps<T>::ls_tol() * jb_abs(d0);
        bool searching = !done && descent && !full && jb_abs(d) > tol;
        for (int it = 0; it < 30 && cm.warp_any(searching); it++) {
#ifdef JBQ_DEBUG_COUNT
          if (lane == 0) g_jbq_ls_iters++;
#endif
          T an = alpha;
          if (searching) {
            if (d < (T)0) lo = alpha; else hi = alpha;
            an = alpha - d / dd;
            if (hi > (T)0 && !(an > lo && an < hi)) an = (T)0.5 * (lo + hi);
            else if (!(an > lo)) an = alpha * (T)2;
            if (an == alpha) searching = false;
          }
          T dl, hl;
          jbq_ls_local(cs, mu, an, &dl, &hl);
          cm.sum2(dl, hl);
          if (searching) {
            alpha = an;
            d = g1 + alpha * g2 + dl; dd = g2 + hl;
            if (jb_abs(d) <= tol) searching = false;
          }
        }
      }
      if (!done && !descent) done = true;
      if (done) alpha = 0;
      else iters++;
      cm.phase_sync(11);
      qv_axpy(a, alpha, search);
      qv_axpy(Ma, alpha, Mv);
      // J a - uref at the new point: only a quad that goes on iterating reads it again (a finished quad takes zero-length
      // steps, a full step inside one active set ends the solve exactly)
      if (!done && !full) {
        for (int i = 0; i < cs.n; i++) {
          QCon<T>& c = cs.c[i];
          c.ww[0] += alpha * c.jv[0]; c.ww[1] += alpha * c.jv[1]; c.ww[2] += alpha * c.jv[2];
        }
      }
      mask_prev = mask; full_prev = full;
      if (!done && full) { done = true; exact = true; qv_zero(grad); }
    }
    // constraint force at the solution: f_c = M a - f_s - grad
    fcon = Ma;
    qv_axpy(fcon, (T)-1, fs); qv_axpy(fcon, (T)-1, grad);
  }
  JBQ_TICK(6);
  if (iter_accum) *iter_accum += iters;
  if (ncon_accum) *ncon_accum += ncon;
  if (overflow_accum) *overflow_accum += cs.overflow;          // per lane; the caller sums

  // ---------------- semi-implicit Euler, implicit in joint damping: (M + hB) a' = f_s + f_c  (mj_Euler)
  QVec<T> aeu;
  {
    QVec<T> rhs = fs;
    qv_axpy(rhs, (T)1, fcon);
    QArrow<T> F = Mx;
    F.aa += h * LL(L_D1); F.bb += h * LL(L_D2); F.mm += h * g[G_DAMPM];
    qarrow_factor<T, Comm>(cm, F, false);
    qarrow_solve(cm, aeu, F, rhs);
  }
  // ---------------- save warm start (MuJoCo coordinates).  Without contacts qacc = M^-1 f_s in MuJoCo; the Euler
  // acceleration (which differs by the O(hB/M) damping term) is kept instead: it is only the next solve's initial guess.
  if (!has_con) a = aeu;
  {
    T lw[3];
    mulMV(lw, fr.Rw, a.r);
    s.wr[0] = lw[0]; s.wr[1] = lw[1]; s.wr[2] = lw[2]; s.wr[3] = a.r[3]; s.wr[4] = a.r[4]; s.wr[5] = a.r[5];
    s.wla = a.la; s.wlb = a.lb; s.wmo = a.mo;
  }
  a = aeu;
  {
    T lw[3];
    mulMV(lw, fr.Rw, a.r);
    s.vlin[0] += h * lw[0]; s.vlin[1] += h * lw[1]; s.vlin[2] += h * lw[2];
    s.w[0] += h * a.r[3]; s.w[1] += h * a.r[4]; s.w[2] += h * a.r[5];
    s.va += h * a.la; s.vb += h * a.lb; s.vm += h * a.mo;
  }
  s.pos[0] += h * s.vlin[0]; s.pos[1] += h * s.vlin[1]; s.pos[2] += h * s.vlin[2];
  {
    T wn = jb_sqrt(s.w[0] * s.w[0] + s.w[1] * s.w[1] + s.w[2] * s.w[2]);
    T ax[3] = {1, 0, 0};
    T ang = 0;
    if (wn >= (T)1e-15) { ax[0] = s.w[0] / wn; ax[1] = s.w[1] / wn; ax[2] = s.w[2] / wn; ang = h * wn; }
    T sh, ch;
    jb_sincos_fast((T)0.5 * ang, &sh, &ch);
    T q[4] = {s.quat[0], s.quat[1], s.quat[2], s.quat[3]};
    T dq[4] = {ch, ax[0] * sh, ax[1] * sh, ax[2] * sh};
    s.quat[0] = q[0] * dq[0] - q[1] * dq[1] - q[2] * dq[2] - q[3] * dq[3];
    s.quat[1] = q[0] * dq[1] + q[1] * dq[0] + q[2] * dq[3] - q[3] * dq[2];
    s.quat[2] = q[0] * dq[2] - q[1] * dq[3] + q[2] * dq[0] + q[3] * dq[1];
    s.quat[3] = q[0] * dq[3] + q[1] * dq[2] - q[2] * dq[1] + q[3] * dq[0];
  }
  s.qa += h * s.va; s.qb += h * s.vb;
  s.qm += h * s.vm;
  if (s.qm > (T)3.14159265358979323846) { s.qm -= (T)6.28318530717958647692; s.rev += 1; }
  else if (s.qm <= -(T)3.14159265358979323846) { s.qm += (T)6.28318530717958647692; s.rev -= 1; }
  JBQ_TICK(7);
#undef LX
#undef LL
}

#endif  // JB_QUAD_CUH
