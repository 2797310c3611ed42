// TEST INFRASTRUCTURE (round-1 v1 kernel body, kept as an independent second implementation and as the flop-count
// instrument; the product kernel is jitterbug_b200/csrc/jb_quad.cuh).
// One MuJoCo-equivalent physics sub-step (mj_step2 ; mj_step1 of dm_control's Physics.step(), reference call
// site jitterbug_dmc/jitterbug.py:84-90 -> control.Environment -> 50 x physics.step()) for ONE Jitterbug, written
// for one GPU thread per environment.  Everything is expressed in the root body's frame with the root origin as
// reference point ("body-frame, product-of-exponentials" formulation):
//
//   * generalized accelerations a' = (R^T a_lin_world, alpha_body, qdd_legs[8], qdd_motor) -- MuJoCo's qacc with the
//     free joint's linear block rotated into the body frame.  The constraint cost 1/2 (a-a_s)' M (a-a_s) + s(J a - aref)
//     is invariant under this orthogonal change of coordinates, so the minimiser maps back exactly.
//   * the joint-space inertia has MuJoCo's tree sparsity (88 non-zeros): a dense 6x6 root block plus, per leg, a 2x2
//     block coupled to the root, plus the motor dof -- an "arrow" matrix that is factorised leaf-first exactly like
//     MuJoCo's sparse L'DL.  The Newton Hessian M + J' D J has the same pattern because every contact Jacobian row only
//     touches the root dofs and the dofs of one leg (or the motor).
//   * bias forces come from a 3-vector Newton-Euler recursion with the fictitious acceleration -g at the root.
//   * contacts: floor plane vs sphere / cylinder / box / ellipsoid in root coordinates; pyramidal friction cone
//     (condim 3): rows n +- mu*t1, n +- mu*t2 with n,t1,t2 = world z,y,x.
//
// The file is plain C++ under a __host__ __device__ macro so tests can instantiate the *same* code in fp64 on the CPU
// (tests/hostsim, test-only) to separate algorithmic parity from fp32 rounding; the product only instantiates
// jb_substep<float> inside the CUDA kernels of jb_kernels.cu.
#ifndef JB_DYNAMICS_CUH
#define JB_DYNAMICS_CUH

#include "../../jitterbug_b200/csrc/jb_model.h"
#include "../../jitterbug_b200/csrc/jb_math.cuh"

#define JB_CAP_LEG 6       /* contact slots per leg (foot 1 + lower cylinder 4 + 1) */
#define JB_CAP_ROOT 8
#define JB_CAP_ROTOR 5
#define JB_NGROUP 6        /* 4 legs, root, rotor */
#define JB_CAP_MAX 8

// ------------------------------------------------------------------------------------------------ state, vectors, arrow matrix
template <typename T> struct JbState {     // MuJoCo coordinates and ordering
  T pos[3], quat[4], ql[JB_NLEG][2], qm;   // qpos (16)
  T vlin[3], w[3], vl[JB_NLEG][2], vm;     // qvel (15): world linear, body angular, joints
  T warm[JB_NV];                           // qacc_warmstart
  int rev;                                 // whole motor revolutions: qpos[15] = qm + 2*pi*rev, qm kept in (-pi, pi]
};

template <typename T> struct Vec15 {
  T r[6]; T la[JB_NLEG], lb[JB_NLEG]; T mo;
};
template <typename T> JB_HD void v15_zero(Vec15<T>& v) {
#pragma unroll
  for (int i = 0; i < 6; i++) v.r[i] = 0;
#pragma unroll
  for (int k = 0; k < JB_NLEG; k++) { v.la[k] = 0; v.lb[k] = 0; }
  v.mo = 0;
}
template <typename T> JB_HD T v15_dot(const Vec15<T>& a, const Vec15<T>& b) {
  T s = a.mo * b.mo;
#pragma unroll
  for (int i = 0; i < 6; i++) s += a.r[i] * b.r[i];
#pragma unroll
  for (int k = 0; k < JB_NLEG; k++) s += a.la[k] * b.la[k] + a.lb[k] * b.lb[k];
  return s;
}
template <typename T> JB_HD void v15_axpy(Vec15<T>& y, T a, const Vec15<T>& x) {
#pragma unroll
  for (int i = 0; i < 6; i++) y.r[i] += a * x.r[i];
#pragma unroll
  for (int k = 0; k < JB_NLEG; k++) { y.la[k] += a * x.la[k]; y.lb[k] += a * x.lb[k]; }
  y.mo += a * x.mo;
}

#define JB_TRI(i, j) ((i) * ((i) + 1) / 2 + (j))   /* lower triangle, i >= j */

template <typename T> struct Arrow {
  T R[21];                                       // root 6x6 block, lower triangle
  T aa[JB_NLEG], ab[JB_NLEG], bb[JB_NLEG];       // leg 2x2 blocks
  T ca[JB_NLEG][6], cb[JB_NLEG][6];              // leg-root coupling
  T mm, cm[6];                                   // motor dof
};

// y = A x
template <typename T> JB_HD void arrow_mul(Vec15<T>& y, const Arrow<T>& A, const Vec15<T>& x) {
#pragma unroll
  for (int i = 0; i < 6; i++) {
    T s = 0;
#pragma unroll
    for (int j = 0; j < 6; j++) s += A.R[i >= j ? JB_TRI(i, j) : JB_TRI(j, i)] * x.r[j];
#pragma unroll
    for (int k = 0; k < JB_NLEG; k++) s += A.ca[k][i] * x.la[k] + A.cb[k][i] * x.lb[k];
    s += A.cm[i] * x.mo;
    y.r[i] = s;
  }
#pragma unroll
  for (int k = 0; k < JB_NLEG; k++) {
    T sa = A.aa[k] * x.la[k] + A.ab[k] * x.lb[k], sb = A.ab[k] * x.la[k] + A.bb[k] * x.lb[k];
#pragma unroll
    for (int i = 0; i < 6; i++) { sa += A.ca[k][i] * x.r[i]; sb += A.cb[k][i] * x.r[i]; }
    y.la[k] = sa; y.lb[k] = sb;
  }
  T sm = A.mm * x.mo;
#pragma unroll
  for (int i = 0; i < 6; i++) sm += A.cm[i] * x.r[i];
  y.mo = sm;
}

// In-place leaf-first L'DL of an arrow matrix (same elimination order as MuJoCo's mj_factorM: last dof first).
// After the call: bb,aa,mm hold reciprocals of the pivots, ab/cb/ca/cm the unit-triangular factors, R the root LDL'
// (unit L below the diagonal, reciprocal pivots on it).
template <typename T> JB_HD void arrow_factor(Arrow<T>& A) {
#pragma unroll
  for (int k = 0; k < JB_NLEG; k++) {
    T ibb = (T)1 / A.bb[k];
    T lab = A.ab[k] * ibb;
    T Lb[6];
#pragma unroll
    for (int i = 0; i < 6; i++) Lb[i] = A.cb[k][i] * ibb;
    A.aa[k] -= lab * A.ab[k];
#pragma unroll
    for (int i = 0; i < 6; i++) A.ca[k][i] -= lab * A.cb[k][i];
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
      for (int j = 0; j <= i; j++) A.R[JB_TRI(i, j)] -= Lb[i] * A.cb[k][j];
    T iaa = (T)1 / A.aa[k];
    T La[6];
#pragma unroll
    for (int i = 0; i < 6; i++) La[i] = A.ca[k][i] * iaa;
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
      for (int j = 0; j <= i; j++) A.R[JB_TRI(i, j)] -= La[i] * A.ca[k][j];
    A.bb[k] = ibb; A.ab[k] = lab; A.aa[k] = iaa;
#pragma unroll
    for (int i = 0; i < 6; i++) { A.cb[k][i] = Lb[i]; A.ca[k][i] = La[i]; }
  }
  {
    T imm = (T)1 / A.mm;
    T Lm[6];
#pragma unroll
    for (int i = 0; i < 6; i++) Lm[i] = A.cm[i] * imm;
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
      for (int j = 0; j <= i; j++) A.R[JB_TRI(i, j)] -= Lm[i] * A.cm[j];
    A.mm = imm;
#pragma unroll
    for (int i = 0; i < 6; i++) A.cm[i] = Lm[i];
  }
  // dense LDL' of the 6x6 Schur complement (unit L below the diagonal, reciprocal pivots on it)
  T dv[6];
#pragma unroll
  for (int j = 0; j < 6; j++) {
    T d = A.R[JB_TRI(j, j)];
    T ld[6];
#pragma unroll
    for (int k = 0; k < j; k++) { ld[k] = A.R[JB_TRI(j, k)] * dv[k]; d -= A.R[JB_TRI(j, k)] * ld[k]; }
    dv[j] = d;
    T id = (T)1 / d;
#pragma unroll
    for (int i = j + 1; i < 6; i++) {
      T s = A.R[JB_TRI(i, j)];
#pragma unroll
      for (int k = 0; k < j; k++) s -= A.R[JB_TRI(i, k)] * ld[k];
      A.R[JB_TRI(i, j)] = s * id;
    }
    A.R[JB_TRI(j, j)] = id;
  }
}

// x = A^-1 r with A factorised by arrow_factor
template <typename T> JB_HD void arrow_solve(Vec15<T>& x, const Arrow<T>& F, const Vec15<T>& r) {
  T rr[6];
#pragma unroll
  for (int i = 0; i < 6; i++) rr[i] = r.r[i];
  T ra[JB_NLEG];
#pragma unroll
  for (int k = 0; k < JB_NLEG; k++) {
    ra[k] = r.la[k] - F.ab[k] * r.lb[k];
#pragma unroll
    for (int i = 0; i < 6; i++) rr[i] -= F.cb[k][i] * r.lb[k] + F.ca[k][i] * ra[k];
  }
#pragma unroll
  for (int i = 0; i < 6; i++) rr[i] -= F.cm[i] * r.mo;
  // root: L y = rr ; z = y / d ; L' x = z
#pragma unroll
  for (int i = 0; i < 6; i++) {
#pragma unroll
    for (int k = 0; k < i; k++) rr[i] -= F.R[JB_TRI(i, k)] * rr[k];
  }
#pragma unroll
  for (int i = 0; i < 6; i++) rr[i] *= F.R[JB_TRI(i, i)];
#pragma unroll
  for (int i = 5; i >= 0; i--) {
#pragma unroll
    for (int k = i + 1; k < 6; k++) rr[i] -= F.R[JB_TRI(k, i)] * rr[k];
  }
#pragma unroll
  for (int i = 0; i < 6; i++) x.r[i] = rr[i];
#pragma unroll
  for (int k = 0; k < JB_NLEG; k++) {
    T xa = ra[k] * F.aa[k], xb = r.lb[k] * F.bb[k];
#pragma unroll
    for (int i = 0; i < 6; i++) { xa -= F.ca[k][i] * rr[i]; xb -= F.cb[k][i] * rr[i]; }
    xb -= F.ab[k] * xa;
    x.la[k] = xa; x.lb[k] = xb;
  }
  T xm = r.mo * F.mm;
#pragma unroll
  for (int i = 0; i < 6; i++) xm -= F.cm[i] * rr[i];
  x.mo = xm;
}

// ------------------------------------------------------------------------------------------------ contacts
template <typename T> struct Contact {
  T x[3];        // contact point, root coordinates relative to the root origin
  T D;           // 1 / Rpy (same for the 4 pyramid rows)
  T ar0;         // -K * imp * dist
  T uref[3];     // -B * (world-frame velocity of the contact point), WORLD components
  T low;         // 1 if the contact is on a lower leg (both leg dofs), 0 for an upper leg (first dof only)
  T ww[3];       // solver scratch: J_c a - uref, WORLD components
  T jv[3];       // solver scratch: J_c search, WORLD components
  T dist; int geom;   // reporting only
};

template <typename T> struct ContactSet {
  Contact<T> c[JB_NGROUP][JB_CAP_MAX];
  int n[JB_NGROUP];
  int overflow;
};
JB_HD int jb_group_cap(int g) { return g < 4 ? JB_CAP_LEG : (g == 4 ? JB_CAP_ROOT : JB_CAP_ROTOR); }

// MuJoCo's solimp sigmoid (getimpedance), impedance only
template <typename T, typename M> JB_HD T jb_impedance(const M& m, T dist) {
  T d0 = (T)m.solimp[0], d1 = (T)m.solimp[1], width = (T)m.solimp[2], mid = (T)m.solimp[3], power = (T)m.solimp[4];
  d0 = jb_min((T)0.9999, jb_max((T)0.0001, d0));
  d1 = jb_min((T)0.9999, jb_max((T)0.0001, d1));
  mid = jb_min((T)0.9999, jb_max((T)0.0001, mid));
  if (d0 == d1 || width <= (T)1e-15) return (T)0.5 * (d0 + d1);
  T x = jb_abs(dist) / width;
  if (x >= (T)1) return d1;
  if (x <= (T)0) return d0;
  T y;
  if (power == (T)2) {                     // the model's value; avoids pow()
    y = x <= mid ? x * x / mid : (T)1 - ((T)1 - x) * ((T)1 - x) / ((T)1 - mid);
  } else if (power <= (T)1) {
    y = x;
  } else {
#if defined(__CUDA_ARCH__)
    y = x <= mid ? powf((float)x, (float)power) / powf((float)mid, (float)power - 1.f)
                 : 1.f - powf(1.f - (float)x, (float)power) / powf(1.f - (float)mid, (float)power - 1.f);
#else
    y = x <= mid ? (T)(pow((double)x, (double)power) / pow((double)mid, (double)power - 1))
                 : (T)(1 - pow(1 - (double)x, (double)power) / pow(1 - (double)mid, (double)power - 1));
#endif
  }
  return d0 + y * (d1 - d0);
}

// Everything the narrow-phase needs to turn a penetrating point into a solver contact.
template <typename T> struct ConCtx {
  T pz;            // world height of the root origin
  T n[3];          // world z axis in root coordinates (floor normal)
  T Rw[9];         // root rotation, body->world
  T vr[3], w[3];   // root linear velocity (root comps), angular velocity (root comps)
};

// add a contact at root-frame point x (already the MuJoCo mid-surface point) with signed distance `dist`
template <typename T, typename M>
JB_HD void jb_add_contact(const M& m, ContactSet<T>& cs, int group, int geom, int slot, const ConCtx<T>& cx,
                          const T* x, T dist, const T* vpt_root, T low) {
  // MuJoCo keeps contacts with dist <= margin (=0) in d->contact but excludes dist >= includemargin from the solver
  if (!(dist < (T)0)) return;
  int i = cs.n[group];
  if (i >= jb_group_cap(group)) { cs.overflow = 1; return; }
  cs.n[group] = i + 1;
  Contact<T>& c = cs.c[group][i];
  c.x[0] = x[0]; c.x[1] = x[1]; c.x[2] = x[2];
  T imp = jb_impedance<T>(m, dist);
  T mu = (T)m.friction;
  T tran = (T)m.invw[slot];
  // diagApprox = tran*(1+mu^2); R0 = (1-imp)/imp*diagApprox; Rpy = 2*mu_reg^2*R0, mu_reg = mu/sqrt(impratio)
  T Rpy = (T)2 * mu * mu / (T)m.impratio * ((T)1 - imp) / imp * tran * ((T)1 + mu * mu);
  Rpy = jb_max(Rpy, (T)1e-15);
  c.D = (T)1 / Rpy;
  c.ar0 = -(T)m.solref_K * imp * dist;
  T vw[3];
  mulMV(vw, cx.Rw, vpt_root);
  c.uref[0] = -(T)m.solref_B * vw[0]; c.uref[1] = -(T)m.solref_B * vw[1]; c.uref[2] = -(T)m.solref_B * vw[2];
  c.low = low;
  c.dist = dist; c.geom = geom;
}

// velocity (root comps) of a point x rigidly attached to: root (nj=0), first joint body (nj=1), second joint body (nj=2)
template <typename T>
JB_HD void jb_point_vel(T* v, const ConCtx<T>& cx, const T* x, int nj, const T* a1, const T* p1, T qd1,
                        const T* a2, const T* p2, T qd2) {
  T t[3];
  cross3(t, cx.w, x);
  v[0] = cx.vr[0] + t[0]; v[1] = cx.vr[1] + t[1]; v[2] = cx.vr[2] + t[2];
  if (nj >= 1) { T r[3]; sub3(r, x, p1); cross3(t, a1, r); addscl3(v, t, qd1); }
  if (nj >= 2) { T r[3]; sub3(r, x, p2); cross3(t, a2, r); addscl3(v, t, qd2); }
}

// joint description handed to the narrow phase so it can evaluate contact-point velocities
template <typename T> struct BodyJ {
  int nj; const T* a1; const T* p1; T qd1; const T* a2; const T* p2; T qd2; T low; int group; int slot;
};

template <typename T, typename M>
JB_HD void jb_emit(const M& m, ContactSet<T>& cs, const ConCtx<T>& cx, const BodyJ<T>& bj, int geom, const T* x, T dist) {
  T v[3];
  jb_point_vel(v, cx, x, bj.nj, bj.a1, bj.p1, bj.qd1, bj.a2, bj.p2, bj.qd2);
  jb_add_contact<T>(m, cs, bj.group, geom, bj.slot, cx, x, dist, v, bj.low);
}

// plane-sphere (mjc_PlaneSphere); c = sphere centre in root coordinates
template <typename T, typename M>
JB_HD void jb_col_sphere(const M& m, ContactSet<T>& cs, const ConCtx<T>& cx, const BodyJ<T>& bj, int geom, const T* c, T r) {
  T dist = cx.pz + dot3(cx.n, c) - r;
  if (dist > (T)0) return;
  T x[3];
  T s = r + (T)0.5 * dist;
  x[0] = c[0] - cx.n[0] * s; x[1] = c[1] - cx.n[1] * s; x[2] = c[2] - cx.n[2] * s;
  jb_emit<T>(m, cs, cx, bj, geom, x, dist);
}

// plane-cylinder (mjc_PlaneCylinder); c centre, ax unit axis (geom z), xax geom x axis, all root coordinates
template <typename T, typename M>
JB_HD void jb_col_cylinder(const M& m, ContactSet<T>& cs, const ConCtx<T>& cx, const BodyJ<T>& bj, int geom,
                           const T* c, const T* ax_in, const T* xax, T r, T hl) {
  T dist0 = cx.pz + dot3(cx.n, c);
  if (dist0 - jb_sqrt(r * r + hl * hl) > (T)0) return;      // bounding sphere
  T ax[3] = {ax_in[0], ax_in[1], ax_in[2]};
  T prjaxis = dot3(cx.n, ax);
  if (prjaxis > (T)0) { ax[0] = -ax[0]; ax[1] = -ax[1]; ax[2] = -ax[2]; prjaxis = -prjaxis; }
  T vec[3] = {ax[0] * prjaxis - cx.n[0], ax[1] * prjaxis - cx.n[1], ax[2] * prjaxis - cx.n[2]};
  T len_sqr = dot3(vec, vec);
  if (len_sqr >= (T)1e-30) {
    T s = r / jb_sqrt(len_sqr);
    vec[0] *= s; vec[1] *= s; vec[2] *= s;
  } else {
    vec[0] = xax[0] * r; vec[1] = xax[1] * r; vec[2] = xax[2] * r;
  }
  T prjvec = dot3(vec, cx.n);
  ax[0] *= hl; ax[1] *= hl; ax[2] *= hl;
  prjaxis *= hl;
  T d1 = dist0 + prjaxis + prjvec;
  if (d1 > (T)0) return;
  T x[3];
#pragma unroll
  for (int k = 0; k < 3; k++) x[k] = c[k] + vec[k] + ax[k] - cx.n[k] * d1 * (T)0.5;
  jb_emit<T>(m, cs, cx, bj, geom, x, d1);
  T d2 = dist0 - prjaxis + prjvec;
  if (d2 <= (T)0) {
#pragma unroll
    for (int k = 0; k < 3; k++) x[k] = c[k] + vec[k] - ax[k] - cx.n[k] * d2 * (T)0.5;
    jb_emit<T>(m, cs, cx, bj, geom, x, d2);
  }
  T prjvec1 = -prjvec * (T)0.5;
  T d3 = dist0 + prjaxis + prjvec1;
  if (d3 <= (T)0) {
    T v1[3];
    cross3(v1, vec, ax);
    T nn = jb_sqrt(dot3(v1, v1));
    if (nn < (T)1e-15) { v1[0] = 1; v1[1] = 0; v1[2] = 0; nn = 1; }
    T s = r * (T)0.8660254037844386 / nn;
    v1[0] *= s; v1[1] *= s; v1[2] *= s;
#pragma unroll
    for (int k = 0; k < 3; k++) x[k] = c[k] + v1[k] + ax[k] - vec[k] * (T)0.5 - cx.n[k] * d3 * (T)0.5;
    jb_emit<T>(m, cs, cx, bj, geom, x, d3);
#pragma unroll
    for (int k = 0; k < 3; k++) x[k] = c[k] - v1[k] + ax[k] - vec[k] * (T)0.5 - cx.n[k] * d3 * (T)0.5;
    jb_emit<T>(m, cs, cx, bj, geom, x, d3);
  }
}

// plane-box (mjc_PlaneBox); Rg = geom->root rotation
template <typename T, typename M>
JB_HD void jb_col_box(const M& m, ContactSet<T>& cs, const ConCtx<T>& cx, const BodyJ<T>& bj, int geom,
                      const T* c, const T* Rg, const T* half) {
  T dist = cx.pz + dot3(cx.n, c);
  if (dist - jb_sqrt(dot3(half, half)) > (T)0) return;
  int cnt = 0;
  for (int i = 0; i < 8; i++) {
    T v[3] = {(i & 1) ? half[0] : -half[0], (i & 2) ? half[1] : -half[1], (i & 4) ? half[2] : -half[2]};
    T corner[3];
    mulMV(corner, Rg, v);
    T ldist = dot3(cx.n, corner);
    if (dist + ldist > (T)0 || ldist > (T)0) continue;
    T cd = dist + ldist;
    T x[3];
#pragma unroll
    for (int k = 0; k < 3; k++) x[k] = corner[k] + c[k] - cx.n[k] * cd * (T)0.5;
    jb_emit<T>(m, cs, cx, bj, geom, x, cd);
    if (++cnt >= 4) return;
  }
}

// plane-ellipsoid (mjc_PlaneConvex with the ellipsoid support mapping)
template <typename T, typename M>
JB_HD void jb_col_ellipsoid(const M& m, ContactSet<T>& cs, const ConCtx<T>& cx, const BodyJ<T>& bj, int geom,
                            const T* c, const T* Rg, const T* size) {
  T d0 = cx.pz + dot3(cx.n, c);
  if (d0 - jb_max(size[0], jb_max(size[1], size[2])) > (T)0) return;
  T dir[3] = {-cx.n[0], -cx.n[1], -cx.n[2]}, loc[3];
  mulMTV(loc, Rg, dir);
  loc[0] *= size[0]; loc[1] *= size[1]; loc[2] *= size[2];
  T nrm = jb_sqrt(dot3(loc, loc));
  if (nrm < (T)1e-15) nrm = (T)1e-15;
  loc[0] = loc[0] / nrm * size[0]; loc[1] = loc[1] / nrm * size[1]; loc[2] = loc[2] / nrm * size[2];
  T wv[3], sup[3];
  mulMV(wv, Rg, loc);
  add3(sup, wv, c);
  T dist = cx.pz + dot3(cx.n, sup);
  if (dist > (T)0) return;
  T x[3];
#pragma unroll
  for (int k = 0; k < 3; k++) x[k] = sup[k] - cx.n[k] * (T)0.5 * dist;
  jb_emit<T>(m, cs, cx, bj, geom, x, dist);
}

// ------------------------------------------------------------------------------------------------ per-substep scratch
template <typename T> struct LegKin {       // what the solver keeps per leg after the position pass
  T a1[3], p1[3], a2[3], p2[3];
};

template <typename T> struct ContactReport { int n; int geom[24]; T dist[24]; };   // parity hook (jb_sim_get_contacts)

template <typename T> struct SubstepDiag {  // optional dump for parity tests
  T M[JB_NV][JB_NV];       // dense inertia, internal coordinates / ordering [lin, ang, (a,b)x4, motor]
  T bias[JB_NV], fsmooth[JB_NV], asmooth[JB_NV], qacc[JB_NV], fcon[JB_NV];
  T Rw[9];
  int ncon, geom[64]; T dist[64]; T cpos[64][3];
  int newton_iters, overflow;
};

// Jacobian-transpose accumulation: out += J_c^T f for a contact in `group`; f in ROOT components
template <typename T>
JB_HD void jb_JT_accum(Vec15<T>& out, int group, const Contact<T>& c, const LegKin<T>* lk, const T* pm, const T* am, const T* f) {
  T t[3];
  out.r[0] += f[0]; out.r[1] += f[1]; out.r[2] += f[2];
  cross3(t, c.x, f);
  out.r[3] += t[0]; out.r[4] += t[1]; out.r[5] += t[2];
  if (group < 4) {
    T r[3];
    sub3(r, c.x, lk[group].p1); cross3(t, r, f); out.la[group] += dot3(lk[group].a1, t);
    sub3(r, c.x, lk[group].p2); cross3(t, r, f); out.lb[group] += c.low * dot3(lk[group].a2, t);
  } else if (group == 5) {
    T r[3];
    sub3(r, c.x, pm); cross3(t, r, f); out.mo += dot3(am, t);
  }
}

// u = J_c a (acceleration of the contact point generated by generalized acceleration a), ROOT components
template <typename T>
JB_HD void jb_J_apply(T* u, int group, const Contact<T>& c, const LegKin<T>* lk, const T* pm, const T* am, const Vec15<T>& a) {
  T t[3];
  cross3(t, a.r + 3, c.x);
  u[0] = a.r[0] + t[0]; u[1] = a.r[1] + t[1]; u[2] = a.r[2] + t[2];
  if (group < 4) {
    T r[3];
    sub3(r, c.x, lk[group].p1); cross3(t, lk[group].a1, r); addscl3(u, t, a.la[group]);
    sub3(r, c.x, lk[group].p2); cross3(t, lk[group].a2, r); addscl3(u, t, c.low * a.lb[group]);
  } else if (group == 5) {
    T r[3];
    sub3(r, c.x, pm); cross3(t, am, r); addscl3(u, t, a.mo);
  }
}

// the four pyramid rows of a contact, given w = (J a - uref) in WORLD components: jar_r = w_z +- mu*w_{y|x} - ar0
template <typename T> JB_HD void jb_rows(T* jar, const T* ww, T mu, T ar0) {
  jar[0] = ww[2] + mu * ww[1] - ar0;
  jar[1] = ww[2] - mu * ww[1] - ar0;
  jar[2] = ww[2] - mu * ww[0] - ar0;    // t2 = n x t1 = -x
  jar[3] = ww[2] + mu * ww[0] - ar0;
}
// row directions in WORLD components: e_r = z + s*mu*(y or -x)
template <typename T> JB_HD void jb_row_dir(T* e, int r, T mu) {
  e[0] = r == 2 ? -mu : (r == 3 ? mu : (T)0);
  e[1] = r == 0 ? mu : (r == 1 ? -mu : (T)0);
  e[2] = (T)1;
}

// constraint cost + (optionally) forces at the current ww of every contact
template <typename T>
JB_HD T jb_constraint_cost(const ContactSet<T>& cs, T mu) {
  T cost = 0;
#pragma unroll
  for (int g = 0; g < JB_NGROUP; g++)
    for (int i = 0; i < cs.n[g]; i++) {
      const Contact<T>& c = cs.c[g][i];
      T jar[4];
      jb_rows(jar, c.ww, mu, c.ar0);
#pragma unroll
      for (int r = 0; r < 4; r++) if (jar[r] < (T)0) cost += (T)0.5 * c.D * jar[r] * jar[r];
    }
  return cost;
}

// set ww = R_w (J_c a) - uref for every contact
template <typename T>
JB_HD void jb_eval_ww(ContactSet<T>& cs, const T* Rw, const LegKin<T>* lk, const T* pm, const T* am, const Vec15<T>& a) {
#pragma unroll
  for (int g = 0; g < JB_NGROUP; g++)
    for (int i = 0; i < cs.n[g]; i++) {
      Contact<T>& c = cs.c[g][i];
      T u[3], uw[3];
      jb_J_apply(u, g, c, lk, pm, am, a);
      mulMV(uw, Rw, u);
      c.ww[0] = uw[0] - c.uref[0]; c.ww[1] = uw[1] - c.uref[1]; c.ww[2] = uw[2] - c.uref[2];
    }
}

// forces at current ww: qfrc = sum J_c^T f_c ; returns packed active mask (4 bits per contact, max 16 contacts -> 64 bits)
template <typename T>
JB_HD unsigned long long jb_forces(const ContactSet<T>& cs, const T* Rw, const LegKin<T>* lk, const T* pm, const T* am,
                                   T mu, Vec15<T>& qfrc, T* cost_out) {
  unsigned long long mask = 0;
  int bit = 0;
  T cost = 0;
  v15_zero(qfrc);
#pragma unroll
  for (int g = 0; g < JB_NGROUP; g++)
    for (int i = 0; i < cs.n[g]; i++) {
      const Contact<T>& c = cs.c[g][i];
      T jar[4], fw[3] = {0, 0, 0};
      jb_rows(jar, c.ww, mu, c.ar0);
#pragma unroll
      for (int r = 0; r < 4; r++) {
        if (jar[r] < (T)0) {
          T f = -c.D * jar[r];
          T e[3];
          jb_row_dir(e, r, mu);
          fw[0] += f * e[0]; fw[1] += f * e[1]; fw[2] += f * e[2];
          cost += (T)0.5 * c.D * jar[r] * jar[r];
          if (bit < 64) mask |= 1ull << bit;
        }
        bit++;
      }
      T fr[3];
      mulMTV(fr, Rw, fw);
      jb_JT_accum(qfrc, g, c, lk, pm, am, fr);
    }
  *cost_out = cost;
  return mask;
}

// H += sum over active rows D j j^T (j = Jacobian row in internal coordinates)
template <typename T>
JB_HD void jb_hessian(Arrow<T>& H, const ContactSet<T>& cs, const T* Rw, const LegKin<T>* lk, const T* pm, const T* am, T mu) {
#pragma unroll
  for (int g = 0; g < JB_NGROUP; g++)
    for (int i = 0; i < cs.n[g]; i++) {
      const Contact<T>& c = cs.c[g][i];
      T jar[4];
      jb_rows(jar, c.ww, mu, c.ar0);
      // joint Jacobian columns (root comps)
      T ja[3] = {0, 0, 0}, jb[3] = {0, 0, 0};
      if (g < 4) {
        T r[3];
        sub3(r, c.x, lk[g].p1); cross3(ja, lk[g].a1, r);
        sub3(r, c.x, lk[g].p2); cross3(jb, lk[g].a2, r);
        jb[0] *= c.low; jb[1] *= c.low; jb[2] *= c.low;
      } else if (g == 5) {
        T r[3];
        sub3(r, c.x, pm); cross3(ja, am, r);
      }
#pragma unroll
      for (int r = 0; r < 4; r++) {
        if (!(jar[r] < (T)0)) continue;
        T ew[3], e[3], j6[6];
        jb_row_dir(ew, r, mu);
        mulMTV(e, Rw, ew);                       // row direction in root comps
        j6[0] = e[0]; j6[1] = e[1]; j6[2] = e[2];
        cross3(j6 + 3, c.x, e);
        T d = c.D;
#pragma unroll
        for (int p = 0; p < 6; p++)
#pragma unroll
          for (int q = 0; q <= p; q++) H.R[JB_TRI(p, q)] += d * j6[p] * j6[q];
        if (g < 4) {
          T jja = dot3(ja, e), jjb = dot3(jb, e);
          H.aa[g] += d * jja * jja; H.ab[g] += d * jja * jjb; H.bb[g] += d * jjb * jjb;
#pragma unroll
          for (int p = 0; p < 6; p++) { H.ca[g][p] += d * jja * j6[p]; H.cb[g][p] += d * jjb * j6[p]; }
        } else if (g == 5) {
          T jjm = dot3(ja, e);
          H.mm += d * jjm * jjm;
#pragma unroll
          for (int p = 0; p < 6; p++) H.cm[p] += d * jjm * j6[p];
        }
      }
    }
}

// derivative of the 1-D cost along the search direction at step alpha: returns phi'(alpha), writes phi''(alpha)
template <typename T>
JB_HD T jb_ls_deriv(const ContactSet<T>& cs, T mu, T g1, T g2, T alpha, T* dd, unsigned long long* mask_out) {
  T d = g1 + alpha * g2, h = g2;
  unsigned long long mask = 0;
  int bit = 0;
#pragma unroll
  for (int g = 0; g < JB_NGROUP; g++)
    for (int i = 0; i < cs.n[g]; i++) {
      const Contact<T>& c = cs.c[g][i];
      T jar[4], jv[4];
      jb_rows(jar, c.ww, mu, c.ar0);
      jb_rows(jv, c.jv, mu, (T)0);
#pragma unroll
      for (int r = 0; r < 4; r++) {
        T x = jar[r] + alpha * jv[r];
        if (x < (T)0) {
          d += c.D * x * jv[r];
          h += c.D * jv[r] * jv[r];
          if (bit < 64) mask |= 1ull << bit;
        }
        bit++;
      }
    }
  *dd = h;
  *mask_out = mask;
  return d;
}

// ------------------------------------------------------------------------------------------------ the sub-step
template <typename T, typename M, bool DIAG>
JB_HD void jb_substep(const M& m, JbState<T>& s, T ctrl, SubstepDiag<T>* diag, int* iter_accum,
                      bool forward_only = false, int* ncon_accum = nullptr, int* overflow_accum = nullptr,
                      ContactReport<T>* report = nullptr) {
  const T h = (T)m.timestep;
  // ---------------- root frame
  ConCtx<T> cx;
  {
    T qn = (T)1 / jb_sqrt(s.quat[0] * s.quat[0] + s.quat[1] * s.quat[1] + s.quat[2] * s.quat[2] + s.quat[3] * s.quat[3]);
    s.quat[0] *= qn; s.quat[1] *= qn; s.quat[2] *= qn; s.quat[3] *= qn;     // mj_kinematics normalises qpos quats
    quat2mat(cx.Rw, s.quat);
  }
  cx.pz = s.pos[2];
  cx.n[0] = cx.Rw[6]; cx.n[1] = cx.Rw[7]; cx.n[2] = cx.Rw[8];
  mulMTV(cx.vr, cx.Rw, s.vlin);
  cx.w[0] = s.w[0]; cx.w[1] = s.w[1]; cx.w[2] = s.w[2];
  const T* w = cx.w;
  // fictitious root acceleration: -g expressed in root comps (gravity = -g*z_world)
  T aO[3] = {(T)m.gravity * cx.n[0], (T)m.gravity * cx.n[1], (T)m.gravity * cx.n[2]};

  Arrow<T> Mx;
  Vec15<T> bias;
  v15_zero(bias);
  ContactSet<T> cs;
#pragma unroll
  for (int g = 0; g < JB_NGROUP; g++) cs.n[g] = 0;
  cs.overflow = 0;
  LegKin<T> lk[JB_NLEG];

  // accumulators for the root block: total mass, first moment, inertia about the origin
  T mt = 0, ht[3] = {0, 0, 0}, It[6] = {0, 0, 0, 0, 0, 0};
  auto accum_body = [&](T mass, const T* c, const T* Ic) {
    mt += mass;
    addscl3(ht, c, mass);
    T cc = dot3(c, c);
    It[0] += Ic[0] + mass * (cc - c[0] * c[0]);
    It[1] += Ic[1] + mass * (cc - c[1] * c[1]);
    It[2] += Ic[2] + mass * (cc - c[2] * c[2]);
    It[3] += Ic[3] - mass * c[0] * c[1];
    It[4] += Ic[4] - mass * c[0] * c[2];
    It[5] += Ic[5] - mass * c[1] * c[2];
  };
  // acceleration of a point fixed in the root: aO + w x (w x p)
  auto acc_root_point = [&](T* a, const T* p) {
    T t[3], u[3];
    cross3(t, w, p); cross3(u, w, t);
    a[0] = aO[0] + u[0]; a[1] = aO[1] + u[1]; a[2] = aO[2] + u[2];
  };
  // Newton-Euler wrench of a body: F = m*acom ; N = I*alpha + om x (I*om)
  auto body_wrench = [&](T* F, T* N, T mass, const T* Ic, const T* acom, const T* om, const T* al) {
    scl3(F, acom, mass);
    T Iw[3], Ia[3], t[3];
    mulSV(Iw, Ic, om); mulSV(Ia, Ic, al); cross3(t, om, Iw);
    add3(N, Ia, t);
  };

  // ---------------- root (core) body
  {
    T c0[3], I0[6];
    ld3(c0, m.c0);
#pragma unroll
    for (int i = 0; i < 6; i++) I0[i] = (T)m.I0[i];
    accum_body((T)m.m0, c0, I0);
    T ac[3], F[3], N[3], zero[3] = {0, 0, 0}, t[3];
    acc_root_point(ac, c0);
    body_wrench(F, N, (T)m.m0, I0, ac, w, zero);
    bias.r[0] += F[0]; bias.r[1] += F[1]; bias.r[2] += F[2];
    cross3(t, c0, F);
    bias.r[3] += t[0] + N[0]; bias.r[4] += t[1] + N[1]; bias.r[5] += t[2] + N[2];
    // root geoms
    BodyJ<T> bj = {0, nullptr, nullptr, 0, nullptr, nullptr, 0, 0, 4, 0};
#pragma unroll
    for (int b = 0; b < 2; b++) {
      T c[3], Rg[9], half[3];
      ld3(c, m.g_box[b].center); ld3(half, m.g_box[b].size);
#pragma unroll
      for (int i = 0; i < 9; i++) Rg[i] = (T)m.g_box[b].rot[i];
      jb_col_box<T>(m, cs, cx, bj, m.g_box[b].geom_id, c, Rg, half);
    }
    {
      T c[3], ax[3] = {(T)m.g_rcyl.rot[2], (T)m.g_rcyl.rot[5], (T)m.g_rcyl.rot[8]};
      T xa[3] = {(T)m.g_rcyl.rot[0], (T)m.g_rcyl.rot[3], (T)m.g_rcyl.rot[6]};
      ld3(c, m.g_rcyl.center);
      jb_col_cylinder<T>(m, cs, cx, bj, m.g_rcyl.geom_id, c, ax, xa, (T)m.g_rcyl.size[0], (T)m.g_rcyl.size[1]);
    }
    {
      T c[3], Rg[9], sz[3];
      ld3(c, m.g_rell.center); ld3(sz, m.g_rell.size);
#pragma unroll
      for (int i = 0; i < 9; i++) Rg[i] = (T)m.g_rell.rot[i];
      jb_col_ellipsoid<T>(m, cs, cx, bj, m.g_rell.geom_id, c, Rg, sz);
    }
  }

  // ---------------- legs
  Vec15<T> fpass;    // passive + actuator generalized forces
  v15_zero(fpass);
#pragma unroll
  for (int k = 0; k < JB_NLEG; k++) {
    T a1[3], p1[3], a2h[3], p2h[3];
    ld3(a1, m.a1[k]); ld3(p1, m.p1[k]); ld3(a2h, m.a2[k]); ld3(p2h, m.p2[k]);
    T qa = s.ql[k][0], qb = s.ql[k][1], qda = s.vl[k][0], qdb = s.vl[k][1];
    T sa, ca, sb, cb;
    jb_sincos(qa, &sa, &ca);
    jb_sincos(qb, &sb, &cb);
    T R1[9], R2[9], R12[9];
    rodrigues(R1, a1, sa, ca);
    rodrigues(R2, a2h, sb, cb);
    mulMM(R12, R1, R2);
    // current hinge-2 axis and anchor
    T a2[3], p2[3], t[3], r[3];
    mulMV(a2, R1, a2h);
    sub3(r, p2h, p1); mulMV(t, R1, r); add3(p2, p1, t);
    // COMs and inertias
    T cu[3], cl[3], Iu[6], Il[6], ch[3];
    ld3(ch, m.cu[k]); sub3(r, ch, p1); mulMV(t, R1, r); add3(cu, p1, t);
    ld3(ch, m.cl[k]); sub3(r, ch, p2h); mulMV(t, R12, r); add3(cl, p2, t);
    rotSym(Iu, R1, m.Iu[k]);
    rotSym(Il, R12, m.Il[k]);
    T mu_ = (T)m.mu[k], ml_ = (T)m.ml[k];
    accum_body(mu_, cu, Iu);
    accum_body(ml_, cl, Il);
    // ---- inertia columns: P = sum m a x (c - p), L = sum (I a + c x P_b)
    T Pb[3], Lb[3], Pa[3], La[3], tmp[3], tmp2[3];
    sub3(r, cl, p2); cross3(tmp, a2, r); scl3(Pb, tmp, ml_);
    mulSV(tmp, Il, a2); cross3(tmp2, cl, Pb); add3(Lb, tmp, tmp2);
    T Pal[3], Pau[3];
    sub3(r, cl, p1); cross3(tmp, a1, r); scl3(Pal, tmp, ml_);
    sub3(r, cu, p1); cross3(tmp, a1, r); scl3(Pau, tmp, mu_);
    add3(Pa, Pal, Pau);
    mulSV(La, Il, a1); mulSV(tmp, Iu, a1); addto3(La, tmp);
    cross3(tmp, cl, Pal); addto3(La, tmp);
    cross3(tmp, cu, Pau); addto3(La, tmp);
#pragma unroll
    for (int i = 0; i < 3; i++) {
      Mx.cb[k][i] = Pb[i]; Mx.cb[k][3 + i] = Lb[i];
      Mx.ca[k][i] = Pa[i]; Mx.ca[k][3 + i] = La[i];
    }
    cross3(tmp, p2, Pb); sub3(tmp2, Lb, tmp); Mx.bb[k] = dot3(a2, tmp2);
    cross3(tmp, p1, Pb); sub3(tmp2, Lb, tmp); Mx.ab[k] = dot3(a1, tmp2);
    cross3(tmp, p1, Pa); sub3(tmp2, La, tmp); Mx.aa[k] = dot3(a1, tmp2);
    // ---- Newton-Euler bias
    T om1[3], al1[3], om2[3], al2[3], ap1[3], ap2[3], acu[3], acl[3];
    om1[0] = w[0] + a1[0] * qda; om1[1] = w[1] + a1[1] * qda; om1[2] = w[2] + a1[2] * qda;
    cross3(al1, w, a1); scl3(al1, al1, qda);
    om2[0] = om1[0] + a2[0] * qdb; om2[1] = om1[1] + a2[1] * qdb; om2[2] = om1[2] + a2[2] * qdb;
    cross3(tmp, om1, a2); al2[0] = al1[0] + tmp[0] * qdb; al2[1] = al1[1] + tmp[1] * qdb; al2[2] = al1[2] + tmp[2] * qdb;
    acc_root_point(ap1, p1);
    auto acc_child = [&](T* out, const T* abase, const T* al, const T* om, const T* rr) {
      T x1[3], x2[3], x3[3];
      cross3(x1, al, rr); cross3(x2, om, rr); cross3(x3, om, x2);
      out[0] = abase[0] + x1[0] + x3[0]; out[1] = abase[1] + x1[1] + x3[1]; out[2] = abase[2] + x1[2] + x3[2];
    };
    T ru[3], r12[3], rl[3];
    sub3(ru, cu, p1); acc_child(acu, ap1, al1, om1, ru);
    sub3(r12, p2, p1); acc_child(ap2, ap1, al1, om1, r12);
    sub3(rl, cl, p2); acc_child(acl, ap2, al2, om2, rl);
    T Fu[3], Nu[3], Fl[3], Nl[3];
    body_wrench(Fu, Nu, mu_, Iu, acu, om1, al1);
    body_wrench(Fl, Nl, ml_, Il, acl, om2, al2);
    // joint b: a2 . ((cl - p2) x Fl + Nl)
    cross3(tmp, rl, Fl); addto3(tmp, Nl); bias.lb[k] = dot3(a2, tmp);
    // joint a: a1 . ((cu - p1) x Fu + Nu + (cl - p1) x Fl + Nl)
    cross3(tmp, ru, Fu); addto3(tmp, Nu);
    sub3(r, cl, p1); cross3(tmp2, r, Fl); addto3(tmp, tmp2); addto3(tmp, Nl);
    bias.la[k] = dot3(a1, tmp);
    // root
    bias.r[0] += Fu[0] + Fl[0]; bias.r[1] += Fu[1] + Fl[1]; bias.r[2] += Fu[2] + Fl[2];
    cross3(tmp, cu, Fu); cross3(tmp2, cl, Fl);
    bias.r[3] += tmp[0] + tmp2[0] + Nu[0] + Nl[0];
    bias.r[4] += tmp[1] + tmp2[1] + Nu[1] + Nl[1];
    bias.r[5] += tmp[2] + tmp2[2] + Nu[2] + Nl[2];
    // ---- passive spring / damper
    fpass.la[k] = -(T)m.stiff[k][0] * qa - (T)m.damp[k][0] * qda;
    fpass.lb[k] = -(T)m.stiff[k][1] * qb - (T)m.damp[k][1] * qdb;
    // ---- keep for the solver
#pragma unroll
    for (int i = 0; i < 3; i++) { lk[k].a1[i] = a1[i]; lk[k].p1[i] = p1[i]; lk[k].a2[i] = a2[i]; lk[k].p2[i] = p2[i]; }
    // ---- collision: upper cylinder, knee sphere (upper body); lower cylinder, foot sphere (lower body)
    {
      BodyJ<T> bu = {1, lk[k].a1, lk[k].p1, qda, lk[k].a2, lk[k].p2, (T)0, (T)0, k, 1 + 2 * k};
      BodyJ<T> bl = {2, lk[k].a1, lk[k].p1, qda, lk[k].a2, lk[k].p2, qdb, (T)1, k, 2 + 2 * k};
      T c[3], g[3], ax[3], xa[3], gh[3];
      // upper cylinder
      ld3(gh, m.g_ucyl[k].center); sub3(r, gh, p1); mulMV(t, R1, r); add3(c, p1, t);
      g[0] = (T)m.g_ucyl[k].rot[2]; g[1] = (T)m.g_ucyl[k].rot[5]; g[2] = (T)m.g_ucyl[k].rot[8]; mulMV(ax, R1, g);
      g[0] = (T)m.g_ucyl[k].rot[0]; g[1] = (T)m.g_ucyl[k].rot[3]; g[2] = (T)m.g_ucyl[k].rot[6]; mulMV(xa, R1, g);
      jb_col_cylinder<T>(m, cs, cx, bu, m.g_ucyl[k].geom_id, c, ax, xa, (T)m.g_ucyl[k].size[0], (T)m.g_ucyl[k].size[1]);
      // knee sphere
      ld3(gh, m.g_knee[k].center); sub3(r, gh, p1); mulMV(t, R1, r); add3(c, p1, t);
      jb_col_sphere<T>(m, cs, cx, bu, m.g_knee[k].geom_id, c, (T)m.g_knee[k].size[0]);
      // lower cylinder
      ld3(gh, m.g_lcyl[k].center); sub3(r, gh, p2h); mulMV(t, R12, r); add3(c, p2, t);
      g[0] = (T)m.g_lcyl[k].rot[2]; g[1] = (T)m.g_lcyl[k].rot[5]; g[2] = (T)m.g_lcyl[k].rot[8]; mulMV(ax, R12, g);
      g[0] = (T)m.g_lcyl[k].rot[0]; g[1] = (T)m.g_lcyl[k].rot[3]; g[2] = (T)m.g_lcyl[k].rot[6]; mulMV(xa, R12, g);
      jb_col_cylinder<T>(m, cs, cx, bl, m.g_lcyl[k].geom_id, c, ax, xa, (T)m.g_lcyl[k].size[0], (T)m.g_lcyl[k].size[1]);
      // foot sphere
      ld3(gh, m.g_foot[k].center); sub3(r, gh, p2h); mulMV(t, R12, r); add3(c, p2, t);
      jb_col_sphere<T>(m, cs, cx, bl, m.g_foot[k].geom_id, c, (T)m.g_foot[k].size[0]);
    }
  }

  // ---------------- rotor
  T pm[3], am[3];
  ld3(pm, m.pm); ld3(am, m.am);
  {
    T sm, cmn, Rm[9], r[3], t[3], cmc[3], Imc[6], ch[3];
    jb_sincos(s.qm, &sm, &cmn);
    rodrigues(Rm, am, sm, cmn);
    ld3(ch, m.cm); sub3(r, ch, pm); mulMV(t, Rm, r); add3(cmc, pm, t);
    rotSym(Imc, Rm, m.Im);
    T mm_ = (T)m.mm;
    accum_body(mm_, cmc, Imc);
    T P[3], L[3], tmp[3], tmp2[3];
    sub3(r, cmc, pm); cross3(tmp, am, r); scl3(P, tmp, mm_);
    mulSV(L, Imc, am); cross3(tmp, cmc, P); addto3(L, tmp);
#pragma unroll
    for (int i = 0; i < 3; i++) { Mx.cm[i] = P[i]; Mx.cm[3 + i] = L[i]; }
    cross3(tmp, pm, P); sub3(tmp2, L, tmp); Mx.mm = dot3(am, tmp2);
    T om[3], al[3], ap[3], ac[3], F[3], N[3];
    om[0] = w[0] + am[0] * s.vm; om[1] = w[1] + am[1] * s.vm; om[2] = w[2] + am[2] * s.vm;
    cross3(al, w, am); scl3(al, al, s.vm);
    acc_root_point(ap, pm);
    {
      T x1[3], x2[3], x3[3];
      cross3(x1, al, r); cross3(x2, om, r); cross3(x3, om, x2);
      ac[0] = ap[0] + x1[0] + x3[0]; ac[1] = ap[1] + x1[1] + x3[1]; ac[2] = ap[2] + x1[2] + x3[2];
    }
    body_wrench(F, N, mm_, Imc, ac, om, al);
    cross3(tmp, r, F); addto3(tmp, N); bias.mo = dot3(am, tmp);
    bias.r[0] += F[0]; bias.r[1] += F[1]; bias.r[2] += F[2];
    cross3(tmp, cmc, F);
    bias.r[3] += tmp[0] + N[0]; bias.r[4] += tmp[1] + N[1]; bias.r[5] += tmp[2] + N[2];
    // actuator + passive
    T u = jb_min((T)m.ctrl_hi, jb_max((T)m.ctrl_lo, ctrl));
    T force = (T)m.gain * u + (T)m.bias_v * ((T)m.gear * s.vm);
    fpass.mo = (T)m.gear * force - (T)m.damp_m * s.vm;
    // rotor geoms
    BodyJ<T> bj = {1, am, pm, s.vm, am, pm, (T)0, (T)0, 5, 9};
    {
      T c[3], gh[3], g[3], ax[3], xa[3];
      ld3(gh, m.g_mcyl.center); sub3(r, gh, pm); mulMV(t, Rm, r); add3(c, pm, t);
      g[0] = (T)m.g_mcyl.rot[2]; g[1] = (T)m.g_mcyl.rot[5]; g[2] = (T)m.g_mcyl.rot[8]; mulMV(ax, Rm, g);
      g[0] = (T)m.g_mcyl.rot[0]; g[1] = (T)m.g_mcyl.rot[3]; g[2] = (T)m.g_mcyl.rot[6]; mulMV(xa, Rm, g);
      jb_col_cylinder<T>(m, cs, cx, bj, m.g_mcyl.geom_id, c, ax, xa, (T)m.g_mcyl.size[0], (T)m.g_mcyl.size[1]);
      T Rg[9], Rh[9], sz[3];
      ld3(gh, m.g_mell.center); sub3(r, gh, pm); mulMV(t, Rm, r); add3(c, pm, t);
#pragma unroll
      for (int i = 0; i < 9; i++) Rh[i] = (T)m.g_mell.rot[i];
      mulMM(Rg, Rm, Rh);
      ld3(sz, m.g_mell.size);
      jb_col_ellipsoid<T>(m, cs, cx, bj, m.g_mell.geom_id, c, Rg, sz);
    }
  }

  // ---------------- root block of M
#pragma unroll
  for (int i = 0; i < 21; i++) Mx.R[i] = 0;
  Mx.R[JB_TRI(0, 0)] = mt; Mx.R[JB_TRI(1, 1)] = mt; Mx.R[JB_TRI(2, 2)] = mt;
  Mx.R[JB_TRI(3, 1)] = -ht[2]; Mx.R[JB_TRI(3, 2)] = ht[1];
  Mx.R[JB_TRI(4, 0)] = ht[2];  Mx.R[JB_TRI(4, 2)] = -ht[0];
  Mx.R[JB_TRI(5, 0)] = -ht[1]; Mx.R[JB_TRI(5, 1)] = ht[0];
  Mx.R[JB_TRI(3, 3)] = It[0]; Mx.R[JB_TRI(4, 4)] = It[1]; Mx.R[JB_TRI(5, 5)] = It[2];
  Mx.R[JB_TRI(4, 3)] = It[3]; Mx.R[JB_TRI(5, 3)] = It[4]; Mx.R[JB_TRI(5, 4)] = It[5];

  // ---------------- smooth dynamics: f_s = passive + actuator - bias ; a_s = M^-1 f_s
  Vec15<T> fs = fpass;
  v15_axpy(fs, (T)-1, bias);
  Vec15<T> as_;
  {
    Arrow<T> F = Mx;
    arrow_factor(F);
    arrow_solve(as_, F, fs);
  }

  int ncon = 0;
#pragma unroll
  for (int g = 0; g < JB_NGROUP; g++) ncon += cs.n[g];

  // ---------------- constraint solve (Newton, primal)
  Vec15<T> a = as_, fcon;
  v15_zero(fcon);
  int iters = 0;
  const T mu = (T)m.friction;
  if (ncon > 0) {
    // warm start: better of qacc_warmstart and qacc_smooth by total cost
    Vec15<T> aw;
    {
      T lw[3] = {s.warm[0], s.warm[1], s.warm[2]}, lr[3];
      mulMTV(lr, cx.Rw, lw);
      aw.r[0] = lr[0]; aw.r[1] = lr[1]; aw.r[2] = lr[2];
      aw.r[3] = s.warm[3]; aw.r[4] = s.warm[4]; aw.r[5] = s.warm[5];
#pragma unroll
      for (int k = 0; k < JB_NLEG; k++) { aw.la[k] = s.warm[6 + 2 * k]; aw.lb[k] = s.warm[7 + 2 * k]; }
      aw.mo = s.warm[14];
    }
    Vec15<T> Ma, tmpv;
    arrow_mul(Ma, Mx, aw);
    jb_eval_ww(cs, cx.Rw, lk, pm, am, aw);
    T cost_w = jb_constraint_cost(cs, mu);
    tmpv = Ma; v15_axpy(tmpv, (T)-1, fs);
    Vec15<T> da = aw; v15_axpy(da, (T)-1, as_);
    cost_w += (T)0.5 * v15_dot(tmpv, da);
    jb_eval_ww(cs, cx.Rw, lk, pm, am, as_);
    T cost_s = jb_constraint_cost(cs, mu);
    if (cost_w > cost_s) {
      a = as_;
      Ma = fs;                       // M a_s = f_s
    } else {
      a = aw;
      jb_eval_ww(cs, cx.Rw, lk, pm, am, a);
    }
    const T scale = (T)1 / ((T)m.meaninertia * (T)JB_NV);
    T ccost;
    unsigned long long mask = jb_forces(cs, cx.Rw, lk, pm, am, mu, fcon, &ccost);
    Vec15<T> grad = Ma;
    v15_axpy(grad, (T)-1, fs); v15_axpy(grad, (T)-1, fcon);
    const int maxit = JbEps<T>::newton_iters();
    while (iters < maxit) {
      // Hessian, Newton direction
      Arrow<T> H = Mx;
      jb_hessian(H, cs, cx.Rw, lk, pm, am, mu);
      arrow_factor(H);
      Vec15<T> search;
      arrow_solve(search, H, grad);
#pragma unroll
      for (int i = 0; i < 6; i++) search.r[i] = -search.r[i];
#pragma unroll
      for (int k = 0; k < JB_NLEG; k++) { search.la[k] = -search.la[k]; search.lb[k] = -search.lb[k]; }
      search.mo = -search.mo;
      Vec15<T> Mv;
      arrow_mul(Mv, Mx, search);
#pragma unroll
      for (int g = 0; g < JB_NGROUP; g++)
        for (int i = 0; i < cs.n[g]; i++) {
          T u[3];
          jb_J_apply(u, g, cs.c[g][i], lk, pm, am, search);
          mulMV(cs.c[g][i].jv, cx.Rw, u);
        }
      Vec15<T> res = Ma; v15_axpy(res, (T)-1, fs);
      T g1 = v15_dot(search, res), g2 = v15_dot(search, Mv);
      // exact line search on the piecewise-quadratic 1-D cost: safeguarded Newton on phi'
      T alpha = 1, dd;
      unsigned long long m1;
      bool full = false;
      {
        T d0 = jb_ls_deriv(cs, mu, g1, g2, (T)0, &dd, &m1);
        if (!(d0 < (T)0)) break;                         // not a descent direction: converged to rounding
        T d = jb_ls_deriv(cs, mu, g1, g2, (T)1, &dd, &m1);
        if (m1 == mask) {
          full = true;                                    // same active set at alpha=1: the Newton step is exact
        } else {
          T lo = 0, hi = -1, tol = JbEps<T>::ls_tol() * jb_abs(d0);
          for (int it = 0; it < 30; it++) {
            if (jb_abs(d) <= tol) break;
            if (d < (T)0) lo = alpha; else hi = alpha;
            T an = alpha - d / dd;
            if (hi > (T)0 && !(an > lo && an < hi)) an = (T)0.5 * (lo + hi);
            else if (!(an > lo)) an = alpha * (T)2;
            if (an == alpha) break;
            alpha = an;
            d = jb_ls_deriv(cs, mu, g1, g2, alpha, &dd, &m1);
          }
        }
      }
      v15_axpy(a, alpha, search);
      v15_axpy(Ma, alpha, Mv);
#pragma unroll
      for (int g = 0; g < JB_NGROUP; g++)
        for (int i = 0; i < cs.n[g]; i++) {
          Contact<T>& c = cs.c[g][i];
          c.ww[0] += alpha * c.jv[0]; c.ww[1] += alpha * c.jv[1]; c.ww[2] += alpha * c.jv[2];
        }
      unsigned long long newmask = jb_forces(cs, cx.Rw, lk, pm, am, mu, fcon, &ccost);
      grad = Ma;
      v15_axpy(grad, (T)-1, fs); v15_axpy(grad, (T)-1, fcon);
      iters++;
      T gn = scale * jb_sqrt(v15_dot(grad, grad));
      bool same = (newmask == mask);
      mask = newmask;
      if ((full && same) || gn < JbEps<T>::newton_tol()) break;
    }
  }
  if (iter_accum) *iter_accum += iters;
  if (ncon_accum) *ncon_accum += ncon;
  if (overflow_accum) *overflow_accum += cs.overflow;

  // ---------------- save warm start (MuJoCo coordinates)
  {
    T lw[3];
    mulMV(lw, cx.Rw, a.r);
    s.warm[0] = lw[0]; s.warm[1] = lw[1]; s.warm[2] = lw[2];
    s.warm[3] = a.r[3]; s.warm[4] = a.r[4]; s.warm[5] = a.r[5];
#pragma unroll
    for (int k = 0; k < JB_NLEG; k++) { s.warm[6 + 2 * k] = a.la[k]; s.warm[7 + 2 * k] = a.lb[k]; }
    s.warm[14] = a.mo;
  }

  if (DIAG) {
    Vec15<T> e, col;
    for (int j = 0; j < JB_NV; j++) {
      v15_zero(e);
      if (j < 6) e.r[j] = 1; else if (j < 14) { if ((j - 6) & 1) e.lb[(j - 6) / 2] = 1; else e.la[(j - 6) / 2] = 1; } else e.mo = 1;
      arrow_mul(col, Mx, e);
      for (int i = 0; i < 6; i++) diag->M[i][j] = col.r[i];
      for (int k = 0; k < JB_NLEG; k++) { diag->M[6 + 2 * k][j] = col.la[k]; diag->M[7 + 2 * k][j] = col.lb[k]; }
      diag->M[14][j] = col.mo;
    }
    auto dump = [&](T* o, const Vec15<T>& v) {
      for (int i = 0; i < 6; i++) o[i] = v.r[i];
      for (int k = 0; k < JB_NLEG; k++) { o[6 + 2 * k] = v.la[k]; o[7 + 2 * k] = v.lb[k]; }
      o[14] = v.mo;
    };
    dump(diag->bias, bias); dump(diag->fsmooth, fs); dump(diag->asmooth, as_); dump(diag->qacc, a); dump(diag->fcon, fcon);
    for (int i = 0; i < 9; i++) diag->Rw[i] = cx.Rw[i];
    int n = 0;
    for (int g = 0; g < JB_NGROUP; g++)
      for (int i = 0; i < cs.n[g]; i++) {
        diag->geom[n] = cs.c[g][i].geom; diag->dist[n] = cs.c[g][i].dist;
        for (int q = 0; q < 3; q++) diag->cpos[n][q] = cs.c[g][i].x[q];
        n++;
      }
    diag->ncon = n; diag->newton_iters = iters; diag->overflow = cs.overflow;
  }

  if (report) {
    int n = 0;
#pragma unroll
    for (int g = 0; g < JB_NGROUP; g++)
      for (int i = 0; i < cs.n[g]; i++)
        if (n < 24) { report->geom[n] = cs.c[g][i].geom; report->dist[n] = cs.c[g][i].dist; n++; }
    report->n = n;
  }
  if (forward_only) return;     // mj_forward: everything up to qacc / qacc_warmstart, no integration

  // ---------------- semi-implicit Euler, implicit in joint damping: (M + h B) a = f_s + f_c
  Vec15<T> rhs = fs;
  v15_axpy(rhs, (T)1, fcon);
  {
    Arrow<T> F = Mx;
#pragma unroll
    for (int k = 0; k < JB_NLEG; k++) { F.aa[k] += h * (T)m.damp[k][0]; F.bb[k] += h * (T)m.damp[k][1]; }
    F.mm += h * (T)m.damp_m;
    arrow_factor(F);
    arrow_solve(a, F, rhs);
  }
  {
    T lw[3];
    mulMV(lw, cx.Rw, a.r);
    s.vlin[0] += h * lw[0]; s.vlin[1] += h * lw[1]; s.vlin[2] += h * lw[2];
    s.w[0] += h * a.r[3]; s.w[1] += h * a.r[4]; s.w[2] += h * a.r[5];
#pragma unroll
    for (int k = 0; k < JB_NLEG; k++) { s.vl[k][0] += h * a.la[k]; s.vl[k][1] += h * a.lb[k]; }
    s.vm += h * a.mo;
  }
  // mj_integratePos with the NEW velocity
  s.pos[0] += h * s.vlin[0]; s.pos[1] += h * s.vlin[1]; s.pos[2] += h * s.vlin[2];
  {
    T wn = jb_sqrt(s.w[0] * s.w[0] + s.w[1] * s.w[1] + s.w[2] * s.w[2]);
    T ax[3] = {1, 0, 0};
    T ang = 0;
    if (wn >= (T)1e-15) { ax[0] = s.w[0] / wn; ax[1] = s.w[1] / wn; ax[2] = s.w[2] / wn; ang = h * wn; }
    T sh, ch;
    jb_sincos((T)0.5 * ang, &sh, &ch);
    T q[4] = {s.quat[0], s.quat[1], s.quat[2], s.quat[3]};
    T dq[4] = {ch, ax[0] * sh, ax[1] * sh, ax[2] * sh};
    s.quat[0] = q[0] * dq[0] - q[1] * dq[1] - q[2] * dq[2] - q[3] * dq[3];
    s.quat[1] = q[0] * dq[1] + q[1] * dq[0] + q[2] * dq[3] - q[3] * dq[2];
    s.quat[2] = q[0] * dq[2] - q[1] * dq[3] + q[2] * dq[0] + q[3] * dq[1];
    s.quat[3] = q[0] * dq[3] + q[1] * dq[2] - q[2] * dq[1] + q[3] * dq[0];
  }
#pragma unroll
  for (int k = 0; k < JB_NLEG; k++) { s.ql[k][0] += h * s.vl[k][0]; s.ql[k][1] += h * s.vl[k][1]; }
  s.qm += h * s.vm;
  // keep the motor angle small so fp32 keeps its resolution over a 10 s episode (150 rad/s); dynamics are periodic in qm
  if (s.qm > (T)3.14159265358979323846) { s.qm -= (T)6.28318530717958647692; s.rev += 1; }
  else if (s.qm <= -(T)3.14159265358979323846) { s.qm += (T)6.28318530717958647692; s.rev -= 1; }
}

#endif  // JB_DYNAMICS_CUH
