/* TEST INFRASTRUCTURE ONLY -- see jb_oracle.h for the scope statement ("parity unpinned").
 *
 * fp64 restatement of dm_control's Physics.step() for the Jitterbug model:
 *     mj_step2 (actuation, smooth acceleration, Newton constraint solve, semi-implicit Euler)
 *     mj_step1 (kinematics, com, CRB, collision, constraint rows, velocity stage, RNE, sensors)
 * Reference call sites: jitterbug_dmc/jitterbug.py:81 (Physics.from_xml_string), :84-90
 * (control.Environment, n_sub_steps = 50), :177 (class Physics(mujoco.Physics)).
 * Function names follow MuJoCo's engine (engine_core_smooth.c, engine_core_constraint.c,
 * engine_collision_primitive.c, engine_solver.c, engine_forward.c) so a reader can diff the
 * restatement against the real thing when one is available.
 */
#include "jb_oracle.h"

#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <pthread.h>

#define MINVAL 1e-15
#define MINIMP 0.0001
#define MAXIMP 0.9999
#define MINMU 1e-5
#define MAXVAL 1e10
#define PI 3.14159265358979323846

/* ------------------------------------------------------------------ small vector / quaternion helpers */
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void cross3(double* r, const double* a, const double* b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static double normalize3(double* v) {
  double n = sqrt(dot3(v, v));
  if (n < MINVAL) { v[0] = 1; v[1] = 0; v[2] = 0; return 0; }
  v[0] /= n; v[1] /= n; v[2] /= n;
  return n;
}
static void normalize4(double* q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
static void mulQuat(double* r, const double* a, const double* b) {
  double t[4];
  t[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  t[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  t[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  t[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  memcpy(r, t, sizeof(t));
}
static void quat2Mat(double* m, const double* q) {   /* row-major 3x3, as mju_quat2Mat */
  double w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
static void mulMatVec3(double* r, const double* m, const double* v) {
  double x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  double y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  double z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static void mulMatTVec3(double* r, const double* m, const double* v) {
  double x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
  double y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
  double z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static void rotVecQuat(double* r, const double* v, const double* q) {
  double m[9];
  quat2Mat(m, q);
  mulMatVec3(r, m, v);
}
static void axisAngle2Quat(double* q, const double* axis, double angle) {
  double s = sin(0.5 * angle);
  q[0] = cos(0.5 * angle); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}

/* ------------------------------------------------------------------ spatial algebra, [rot; lin] ordering (MuJoCo) */
static void mulInertVec(double* r, const double* i, const double* v) {
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
static void crossMotion(double* r, const double* vel, const double* v) {
  double a[3], b[3], c[3];
  cross3(a, vel, v);          /* w x v_rot */
  cross3(b, vel, v + 3);      /* w x v_lin */
  cross3(c, vel + 3, v);      /* vlin x v_rot */
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2];
  r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
static void crossForce(double* r, const double* vel, const double* f) {
  double a[3], b[3], c[3];
  cross3(a, vel, f);          /* w x torque */
  cross3(b, vel + 3, f + 3);  /* vlin x force */
  cross3(c, vel, f + 3);      /* w x force */
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2];
  r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}

/* ------------------------------------------------------------------ mj_kinematics */
static void kinematics(const JboModel* m, JboData* d) {
  int b, j, g;
  memset(d->xpos[0], 0, sizeof(d->xpos[0]));
  d->xquat[0][0] = 1; d->xquat[0][1] = d->xquat[0][2] = d->xquat[0][3] = 0;
  quat2Mat(d->xmat[0], d->xquat[0]);
  /* normalize quaternions stored in qpos */
  for (j = 0; j < m->njnt; j++)
    if (m->jnt_type[j] == JBO_FREE) normalize4(d->qpos + m->jnt_qposadr[j] + 3);
  for (b = 1; b < m->nbody; b++) {
    int p = m->body_parentid[b], ja = m->body_jntadr[b], jn = m->body_jntnum[b];
    double xp[3], xq[4];
    if (jn == 1 && m->jnt_type[ja] == JBO_FREE) {
      int a = m->jnt_qposadr[ja];
      memcpy(xp, d->qpos + a, 3 * sizeof(double));
      memcpy(xq, d->qpos + a + 3, 4 * sizeof(double));
      memcpy(d->xanchor[ja], xp, sizeof(xp));
      d->xaxis[ja][0] = 0; d->xaxis[ja][1] = 0; d->xaxis[ja][2] = 1;
    } else {
      double t[3];
      mulMatVec3(t, d->xmat[p], m->body_pos[b]);
      xp[0] = d->xpos[p][0] + t[0]; xp[1] = d->xpos[p][1] + t[1]; xp[2] = d->xpos[p][2] + t[2];
      mulQuat(xq, d->xquat[p], m->body_quat[b]);
      for (j = ja; j < ja + jn; j++) {
        double qloc[4], vec[3];
        rotVecQuat(d->xaxis[j], m->jnt_axis[j], xq);
        rotVecQuat(vec, m->jnt_pos[j], xq);
        d->xanchor[j][0] = vec[0] + xp[0]; d->xanchor[j][1] = vec[1] + xp[1]; d->xanchor[j][2] = vec[2] + xp[2];
        /* hinge: rotate about local axis, then correct for off-centre rotation */
        axisAngle2Quat(qloc, m->jnt_axis[j], d->qpos[m->jnt_qposadr[j]] - m->qpos0[m->jnt_qposadr[j]]);
        mulQuat(xq, xq, qloc);
        rotVecQuat(vec, m->jnt_pos[j], xq);
        xp[0] = d->xanchor[j][0] - vec[0]; xp[1] = d->xanchor[j][1] - vec[1]; xp[2] = d->xanchor[j][2] - vec[2];
      }
    }
    normalize4(xq);
    memcpy(d->xpos[b], xp, sizeof(xp));
    memcpy(d->xquat[b], xq, sizeof(xq));
    quat2Mat(d->xmat[b], xq);
  }
  for (b = 1; b < m->nbody; b++) {
    double t[3], q[4];
    mulMatVec3(t, d->xmat[b], m->body_ipos[b]);
    d->xipos[b][0] = d->xpos[b][0] + t[0]; d->xipos[b][1] = d->xpos[b][1] + t[1]; d->xipos[b][2] = d->xpos[b][2] + t[2];
    mulQuat(q, d->xquat[b], m->body_iquat[b]);
    quat2Mat(d->ximat[b], q);
  }
  for (g = 0; g < m->ngeom; g++) {
    int bb = m->geom_bodyid[g];
    double t[3], q[4];
    mulMatVec3(t, d->xmat[bb], m->geom_pos[g]);
    d->geom_xpos[g][0] = d->xpos[bb][0] + t[0]; d->geom_xpos[g][1] = d->xpos[bb][1] + t[1];
    d->geom_xpos[g][2] = d->xpos[bb][2] + t[2];
    mulQuat(q, d->xquat[bb], m->geom_quat[g]);
    normalize4(q);
    quat2Mat(d->geom_xmat[g], q);
  }
}

/* ------------------------------------------------------------------ mj_comPos: subtree com, cinert, cdof */
static void comPos(const JboModel* m, JboData* d) {
  int b, j, k;
  double mass_subtree[JBO_MAXB];
  for (b = 0; b < m->nbody; b++) {
    mass_subtree[b] = m->body_mass[b];
    for (k = 0; k < 3; k++) d->subtree_com[b][k] = m->body_mass[b] * d->xipos[b][k];
  }
  for (b = m->nbody - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    mass_subtree[p] += mass_subtree[b];
    for (k = 0; k < 3; k++) d->subtree_com[p][k] += d->subtree_com[b][k];
  }
  for (b = 0; b < m->nbody; b++) {
    if (mass_subtree[b] < MINVAL) memcpy(d->subtree_com[b], d->xipos[b], 3 * sizeof(double));
    else for (k = 0; k < 3; k++) d->subtree_com[b][k] /= mass_subtree[b];
  }
  /* cinert: inertia about the com-frame origin (subtree com of the tree root), world axes */
  for (b = 1; b < m->nbody; b++) {
    const double* mat = d->ximat[b];
    const double* in = m->body_inertia[b];
    const double* O = d->subtree_com[m->body_rootid[b]];
    double dif[3] = {d->xipos[b][0] - O[0], d->xipos[b][1] - O[1], d->xipos[b][2] - O[2]};
    double mass = m->body_mass[b];
    double* r = d->cinert[b];
    /* mat * diag(in) * mat' */
    r[0] = mat[0] * mat[0] * in[0] + mat[1] * mat[1] * in[1] + mat[2] * mat[2] * in[2];
    r[1] = mat[3] * mat[3] * in[0] + mat[4] * mat[4] * in[1] + mat[5] * mat[5] * in[2];
    r[2] = mat[6] * mat[6] * in[0] + mat[7] * mat[7] * in[1] + mat[8] * mat[8] * in[2];
    r[3] = mat[0] * mat[3] * in[0] + mat[1] * mat[4] * in[1] + mat[2] * mat[5] * in[2];
    r[4] = mat[0] * mat[6] * in[0] + mat[1] * mat[7] * in[1] + mat[2] * mat[8] * in[2];
    r[5] = mat[3] * mat[6] * in[0] + mat[4] * mat[7] * in[1] + mat[5] * mat[8] * in[2];
    r[0] += mass * (dif[1] * dif[1] + dif[2] * dif[2]);
    r[1] += mass * (dif[0] * dif[0] + dif[2] * dif[2]);
    r[2] += mass * (dif[0] * dif[0] + dif[1] * dif[1]);
    r[3] -= mass * dif[0] * dif[1];
    r[4] -= mass * dif[0] * dif[2];
    r[5] -= mass * dif[1] * dif[2];
    r[6] = mass * dif[0]; r[7] = mass * dif[1]; r[8] = mass * dif[2];
    r[9] = mass;
  }
  memset(d->cinert[0], 0, sizeof(d->cinert[0]));
  /* cdof */
  for (j = 0; j < m->njnt; j++) {
    int b2 = m->jnt_bodyid[j], da = m->jnt_dofadr[j];
    const double* O = d->subtree_com[m->body_rootid[b2]];
    double off[3] = {O[0] - d->xanchor[j][0], O[1] - d->xanchor[j][1], O[2] - d->xanchor[j][2]};
    if (m->jnt_type[j] == JBO_FREE) {
      for (k = 0; k < 3; k++) {
        double ax[3] = {d->xmat[b2][k], d->xmat[b2][3 + k], d->xmat[b2][6 + k]};   /* body axis k in world */
        memset(d->cdof[da + k], 0, 6 * sizeof(double));
        d->cdof[da + k][3 + k] = 1;
        memcpy(d->cdof[da + 3 + k], ax, sizeof(ax));
        cross3(d->cdof[da + 3 + k] + 3, ax, off);
      }
    } else {
      memcpy(d->cdof[da], d->xaxis[j], 3 * sizeof(double));
      cross3(d->cdof[da] + 3, d->xaxis[j], off);
    }
  }
}

/* ------------------------------------------------------------------ mj_crb: composite rigid body -> qM */
static void crb(const JboModel* m, JboData* d) {
  int b, i, j, k;
  for (b = 0; b < m->nbody; b++) memcpy(d->crb[b], d->cinert[b], sizeof(d->crb[b]));
  for (b = m->nbody - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    if (p > 0) for (k = 0; k < 10; k++) d->crb[p][k] += d->crb[b][k];
  }
  memset(d->qM, 0, sizeof(d->qM));
  for (i = 0; i < m->nv; i++) {
    double buf[6];
    mulInertVec(buf, d->crb[m->dof_bodyid[i]], d->cdof[i]);
    for (j = i; j >= 0; j = m->dof_parentid[j]) {
      double s = 0;
      for (k = 0; k < 6; k++) s += d->cdof[j][k] * buf[k];
      d->qM[i][j] = s;
      d->qM[j][i] = s;
    }
  }
}

/* ------------------------------------------------------------------ dense Cholesky (A = L L') */
static int cholFactor(double* A, int n, int ld) {   /* in place, lower triangle */
  int i, j, k;
  for (j = 0; j < n; j++) {
    double s = A[j * ld + j];
    for (k = 0; k < j; k++) s -= A[j * ld + k] * A[j * ld + k];
    if (s < MINVAL) s = MINVAL;
    s = sqrt(s);
    A[j * ld + j] = s;
    for (i = j + 1; i < n; i++) {
      double t = A[i * ld + j];
      for (k = 0; k < j; k++) t -= A[i * ld + k] * A[j * ld + k];
      A[i * ld + j] = t / s;
    }
  }
  return 0;
}
static void cholSolve(const double* L, int n, int ld, double* x, const double* bvec) {
  int i, k;
  for (i = 0; i < n; i++) {
    double s = bvec[i];
    for (k = 0; k < i; k++) s -= L[i * ld + k] * x[k];
    x[i] = s / L[i * ld + i];
  }
  for (i = n - 1; i >= 0; i--) {
    double s = x[i];
    for (k = i + 1; k < n; k++) s -= L[k * ld + i] * x[k];
    x[i] = s / L[i * ld + i];
  }
}

/* ------------------------------------------------------------------ collision: plane vs {sphere, cylinder, box, ellipsoid} */
typedef struct { double dist, pos[3]; } RawCon;

static int planeSphere(const double* ppos, const double* pmat, const double* pos2, const double* size2,
                       double margin, RawCon* out) {
  double n[3] = {pmat[2], pmat[5], pmat[8]};
  double t[3] = {pos2[0] - ppos[0], pos2[1] - ppos[1], pos2[2] - ppos[2]};
  double dist = dot3(t, n) - size2[0];
  int k;
  if (dist > margin) return 0;
  out->dist = dist;
  for (k = 0; k < 3; k++) out->pos[k] = pos2[k] - n[k] * (size2[0] + 0.5 * dist);
  return 1;
}

static int planeCylinder(const double* ppos, const double* pmat, const double* pos2, const double* mat2,
                         const double* size2, double margin, RawCon* out) {
  double n[3] = {pmat[2], pmat[5], pmat[8]};
  double axis[3] = {mat2[2], mat2[5], mat2[8]};
  double vec[3], prjaxis, prjvec, dist0, len_sqr, prjvec1;
  int cnt = 0, k;
  prjaxis = dot3(n, axis);
  if (prjaxis > 0) { axis[0] = -axis[0]; axis[1] = -axis[1]; axis[2] = -axis[2]; prjaxis = -prjaxis; }
  for (k = 0; k < 3; k++) vec[k] = pos2[k] - ppos[k];
  dist0 = dot3(vec, n);
  /* remove component of -normal along the axis */
  for (k = 0; k < 3; k++) vec[k] = axis[k] * prjaxis - n[k];
  len_sqr = dot3(vec, vec);
  if (len_sqr >= MINVAL * MINVAL) {
    double s = size2[0] / sqrt(len_sqr);
    for (k = 0; k < 3; k++) vec[k] *= s;
  } else {
    /* disk parallel to plane: pick the cylinder's x axis */
    vec[0] = mat2[0] * size2[0]; vec[1] = mat2[3] * size2[0]; vec[2] = mat2[6] * size2[0];
  }
  prjvec = dot3(vec, n);
  for (k = 0; k < 3; k++) axis[k] *= size2[1];
  prjaxis *= size2[1];
  /* first point: lowest rim point of the lower cap */
  if (dist0 + prjaxis + prjvec <= margin) {
    out[cnt].dist = dist0 + prjaxis + prjvec;
    for (k = 0; k < 3; k++) out[cnt].pos[k] = pos2[k] + vec[k] + axis[k] - n[k] * out[cnt].dist * 0.5;
    cnt++;
  } else {
    return 0;
  }
  /* second point: same rim direction on the other cap */
  if (dist0 - prjaxis + prjvec <= margin) {
    out[cnt].dist = dist0 - prjaxis + prjvec;
    for (k = 0; k < 3; k++) out[cnt].pos[k] = pos2[k] + vec[k] - axis[k] - n[k] * out[cnt].dist * 0.5;
    cnt++;
  }
  /* triangle points on the lower cap */
  prjvec1 = -prjvec * 0.5;
  if (dist0 + prjaxis + prjvec1 <= margin) {
    double vec1[3];
    cross3(vec1, vec, axis);
    normalize3(vec1);
    for (k = 0; k < 3; k++) vec1[k] *= size2[0] * sqrt(3.0) / 2;
    out[cnt].dist = dist0 + prjaxis + prjvec1;
    for (k = 0; k < 3; k++)
      out[cnt].pos[k] = pos2[k] + vec1[k] + axis[k] - vec[k] * 0.5 - n[k] * out[cnt].dist * 0.5;
    cnt++;
    out[cnt].dist = dist0 + prjaxis + prjvec1;
    for (k = 0; k < 3; k++)
      out[cnt].pos[k] = pos2[k] - vec1[k] + axis[k] - vec[k] * 0.5 - n[k] * out[cnt].dist * 0.5;
    cnt++;
  }
  return cnt;
}

static int planeBox(const double* ppos, const double* pmat, const double* pos2, const double* mat2,
                    const double* size2, double margin, RawCon* out) {
  double n[3] = {pmat[2], pmat[5], pmat[8]};
  double dif[3] = {pos2[0] - ppos[0], pos2[1] - ppos[1], pos2[2] - ppos[2]};
  double dist = dot3(dif, n);
  int cnt = 0, i, k;
  for (i = 0; i < 8; i++) {
    double vec[3], corner[3], ldist, cdist;
    vec[0] = (i & 1 ? size2[0] : -size2[0]);
    vec[1] = (i & 2 ? size2[1] : -size2[1]);
    vec[2] = (i & 4 ? size2[2] : -size2[2]);
    mulMatVec3(corner, mat2, vec);
    ldist = dot3(n, corner);
    if (dist + ldist > margin || ldist > 0) continue;
    cdist = dist + ldist;
    out[cnt].dist = cdist;
    for (k = 0; k < 3; k++) out[cnt].pos[k] = corner[k] + pos2[k] - n[k] * cdist * 0.5;
    if (++cnt >= 4) return 4;
  }
  return cnt;
}

static int planeEllipsoid(const double* ppos, const double* pmat, const double* pos2, const double* mat2,
                          const double* size2, double margin, RawCon* out) {
  /* mjc_PlaneConvex with the ellipsoid support function: support point in direction -normal */
  double n[3] = {pmat[2], pmat[5], pmat[8]};
  double dir[3] = {-n[0], -n[1], -n[2]}, loc[3], sup[3], w[3], dif[3], dist, nrm;
  int k;
  mulMatTVec3(loc, mat2, dir);
  for (k = 0; k < 3; k++) loc[k] *= size2[k];
  nrm = sqrt(dot3(loc, loc));
  if (nrm < MINVAL) nrm = MINVAL;
  for (k = 0; k < 3; k++) loc[k] = loc[k] / nrm * size2[k];
  mulMatVec3(w, mat2, loc);
  for (k = 0; k < 3; k++) sup[k] = w[k] + pos2[k];
  for (k = 0; k < 3; k++) dif[k] = sup[k] - ppos[k];
  dist = dot3(n, dif);
  if (dist > margin) return 0;
  out->dist = dist;
  for (k = 0; k < 3; k++) out->pos[k] = sup[k] - n[k] * 0.5 * dist;
  return 1;
}

static void makeFrame(double* frame) {   /* mju_makeFrame: complete the basis given frame[0..2] = normal */
  double tmp[3], s;
  int k;
  normalize3(frame);
  if (frame[1] < 0.5 && frame[1] > -0.5) { tmp[0] = 0; tmp[1] = 1; tmp[2] = 0; }
  else { tmp[0] = 0; tmp[1] = 0; tmp[2] = 1; }
  s = dot3(frame, tmp);
  for (k = 0; k < 3; k++) frame[3 + k] = tmp[k] - s * frame[k];
  normalize3(frame + 3);
  cross3(frame + 6, frame, frame + 3);
}

static void collision(const JboModel* m, JboData* d) {
  int g1, g2, k, c;
  d->ncon = 0;
  /* pairs come out in (body pair, geom id) order; only plane-on-world vs moving-body geoms can touch in this
   * model (no self-contact is geometrically reachable) */
  for (g1 = 0; g1 < m->ngeom; g1++) {
    if (m->geom_type[g1] != JBO_PLANE || m->geom_bodyid[g1] != 0) continue;
    for (g2 = 0; g2 < m->ngeom; g2++) {
      RawCon rc[4];
      int n = 0, b2 = m->geom_bodyid[g2];
      double margin, gap, mix;
      if (g2 == g1 || b2 == 0 || m->geom_type[g2] == JBO_PLANE) continue;
      if (!((m->geom_contype[g1] & m->geom_conaffinity[g2]) || (m->geom_contype[g2] & m->geom_conaffinity[g1])))
        continue;
      margin = fmax(m->geom_margin[g1], m->geom_margin[g2]);
      gap = fmax(m->geom_gap[g1], m->geom_gap[g2]);
      switch (m->geom_type[g2]) {
        case JBO_SPHERE:
          n = planeSphere(d->geom_xpos[g1], d->geom_xmat[g1], d->geom_xpos[g2], m->geom_size[g2], margin, rc); break;
        case JBO_CYLINDER:
          n = planeCylinder(d->geom_xpos[g1], d->geom_xmat[g1], d->geom_xpos[g2], d->geom_xmat[g2],
                            m->geom_size[g2], margin, rc); break;
        case JBO_BOX:
          n = planeBox(d->geom_xpos[g1], d->geom_xmat[g1], d->geom_xpos[g2], d->geom_xmat[g2],
                       m->geom_size[g2], margin, rc); break;
        case JBO_ELLIPSOID:
          n = planeEllipsoid(d->geom_xpos[g1], d->geom_xmat[g1], d->geom_xpos[g2], d->geom_xmat[g2],
                             m->geom_size[g2], margin, rc); break;
        default: n = 0;
      }
      mix = m->geom_solmix[g1] / (m->geom_solmix[g1] + m->geom_solmix[g2]);
      for (c = 0; c < n && d->ncon < JBO_MAXCON; c++) {
        int id = d->ncon++;
        d->con_geom1[id] = g1; d->con_geom2[id] = g2;
        d->con_dist[id] = rc[c].dist;
        memcpy(d->con_pos[id], rc[c].pos, sizeof(rc[c].pos));
        d->con_frame[id][0] = d->geom_xmat[g1][2]; d->con_frame[id][1] = d->geom_xmat[g1][5];
        d->con_frame[id][2] = d->geom_xmat[g1][8];
        makeFrame(d->con_frame[id]);
        /* includemargin = margin - gap; contacts inside the gap are excluded from the solver */
        d->con_exclude[id] = (rc[c].dist >= margin - gap);
        for (k = 0; k < 3; k++) {
          double f = fmax(m->geom_friction[g1][k], m->geom_friction[g2][k]);
          if (k == 0) { d->con_friction[id][0] = d->con_friction[id][1] = fmax(MINMU, f); }
          else if (k == 1) d->con_friction[id][2] = fmax(MINMU, f);
          else { d->con_friction[id][3] = d->con_friction[id][4] = fmax(MINMU, f); }
        }
        for (k = 0; k < 2; k++)
          d->con_solref[id][k] = mix * m->geom_solref[g1][k] + (1 - mix) * m->geom_solref[g2][k];
        for (k = 0; k < 5; k++)
          d->con_solimp[id][k] = mix * m->geom_solimp[g1][k] + (1 - mix) * m->geom_solimp[g2][k];
      }
    }
  }
}

/* ------------------------------------------------------------------ mj_jac: translational Jacobian of a world point on a body */
static void jacPoint(const JboModel* m, const JboData* d, double jacp[3][JBO_MAXV], const double* point, int body) {
  int b = body, j, k, r;
  for (r = 0; r < 3; r++) memset(jacp[r], 0, sizeof(double) * JBO_MAXV);
  while (b > 0) {
    for (j = m->body_jntadr[b]; j >= 0 && j < m->body_jntadr[b] + m->body_jntnum[b]; j++) {
      int da = m->jnt_dofadr[j];
      if (m->jnt_type[j] == JBO_FREE) {
        double off[3] = {point[0] - d->xpos[b][0], point[1] - d->xpos[b][1], point[2] - d->xpos[b][2]};
        for (k = 0; k < 3; k++) {
          double ax[3] = {d->xmat[b][k], d->xmat[b][3 + k], d->xmat[b][6 + k]}, c[3];
          jacp[k][da + k] = 1;
          cross3(c, ax, off);
          for (r = 0; r < 3; r++) jacp[r][da + 3 + k] = c[r];
        }
      } else {
        double off[3] = {point[0] - d->xanchor[j][0], point[1] - d->xanchor[j][1], point[2] - d->xanchor[j][2]}, c[3];
        cross3(c, d->xaxis[j], off);
        for (r = 0; r < 3; r++) jacp[r][da] = c[r];
      }
    }
    b = m->body_parentid[b];
  }
}

/* ------------------------------------------------------------------ solimp sigmoid (getimpedance) */
static void getImpedance(const double* solimp_in, double pos, double margin, double* imp, double* impP) {
  double s[5], x, y, yP;
  memcpy(s, solimp_in, sizeof(s));
  s[0] = fmin(MAXIMP, fmax(MINIMP, s[0]));
  s[1] = fmin(MAXIMP, fmax(MINIMP, s[1]));
  s[2] = fmax(0, s[2]);
  s[3] = fmin(MAXIMP, fmax(MINIMP, s[3]));
  s[4] = fmax(1, s[4]);
  if (s[0] == s[1] || s[2] <= MINVAL) { *imp = 0.5 * (s[0] + s[1]); *impP = 0; return; }
  x = (pos - margin) / s[2];
  if (x < 0) x = -x;
  if (x >= 1) { *imp = s[1]; *impP = 0; return; }
  if (x <= 0) { *imp = s[0]; *impP = 0; return; }
  if (s[4] == 1) { y = x; yP = 1; }
  else if (x <= s[3]) {
    double a = 1 / pow(s[3], s[4] - 1);
    y = a * pow(x, s[4]);
    yP = s[4] * a * pow(x, s[4] - 1);
  } else {
    double bb = 1 / pow(1 - s[3], s[4] - 1);
    y = 1 - bb * pow(1 - x, s[4]);
    yP = s[4] * bb * pow(1 - x, s[4] - 1);
  }
  *imp = s[0] + y * (s[1] - s[0]);
  *impP = yP * (s[1] - s[0]) / s[2];
}

/* ------------------------------------------------------------------ mj_makeConstraint (pyramidal contacts, condim 3) */
static void makeConstraint(const JboModel* m, JboData* d) {
  int c, i, k, r, nv = m->nv;
  d->nefc = 0;
  for (c = 0; c < d->ncon; c++) {
    double jacp[3][JBO_MAXV], jn[JBO_MAXV], jt[2][JBO_MAXV];
    int b2 = m->geom_bodyid[d->con_geom2[c]], b1 = m->geom_bodyid[d->con_geom1[c]];
    double tran, imp, impP, R0, Rpy, timeconst, dampratio, dmax, K, B, mu;
    int row0;
    if (d->con_exclude[c]) continue;
    if (d->nefc + 4 > JBO_MAXEFC) break;
    jacPoint(m, d, jacp, d->con_pos[c], b2);        /* body1 is the world: zero Jacobian */
    for (i = 0; i < nv; i++) {
      jn[i] = 0; jt[0][i] = 0; jt[1][i] = 0;
      for (r = 0; r < 3; r++) {
        jn[i] += d->con_frame[c][r] * jacp[r][i];
        jt[0][i] += d->con_frame[c][3 + r] * jacp[r][i];
        jt[1][i] += d->con_frame[c][6 + r] * jacp[r][i];
      }
    }
    row0 = d->nefc;
    tran = m->body_invweight0[b1][0] + m->body_invweight0[b2][0];
    for (k = 0; k < 2; k++) {
      double fri = d->con_friction[c][k];
      for (r = 0; r < 2; r++) {
        int row = d->nefc++;
        double sgn = r == 0 ? 1.0 : -1.0;
        for (i = 0; i < nv; i++) d->efc_J[row][i] = jn[i] + sgn * fri * jt[k][i];
        d->efc_id[row] = c;
        d->efc_pos[row] = d->con_dist[c];
        d->efc_margin[row] = 0;   /* includemargin = margin - gap = 0 for this model */
        d->efc_diagApprox[row] = tran + fri * fri * tran;
      }
    }
    /* mj_makeImpedance */
    getImpedance(d->con_solimp[c], d->con_dist[c], 0, &imp, &impP);
    timeconst = d->con_solref[c][0]; dampratio = d->con_solref[c][1];
    dmax = fmin(MAXIMP, fmax(MINIMP, d->con_solimp[c][1]));
    if (timeconst > 0) {
      /* refsafe: time constant at least 2*timestep */
      timeconst = fmax(timeconst, 2 * m->timestep);
      K = 1 / fmax(MINVAL, dmax * dmax * timeconst * timeconst * dampratio * dampratio);
      B = 2 / fmax(MINVAL, dmax * timeconst);
    } else {
      K = -timeconst / fmax(MINVAL, dmax * dmax);
      B = -dampratio / fmax(MINVAL, dmax);
    }
    R0 = fmax(MINVAL, (1 - imp) * d->efc_diagApprox[row0] / imp);
    /* regularised cone: R[1] = R[0]/impratio; mu = friction[0]*sqrt(R[1]/R[0]); Rpy = 2*mu^2*R[0] */
    mu = d->con_friction[c][0] * sqrt(1.0 / fmax(MINVAL, m->impratio));
    d->con_mu[c] = mu;
    Rpy = 2 * mu * mu * R0;
    for (r = row0; r < row0 + 4; r++) {
      d->efc_R[r] = Rpy;
      d->efc_D[r] = 1 / Rpy;
      d->efc_KBIP[r][0] = K; d->efc_KBIP[r][1] = B; d->efc_KBIP[r][2] = imp; d->efc_KBIP[r][3] = impP;
    }
  }
}

/* ------------------------------------------------------------------ velocity stage */
static void comVel(const JboModel* m, JboData* d) {
  int b, j, k;
  memset(d->cvel[0], 0, sizeof(d->cvel[0]));
  for (b = 1; b < m->nbody; b++) {
    double cvel[6];
    int ja = m->body_jntadr[b], jn = m->body_jntnum[b];
    memcpy(cvel, d->cvel[m->body_parentid[b]], sizeof(cvel));
    for (j = ja; j >= 0 && j < ja + jn; j++) {
      int da = m->jnt_dofadr[j];
      if (m->jnt_type[j] == JBO_FREE) {
        int t;
        for (t = 0; t < 3; t++) {
          memset(d->cdof_dot[da + t], 0, 6 * sizeof(double));
          for (k = 0; k < 6; k++) cvel[k] += d->cdof[da + t][k] * d->qvel[da + t];
        }
        for (t = 3; t < 6; t++) crossMotion(d->cdof_dot[da + t], cvel, d->cdof[da + t]);
        for (t = 3; t < 6; t++)
          for (k = 0; k < 6; k++) cvel[k] += d->cdof[da + t][k] * d->qvel[da + t];
      } else {
        crossMotion(d->cdof_dot[da], cvel, d->cdof[da]);
        for (k = 0; k < 6; k++) cvel[k] += d->cdof[da][k] * d->qvel[da];
      }
    }
    memcpy(d->cvel[b], cvel, sizeof(cvel));
  }
}

static void passive(const JboModel* m, JboData* d) {
  int j, i;
  for (i = 0; i < m->nv; i++) d->qfrc_passive[i] = 0;
  for (j = 0; j < m->njnt; j++) {
    if (m->jnt_type[j] == JBO_HINGE && m->jnt_stiffness[j] != 0) {
      int qa = m->jnt_qposadr[j];
      d->qfrc_passive[m->jnt_dofadr[j]] = -m->jnt_stiffness[j] * (d->qpos[qa] - m->qpos_spring[qa]);
    }
  }
  for (i = 0; i < m->nv; i++) d->qfrc_passive[i] -= m->dof_damping[i] * d->qvel[i];
}

static void referenceConstraint(const JboModel* m, JboData* d) {
  int r, i;
  for (r = 0; r < d->nefc; r++) {
    double v = 0;
    for (i = 0; i < m->nv; i++) v += d->efc_J[r][i] * d->qvel[i];
    d->efc_vel[r] = v;
    d->efc_aref[r] = -d->efc_KBIP[r][0] * d->efc_KBIP[r][2] * (d->efc_pos[r] - d->efc_margin[r])
                     - d->efc_KBIP[r][1] * v;
  }
}

static void rneBias(const JboModel* m, JboData* d) {   /* mj_rne(flg_acc = 0) */
  double cacc[JBO_MAXB][6], cfrc[JBO_MAXB][6];
  int b, i, k;
  cacc[0][0] = cacc[0][1] = cacc[0][2] = 0;
  cacc[0][3] = -m->gravity[0]; cacc[0][4] = -m->gravity[1]; cacc[0][5] = -m->gravity[2];
  memset(cfrc[0], 0, sizeof(cfrc[0]));
  for (b = 1; b < m->nbody; b++) {
    double t1[6], t2[6];
    int da = m->body_dofadr[b];
    memcpy(cacc[b], cacc[m->body_parentid[b]], sizeof(cacc[b]));
    for (i = 0; da >= 0 && i < m->body_dofnum[b]; i++)
      for (k = 0; k < 6; k++) cacc[b][k] += d->cdof_dot[da + i][k] * d->qvel[da + i];
    mulInertVec(t1, d->cinert[b], cacc[b]);
    mulInertVec(t2, d->cinert[b], d->cvel[b]);
    crossForce(cfrc[b], d->cvel[b], t2);
    for (k = 0; k < 6; k++) cfrc[b][k] += t1[k];
  }
  for (b = m->nbody - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    if (p > 0) for (k = 0; k < 6; k++) cfrc[p][k] += cfrc[b][k];
  }
  for (i = 0; i < m->nv; i++) {
    double s = 0;
    for (k = 0; k < 6; k++) s += d->cdof[i][k] * cfrc[m->dof_bodyid[i]][k];
    d->qfrc_bias[i] = s;
  }
}

static void sensorVel(const JboModel* m, JboData* d) {
  int b = m->sensor_body, k;
  if (b < 0) return;
  {
    const double* O = d->subtree_com[m->body_rootid[b]];
    double off[3] = {d->xipos[b][0] - O[0], d->xipos[b][1] - O[1], d->xipos[b][2] - O[2]}, c[3];
    cross3(c, d->cvel[b], off);
    for (k = 0; k < 3; k++) d->sensordata[k] = d->cvel[b][3 + k] + c[k];
  }
}

/* ------------------------------------------------------------------ state checks (mj_checkPos / Vel / Acc) */
static int isBad(double x) { return x != x || x > MAXVAL || x < -MAXVAL; }

static void checkReset(const JboModel* m, JboData* d, const double* v, int n, int flag) {
  int i;
  for (i = 0; i < n; i++)
    if (isBad(v[i])) {
      int keep = d->warning_bad | flag;
      jbo_reset_data(m, d);
      d->warning_bad = keep;
      return;
    }
}

/* ------------------------------------------------------------------ public stage drivers */
void jbo_reset_data(const JboModel* m, JboData* d) {
  memset(d, 0, sizeof(*d));
  memcpy(d->qpos, m->qpos0, sizeof(double) * m->nq);
}

static void fwdPosition(const JboModel* m, JboData* d) {
  int u;
  kinematics(m, d);
  comPos(m, d);
  crb(m, d);
  collision(m, d);
  makeConstraint(m, d);
  for (u = 0; u < m->nu; u++)   /* mj_transmission, joint type */
    d->actuator_length[u] = m->actuator_gear[u] * d->qpos[m->jnt_qposadr[m->actuator_trnid[u]]];
}

static void fwdVelocity(const JboModel* m, JboData* d) {
  int u;
  for (u = 0; u < m->nu; u++)
    d->actuator_velocity[u] = m->actuator_gear[u] * d->qvel[m->jnt_dofadr[m->actuator_trnid[u]]];
  comVel(m, d);
  passive(m, d);
  referenceConstraint(m, d);
  rneBias(m, d);
}

void jbo_step1(const JboModel* m, JboData* d) {
  checkReset(m, d, d->qpos, m->nq, 1);
  checkReset(m, d, d->qvel, m->nv, 2);
  fwdPosition(m, d);
  fwdVelocity(m, d);
  sensorVel(m, d);
}

static void fwdActuation(const JboModel* m, JboData* d) {
  int u, i;
  for (i = 0; i < m->nv; i++) d->qfrc_actuator[i] = 0;
  for (u = 0; u < m->nu; u++) {
    double ctrl = d->ctrl[u], gain, bias;
    if (m->actuator_ctrllimited[u])
      ctrl = fmin(m->actuator_ctrlrange[u][1], fmax(m->actuator_ctrlrange[u][0], ctrl));
    gain = m->actuator_gainprm[u][0];
    bias = m->actuator_biasprm[u][0] + m->actuator_biasprm[u][1] * d->actuator_length[u]
           + m->actuator_biasprm[u][2] * d->actuator_velocity[u];
    d->actuator_force[u] = gain * ctrl + bias;
    d->qfrc_actuator[m->jnt_dofadr[m->actuator_trnid[u]]] += m->actuator_gear[u] * d->actuator_force[u];
  }
}

static void fwdAcceleration(const JboModel* m, JboData* d) {
  double L[JBO_MAXV * JBO_MAXV];
  int i, j, nv = m->nv;
  for (i = 0; i < nv; i++) d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_actuator[i];
  for (i = 0; i < nv; i++) for (j = 0; j < nv; j++) L[i * nv + j] = d->qM[i][j];
  cholFactor(L, nv, nv);
  cholSolve(L, nv, nv, d->qacc_smooth, d->qfrc_smooth);
}

/* constraint update: forces, active set, cost (mj_constraintUpdate for inequality rows) */
static double constraintUpdate(const JboData* d, const double* jar, double* force, int* active) {
  double cost = 0;
  int r;
  for (r = 0; r < d->nefc; r++) {
    if (jar[r] < 0) {
      force[r] = -d->efc_D[r] * jar[r];
      cost += 0.5 * d->efc_D[r] * jar[r] * jar[r];
      active[r] = 1;
    } else {
      force[r] = 0;
      active[r] = 0;
    }
  }
  return cost;
}

typedef struct { double alpha; int row; } Brk;
static int brkCmp(const void* a, const void* b) {
  double x = ((const Brk*)a)->alpha, y = ((const Brk*)b)->alpha;
  return (x > y) - (x < y);
}

/* exact minimiser of the piecewise-quadratic 1-D cost along the search direction */
static double exactLineSearch(const JboData* d, int nv, const double* jar, const double* Jv, double g1, double g2) {
  Brk brk[JBO_MAXEFC];
  int nb = 0, r, k;
  double A = g1, B = g2, lo = 0;
  (void)nv;
  for (r = 0; r < d->nefc; r++) {
    int act = jar[r] < 0 || (jar[r] == 0 && Jv[r] < 0);
    if (act) { A += d->efc_D[r] * jar[r] * Jv[r]; B += d->efc_D[r] * Jv[r] * Jv[r]; }
    if (Jv[r] != 0) {
      double a = -jar[r] / Jv[r];
      if (a > 0) { brk[nb].alpha = a; brk[nb].row = r; nb++; }
    }
  }
  if (A >= 0) return 0;     /* not a descent direction */
  qsort(brk, nb, sizeof(Brk), brkCmp);
  for (k = 0; k <= nb; k++) {
    double hi = k < nb ? brk[k].alpha : INFINITY;
    double root = B > 0 ? -A / B : INFINITY;
    if (root >= lo && root <= hi) return root;
    if (k < nb) {
      r = brk[k].row;
      if (Jv[r] > 0) { A -= d->efc_D[r] * jar[r] * Jv[r]; B -= d->efc_D[r] * Jv[r] * Jv[r]; }   /* leaves the active set */
      else { A += d->efc_D[r] * jar[r] * Jv[r]; B += d->efc_D[r] * Jv[r] * Jv[r]; }            /* enters */
      lo = hi;
    }
  }
  return lo;
}

static void solveNewton(const JboModel* m, JboData* d) {
  int nv = m->nv, nefc = d->nefc, i, j, r, iter;
  double Ma[JBO_MAXV], grad[JBO_MAXV], search[JBO_MAXV], Mv[JBO_MAXV], Mgrad[JBO_MAXV];
  double jar[JBO_MAXEFC] = {0}, Jv[JBO_MAXEFC];
  int active[JBO_MAXEFC] = {0};
  double H[JBO_MAXV * JBO_MAXV];
  double scale = 1 / (m->meaninertia * (nv > 1 ? nv : 1));
  double cost, gauss;

  for (i = 0; i < nv; i++) { Ma[i] = 0; for (j = 0; j < nv; j++) Ma[i] += d->qM[i][j] * d->qacc[j]; }
  for (r = 0; r < nefc; r++) {
    double s = 0;
    for (i = 0; i < nv; i++) s += d->efc_J[r][i] * d->qacc[i];
    jar[r] = s - d->efc_aref[r];
  }
  cost = constraintUpdate(d, jar, d->efc_force, active);
  gauss = 0;
  for (i = 0; i < nv; i++) gauss += 0.5 * (Ma[i] - d->qfrc_smooth[i]) * (d->qacc[i] - d->qacc_smooth[i]);
  cost += gauss;
  for (i = 0; i < nv; i++) {
    double s = 0;
    for (r = 0; r < nefc; r++) s += d->efc_J[r][i] * d->efc_force[r];
    d->qfrc_constraint[i] = s;
    grad[i] = Ma[i] - d->qfrc_smooth[i] - s;
  }
  iter = 0;
  while (iter < m->iterations) {
    double g1 = 0, g2 = 0, alpha, oldcost, improvement, gradient;
    /* Hessian H = M + J' diag(D*active) J */
    for (i = 0; i < nv; i++) for (j = 0; j < nv; j++) H[i * nv + j] = d->qM[i][j];
    for (r = 0; r < nefc; r++) if (active[r])
      for (i = 0; i < nv; i++) {
        double t = d->efc_D[r] * d->efc_J[r][i];
        if (t != 0) for (j = 0; j < nv; j++) H[i * nv + j] += t * d->efc_J[r][j];
      }
    cholFactor(H, nv, nv);
    cholSolve(H, nv, nv, Mgrad, grad);
    for (i = 0; i < nv; i++) search[i] = -Mgrad[i];
    for (i = 0; i < nv; i++) { Mv[i] = 0; for (j = 0; j < nv; j++) Mv[i] += d->qM[i][j] * search[j]; }
    for (r = 0; r < nefc; r++) {
      double s = 0;
      for (i = 0; i < nv; i++) s += d->efc_J[r][i] * search[i];
      Jv[r] = s;
    }
    for (i = 0; i < nv; i++) { g1 += search[i] * (Ma[i] - d->qfrc_smooth[i]); g2 += search[i] * Mv[i]; }
    alpha = exactLineSearch(d, nv, jar, Jv, g1, g2);
    if (alpha == 0) break;
    for (i = 0; i < nv; i++) { d->qacc[i] += alpha * search[i]; Ma[i] += alpha * Mv[i]; }
    for (r = 0; r < nefc; r++) jar[r] += alpha * Jv[r];
    oldcost = cost;
    cost = constraintUpdate(d, jar, d->efc_force, active);
    gauss = 0;
    for (i = 0; i < nv; i++) gauss += 0.5 * (Ma[i] - d->qfrc_smooth[i]) * (d->qacc[i] - d->qacc_smooth[i]);
    cost += gauss;
    gradient = 0;
    for (i = 0; i < nv; i++) {
      double s = 0;
      for (r = 0; r < nefc; r++) s += d->efc_J[r][i] * d->efc_force[r];
      d->qfrc_constraint[i] = s;
      grad[i] = Ma[i] - d->qfrc_smooth[i] - s;
      gradient += grad[i] * grad[i];
    }
    improvement = scale * (oldcost - cost);
    gradient = scale * sqrt(gradient);
    iter++;
    if (improvement < m->tolerance || gradient < m->tolerance) break;
  }
  d->solver_iter = iter;
}

static void fwdConstraint(const JboModel* m, JboData* d) {
  int nv = m->nv, nefc = d->nefc, i, j, r;
  d->solver_iter = 0;
  if (nefc == 0) {
    memcpy(d->qacc, d->qacc_smooth, sizeof(double) * nv);
    memcpy(d->qacc_warmstart, d->qacc_smooth, sizeof(double) * nv);
    memset(d->qfrc_constraint, 0, sizeof(double) * nv);
    return;
  }
  /* warmstart: better of qacc_warmstart and qacc_smooth by total cost */
  {
    double jar[JBO_MAXEFC] = {0}, force[JBO_MAXEFC], Ma[JBO_MAXV], cost_warm, cost_smooth;
    int active[JBO_MAXEFC];
    for (r = 0; r < nefc; r++) {
      double s = 0;
      for (i = 0; i < nv; i++) s += d->efc_J[r][i] * d->qacc_warmstart[i];
      jar[r] = s - d->efc_aref[r];
    }
    cost_warm = constraintUpdate(d, jar, force, active);
    for (i = 0; i < nv; i++) { Ma[i] = 0; for (j = 0; j < nv; j++) Ma[i] += d->qM[i][j] * d->qacc_warmstart[j]; }
    for (i = 0; i < nv; i++) cost_warm += 0.5 * (Ma[i] - d->qfrc_smooth[i]) * (d->qacc_warmstart[i] - d->qacc_smooth[i]);
    for (r = 0; r < nefc; r++) {
      double s = 0;
      for (i = 0; i < nv; i++) s += d->efc_J[r][i] * d->qacc_smooth[i];
      jar[r] = s - d->efc_aref[r];
    }
    cost_smooth = constraintUpdate(d, jar, force, active);
    if (cost_warm > cost_smooth) memcpy(d->qacc, d->qacc_smooth, sizeof(double) * nv);
    else memcpy(d->qacc, d->qacc_warmstart, sizeof(double) * nv);
  }
  solveNewton(m, d);
  memcpy(d->qacc_warmstart, d->qacc, sizeof(double) * nv);
}

static void integratePos(const JboModel* m, double* qpos, const double* qvel, double h) {
  int j, k;
  for (j = 0; j < m->njnt; j++) {
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    if (m->jnt_type[j] == JBO_FREE) {
      double ax[3] = {qvel[da + 3], qvel[da + 4], qvel[da + 5]}, qrot[4], angle;
      for (k = 0; k < 3; k++) qpos[qa + k] += h * qvel[da + k];
      /* mju_quatIntegrate: right-multiply by the local rotation exp(h*w) */
      angle = h * normalize3(ax);
      axisAngle2Quat(qrot, ax, angle);
      normalize4(qpos + qa + 3);
      mulQuat(qpos + qa + 3, qpos + qa + 3, qrot);
    } else {
      qpos[qa] += h * qvel[da];
    }
  }
}

static void euler(const JboModel* m, JboData* d) {
  int nv = m->nv, i, j, damped = 0;
  double qacc[JBO_MAXV];
  for (i = 0; i < nv; i++) if (m->dof_damping[i] > 0) damped = 1;
  if (!damped) {
    memcpy(qacc, d->qacc, sizeof(double) * nv);
  } else {
    /* implicit in joint damping: (M + h*diag(B)) qacc = qfrc_smooth + qfrc_constraint */
    double L[JBO_MAXV * JBO_MAXV], rhs[JBO_MAXV];
    for (i = 0; i < nv; i++) for (j = 0; j < nv; j++) L[i * nv + j] = d->qM[i][j];
    for (i = 0; i < nv; i++) L[i * nv + i] += m->timestep * m->dof_damping[i];
    for (i = 0; i < nv; i++) rhs[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i];
    cholFactor(L, nv, nv);
    cholSolve(L, nv, nv, qacc, rhs);
  }
  for (i = 0; i < nv; i++) d->qvel[i] += m->timestep * qacc[i];
  integratePos(m, d->qpos, d->qvel, m->timestep);
  d->time += m->timestep;
}

void jbo_step2(const JboModel* m, JboData* d) {
  fwdActuation(m, d);
  fwdAcceleration(m, d);
  fwdConstraint(m, d);
  checkReset(m, d, d->qacc, m->nv, 4);
  d->total_solver_iter += d->solver_iter;
  d->total_contacts += d->nefc / 4;
  d->total_substeps += 1;
  euler(m, d);
}

void jbo_forward(const JboModel* m, JboData* d) {
  jbo_step1(m, d);
  fwdActuation(m, d);
  fwdAcceleration(m, d);
  fwdConstraint(m, d);
}

void jbo_physics_step(const JboModel* m, JboData* d) {
  jbo_step2(m, d);
  jbo_step1(m, d);
}

void jbo_control_step(const JboModel* m, JboData* d, const double* ctrl, int n_sub_steps) {
  int s, u;
  for (u = 0; u < m->nu; u++) d->ctrl[u] = ctrl[u];
  for (s = 0; s < n_sub_steps; s++) jbo_physics_step(m, d);
}

int jbo_sizeof_model(void) { return (int)sizeof(JboModel); }
int jbo_sizeof_data(void) { return (int)sizeof(JboData); }

/* ------------------------------------------------------------------ task layer (jitterbug.py:180-317, :673-763, :840-925) */
static double wrapPi(double a) {
  while (a > PI) a -= 2 * PI;
  while (a <= -PI) a += 2 * PI;
  return a;
}
static double normObs(double v, double lo, double hi) { return (v - lo) / (hi - lo) * 2.0 - 1.0; }

/* dm_control.utils.rewards.tolerance with bounds=(lo,hi); sigmoid: 0 gaussian, 1 linear, 2 cosine */
static double tolerance(double x, double lo, double hi, double margin, int sigmoid, double vam) {
  int in_bounds = lo <= x && x <= hi;
  double dd;
  if (in_bounds) return 1;
  if (margin == 0) return 0;
  dd = (x < lo ? lo - x : x - hi) / margin;
  if (sigmoid == 0) {
    double scale = sqrt(-2 * log(vam));
    return exp(-0.5 * (dd * scale) * (dd * scale));
  } else if (sigmoid == 1) {
    double sx = dd * (1 - vam);
    return fabs(sx) < 1 ? 1 - sx : 0;
  } else {
    double sx = dd * acos(2 * vam - 1) / PI;
    return fabs(sx) < 1 ? (1 + cos(PI * sx)) / 2 : 0;
  }
}

int jbo_task_obs_reward(const JboModel* m, const JboData* d, int task, double* obs, double* reward) {
  const double* q = d->qpos;
  const double* v = d->qvel;
  double Rj[9], Rt[9], jb_yaw, tg_yaw, angle, motor, rel[3], tib[3], vit[3], dist;
  double r_pos, r_head, r_vel, r_up, r = 0;
  int n = 0, k;
  const double* tq = d->xquat[m->target_body];
  const double* tp = d->geom_xpos[m->target_geom];
  quat2Mat(Rj, q + 3);
  quat2Mat(Rt, tq);
  jb_yaw = atan2(Rj[3], Rj[0]) - PI / 2;
  tg_yaw = atan2(Rt[3], Rt[0]);
  angle = wrapPi(tg_yaw - jb_yaw);
  motor = wrapPi(q[15] + PI / 2);
  for (k = 0; k < 3; k++) rel[k] = tp[k] - q[k];
  mulMatTVec3(tib, Rj, rel);                 /* inv(R) == R' */
  mulMatTVec3(vit, Rt, d->sensordata);
  obs[n++] = normObs(q[0], -2, 2); obs[n++] = normObs(q[1], -2, 2); obs[n++] = normObs(q[2], 0, 0.1);
  for (k = 3; k < 7; k++) obs[n++] = normObs(q[k], -1, 1);
  for (k = 0; k < 3; k++) obs[n++] = normObs(v[k], -1, 1);
  for (k = 3; k < 6; k++) obs[n++] = normObs(v[k], -35, 35);
  obs[n++] = normObs(motor, -PI, PI);
  obs[n++] = normObs(v[14], -180, 180);
  dist = sqrt(dot3(tib, tib));
  r_pos = tolerance(dist, 0, 0, 0.05, 0, 0.1);
  r_head = tolerance(angle, 0, 0, PI / 2, 2, 0);
  r_vel = tolerance(vit[0], 0.1, INFINITY, 0.1, 1, 0);
  r_up = tolerance(d->xmat[1][8], 1, 1, 0.5, 0, 0.1);
  switch (task) {
    case 0: r = 1 - r_pos; break;
    case 1: obs[n++] = normObs(angle, -PI, PI); r = r_head; break;
    case 2:
      obs[n++] = normObs(angle, -PI, PI);
      for (k = 0; k < 3; k++) obs[n++] = normObs(vit[k], -1, 1);
      r = r_vel; break;
    case 3:
      obs[n++] = normObs(tib[0], -3, 3); obs[n++] = normObs(tib[1], -3, 3); obs[n++] = normObs(tib[2], -0.1, 0.1);
      r = r_pos; break;
    case 4:
      obs[n++] = normObs(tib[0], -3, 3); obs[n++] = normObs(tib[1], -3, 3); obs[n++] = normObs(tib[2], -0.1, 0.1);
      obs[n++] = normObs(angle, -PI, PI);
      r = r_pos * r_head; break;
    default: return -1;
  }
  *reward = r * r_up;
  return n;
}

typedef struct {
  const JboModel* models; JboData* datas; int nenv, task, nsteps, n_sub_steps, tid, nthreads;
  const double* actions; double* rewards_out; double* obs_out;
} RolloutJob;

static void* rolloutWorker(void* arg) {
  RolloutJob* j = (RolloutJob*)arg;
  int e, s;
  for (e = j->tid; e < j->nenv; e += j->nthreads) {
    double obs[32], rew = 0;
    for (s = 0; s < j->nsteps; s++) {
      int n;
      jbo_control_step(&j->models[e], &j->datas[e], &j->actions[(long)s * j->nenv + e], j->n_sub_steps);
      n = jbo_task_obs_reward(&j->models[e], &j->datas[e], j->task, obs, &rew);
      if (j->rewards_out) j->rewards_out[(long)s * j->nenv + e] = rew;
      if (j->obs_out && s == j->nsteps - 1) memcpy(j->obs_out + (long)e * 19, obs, sizeof(double) * (n > 0 ? n : 0));
    }
  }
  return 0;
}

/* heuristic_policies.py:6-136 on the normalised observation vector (layout of jbo_task_obs_reward), fp64 */
static double hFace(double a) { double v = 3 * a / PI; return 0.9 * (v > 1 ? 1 : (v < -1 ? -1 : v)); }
static double hForward(double motor_angle, double motor_vel, double offset) {
  const double kick = PI / 4, speed = 0.3;                               /* move_forward(), :29-56 */
  if (motor_angle < offset - kick) return speed;
  if (motor_angle > offset + kick) return -speed;
  return motor_vel > 0 ? speed : -speed;
}
static double hOrient(double a, double* offset) {                        /* optimal_orientation_to_move(), :119-136 */
  *offset = 0;
  if (a >= -45 * PI / 180 && a <= 45 * PI / 180) return a;
  if (a > 45 * PI / 180 && a <= PI) { *offset = PI / 2; return fabs(fabs(a) - 90 * PI / 180); }
  if (a >= -PI && a < -45 * PI / 180) { *offset = -PI / 2; return -fabs(fabs(a) - 90 * PI / 180); }
  return a;
}
double jbo_heuristic(int task, const double* obs) {
  const double thr = 20 * PI / 180;
  double ma = obs[13], mv = obs[14], a, off;
  if (task == 0) return hForward(ma, mv, 0);                             /* move_from_origin, :59-61 */
  if (task == 1) return hFace(obs[15]);                                  /* face_direction, :6-26 */
  if (task == 2) {                                                       /* move_in_direction, :64-78 */
    a = hOrient(obs[15], &off);
    return fabs(a) > thr ? hFace(a) : hForward(ma, mv, off);
  }
  a = atan2(obs[15], -obs[16]);
  if (task == 3) {                                                       /* move_to_position, :81-95 */
    a = hOrient(a, &off);
    return fabs(a) > thr ? hFace(a) : hForward(ma, mv, off);
  }
  if (fabs(a) > thr) return hFace(a);                                    /* move_to_pose, :98-117 */
  if (sqrt(obs[15] * obs[15] + obs[16] * obs[16]) > 0.01) return hForward(ma, mv, 0);
  return hFace(obs[18]);
}

/* closed-loop rollouts under the heuristic policy of `task` (benchmarks/evaluate_policy.py:10-35: act on the current
 * observation, step, record the reward); the datas must have been forwarded (jbo_forward) */
static void* heuristicWorker(void* arg) {
  RolloutJob* j = (RolloutJob*)arg;
  int e, s;
  for (e = j->tid; e < j->nenv; e += j->nthreads) {
    double obs[32], rew = 0, act;
    jbo_task_obs_reward(&j->models[e], &j->datas[e], j->task, obs, &rew);
    for (s = 0; s < j->nsteps; s++) {
      act = jbo_heuristic(j->task, obs);
      jbo_control_step(&j->models[e], &j->datas[e], &act, j->n_sub_steps);
      jbo_task_obs_reward(&j->models[e], &j->datas[e], j->task, obs, &rew);
      if (j->rewards_out) j->rewards_out[(long)s * j->nenv + e] = rew;
    }
  }
  return 0;
}

void jbo_rollout_heuristic(const JboModel* models, JboData* datas, int nenv, int task, int nsteps, int n_sub_steps,
                           int nthreads, double* rewards_out) {
  pthread_t th[256];
  RolloutJob jobs[256];
  int t;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  for (t = 0; t < nthreads; t++) {
    RolloutJob jb = {models, datas, nenv, task, nsteps, n_sub_steps, t, nthreads, 0, rewards_out, 0};
    jobs[t] = jb;
    if (nthreads > 1) pthread_create(&th[t], 0, heuristicWorker, &jobs[t]);
  }
  if (nthreads == 1) heuristicWorker(&jobs[0]);
  else for (t = 0; t < nthreads; t++) pthread_join(th[t], 0);
}

void jbo_rollout(const JboModel* models, JboData* datas, int nenv, int task, const double* actions, int nsteps,
                 int n_sub_steps, int nthreads, double* rewards_out, double* obs_out) {
  pthread_t th[256];
  RolloutJob jobs[256];
  int t;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  for (t = 0; t < nthreads; t++) {
    RolloutJob jb = {models, datas, nenv, task, nsteps, n_sub_steps, t, nthreads, actions, rewards_out, obs_out};
    jobs[t] = jb;
    if (nthreads > 1) pthread_create(&th[t], 0, rolloutWorker, &jobs[t]);
  }
  if (nthreads == 1) rolloutWorker(&jobs[0]);
  else for (t = 0; t < nthreads; t++) pthread_join(th[t], 0);
}
