// One augmented Jitterbug: perturbation vector -> model (fp64).  Plain C++ under JBA_HD so the test-only host build
// (tests/hostsim/augsim.cpp) runs the very code the CUDA kernel in jb_augment.cu runs.  See jb_augment.cu for the story.
#ifndef JB_AUGMENT_CUH
#define JB_AUGMENT_CUH

#include <math.h>
#include "../../include/jbsim.h"

#if defined(__CUDACC__)
#define JBA_HD __host__ __device__
#else
#define JBA_HD
#endif

namespace jba {


struct V3 { double x, y, z; };
JBA_HD inline V3 v3(const double* p) { return {p[0], p[1], p[2]}; }
JBA_HD inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
JBA_HD inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
JBA_HD inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
JBA_HD inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
JBA_HD inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
JBA_HD inline V3 unit(V3 a) { return (1.0 / sqrt(dot(a, a))) * a; }
JBA_HD inline void put(double* p, V3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }

// "fromto" cylinder -> centre, half length, frame.  MuJoCo points the geom z axis from "to" towards "from" and builds
// the frame as the rotation of z onto it about z x v (mjcf.py add_geom / z2quat).
JBA_HD inline void cyl_fromto(JbGeomD& g, V3 from, V3 to) {
  V3 d = from - to;
  double len = sqrt(dot(d, d));
  V3 v = (1.0 / len) * d;
  put(g.center, 0.5 * (from + to));
  g.size[1] = 0.5 * len;
  double s = sqrt(v.x * v.x + v.y * v.y), c = v.z;
  double kx = 1.0, ky = 0.0;
  if (s >= 1e-10) { kx = -v.y / s; ky = v.x / s; }
  // R = c I + s [k]x + (1 - c) k k^T, k = (kx, ky, 0)
  double r[9] = {c + (1 - c) * kx * kx, (1 - c) * kx * ky, s * ky,
                 (1 - c) * kx * ky, c + (1 - c) * ky * ky, -s * kx,
                 -s * ky, s * kx, c};
  for (int i = 0; i < 9; i++) g.rot[i] = r[i];
  g.rbound = sqrt(g.size[0] * g.size[0] + g.size[1] * g.size[1]);
}

enum { T_SPHERE, T_CYL, T_BOX, T_ELL };

// mass and diagonal inertia in the geom frame (MuJoCo compiler formulas; mjcf.py geom_mass_inertia)
JBA_HD inline void geom_mass(int type, const double* size, double density, double& m, double* i) {
  const double PI = 3.14159265358979323846;
  if (type == T_SPHERE) {
    double r = size[0];
    m = density * 4.0 / 3.0 * PI * r * r * r;
    i[0] = i[1] = i[2] = 0.4 * m * r * r;
  } else if (type == T_CYL) {
    double r = size[0], h = 2.0 * size[1];
    m = density * PI * r * r * h;
    i[0] = i[1] = m * (3.0 * r * r + h * h) / 12.0;
    i[2] = 0.5 * m * r * r;
  } else {
    double a = size[0], b = size[1], c = size[2];
    double k = type == T_ELL ? 1.0 / 5.0 : 1.0 / 3.0;
    m = density * (type == T_ELL ? 4.0 / 3.0 * PI : 8.0) * a * b * c;
    i[0] = k * m * (b * b + c * c); i[1] = k * m * (a * a + c * c); i[2] = k * m * (a * a + b * b);
  }
}

struct BodyGeom { int type; const JbGeomD* g; double density; };

// body mass, COM, inertia about the COM in root axes (xx, yy, zz, xy, xz, yz) from its geoms: two passes like the
// host compiler (COM first, then inertias shifted to it)
JBA_HD inline void body_inertial(const BodyGeom* bg, int n, double& mass, double* com, double* I6) {
  double gm[4], gi[4][3];
  mass = 0;
  V3 mc = {0, 0, 0};
  for (int k = 0; k < n; k++) {
    geom_mass(bg[k].type, bg[k].g->size, bg[k].density, gm[k], gi[k]);
    mass += gm[k];
    mc = mc + gm[k] * v3(bg[k].g->center);
  }
  V3 c = (1.0 / mass) * mc;
  put(com, c);
  double I[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int k = 0; k < n; k++) {
    const double* R = bg[k].g->rot;
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++)
        I[a][b] += R[3 * a] * gi[k][0] * R[3 * b] + R[3 * a + 1] * gi[k][1] * R[3 * b + 1] + R[3 * a + 2] * gi[k][2] * R[3 * b + 2];
    V3 d = v3(bg[k].g->center) - c;
    double dd = dot(d, d), dv[3] = {d.x, d.y, d.z};
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) I[a][b] += gm[k] * ((a == b ? dd : 0.0) - dv[a] * dv[b]);
  }
  I6[0] = I[0][0]; I6[1] = I[1][1]; I6[2] = I[2][2]; I6[3] = I[0][1]; I6[4] = I[0][2]; I6[5] = I[1][2];
}

// mj_setConst at qpos0 (mjcf.py _set_const): M from body Jacobians, Cholesky inverse, translational body_invweight0
// per moving body, meaninertia.  Dofs: 0-2 root linear, 3-5 root angular (body axes = root axes at home), 6+2k / 7+2k
// hip / knee of leg k, 14 rotor.
struct BodyRef { double m; const double* c; const double* I6; int h0, h1; };   // h0/h1: hinge dofs moving it, -1 = none

JBA_HD inline void body_jac(const JbModelD& M, const BodyRef& b, double jp[3][15], double jr[3][15]) {
  for (int r = 0; r < 3; r++)
    for (int j = 0; j < 15; j++) { jp[r][j] = 0; jr[r][j] = 0; }
  V3 c = v3(b.c);
  for (int j = 0; j < 3; j++) {
    jp[j][j] = 1;
    jr[j][3 + j] = 1;
    V3 e = {j == 0 ? 1.0 : 0.0, j == 1 ? 1.0 : 0.0, j == 2 ? 1.0 : 0.0};
    V3 col = cross(e, c);
    jp[0][3 + j] = col.x; jp[1][3 + j] = col.y; jp[2][3 + j] = col.z;
  }
  const int hs[2] = {b.h0, b.h1};
  for (int q = 0; q < 2; q++) {
    int h = hs[q];
    if (h < 0) continue;
    V3 a, p;
    if (h == 14) { a = v3(M.am); p = v3(M.pm); }
    else {
      int k = (h - 6) >> 1;
      if ((h - 6) & 1) { a = v3(M.a2[k]); p = v3(M.p2[k]); } else { a = v3(M.a1[k]); p = v3(M.p1[k]); }
    }
    V3 col = cross(a, c - p);
    jp[0][h] = col.x; jp[1][h] = col.y; jp[2][h] = col.z;
    jr[0][h] = a.x; jr[1][h] = a.y; jr[2][h] = a.z;
  }
}

JBA_HD inline void set_const(JbModelD& M) {
  BodyRef bodies[JB_NSLOT];
  bodies[0] = {M.m0, M.c0, M.I0, -1, -1};
  for (int k = 0; k < JB_NLEG; k++) {
    bodies[1 + 2 * k] = {M.mu[k], M.cu[k], M.Iu[k], 6 + 2 * k, -1};
    bodies[2 + 2 * k] = {M.ml[k], M.cl[k], M.Il[k], 6 + 2 * k, 7 + 2 * k};
  }
  bodies[9] = {M.mm, M.cm, M.Im, 14, -1};
  double A[15][15], Li[15][15], jp[3][15], jr[3][15];
  for (int i = 0; i < 15; i++)
    for (int j = 0; j < 15; j++) A[i][j] = 0;
  for (int b = 0; b < JB_NSLOT; b++) {
    body_jac(M, bodies[b], jp, jr);
    const double* I6 = bodies[b].I6;
    const double I[3][3] = {{I6[0], I6[3], I6[4]}, {I6[3], I6[1], I6[5]}, {I6[4], I6[5], I6[2]}};
    for (int i = 0; i < 15; i++) {
      double Ijr[3];
      for (int r = 0; r < 3; r++) Ijr[r] = I[r][0] * jr[0][i] + I[r][1] * jr[1][i] + I[r][2] * jr[2][i];
      for (int j = 0; j <= i; j++) {
        double s = bodies[b].m * (jp[0][i] * jp[0][j] + jp[1][i] * jp[1][j] + jp[2][i] * jp[2][j]);
        s += Ijr[0] * jr[0][j] + Ijr[1] * jr[1][j] + Ijr[2] * jr[2][j];
        A[i][j] += s;
      }
    }
  }
  double tr = 0;
  for (int i = 0; i < 15; i++) tr += A[i][i];
  M.meaninertia = tr / 15.0;
  // Cholesky A = L L^T in place (lower), then Li = L^-1; M^-1 = Li^T Li
  for (int j = 0; j < 15; j++) {
    double d = A[j][j];
    for (int k = 0; k < j; k++) d -= A[j][k] * A[j][k];
    d = sqrt(d);
    A[j][j] = d;
    for (int i = j + 1; i < 15; i++) {
      double s = A[i][j];
      for (int k = 0; k < j; k++) s -= A[i][k] * A[j][k];
      A[i][j] = s / d;
    }
  }
  for (int c = 0; c < 15; c++)
    for (int i = 0; i < 15; i++) {
      double s = i == c ? 1.0 : 0.0;
      for (int k = c; k < i; k++) s -= A[i][k] * Li[k][c];
      Li[i][c] = i < c ? 0.0 : s / A[i][i];
    }
  for (int b = 0; b < JB_NSLOT; b++) {
    body_jac(M, bodies[b], jp, jr);
    double acc = 0;
    for (int r = 0; r < 3; r++)
      for (int i = 0; i < 15; i++) {          // | Li jp_r |^2 = jp_r M^-1 jp_r^T
        double s = 0;
        for (int j = 0; j <= i; j++) s += Li[i][j] * jp[r][j];
        acc += s * s;
      }
    M.invw[b] = acc / 3.0;
  }
}


// deltas: JB_AUG_NDELTA doubles in include/jbsim.h's order
JBA_HD inline void augment_model(const JbAugBase& base, const double* dl, JbModelD& M) {
  M = base.model;
  double dens[JB_AUG_NGEOM];
  for (int g = 0; g < JB_AUG_NGEOM; g++) dens[g] = base.density[g] + (base.follows_default[g] ? dl[0] : 0.0);
  dens[0] += dl[1];                                                     // coreBody1 (augmented_jitterbug.py:146-149)
  dens[1] += dl[2];                                                     // coreBody2 (:150-153)
  const V3 ez = {0, 0, 1};
  for (int k = 0; k < JB_NLEG; k++) {
    V3 du = v3(dl + 3 + 6 * k), dlo = v3(dl + 6 + 6 * k);
    // upper leg (:161-186): the cylinder's "from" end and the knee sphere move by du; hinge 1 keeps its anchor and
    // gets the axis cross(v_upper, z)
    JbGeomD& uc = M.g_ucyl[k];
    V3 ax = {uc.rot[2], uc.rot[5], uc.rot[8]};
    V3 from = v3(uc.center) + uc.size[1] * ax + du, to = v3(uc.center) - uc.size[1] * ax;
    cyl_fromto(uc, from, to);
    V3 vu = unit(to - from);
    put(M.g_knee[k].center, v3(M.g_knee[k].center) + du);
    put(M.a1[k], unit(cross(vu, ez)));
    // lower leg (:188-216): "from" (foot) end moves by dlo, "to" (knee) end by du; foot sphere by dlo; hinge 2 anchor
    // by du, axis cross(v_upper, v_lower)
    JbGeomD& lc = M.g_lcyl[k];
    V3 al = {lc.rot[2], lc.rot[5], lc.rot[8]};
    V3 lfrom = v3(lc.center) + lc.size[1] * al + dlo, lto = v3(lc.center) - lc.size[1] * al + du;
    cyl_fromto(lc, lfrom, lto);
    V3 vl = unit(lto - lfrom);
    put(M.g_foot[k].center, v3(M.g_foot[k].center) + dlo);
    put(M.p2[k], v3(M.p2[k]) + du);
    put(M.a2[k], unit(cross(vu, vl)));
    BodyGeom ub[2] = {{T_CYL, &uc, dens[4 + 4 * k]}, {T_SPHERE, &M.g_knee[k], dens[5 + 4 * k]}};
    body_inertial(ub, 2, M.mu[k], M.cu[k], M.Iu[k]);
    BodyGeom lb[2] = {{T_CYL, &lc, dens[6 + 4 * k]}, {T_SPHERE, &M.g_foot[k], dens[7 + 4 * k]}};
    body_inertial(lb, 2, M.ml[k], M.cl[k], M.Il[k]);
  }
  {
    // rotor (:218-238): threadMass fromto += (dx, dy, 0, dx, dy, dz); hinge anchor += d; the ellipsoid stays (the
    // reference's "mass_geom" branch never matches)
    V3 dp = v3(dl + 27);
    JbGeomD& mc = M.g_mcyl;
    V3 ax = {mc.rot[2], mc.rot[5], mc.rot[8]};
    V3 from = v3(mc.center) + mc.size[1] * ax + V3{dp.x, dp.y, 0.0}, to = v3(mc.center) - mc.size[1] * ax + dp;
    cyl_fromto(mc, from, to);
    put(M.pm, v3(M.pm) + dp);
    BodyGeom mb[2] = {{T_CYL, &mc, dens[20]}, {T_ELL, &M.g_mell, dens[21]}};
    body_inertial(mb, 2, M.mm, M.cm, M.Im);
  }
  {
    BodyGeom rb[4] = {{T_BOX, &M.g_box[0], dens[0]}, {T_BOX, &M.g_box[1], dens[1]}, {T_CYL, &M.g_rcyl, dens[2]},
                      {T_ELL, &M.g_rell, dens[3]}};
    body_inertial(rb, 4, M.m0, M.c0, M.I0);
  }
  M.gear += dl[30];                                                     // (:240-243)
  set_const(M);
}

// the same model in the float struct of the C-ABI (inspection / parity tests)
JBA_HD inline void to_float(const JbModelD& M, JbModel& F) {
#define CPY(f) do { const double* s_ = (const double*)&M.f; float* d_ = (float*)&F.f; \
                  for (unsigned i_ = 0; i_ < sizeof(M.f) / sizeof(double); i_++) d_[i_] = (float)s_[i_]; } while (0)
  CPY(timestep); CPY(gravity); CPY(m0); CPY(c0); CPY(I0); CPY(p1); CPY(a1); CPY(p2); CPY(a2); CPY(mu); CPY(cu); CPY(Iu);
  CPY(ml); CPY(cl); CPY(Il); CPY(stiff); CPY(damp); CPY(pm); CPY(am); CPY(mm); CPY(cm); CPY(Im); CPY(damp_m); CPY(gear);
  CPY(gain); CPY(bias_v); CPY(ctrl_lo); CPY(ctrl_hi); CPY(invw); CPY(friction); CPY(solref_K); CPY(solref_B); CPY(solimp);
  CPY(impratio); CPY(meaninertia); CPY(root_pos0); CPY(target_pos0);
#undef CPY
  F.pad_ = 0.f;
  const JbGeomD* gs[JB_AUG_NGEOM]; JbGeom* gd[JB_AUG_NGEOM];
  int n = 0;
  gs[n] = &M.g_box[0]; gd[n++] = &F.g_box[0]; gs[n] = &M.g_box[1]; gd[n++] = &F.g_box[1];
  gs[n] = &M.g_rcyl; gd[n++] = &F.g_rcyl; gs[n] = &M.g_rell; gd[n++] = &F.g_rell;
  for (int k = 0; k < JB_NLEG; k++) {
    gs[n] = &M.g_ucyl[k]; gd[n++] = &F.g_ucyl[k]; gs[n] = &M.g_knee[k]; gd[n++] = &F.g_knee[k];
    gs[n] = &M.g_lcyl[k]; gd[n++] = &F.g_lcyl[k]; gs[n] = &M.g_foot[k]; gd[n++] = &F.g_foot[k];
  }
  gs[n] = &M.g_mcyl; gd[n++] = &F.g_mcyl; gs[n] = &M.g_mell; gd[n++] = &F.g_mell;
  for (int g = 0; g < n; g++) {
    for (int i = 0; i < 3; i++) { gd[g]->center[i] = (float)gs[g]->center[i]; gd[g]->size[i] = (float)gs[g]->size[i]; }
    for (int i = 0; i < 9; i++) gd[g]->rot[i] = (float)gs[g]->rot[i];
    gd[g]->rbound = (float)gs[g]->rbound; gd[g]->geom_id = gs[g]->geom_id;
    gd[g]->pad[0] = gd[g]->pad[1] = gd[g]->pad[2] = 0;
  }
}

}  // namespace jba
#endif
