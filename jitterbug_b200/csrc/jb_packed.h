// Packed, lane-friendly float layout of one JbModel for the quad kernel (4 lanes per env, lane l = leg l).
//   [0, JBP_GLOBAL)                      : model-wide scalars, root body, rotor, contact parameters (uniform access)
//   JBP_X + 4*field + lane               : the "extra" geoms each lane also collides: lane 0/1 = the two core boxes,
//                                          lane 2 = root cylinder + ellipsoid, lane 3 = rotor cylinder + ellipsoid
//   JBP_L + 4*field + lane               : per-leg constants of leg `lane`
// The stride-4 interleave makes a quad's loads hit 4 consecutive words (conflict-free shared-memory wavefront).
// Built on the host by jb_pack_model(); lives in shared memory inside the kernel.
#ifndef JB_PACKED_H
#define JB_PACKED_H

#include "jb_model.h"

enum {
  G_H = 0, G_GRAV = 1, G_M0 = 2, G_C0 = 3, G_I0 = 6, G_PM = 12, G_AM = 15, G_MM = 18, G_CM = 19, G_IM = 22, G_DAMPM = 28,
  G_GEAR = 29, G_GAIN = 30, G_BIASV = 31, G_CLO = 32, G_CHI = 33, G_FRIC = 34, G_SOLK = 35, G_SOLB = 36, G_SOLIMP = 37,
  G_IMPRATIO = 42, G_MEANINERTIA = 43, G_ROOTPOS0 = 44, G_TGT0 = 47, G_INVW0 = 50, G_INVWM = 51, G_MTOT = 52,
  JBP_GLOBAL = 56
};
// extra-geom block, per lane.  Lanes 0,1: box (center 3, rot 9, half 3, geom id).  Lanes 2,3: cylinder (center 3, axis 3,
// x axis 3, r, half length, geom id) then ellipsoid (center 3, rot 9, size 3, geom id).
enum {
  X_BOX_C = 0, X_BOX_R = 3, X_BOX_H = 12, X_BOX_ID = 15,
  X_CYL_C = 0, X_CYL_AX = 3, X_CYL_XA = 6, X_CYL_R = 9, X_CYL_HL = 10, X_CYL_ID = 11,
  X_ELL_C = 12, X_ELL_R = 15, X_ELL_S = 24, X_ELL_ID = 27,
  JBP_XN = 28
};
enum {
  L_P1 = 0, L_A1 = 3, L_P2 = 6, L_A2 = 9, L_MU = 12, L_CU = 13, L_IU = 16, L_ML = 22, L_CL = 23, L_IL = 26,
  L_K1 = 32, L_K2 = 33, L_D1 = 34, L_D2 = 35,
  L_UC_C = 36, L_UC_AX = 39, L_UC_XA = 42, L_UC_R = 45, L_UC_HL = 46,
  L_KN_C = 47, L_KN_R = 50,
  L_LC_C = 51, L_LC_AX = 54, L_LC_XA = 57, L_LC_R = 60, L_LC_HL = 61,
  L_FT_C = 62, L_FT_R = 65,
  L_INVWU = 66, L_INVWL = 67,
  L_UC_ID = 68, L_KN_ID = 69, L_LC_ID = 70, L_FT_ID = 71,
  JBP_LN = 72
};
#define JBP_X JBP_GLOBAL
#define JBP_L (JBP_X + 4 * JBP_XN)
#define JBP_SIZE (JBP_L + 4 * JBP_LN)     /* 456 floats */

#ifdef __cplusplus
// REAL: float (product) or double (test-only host build).  Geom ids are stored as exact small integers in REAL.
#ifdef __CUDACC__
#define JBP_HD __host__ __device__
#else
#define JBP_HD
#endif
template <typename REAL, typename MODEL>
JBP_HD static inline void jb_pack_model(const MODEL& m, REAL* p) {
  for (int i = 0; i < JBP_SIZE; i++) p[i] = 0;
  p[G_H] = m.timestep; p[G_GRAV] = m.gravity; p[G_M0] = m.m0;
  for (int i = 0; i < 3; i++) { p[G_C0 + i] = m.c0[i]; p[G_PM + i] = m.pm[i]; p[G_AM + i] = m.am[i]; p[G_CM + i] = m.cm[i];
                                p[G_ROOTPOS0 + i] = m.root_pos0[i]; p[G_TGT0 + i] = m.target_pos0[i]; }
  for (int i = 0; i < 6; i++) { p[G_I0 + i] = m.I0[i]; p[G_IM + i] = m.Im[i]; }
  p[G_MM] = m.mm; p[G_DAMPM] = m.damp_m; p[G_GEAR] = m.gear; p[G_GAIN] = m.gain; p[G_BIASV] = m.bias_v;
  p[G_CLO] = m.ctrl_lo; p[G_CHI] = m.ctrl_hi; p[G_FRIC] = m.friction; p[G_SOLK] = m.solref_K; p[G_SOLB] = m.solref_B;
  for (int i = 0; i < 5; i++) p[G_SOLIMP + i] = m.solimp[i];
  p[G_IMPRATIO] = m.impratio; p[G_MEANINERTIA] = m.meaninertia; p[G_INVW0] = m.invw[0]; p[G_INVWM] = m.invw[9];
  REAL mt = m.m0 + m.mm;
  for (int k = 0; k < JB_NLEG; k++) mt += m.mu[k] + m.ml[k];
  p[G_MTOT] = mt;
  // extra geoms
  for (int l = 0; l < 2; l++) {
    REAL* x = p + JBP_X + l;
    for (int i = 0; i < 3; i++) { x[4 * (X_BOX_C + i)] = m.g_box[l].center[i]; x[4 * (X_BOX_H + i)] = m.g_box[l].size[i]; }
    for (int i = 0; i < 9; i++) x[4 * (X_BOX_R + i)] = m.g_box[l].rot[i];
    x[4 * X_BOX_ID] = (REAL)m.g_box[l].geom_id;
  }
  for (int l = 2; l < 4; l++) {
    REAL* x = p + JBP_X + l;
    const auto& cy = l == 2 ? m.g_rcyl : m.g_mcyl;
    const auto& el = l == 2 ? m.g_rell : m.g_mell;
    for (int i = 0; i < 3; i++) {
      x[4 * (X_CYL_C + i)] = cy.center[i]; x[4 * (X_CYL_AX + i)] = cy.rot[3 * i + 2]; x[4 * (X_CYL_XA + i)] = cy.rot[3 * i];
      x[4 * (X_ELL_C + i)] = el.center[i]; x[4 * (X_ELL_S + i)] = el.size[i];
    }
    x[4 * X_CYL_R] = cy.size[0]; x[4 * X_CYL_HL] = cy.size[1]; x[4 * X_CYL_ID] = (REAL)cy.geom_id;
    for (int i = 0; i < 9; i++) x[4 * (X_ELL_R + i)] = el.rot[i];
    x[4 * X_ELL_ID] = (REAL)el.geom_id;
  }
  for (int k = 0; k < JB_NLEG; k++) {
    REAL* q = p + JBP_L + k;
    for (int i = 0; i < 3; i++) {
      q[4 * (L_P1 + i)] = m.p1[k][i]; q[4 * (L_A1 + i)] = m.a1[k][i]; q[4 * (L_P2 + i)] = m.p2[k][i]; q[4 * (L_A2 + i)] = m.a2[k][i];
      q[4 * (L_CU + i)] = m.cu[k][i]; q[4 * (L_CL + i)] = m.cl[k][i];
      q[4 * (L_UC_C + i)] = m.g_ucyl[k].center[i]; q[4 * (L_UC_AX + i)] = m.g_ucyl[k].rot[3 * i + 2]; q[4 * (L_UC_XA + i)] = m.g_ucyl[k].rot[3 * i];
      q[4 * (L_LC_C + i)] = m.g_lcyl[k].center[i]; q[4 * (L_LC_AX + i)] = m.g_lcyl[k].rot[3 * i + 2]; q[4 * (L_LC_XA + i)] = m.g_lcyl[k].rot[3 * i];
      q[4 * (L_KN_C + i)] = m.g_knee[k].center[i]; q[4 * (L_FT_C + i)] = m.g_foot[k].center[i];
    }
    for (int i = 0; i < 6; i++) { q[4 * (L_IU + i)] = m.Iu[k][i]; q[4 * (L_IL + i)] = m.Il[k][i]; }
    q[4 * L_MU] = m.mu[k]; q[4 * L_ML] = m.ml[k];
    q[4 * L_K1] = m.stiff[k][0]; q[4 * L_K2] = m.stiff[k][1]; q[4 * L_D1] = m.damp[k][0]; q[4 * L_D2] = m.damp[k][1];
    q[4 * L_UC_R] = m.g_ucyl[k].size[0]; q[4 * L_UC_HL] = m.g_ucyl[k].size[1];
    q[4 * L_LC_R] = m.g_lcyl[k].size[0]; q[4 * L_LC_HL] = m.g_lcyl[k].size[1];
    q[4 * L_KN_R] = m.g_knee[k].size[0]; q[4 * L_FT_R] = m.g_foot[k].size[0];
    q[4 * L_INVWU] = m.invw[1 + 2 * k]; q[4 * L_INVWL] = m.invw[2 + 2 * k];
    q[4 * L_UC_ID] = (REAL)m.g_ucyl[k].geom_id; q[4 * L_KN_ID] = (REAL)m.g_knee[k].geom_id;
    q[4 * L_LC_ID] = (REAL)m.g_lcyl[k].geom_id; q[4 * L_FT_ID] = (REAL)m.g_foot[k].geom_id;
  }
}
#endif
/* contact slots per lane of the quad kernel (jb_quad.cuh); also sizes the contact-report staging of the C-ABI */
#ifndef JBQ_CAP
#define JBQ_CAP 15                      /* contact slots per lane = the most a lane can produce: its leg's foot and knee
                                           spheres (1 + 1) and two cylinders (4 + 4), plus this lane's extra geoms (a core box:
                                           4 corners; or a cylinder + an ellipsoid: 4 + 1), so no contact is ever dropped.
                                           The active-row masks pack 4 bits per slot. */
#endif

#endif
