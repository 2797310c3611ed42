// Device-side model of one Jitterbug: the compiled MJCF re-expressed in the ROOT BODY FRAME at the home
// configuration (all joint angles zero), i.e. a product-of-exponentials parametrisation.  Every leg joint is a
// rotation about a constant axis through a constant anchor in its parent's home coordinates, so the kernel never
// touches body-local frames, quaternions of child bodies or MuJoCo's com-based spatial vectors.
//
// Produced on the host by jitterbug_b200/model.py::flatten_model from the MJCF-subset compiler output
// (reference model: jitterbug_dmc/jitterbug.xml:37-110; augmented family: jitterbug_dmc/augmented_jitterbug.py:95-267).
// Plain C layout, floats only: this struct crosses the C-ABI (include/jbsim.h) verbatim.
#ifndef JB_MODEL_H
#define JB_MODEL_H

#define JB_NLEG 4
#define JB_NQ 16
#define JB_NV 15
#define JB_OBS_MAX 19
#define JB_NSLOT 10       /* moving bodies: root, (upper, lower) x 4, rotor */

/* The two structs are generated from field macros so the test-only fp64 host build (tests/hostsim) can
 * instantiate the same layout with REAL = double. */
#define JB_GEOM_FIELDS(REAL)                                                                              \
  REAL center[3]; /* home position, root frame */                                                         \
  REAL rot[9];    /* home orientation (row-major, geom->root); cylinders use column 2 (axis) and 0 */     \
  REAL size[3];   /* sphere: r; cylinder: r, half-length; box/ellipsoid: half-sizes */                    \
  REAL rbound;    /* bounding-sphere radius */                                                            \
  int geom_id;    /* MuJoCo geom id, used only to report contacts for parity checks */                    \
  int pad[3];

#define JB_MODEL_FIELDS(REAL, GEOM)                                                                       \
  /* ---- options */                                                                                      \
  REAL timestep, gravity; /* gravity magnitude along world -z */                                          \
  /* ---- root (core) body: mass, COM, inertia about COM (xx,yy,zz,xy,xz,yz), root axes */                \
  REAL m0, c0[3], I0[6];                                                                                  \
  /* ---- legs, in dof order (MuJoCo bodies leg2, leg3, leg1, leg4) */                                    \
  REAL p1[JB_NLEG][3], a1[JB_NLEG][3]; /* hinge 1 anchor / unit axis (root frame, constant) */           \
  REAL p2[JB_NLEG][3], a2[JB_NLEG][3]; /* hinge 2 anchor / unit axis at home */                          \
  REAL mu[JB_NLEG], cu[JB_NLEG][3], Iu[JB_NLEG][6]; /* upper leg mass / COM / inertia at home */         \
  REAL ml[JB_NLEG], cl[JB_NLEG][3], Il[JB_NLEG][6]; /* lower leg */                                      \
  REAL stiff[JB_NLEG][2], damp[JB_NLEG][2];                                                               \
  /* ---- rotor ("mass" body) */                                                                          \
  REAL pm[3], am[3], mm, cm[3], Im[6];                                                                    \
  REAL damp_m;                                                                                            \
  /* ---- actuator: force = gain*clamp(ctrl) + bias_v*(gear*qvel_motor); torque = gear*force */           \
  REAL gear, gain, bias_v, ctrl_lo, ctrl_hi;                                                              \
  /* ---- contact parameters (floor vs body geoms; identical for all pairs in this model family) */       \
  REAL invw[JB_NSLOT]; /* body_invweight0 (translational) per moving body slot */                         \
  REAL friction;       /* sliding friction mu */                                                          \
  REAL solref_K, solref_B; /* 1/(dmax^2 tc^2 dr^2), 2/(dmax tc) with refsafe applied */                   \
  REAL solimp[5];                                                                                         \
  REAL impratio;                                                                                          \
  /* ---- geoms */                                                                                        \
  GEOM g_box[2]; /* root: coreBody1, coreBody2 */                                                         \
  GEOM g_rcyl;   /* root: screw1 */                                                                       \
  GEOM g_rell;   /* root: screw2 */                                                                       \
  GEOM g_ucyl[JB_NLEG], g_knee[JB_NLEG], g_lcyl[JB_NLEG], g_foot[JB_NLEG];                                \
  GEOM g_mcyl, g_mell; /* rotor: threadMass, mass */                                                      \
  /* ---- solver statistics scale; reset pose of the root body; home position of the target body */     \
  REAL meaninertia;                                                                                       \
  REAL root_pos0[3];                                                                                      \
  REAL target_pos0[3];                                                                                    \
  REAL pad_;

typedef struct { JB_GEOM_FIELDS(float) } JbGeom;
typedef struct { JB_MODEL_FIELDS(float, JbGeom) } JbModel;

/* task ids (reference task names, jitterbug_dmc/jitterbug.py:72-174) */
enum { JB_TASK_MOVE_FROM_ORIGIN = 0, JB_TASK_FACE_DIRECTION = 1, JB_TASK_MOVE_IN_DIRECTION = 2,
       JB_TASK_MOVE_TO_POSITION = 3, JB_TASK_MOVE_TO_POSE = 4 };

#endif
