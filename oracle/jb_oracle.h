/* TEST INFRASTRUCTURE ONLY -- fp64 CPU restatement ("oracle") of the MuJoCo step pipeline the
 * Jitterbug reference runs through dm_control (jitterbug_dmc/jitterbug.py:81-90 builds
 * control.Environment(physics, task, control_timestep=0.01); each env.step() runs
 * 50 x physics.step() == mj_step2; mj_step1 of MuJoCo 2.0 on jitterbug_dmc/jitterbug.xml).
 *
 * PINNING STATUS.  Pinned to outputs of the reference's OWN code (tests/golden/reference_*.npz, produced by
 * tests/golden/make_reference_golden.py executing /root/reference's jitterbug.py / heuristic_policies.py /
 * augmented_jitterbug.py; checked by tests/test_reference_golden.py): the task layer of this file
 * (jbo_task_obs_reward: Physics helpers, observation, rewards) and the heuristic policies (jbo_heuristic).
 * PARITY UNPINNED for the step pipeline itself (mj_step2; mj_step1): MuJoCo / dm_control are third-party,
 * un-vendored dependencies of the reference (setup.py:10-14 pins dm_control @ 09b27a0 => MuJoCo 2.0) and are
 * not installed here, and the reference holds no tests or golden vectors for it.  That
 * part restates MuJoCo's published algorithm (Computation chapter + the open-sourced engine_*.c structure)
 * from memory and is anchored by analytic known-answer tests and an independent numpy finite-difference
 * dynamics check (tests/test_oracle_*.py); tests/test_mujoco_parity.py diffs it against a real MuJoCo
 * wherever one is importable (it skips in this project's containers).  The only MuJoCo-EXECUTED numbers
 * the reference holds -- the left edge of its published PPO2 training curves (fig-rl-perf.ipynb) -- anchor it
 * at the level of return distributions under a random policy (tests/test_published_curves.py), no finer.
 *
 * Nothing in the product path (jitterbug_b200/) links or calls this file.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 */
#ifndef JB_ORACLE_H
#define JB_ORACLE_H

#define JBO_MAXB 16      /* bodies */
#define JBO_MAXJ 16      /* joints */
#define JBO_MAXV 16      /* dofs */
#define JBO_MAXQ 20      /* qpos */
#define JBO_MAXG 32      /* geoms */
#define JBO_MAXU 2       /* actuators */
#define JBO_MAXCON 64    /* contacts */
#define JBO_MAXEFC (4 * JBO_MAXCON)

enum { JBO_PLANE = 0, JBO_SPHERE = 2, JBO_ELLIPSOID = 4, JBO_CYLINDER = 5, JBO_BOX = 6 };
enum { JBO_FREE = 0, JBO_HINGE = 3 };

typedef struct {
  int nq, nv, nu, nbody, njnt, ngeom;
  int iterations, sensor_body;
  double timestep, gravity[3], tolerance, impratio, meaninertia;
  int body_parentid[JBO_MAXB], body_rootid[JBO_MAXB], body_jntnum[JBO_MAXB], body_jntadr[JBO_MAXB];
  int body_dofnum[JBO_MAXB], body_dofadr[JBO_MAXB];
  double body_pos[JBO_MAXB][3], body_quat[JBO_MAXB][4], body_ipos[JBO_MAXB][3], body_iquat[JBO_MAXB][4];
  double body_mass[JBO_MAXB], body_inertia[JBO_MAXB][3], body_invweight0[JBO_MAXB][2];
  int jnt_type[JBO_MAXJ], jnt_qposadr[JBO_MAXJ], jnt_dofadr[JBO_MAXJ], jnt_bodyid[JBO_MAXJ];
  double jnt_pos[JBO_MAXJ][3], jnt_axis[JBO_MAXJ][3], jnt_stiffness[JBO_MAXJ];
  double qpos0[JBO_MAXQ], qpos_spring[JBO_MAXQ];
  int dof_bodyid[JBO_MAXV], dof_jntid[JBO_MAXV], dof_parentid[JBO_MAXV];
  double dof_damping[JBO_MAXV];
  int geom_type[JBO_MAXG], geom_bodyid[JBO_MAXG], geom_contype[JBO_MAXG], geom_conaffinity[JBO_MAXG];
  double geom_size[JBO_MAXG][3], geom_pos[JBO_MAXG][3], geom_quat[JBO_MAXG][4];
  double geom_friction[JBO_MAXG][3], geom_solref[JBO_MAXG][2], geom_solimp[JBO_MAXG][5];
  double geom_margin[JBO_MAXG], geom_gap[JBO_MAXG], geom_solmix[JBO_MAXG];
  int actuator_trnid[JBO_MAXU], actuator_ctrllimited[JBO_MAXU];
  double actuator_gear[JBO_MAXU], actuator_gainprm[JBO_MAXU][3], actuator_biasprm[JBO_MAXU][3];
  double actuator_ctrlrange[JBO_MAXU][2];
  /* task-level model mutation (reference: jitterbug.py:613-651 writes model.body_pos/body_quat["target"]) */
  int target_body, target_geom;
} JboModel;

typedef struct {
  /* state */
  double qpos[JBO_MAXQ], qvel[JBO_MAXV], ctrl[JBO_MAXU], qacc_warmstart[JBO_MAXV], time;
  /* position stage */
  double xpos[JBO_MAXB][3], xquat[JBO_MAXB][4], xmat[JBO_MAXB][9], xipos[JBO_MAXB][3], ximat[JBO_MAXB][9];
  double xanchor[JBO_MAXJ][3], xaxis[JBO_MAXJ][3];
  double geom_xpos[JBO_MAXG][3], geom_xmat[JBO_MAXG][9];
  double subtree_com[JBO_MAXB][3], cinert[JBO_MAXB][10], crb[JBO_MAXB][10], cdof[JBO_MAXV][6];
  double qM[JBO_MAXV][JBO_MAXV];           /* dense joint-space inertia */
  /* contacts */
  int ncon, con_geom1[JBO_MAXCON], con_geom2[JBO_MAXCON], con_exclude[JBO_MAXCON];
  double con_dist[JBO_MAXCON], con_pos[JBO_MAXCON][3], con_frame[JBO_MAXCON][9];
  double con_friction[JBO_MAXCON][5], con_solref[JBO_MAXCON][2], con_solimp[JBO_MAXCON][5], con_mu[JBO_MAXCON];
  /* constraints */
  int nefc, efc_id[JBO_MAXEFC];
  double efc_J[JBO_MAXEFC][JBO_MAXV], efc_pos[JBO_MAXEFC], efc_margin[JBO_MAXEFC], efc_diagApprox[JBO_MAXEFC];
  double efc_R[JBO_MAXEFC], efc_D[JBO_MAXEFC], efc_KBIP[JBO_MAXEFC][4], efc_vel[JBO_MAXEFC], efc_aref[JBO_MAXEFC];
  double efc_force[JBO_MAXEFC];
  /* velocity stage */
  double cvel[JBO_MAXB][6], cdof_dot[JBO_MAXV][6];
  double actuator_length[JBO_MAXU], actuator_velocity[JBO_MAXU], actuator_force[JBO_MAXU];
  double qfrc_passive[JBO_MAXV], qfrc_bias[JBO_MAXV], qfrc_actuator[JBO_MAXV], qfrc_smooth[JBO_MAXV];
  double qacc_smooth[JBO_MAXV], qfrc_constraint[JBO_MAXV], qacc[JBO_MAXV];
  double sensordata[3];
  /* diagnostics */
  int solver_iter, warning_bad;            /* warning_bad: BADQPOS=1 BADQVEL=2 BADQACC=4 */
  long total_solver_iter, total_substeps, total_contacts;
} JboData;

#ifdef __cplusplus
extern "C" {
#endif
void jbo_reset_data(const JboModel* m, JboData* d);      /* mj_resetData: qpos=qpos0, everything else 0 */
void jbo_forward(const JboModel* m, JboData* d);         /* mj_forward (position, velocity, acceleration stages) */
void jbo_step1(const JboModel* m, JboData* d);           /* mj_step1 */
void jbo_step2(const JboModel* m, JboData* d);           /* mj_step2 (Euler) */
void jbo_physics_step(const JboModel* m, JboData* d);    /* dm_control Physics.step(): mj_step2; mj_step1 */
void jbo_control_step(const JboModel* m, JboData* d, const double* ctrl, int n_sub_steps);
int jbo_sizeof_model(void);
int jbo_sizeof_data(void);
/* task layer: obs/reward restated in C for the multi-threaded CPU baseline (jitterbug.py:673-763, :840-925) */
int jbo_task_obs_reward(const JboModel* m, const JboData* d, int task, double* obs, double* reward);
/* run `nsteps` control steps for `nenv` independent envs on `nthreads` threads; actions[nsteps][nenv] */
double jbo_heuristic(int task, const double* obs);     /* heuristic_policies.py:6-136 on a normalised observation */
void jbo_rollout_heuristic(const JboModel* models, JboData* datas, int nenv, int task, int nsteps, int n_sub_steps,
                           int nthreads, double* rewards_out);   /* closed loop under that policy (evaluate_policy.py:10-35) */
void jbo_rollout(const JboModel* models, JboData* datas, int nenv, int task, const double* actions, int nsteps,
                 int n_sub_steps, int nthreads, double* rewards_out, double* obs_out);
#ifdef __cplusplus
}
#endif
#endif
