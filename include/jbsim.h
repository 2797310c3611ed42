/* libjbsim -- C-ABI of the B200-native batched Jitterbug simulator.
 *
 * Drop-in boundary.  The reference reaches its physics through dm_control's ctypes binding of MuJoCo
 * (jitterbug_dmc/jitterbug.py:46 `mjlib`, :81 `Physics.from_xml_string`, :84-90 `control.Environment`); one
 * `env.step(a)` is `task.before_step` -> 50 x (`mj_step2`; `mj_step1`) -> `task.get_reward` -> `task.get_observation`.
 * Each entry point below names the reference interface it replaces.  All functions return 0 on success, non-zero on
 * error (message via jb_last_error()).  Pointers with the `_dev` suffix are device pointers on the simulator's
 * device that the caller owns (e.g. torch tensor storage); `_host` pointers are host memory.  Work is enqueued on the
 * `stream` argument (a cudaStream_t passed as void*; NULL = default stream); there are no hidden synchronisations
 * except in the `_host` convenience calls, which synchronise the stream before returning.  A JbSim is bound to one
 * device and is not thread-safe.  There is no CPU fallback: without a CUDA device jb_sim_create fails.
 */
#ifndef JBSIM_H
#define JBSIM_H

#include <stddef.h>
#include <stdint.h>
#include "../jitterbug_b200/csrc/jb_model.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct JbSim JbSim;

/* dm_env.StepType values written to step_type arrays */
enum { JB_STEP_FIRST = 0, JB_STEP_MID = 1, JB_STEP_LAST = 2 };

typedef struct {
  int num_envs;            /* environments simulated by this handle (this rank's shard) */
  int device;              /* CUDA device ordinal */
  int n_sub_steps;         /* physics sub-steps per control step: round(control_timestep / timestep) = 50 */
  int step_limit;          /* control steps per episode: time_limit / control_timestep = 1000; <= 0: no limit */
  int random_pose;         /* Jitterbug(random_pose=True): random initial tilt (jitterbug.py:653-664) */
  int obs_stride;          /* floats per env in observation buffers (>= widest task in the batch, max 19) */
  uint64_t seed;           /* Philox key */
  uint32_t env_id_offset;  /* global id of env 0 of this shard: Philox streams are keyed by global env id */
  uint32_t pad;
} JbSimConfig;

/* Replaces Physics.from_xml_string + control.Environment construction (jitterbug.py:81-90).
 * `models`: n_models == 1 (shared model, staged in shared memory once per CTA) or n_models == num_envs (per-env augmented models,
 * augmented_jitterbug.py:95-267).  `task_ids_host`: num_envs task ids (JB_TASK_*), or NULL for all move_from_origin. */
int jb_sim_create(JbSim** out, const JbSimConfig* cfg, const JbModel* models_host, int n_models,
                  const int32_t* task_ids_host);
/* Same with a population of n_models distinct models and a per-env index into it (model_index_host[num_envs]); the
 * per-env array is expanded on the device, so the host never holds num_envs models (config 4: 1 Mi augmented envs). */
int jb_sim_create_indexed(JbSim** out, const JbSimConfig* cfg, const JbModel* models_host, int n_models,
                          const int32_t* model_index_host, const int32_t* task_ids_host);
/* Same from packed models already resident on the device (jb_augment_compile's output): packed_dev holds n_models *
 * jb_packed_model_size() floats, n_models == 1 or num_envs.  The array is copied; the caller keeps its buffer. */
int jb_sim_create_packed(JbSim** out, const JbSimConfig* cfg, const float* packed_dev, int n_models,
                         const int32_t* task_ids_host);
void jb_sim_destroy(JbSim* sim);

/* Replaces control.Environment.reset() -> task.initialize_episode (jitterbug.py:601-666) + get_observation.
 * `mask_dev`: num_envs bytes, non-zero = reset that env; NULL = reset all.  Writes obs for the reset envs. */
int jb_sim_reset(JbSim* sim, const uint8_t* mask_dev, float* obs_dev, void* stream);

/* Replaces control.Environment.step(action) for every env in lock-step: set ctrl, n_sub_steps physics steps,
 * reward, observation, step counting.  An env whose previous step returned LAST (or that was never reset) is reset
 * instead, its action ignored, FIRST returned with reward 0 (dm_control returns reward None there).
 * actions_dev[num_envs]; obs_dev[num_envs * obs_stride]; reward_dev[num_envs]; step_type_dev[num_envs]. */
int jb_sim_step(JbSim* sim, const float* actions_dev, float* obs_dev, float* reward_dev, int32_t* step_type_dev,
                void* stream);

/* Same through host buffers (the reference's numpy in / numpy out call shape): H2D of actions, step, D2H of
 * obs/reward/step_type, then stream synchronise.  Page-locked caller buffers (cudaHostAlloc / cudaHostRegister) are the
 * DMA end points themselves; pageable ones go through the simulator's pinned staging. */
int jb_sim_step_host(JbSim* sim, const float* actions_host, float* obs_host, float* reward_host,
                     int32_t* step_type_host, void* stream);

/* task.get_observation(physics) and task.get_reward(physics) (jitterbug.py:673-763, :891-925) evaluated at the CURRENT
 * state of every env, without stepping.  obs_dev[num_envs * obs_stride] and / or reward_dev[num_envs] (either may be
 * NULL). */
int jb_sim_observe(JbSim* sim, float* obs_dev, float* reward_dev, void* stream);

/* heuristic_policies.* (heuristic_policies.py:6-136) evaluated on device from an observation buffer. */
int jb_sim_heuristic_actions(JbSim* sim, const float* obs_dev, float* actions_dev, void* stream);

/* physics.data.qpos / qvel / qacc_warmstart access (`physics.named.data.qpos[...]`, jitterbug.py:182, :212, :661).
 * Row-major [num_envs][16|15|15] device buffers, MuJoCo ordering.  NULL pointers are skipped. */
int jb_sim_get_state(JbSim* sim, float* qpos_dev, float* qvel_dev, float* warm_dev, void* stream);
int jb_sim_set_state(JbSim* sim, const float* qpos_dev, const float* qvel_dev, const float* warm_dev, void* stream);

/* Exact checkpoint / resume of the episode state (the reference checkpoints the agent only,
 * benchmark.py:234-237; an env has no state file -- a batched simulator needs one).  The blob holds everything
 * jb_sim_step reads and writes per env: positions with the motor angle in its wrapped + revolution-counter form,
 * velocities, warm start, target pose, step count, episode index (which keys the Philox stream of the next reset) and the
 * pending-reset flag; a simulator created with the same config, models and tasks and restored from it continues bit
 * for bit, through episode boundaries.  jb_sim_get_state / jb_sim_set_state are the MuJoCo-layout interchange (fp32
 * motor angle, no RNG state) and are not exact in that sense.  `blob_dev`: jb_sim_snapshot_size(sim) bytes on the device. */
size_t jb_sim_snapshot_size(const JbSim* sim);
int jb_sim_snapshot(JbSim* sim, void* blob_dev, void* stream);
int jb_sim_restore(JbSim* sim, const void* blob_dev, void* stream);
/* target pose per env: [num_envs][4] = x, y, quat_w, quat_z (model.body_pos / body_quat["target"], jitterbug.py:613-648) */
int jb_sim_get_target(JbSim* sim, float* target_dev, void* stream);
int jb_sim_set_target(JbSim* sim, const float* target_dev, void* stream);
/* episode bookkeeping per env: step counts; setting also clears the "reset on next step" flag */
int jb_sim_get_step_count(JbSim* sim, int32_t* count_dev, void* stream);
int jb_sim_set_step_count(JbSim* sim, const int32_t* count_dev, void* stream);

/* Parity hook (d->contact after mj_step1 at the current state): per env the number of active contacts, and for up
 * to `max_con` of them the MuJoCo geom id of the non-floor geom and the signed distance, in ascending geom id.
 * ncon_dev[num_envs]; geom_dev, dist_dev: [num_envs][max_con]. */
int jb_sim_get_contacts(JbSim* sim, int32_t* ncon_dev, int32_t* geom_dev, float* dist_dev, int max_con, void* stream);

/* Slot order of the step kernel.  The kernel hands the envs to its quads in a load-sorted order (heavy envs -- many
 * contacts in one leg's lane, extra Newton trips -- share warps and CTAs instead of stalling light ones); the order is
 * recomputed on the device after every jb_sim_step from that step's load records.  It is a scheduling decision only:
 * per-env results are bit-identical for any order.  mode: JB_ORDER_AUTO (default), JB_ORDER_IDENTITY (slot i = env i),
 * JB_ORDER_EXPLICIT (perm_dev[num_envs], a permutation of the env ids; copied). */
enum { JB_ORDER_AUTO = 0, JB_ORDER_IDENTITY = 1, JB_ORDER_EXPLICIT = 2 };
int jb_sim_set_slot_order(JbSim* sim, int mode, const int32_t* perm_dev, void* stream);
int jb_sim_get_slot_order(JbSim* sim, int32_t* perm_dev, void* stream);
/* load record of the last jb_sim_step per env, load_dev[num_envs][4]: contacts summed over the sub-steps in the fullest
 * lane of the env's quad, sub-steps with >= 3 / >= 2 contacts in one lane, Newton trips. */
int jb_sim_get_load(JbSim* sim, int32_t* load_dev, void* stream);

/* counters since creation: [0] physics sub-steps, [1] Newton iterations, [2] active contacts summed over sub-steps,
 * [3] contact-slot overflows (always 0: a lane has a slot for every contact its geoms can produce), [4] diverged-state
 * resets.  Synchronises the device. */
int jb_sim_get_counters(JbSim* sim, unsigned long long* out5_host);

/* Self-clearance monitor.  MuJoCo would also test the geom pairs of non-adjacent bodies (leg - leg, leg - rotor, lower leg -
 * core; reference model jitterbug.xml); the step kernel generates only floor contacts, on the
 * argument that those pairs never touch.  The monitor measures it: a conservative lower bound (cylinders as capsules,
 * boxes / ellipsoids as bounding spheres) of the smallest distance over all such pairs, per env, at the current state.
 * jb_sim_step runs it every 8th control step into running counters.  clearance_dev[num_envs] (optional): evaluate now
 * and write the per-env bound in metres.  out3_host (optional; synchronises): env-states checked so far, how many of
 * them had a bound below zero, and the smallest bound seen, in metres. */
int jb_sim_self_clearance(JbSim* sim, float* clearance_dev, double* out3_host, void* stream);

/* Profiling builds only (-DJBQ_PHASE_TIMERS): clock64 cycles summed over warps for the 7 phases of the sub-step
 * ([1] leg kinematics+collision, [2] root/rotor, [3] extra geoms, [4] factor M + smooth solve, [5] warm-start
 * selection, [6] Newton loop, [7] Euler); zeros otherwise.  Synchronises the device. */
int jb_sim_get_phase_cycles(JbSim* sim, unsigned long long* out8_host);

/* FP32 FMA-pipe peak of `device` in TFLOP/s (FFMA-chain microbenchmark): the roofline denominator of this
 * compute-bound path.  Synchronises the device. */
int jb_measure_fp32_peak(int device, double* tflops_out);

/* ---- Batched on-device compilation of augmented Jitterbugs.  Replaces
 * augment_Jitterbug + MuJoCo's XML compile, once per env (jitterbug_dmc/augmented_jitterbug.py:95-267): a kernel
 * applies the perturbation to leg end points / rotor offset / densities / gear and recomputes geom and body inertias,
 * hinge axes and mj_setConst's invweights in fp64, writing the packed float layout the step kernel consumes. */
typedef struct { JB_GEOM_FIELDS(double) } JbGeomD;
typedef struct { JB_MODEL_FIELDS(double, JbGeomD) } JbModelD;
#define JB_AUG_NGEOM 22   /* slots: coreBody1, coreBody2, screw1, screw2, 4 x (upper cyl, knee, lower cyl, foot), threadMass, mass */
#define JB_AUG_NDELTA 31  /* [0] default density, [1] coreBody1 density, [2] coreBody2 density, [3 + 6k ..] leg k: upper
                             end offset (3), lower end offset (3), [27..29] rotor offset, [30] gear */
typedef struct {
  JbModelD model;                        /* the unperturbed model (root frame, fp64) */
  double density[JB_AUG_NGEOM];
  int follows_default[JB_AUG_NGEOM];     /* 1: the geom has no density of its own (moves with the default density) */
} JbAugBase;
typedef struct {                         /* augment_Jitterbug's keyword arguments (augmented_jitterbug.py:95-106) */
  int modify_legs, modify_mass, modify_core1, modify_core2, modify_global_density, modify_gear;
  double sd_legs[3], sd_mass_pos[3], sd_core1_density, sd_core2_density, sd_global_density, sd_gear;
} JbAugSpec;
/* floats per packed model */
int jb_packed_model_size(void);
/* deltas_dev[n][JB_AUG_NDELTA] (fp64): one Gaussian per parameter like the reference draws them (rotor y / z offsets
 * clipped to [-0.001, 1], :218-222), Philox streams keyed by (seed, env_id_offset + env). */
int jb_augment_sample(const JbAugSpec* spec_host, uint64_t seed, uint32_t env_id_offset, int n, double* deltas_dev,
                      void* stream);
/* deltas_dev[n][JB_AUG_NDELTA] -> packed_dev[n][jb_packed_model_size()]; models_dev: optional n JbModel structs (the same
 * models in the C-ABI struct layout, for inspection) or NULL.  Synchronises the stream. */
int jb_augment_compile(const JbAugBase* base_host, const double* deltas_dev, int n, float* packed_dev,
                       JbModel* models_dev, void* stream);

/* ---- Observation-encoder hook.  Replaces Jitterbug.encode_obs (jitterbug_dmc/jitterbug.py:927-993):
 * the first `k_encode` observation features go through a small dense network, the remaining obs_dim - k_encode are
 * appended unchanged (the reference's use_autoencoder15 / use_autoencoder13 modes, :957-981); k_encode == obs_dim
 * encodes everything (:949-951).  Network = the hidden layer of the reference's tanh (denoising) auto-encoder
 * (benchmarks/denoising_autoencoder.py:66-98, encode() :195-205) or its VAE encoder (benchmarks/VAE.py:13-46: ReLU
 * layers, then mean and std heads; encode() :134-146 returns mean + std * eps with the std head used as is).
 * `normalize` applies the reference's per-feature extrema scaling first (denoising_autoencoder.py:218-232).
 * n_models > 1 averages an ensemble of identically shaped encoders (jitterbug.py:983-991). */
enum { JB_ACT_LINEAR = 0, JB_ACT_TANH = 1, JB_ACT_RELU = 2 };
enum { JB_ENC_NORM_NONE = 0, JB_ENC_NORM_PM1 = 1, JB_ENC_NORM_01 = 2 };
#define JB_ENC_MAX_LAYERS 4
#define JB_ENC_MAX_WIDTH 64

typedef struct {
  int obs_dim;                        /* features read per env */
  int k_encode;                       /* leading features fed to the network */
  int n_layers;                       /* dense layers, 1..JB_ENC_MAX_LAYERS */
  int dims[JB_ENC_MAX_LAYERS + 1];    /* dims[0] == k_encode, dims[n_layers] = latent size */
  int act[JB_ENC_MAX_LAYERS];         /* JB_ACT_* per layer (ignored on a VAE's head layer) */
  int vae;                            /* 1: the last layer is the (mean | std) head pair, weights [dims[n-1]][2 * latent] */
  int normalize;                      /* JB_ENC_NORM_* */
  int n_models;                       /* ensemble size, >= 1 */
  uint32_t env_id_offset;             /* global id of env 0 (keys the VAE noise stream, like JbSimConfig) */
} JbEncoderDesc;

/* floats per ensemble member in `params_dev`: for each layer W[in][out] row-major, then b[out]; -1 on a bad desc */
int jb_encoder_num_params(const JbEncoderDesc* desc);
/* floats written per env: latent + obs_dim - k_encode; -1 on a bad desc */
int jb_encoder_out_dim(const JbEncoderDesc* desc);
/* obs_dev[n][obs_stride] -> out_dev[n][out_stride].  params_dev: n_models * jb_encoder_num_params floats;
 * extrema_dev: [obs_dim][2] (min, max) or NULL when normalize == JB_ENC_NORM_NONE; (seed, counter) key the VAE noise. */
int jb_encode_obs(const JbEncoderDesc* desc, const float* params_dev, const float* extrema_dev, const float* obs_dev,
                  int n, int obs_stride, float* out_dev, int out_stride, uint64_t seed, uint64_t counter, void* stream);

int jb_sim_num_envs(const JbSim* sim);
/* kernels of this library launched so far on behalf of `sim` by jb_sim_step / _reset / _get_contacts (step kernel
 * segments and the three slot-order sort kernels after each), jb_sim_observe and jb_sim_heuristic_actions */
unsigned long long jb_sim_launch_count(const JbSim* sim);
int jb_sim_obs_stride(const JbSim* sim);
const char* jb_last_error(void);
const char* jb_version(void);

#ifdef __cplusplus
}
#endif
#endif
