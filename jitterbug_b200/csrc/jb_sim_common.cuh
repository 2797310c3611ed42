// Shared declarations of the simulator's translation units: device-side state layout and the launcher of the fused
// step / reset / contact-report kernel.
#ifndef JB_SIM_COMMON_CUH
#define JB_SIM_COMMON_CUH
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/jbsim.h"
#include "jb_packed.h"

#ifndef JB_BLOCK
#define JB_BLOCK 384                 /* threads per CTA; 4 threads (one quad) per env.  One 12-warp CTA per SM (168
                                        registers, a few hundred bytes of spills) measured marginally ahead of 256
                                        threads x 255 registers; see profiles/r01_phase_timers.md */
#endif
#define JB_ENVS_PER_BLOCK (JB_BLOCK / 4)
/* Per-env models in shared memory are JBP_ENV_STRIDE floats apart: 460 = 12 (mod 32), so the 8 quads of a warp, each
 * reading 4 consecutive words (lane-interleaved constants) or one broadcast word of its own model, land on 32
 * distinct banks; the natural stride 456 = 8 (mod 32) would make every model load a 2-way bank conflict. */
#ifndef JBP_ENV_STRIDE
#define JBP_ENV_STRIDE (JBP_SIZE + 4)
#endif

// HBM layout: struct-of-arrays, env index fastest: qpos[16][N] (motor angle wrapped, revolutions in rev[N]),
// qvel[15][N], warm[15][N], target[4][N], task[N], step_count[N], episode[N], reset_next[N].  A warp holds 8 envs, so
// each replicated (root / motor) component is one 32-byte sector per warp and each leg component four of them.
struct JbDev {
  float *qpos, *qvel, *warm, *target;
  int *rev, *task, *step_count, *episode;
  unsigned char* reset_next;        // 0: stepping, 1: the next STEP resets the env, 2: reset earlier in the running control step
  unsigned long long* counters;     // 16: [0-4] see jb_sim_get_counters, [5-7] self-clearance monitor, [8-15] phase timers
  const int* perm;                  // slot -> env: quad `slot` of the grid simulates env perm[slot] (NULL = identity)
  int4* load;                       // per env, written by every STEP launch: how heavy the env's control step was
  const float* packed;              // [JBP_SIZE] (shared model) or [N][JBP_SIZE] (per-env models)
  int per_env;
  int N, n_sub, step_limit, random_pose, obs_stride;
  unsigned long long seed;
  unsigned int env_id_offset;
};

enum { JB_MODE_STEP = 0, JB_MODE_RESET = 1, JB_MODE_CONTACTS = 2 };

struct JbLaunchArgs {
  int mode;
  const unsigned char* mask;        // RESET: optional per-env mask
  const float* actions;             // STEP
  float* obs; float* reward; int* step_type;
  int* rep_geom; float* rep_dist;   // CONTACTS: staging [N][4*JBQ_CAP]
  // STEP: a control step is run as one or more launches ("segments") of sub_count sub-steps each, so that the slot order
  // can be re-sorted in between (jb_sort.cu).  `first`: this launch starts the control step (pending episode resets happen
  // here); `last`: it ends it (step counting, divergence check, reward, observation).  RESET launches set both.
  int sub_count, first, last;
};

// two instantiations of jb_sim_kernel.cu: shared model / per-env models (see the head of that file)
cudaError_t jb_launch_sim(const JbDev& d, const JbLaunchArgs& a, cudaStream_t st);
cudaError_t jb_configure_kernels(void);
cudaError_t jb_launch_sim_perenv(const JbDev& d, const JbLaunchArgs& a, cudaStream_t st);
cudaError_t jb_configure_kernels_perenv(void);   // per device: opt in to the dynamic shared memory the per-env mode needs
// jb_monitor.cu: conservative self-clearance of every env at its current state (counters 5-7; optional per-env output)
#define JB_MONITOR_PERIOD 8          /* jb_sim_step runs the monitor every this many control steps */
cudaError_t jb_launch_self_clearance(const JbDev& d, float* out_dev, cudaStream_t st);
// jb_sort.cu: stable counting sort of the env ids by a small key of their last load record (the step kernel's slot order)
#define JB_SORT_BUCKETS 16
#define JB_SORT_MIN_ENVS (4 * 148 * JB_ENVS_PER_BLOCK)   /* four waves of full CTAs on a 148-SM part */
#define JB_SEGMENT_SUBSTEPS 10
int jb_sort_scratch_ints(int N);
cudaError_t jb_sort_by_load(const int4* load, int N, int n_sub, int* perm_out, int* scratch, cudaStream_t st);
#endif
