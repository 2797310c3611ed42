// Shared declarations of the simulator's translation units: device-side state layout and the launcher of the
// fused step / reset / contact-report kernel (instantiated once per model mode in its own .cu so the two heavy
// instantiations compile in parallel).
#ifndef JB_SIM_COMMON_CUH
#define JB_SIM_COMMON_CUH
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/jbsim.h"
#include "jb_task.cuh"

#define JB_BLOCK 128

// HBM layout: struct-of-arrays, env index fastest, so a warp's 32 envs load/store each state component as one
// 128-byte line:  qpos[16][N] (motor angle wrapped, revolutions in rev[N]), qvel[15][N], warm[15][N],
// target[4][N], task[N], step_count[N], episode[N], reset_next[N].
struct JbDev {
  float *qpos, *qvel, *warm, *target;
  int *rev, *task, *step_count, *episode;
  unsigned char* reset_next;
  unsigned long long* counters;     // 5
  const JbModel* models;            // per-env models or nullptr
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
  int* ncon; int* geom; float* dist; int max_con;   // CONTACTS
};

cudaError_t jb_launch_shared(const JbModel& shared, const JbDev& d, const JbLaunchArgs& a, cudaStream_t st);
cudaError_t jb_launch_perenv(const JbModel& shared, const JbDev& d, const JbLaunchArgs& a, cudaStream_t st);
#endif
