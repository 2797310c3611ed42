// The fused kernel: one thread per environment runs a whole control step (n_sub_steps physics sub-steps + reward +
// observation + episode bookkeeping) without touching HBM in between.  Included by jb_kernel_shared.cu and
// jb_kernel_perenv.cu with JB_PER_ENV defined to 0 / 1.
#include "jb_sim_common.cuh"

// ------------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ void load_state(const JbDev& d, int e, JbState<float>& s) {
  const int N = d.N;
  s.pos[0] = d.qpos[0 * N + e]; s.pos[1] = d.qpos[1 * N + e]; s.pos[2] = d.qpos[2 * N + e];
#pragma unroll
  for (int i = 0; i < 4; i++) s.quat[i] = d.qpos[(3 + i) * N + e];
#pragma unroll
  for (int k = 0; k < 4; k++) { s.ql[k][0] = d.qpos[(7 + 2 * k) * N + e]; s.ql[k][1] = d.qpos[(8 + 2 * k) * N + e]; }
  s.qm = d.qpos[15 * N + e];
  s.rev = d.rev[e];
#pragma unroll
  for (int i = 0; i < 3; i++) { s.vlin[i] = d.qvel[i * N + e]; s.w[i] = d.qvel[(3 + i) * N + e]; }
#pragma unroll
  for (int k = 0; k < 4; k++) { s.vl[k][0] = d.qvel[(6 + 2 * k) * N + e]; s.vl[k][1] = d.qvel[(7 + 2 * k) * N + e]; }
  s.vm = d.qvel[14 * N + e];
#pragma unroll
  for (int i = 0; i < 15; i++) s.warm[i] = d.warm[i * N + e];
}

__device__ __forceinline__ void store_state(const JbDev& d, int e, const JbState<float>& s) {
  const int N = d.N;
  d.qpos[0 * N + e] = s.pos[0]; d.qpos[1 * N + e] = s.pos[1]; d.qpos[2 * N + e] = s.pos[2];
#pragma unroll
  for (int i = 0; i < 4; i++) d.qpos[(3 + i) * N + e] = s.quat[i];
#pragma unroll
  for (int k = 0; k < 4; k++) { d.qpos[(7 + 2 * k) * N + e] = s.ql[k][0]; d.qpos[(8 + 2 * k) * N + e] = s.ql[k][1]; }
  d.qpos[15 * N + e] = s.qm;
  d.rev[e] = s.rev;
#pragma unroll
  for (int i = 0; i < 3; i++) { d.qvel[i * N + e] = s.vlin[i]; d.qvel[(3 + i) * N + e] = s.w[i]; }
#pragma unroll
  for (int k = 0; k < 4; k++) { d.qvel[(6 + 2 * k) * N + e] = s.vl[k][0]; d.qvel[(7 + 2 * k) * N + e] = s.vl[k][1]; }
  d.qvel[14 * N + e] = s.vm;
#pragma unroll
  for (int i = 0; i < 15; i++) d.warm[i * N + e] = s.warm[i];
}

__device__ __forceinline__ void normalize_quat(JbState<float>& s) {
  float qn = rsqrtf(s.quat[0] * s.quat[0] + s.quat[1] * s.quat[1] + s.quat[2] * s.quat[2] + s.quat[3] * s.quat[3]);
  s.quat[0] *= qn; s.quat[1] *= qn; s.quat[2] *= qn; s.quat[3] *= qn;
}

__device__ __forceinline__ bool state_bad(const JbState<float>& s) {
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 3; i++) acc += fabsf(s.pos[i]) + fabsf(s.vlin[i]) + fabsf(s.w[i]);
#pragma unroll
  for (int i = 0; i < 4; i++) acc += fabsf(s.quat[i]);
#pragma unroll
  for (int k = 0; k < 4; k++) acc += fabsf(s.ql[k][0]) + fabsf(s.ql[k][1]) + fabsf(s.vl[k][0]) + fabsf(s.vl[k][1]);
  acc += fabsf(s.qm) + fabsf(s.vm);
  return !(acc < 1e10f);      // NaN or huge (mj_checkPos / mj_checkVel threshold)
}

template <typename M>
__device__ __forceinline__ void do_reset(const M& m, const JbDev& d, int e, JbState<float>& s, JbTarget<float>& tg, int task) {
  int ep = d.episode[e];
  jb_reset_env<float>(m, s, tg, task, d.random_pose, d.seed, d.env_id_offset + (unsigned)e, (unsigned)ep);
  d.episode[e] = ep + 1;
  d.step_count[e] = 0;
  d.reset_next[e] = 0;
  const int N = d.N;
  d.target[0 * N + e] = tg.x; d.target[1 * N + e] = tg.y; d.target[2 * N + e] = tg.qw; d.target[3 * N + e] = tg.qz;
}

template <typename M>
__device__ __forceinline__ void write_obs(const M& m, const JbDev& d, int e, const JbState<float>& s, int task,
                                          const JbTarget<float>& tg, float* obs_out, float* reward_out) {
  float obs[JB_OBS_MAX], rew;
  int n = jb_obs_reward<float>(m, s, task, tg, obs, &rew);
  if (obs_out) {
    float* o = obs_out + (size_t)e * d.obs_stride;
#pragma unroll
    for (int i = 0; i < JB_OBS_MAX; i++)
      if (i < d.obs_stride) o[i] = i < n ? obs[i] : 0.f;
  }
  if (reward_out) reward_out[e] = rew;
}


#if JB_PER_ENV
#define JB_SIM_KERNEL jb_sim_kernel_perenv
#else
#define JB_SIM_KERNEL jb_sim_kernel_shared
#endif

__global__ void __launch_bounds__(JB_BLOCK)
JB_SIM_KERNEL(const __grid_constant__ JbModel shared, const __grid_constant__ JbDev d, const __grid_constant__ JbLaunchArgs a) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.N) return;
#if JB_PER_ENV
  const JbModel& m = d.models[e];
#else
  const JbModel& m = shared;
#endif
  const int N = d.N;
  int task = d.task[e];
  JbState<float> s;
  JbTarget<float> tg;
  ContactReport<float> rep;
  const bool reset_now = a.mode == JB_MODE_RESET ? (!a.mask || a.mask[e]) : (a.mode == JB_MODE_STEP && d.reset_next[e]);
  if (a.mode == JB_MODE_RESET && !reset_now) return;
  int iters = 0, ncon_sum = 0, overflow = 0, nsub = 0;
  float ctrl = 0.f;
  if (reset_now) {
    do_reset(m, d, e, s, tg, task);
  } else {
    load_state(d, e, s);
    tg.x = d.target[0 * N + e]; tg.y = d.target[1 * N + e]; tg.qw = d.target[2 * N + e]; tg.qz = d.target[3 * N + e];
    if (a.mode == JB_MODE_STEP) { ctrl = a.actions[e]; nsub = d.n_sub; }
  }
  // sub-steps; a reset / contact report runs one forward-only pass (mj_forward) for the warm start / contact list
  const bool fwd_only = nsub == 0;
  const int trips = fwd_only ? 1 : nsub;
  for (int k = 0; k < trips; k++)
    jb_substep<float, JbModel, false>(m, s, ctrl, nullptr, &iters, fwd_only, &ncon_sum, &overflow,
                                      a.mode == JB_MODE_CONTACTS ? &rep : nullptr);
  if (a.mode == JB_MODE_CONTACTS) {
    int n = rep.n;
    for (int i = 1; i < n; i++) {        // ascending geom id (MuJoCo order)
      int g = rep.geom[i]; float di = rep.dist[i]; int j = i - 1;
      while (j >= 0 && rep.geom[j] > g) { rep.geom[j + 1] = rep.geom[j]; rep.dist[j + 1] = rep.dist[j]; j--; }
      rep.geom[j + 1] = g; rep.dist[j + 1] = di;
    }
    a.ncon[e] = n;
    for (int i = 0; i < a.max_con; i++) {
      a.geom[(size_t)e * a.max_con + i] = i < n ? rep.geom[i] : -1;
      a.dist[(size_t)e * a.max_con + i] = i < n ? rep.dist[i] : 0.f;
    }
    return;
  }
  if (reset_now) {
    store_state(d, e, s);
    write_obs(m, d, e, s, task, tg, a.obs, nullptr);
    if (a.reward) a.reward[e] = 0.f;
    if (a.step_type) a.step_type[e] = JB_STEP_FIRST;
    return;
  }
  normalize_quat(s);                      // mj_step1 of the last sub-step normalises qpos quaternions in place
  unsigned long long bad = 0;
  if (state_bad(s)) {                     // mj_checkPos/Vel: MuJoCo resets the data; dm_control raises PhysicsError
    bad = 1;
    do_reset(m, d, e, s, tg, task);
  }
  int sc = d.step_count[e] + 1;
  d.step_count[e] = sc;
  int st = JB_STEP_MID;
  if (d.step_limit > 0 && sc >= d.step_limit) { st = JB_STEP_LAST; d.reset_next[e] = 1; }
  store_state(d, e, s);
  write_obs(m, d, e, s, task, tg, a.obs, a.reward);
  if (a.step_type) a.step_type[e] = st;
  // counters: warp-aggregated
  unsigned long long v1 = iters, v2 = ncon_sum, v3 = overflow;
  unsigned act = __activemask();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v1 += __shfl_down_sync(act, v1, o); v2 += __shfl_down_sync(act, v2, o);
    v3 += __shfl_down_sync(act, v3, o); bad += __shfl_down_sync(act, bad, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&d.counters[1], v1); atomicAdd(&d.counters[2], v2);
    if (v3) atomicAdd(&d.counters[3], v3);
    if (bad) atomicAdd(&d.counters[4], bad);
  }
  if (e == 0) atomicAdd(&d.counters[0], (unsigned long long)d.n_sub * (unsigned long long)N);
}

#if JB_PER_ENV
cudaError_t jb_launch_perenv(const JbModel& shared, const JbDev& d, const JbLaunchArgs& a, cudaStream_t st) {
#else
cudaError_t jb_launch_shared(const JbModel& shared, const JbDev& d, const JbLaunchArgs& a, cudaStream_t st) {
#endif
  JB_SIM_KERNEL<<<(d.N + JB_BLOCK - 1) / JB_BLOCK, JB_BLOCK, 0, st>>>(shared, d, a);
  return cudaGetLastError();
}
