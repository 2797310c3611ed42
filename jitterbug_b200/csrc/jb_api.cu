// C-ABI of the batched Jitterbug simulator + the light-weight utility kernels.  See include/jbsim.h for the contract.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>

#include "jb_sim_common.cuh"
#include "jb_task.cuh"
#include <vector>

static thread_local std::string g_err;
static int fail(const std::string& msg) { g_err = msg; return 1; }
int jb_set_error(const std::string& msg) { return fail(msg); }   // for the other translation units
#define CUDA_TRY(x)                                                                                   \
  do {                                                                                                \
    cudaError_t e_ = (x);                                                                             \
    if (e_ != cudaSuccess) return fail(std::string(#x) + ": " + cudaGetErrorString(e_));              \
  } while (0)

// A JbSim is bound to one device: every entry point makes that device current for the duration of the call and restores
// the caller's device afterwards (two simulators on two GPUs in one process, or calls from another host thread).
struct DeviceGuard {
  int prev = -1; bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
    if (prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() { if (switched && prev >= 0) cudaSetDevice(prev); }
};
#define JB_GUARD(s) DeviceGuard guard_((s)->cfg.device)
struct DevBuf {                               // scope-owned device allocation
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  void* release() { void* q = p; p = nullptr; return q; }
};

struct JbSim {
  JbSimConfig cfg;
  JbDev d;
  bool per_env;
  int* rep_geom; float* rep_dist;   // contact-report staging [N][4*JBQ_CAP], allocated on first use
  // pinned staging for the host-buffer entry point
  float *h_act, *h_obs, *h_rew; int32_t* h_st;
  float *s_act, *s_obs, *s_rew; int32_t* s_st;
  // load-sorted slot order (jb_sort.cu): perm buffers [2][N] (double-buffered), the scan scratch, and whether a STEP
  // launch re-sorts for the next one
  int* perm_buf; int* sort_scratch; int perm_cur; int auto_sort;
  int seg_sub;                      // sub-steps per launch of a control step when the slot order is being sorted
  unsigned long long launches;      // kernels of this library launched for this simulator (step path + utilities)
  unsigned long long steps;         // control steps taken (the self-clearance monitor runs every JB_MONITOR_PERIOD-th)
  int monitor;                      // self-clearance monitor on / off
};

static cudaError_t launch(JbSim* s, const JbLaunchArgs& a, cudaStream_t st) {
  static int force256 = -1;                  // tuning: JBSIM_KERNEL=256 runs the shared model through the 256-thread build
  if (force256 < 0) { const char* e = getenv("JBSIM_KERNEL"); force256 = (e && atoi(e) == 256) ? 1 : 0; }
  s->launches++;
  return (s->per_env || force256) ? jb_launch_sim_perenv(s->d, a, st) : jb_launch_sim(s->d, a, st);
}

#define JB_REP (4 * JBQ_CAP)   /* 4 lanes x JBQ_CAP staging slots per env (same macro as jb_quad.cuh) */

// staging [N][32] (-1 = empty) -> compact lists in ascending MuJoCo geom id
__global__ void jb_compact_contacts_kernel(int N, const int* rep_geom, const float* rep_dist, int* ncon, int* geom,
                                           float* dist, int max_con) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  int gl[JB_REP]; float dl[JB_REP];
  int n = 0;
  for (int i = 0; i < JB_REP; i++) {
    int gg = rep_geom[(size_t)e * JB_REP + i];
    if (gg < 0) continue;
    float dd = rep_dist[(size_t)e * JB_REP + i];
    int j = n - 1;
    while (j >= 0 && gl[j] > gg) { gl[j + 1] = gl[j]; dl[j + 1] = dl[j]; j--; }
    gl[j + 1] = gg; dl[j + 1] = dd;
    n++;
  }
  ncon[e] = n;
  for (int i = 0; i < max_con; i++) {
    geom[(size_t)e * max_con + i] = i < n ? gl[i] : -1;
    dist[(size_t)e * max_con + i] = i < n ? dl[i] : 0.f;
  }
}

__global__ void jb_heuristic_kernel(const __grid_constant__ JbDev d, const float* __restrict__ obs, float* __restrict__ act) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.N) return;
  float o[JB_OBS_MAX];
  const float* p = obs + (size_t)e * d.obs_stride;
#pragma unroll
  for (int i = 0; i < JB_OBS_MAX; i++) o[i] = i < d.obs_stride ? p[i] : 0.f;
  act[e] = jb_heuristic<float>(d.task[e], o);
}

// AoS <-> SoA state transfer (row-major [N][16|15|15] user buffers)
__global__ void jb_get_state_kernel(const __grid_constant__ JbDev d, float* qpos, float* qvel, float* warm) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.N) return;
  const int N = d.N;
  if (qpos) {
    for (int i = 0; i < 15; i++) qpos[(size_t)e * 16 + i] = d.qpos[i * N + e];
    qpos[(size_t)e * 16 + 15] = (float)((double)d.qpos[15 * N + e] + 6.28318530717958647692 * (double)d.rev[e]);
  }
  if (qvel) for (int i = 0; i < 15; i++) qvel[(size_t)e * 15 + i] = d.qvel[i * N + e];
  if (warm) for (int i = 0; i < 15; i++) warm[(size_t)e * 15 + i] = d.warm[i * N + e];
}
__global__ void jb_set_state_kernel(const __grid_constant__ JbDev d, const float* qpos, const float* qvel, const float* warm) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.N) return;
  const int N = d.N;
  if (qpos) {
    for (int i = 0; i < 15; i++) d.qpos[i * N + e] = qpos[(size_t)e * 16 + i];
    double qm = qpos[(size_t)e * 16 + 15];
    double r = ceil((qm - 3.14159265358979323846) / 6.28318530717958647692);
    d.qpos[15 * N + e] = (float)(qm - 6.28318530717958647692 * r);
    d.rev[e] = (int)r;
  }
  if (qvel) for (int i = 0; i < 15; i++) d.qvel[i * N + e] = qvel[(size_t)e * 15 + i];
  if (warm) for (int i = 0; i < 15; i++) d.warm[i * N + e] = warm[(size_t)e * 15 + i];
}
__global__ void jb_target_kernel(const __grid_constant__ JbDev d, float* tgt, int set) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.N) return;
  for (int i = 0; i < 4; i++) {
    if (set) d.target[i * d.N + e] = tgt[(size_t)e * 4 + i];
    else tgt[(size_t)e * 4 + i] = d.target[i * d.N + e];
  }
}
// task.get_observation(physics) / task.get_reward(physics) at the current state, one thread per env
struct JbRootView { float pos[3], quat[4], vlin[3], w[3], qm, vm; };
__global__ void jb_observe_kernel(const __grid_constant__ JbDev d, float* __restrict__ obs_out, float* __restrict__ reward_out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.N) return;
  const int N = d.N;
  JbRootView s;
#pragma unroll
  for (int i = 0; i < 3; i++) { s.pos[i] = d.qpos[i * N + e]; s.vlin[i] = d.qvel[i * N + e]; s.w[i] = d.qvel[(3 + i) * N + e]; }
#pragma unroll
  for (int i = 0; i < 4; i++) s.quat[i] = d.qpos[(3 + i) * N + e];
  s.qm = d.qpos[15 * N + e]; s.vm = d.qvel[14 * N + e];       // wrapped motor angle: the observation wraps it anyway
  JbTarget<float> tg = {d.target[e], d.target[N + e], d.target[2 * N + e], d.target[3 * N + e]};
  const float* g = d.packed + (d.per_env ? (size_t)e * JBP_SIZE : 0);
  float obs[JB_OBS_MAX], rew;
  const int n = jb_obs_reward<float>(g, s, d.task[e], tg, obs, &rew);
  if (obs_out) {
    float* o = obs_out + (size_t)e * d.obs_stride;
    for (int i = 0; i < d.obs_stride; i++) o[i] = i < n ? obs[i] : 0.f;
  }
  if (reward_out) reward_out[e] = rew;
}

__global__ void jb_iota_kernel(int* p, int N) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < N) p[e] = e;
}
__global__ void jb_count_kernel(const __grid_constant__ JbDev d, int* cnt, int set) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.N) return;
  if (set) { d.step_count[e] = cnt[e]; d.reset_next[e] = 0; }
  else cnt[e] = d.step_count[e];
}

// expand a population of K packed models into the per-env array: dst[e] = src[index[e]]
__global__ void jb_expand_models_kernel(float* dst, const float* src, const int* index, int N) {
  const int e = blockIdx.x;
  const float* sp = src + (size_t)index[e] * JBP_SIZE;
  float* dp = dst + (size_t)e * JBP_SIZE;
  for (int i = threadIdx.x; i < JBP_SIZE; i += blockDim.x) dp[i] = sp[i];
  (void)N;
}

// FP32 FMA-pipe peak: 16 independent FFMA chains per thread, enough CTAs to fill the chip (roofline denominator
// for this compute-bound path; MEASURED_PEAKS.json only carries HBM and tensor numbers).
__global__ void jb_fma_peak_kernel(float* out, int iters, float a, float b) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = (float)(threadIdx.x + i) * 1e-3f;
  for (int k = 0; k < iters; k++) {
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = fmaf(x[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; i++) s += x[i];
  if (s == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true; keeps the chains alive
}

// ------------------------------------------------------------------------------------------------ C-ABI
extern "C" {

const char* jb_last_error(void) { return g_err.c_str(); }
const char* jb_version(void) { return "jbsim 0.2 (sm_100a, quad-per-env)"; }
int jb_sim_num_envs(const JbSim* sim) { return sim ? sim->cfg.num_envs : 0; }
unsigned long long jb_sim_launch_count(const JbSim* sim) { return sim ? sim->launches : 0; }
int jb_sim_obs_stride(const JbSim* sim) { return sim ? sim->cfg.obs_stride : 0; }

static int grid_of(int n) { return (n + JB_BLOCK - 1) / JB_BLOCK; }

int jb_sim_create(JbSim** out, const JbSimConfig* cfg, const JbModel* models_host, int n_models, const int32_t* task_ids_host) {
  return jb_sim_create_indexed(out, cfg, models_host, n_models, nullptr, task_ids_host);
}

// shared tail of the creators: everything but the model.  Takes ownership of `packed_dev`.
static int sim_fill(JbSim* s, const JbSimConfig* cfg, const int32_t* task_ids_host, float* packed_dev, bool per_env) {
  s->cfg = *cfg;
  s->per_env = per_env;
  const size_t N = cfg->num_envs;
  JbDev& d = s->d;
  d.packed = packed_dev;
  d.per_env = per_env ? 1 : 0;
  d.N = (int)N; d.n_sub = cfg->n_sub_steps; d.step_limit = cfg->step_limit; d.random_pose = cfg->random_pose;
  d.obs_stride = cfg->obs_stride; d.seed = cfg->seed; d.env_id_offset = cfg->env_id_offset;
  CUDA_TRY(cudaMalloc(&d.qpos, 16 * N * sizeof(float)));
  CUDA_TRY(cudaMalloc(&d.qvel, 15 * N * sizeof(float)));
  CUDA_TRY(cudaMalloc(&d.warm, 15 * N * sizeof(float)));
  CUDA_TRY(cudaMalloc(&d.target, 4 * N * sizeof(float)));
  CUDA_TRY(cudaMalloc(&d.rev, N * sizeof(int)));
  CUDA_TRY(cudaMalloc(&d.task, N * sizeof(int)));
  CUDA_TRY(cudaMalloc(&d.step_count, N * sizeof(int)));
  CUDA_TRY(cudaMalloc(&d.episode, N * sizeof(int)));
  CUDA_TRY(cudaMalloc(&d.reset_next, N));
  CUDA_TRY(cudaMalloc(&d.counters, 16 * sizeof(unsigned long long)));
  CUDA_TRY(cudaMemset(d.qpos, 0, 16 * N * sizeof(float)));
  CUDA_TRY(cudaMemset(d.qvel, 0, 15 * N * sizeof(float)));
  CUDA_TRY(cudaMemset(d.warm, 0, 15 * N * sizeof(float)));
  CUDA_TRY(cudaMemset(d.target, 0, 4 * N * sizeof(float)));
  CUDA_TRY(cudaMemset(d.rev, 0, N * sizeof(int)));
  CUDA_TRY(cudaMemset(d.step_count, 0, N * sizeof(int)));
  CUDA_TRY(cudaMemset(d.episode, 0, N * sizeof(int)));
  CUDA_TRY(cudaMemset(d.reset_next, 1, N));           // a step() before any reset() is a reset (dm_control)
  CUDA_TRY(cudaMemset(d.counters, 0, 16 * sizeof(unsigned long long)));
  {
    const long long big = 0x7fffffffffffffffLL;                        // counters[7]: running minimum
    CUDA_TRY(cudaMemcpy(d.counters + 7, &big, sizeof(big), cudaMemcpyHostToDevice));
  }
  s->monitor = 1;
  if (const char* e = getenv("JBSIM_MONITOR")) s->monitor = atoi(e) != 0;
  if (task_ids_host) CUDA_TRY(cudaMemcpy(d.task, task_ids_host, N * sizeof(int), cudaMemcpyHostToDevice));
  else CUDA_TRY(cudaMemset(d.task, 0, N * sizeof(int)));
  CUDA_TRY(jb_configure_kernels_perenv());
  CUDA_TRY(jb_configure_kernels());
  CUDA_TRY(cudaMalloc(&d.load, N * sizeof(int4)));
  CUDA_TRY(cudaMemset(d.load, 0, N * sizeof(int4)));
  CUDA_TRY(cudaMalloc(&s->perm_buf, 2 * N * sizeof(int)));
  CUDA_TRY(cudaMalloc(&s->sort_scratch, (size_t)jb_sort_scratch_ints((int)N) * sizeof(int)));
  jb_iota_kernel<<<grid_of((int)N), JB_BLOCK>>>(s->perm_buf, (int)N);      // identity until the first sort
  CUDA_TRY(cudaGetLastError());
  s->perm_cur = 0;
  // Sorting pays when the grid is several waves deep (the heavy envs then fill whole CTAs of their own); small batches
  // (parity configurations, the single-env adapter) keep slot i = env i and one launch per control step.
  s->auto_sort = N >= (size_t)JB_SORT_MIN_ENVS ? 1 : 0;
  d.perm = s->auto_sort ? s->perm_buf : nullptr;
  // An env's load is persistent over ~10 sub-steps, much less over 50 (measured: profiles/README.md, round 2): the
  // control step is cut into launches of JB_SEGMENT_SUBSTEPS sub-steps and the order re-sorted after each.
  s->seg_sub = JB_SEGMENT_SUBSTEPS;
  if (const char* e = getenv("JBSIM_SEGMENT")) s->seg_sub = atoi(e) > 0 ? atoi(e) : (1 << 30);
  CUDA_TRY(cudaDeviceSynchronize());      // the initialisation above ran on the default stream; callers bring their own
  return 0;
}

extern "C" void jb_sim_destroy(JbSim* s);
static int sim_create_common(JbSim** out, const JbSimConfig* cfg, const int32_t* task_ids_host, float* packed_dev,
                             bool per_env) {
  JbSim* s = new JbSim();
  memset(s, 0, sizeof(*s));
  s->d.packed = packed_dev;                   // owned from here on: released by jb_sim_destroy on any failure below
  s->cfg.device = cfg->device;
  const int rc = sim_fill(s, cfg, task_ids_host, packed_dev, per_env);
  if (rc != 0) { jb_sim_destroy(s); return rc; }
  *out = s;
  return 0;
}

static int check_config(const JbSimConfig* cfg) {
  if (cfg->num_envs <= 0) return fail("jb_sim_create: num_envs must be positive");
  if (cfg->obs_stride < 15 || cfg->obs_stride > JB_OBS_MAX) return fail("jb_sim_create: obs_stride must be in [15, 19]");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) return fail("jb_sim_create: no CUDA device (this library has no CPU fallback)");
  if (cfg->device < 0 || cfg->device >= ndev) return fail("jb_sim_create: no such CUDA device");
  return 0;
}

int jb_sim_create_indexed(JbSim** out, const JbSimConfig* cfg, const JbModel* models_host, int n_models,
                          const int32_t* model_index_host, const int32_t* task_ids_host) {
  if (!out || !cfg || !models_host) return fail("jb_sim_create: null argument");
  if (cfg->num_envs <= 0) return fail("jb_sim_create: num_envs must be positive");
  if (n_models <= 0) return fail("jb_sim_create: n_models must be positive");
  if (!model_index_host && n_models != 1 && n_models != cfg->num_envs)
    return fail("jb_sim_create: n_models must be 1 or num_envs (or pass a model index)");
  if (model_index_host)
    for (int e = 0; e < cfg->num_envs; e++)
      if (model_index_host[e] < 0 || model_index_host[e] >= n_models) return fail("jb_sim_create: model index out of range");
  if (int rc = check_config(cfg)) return rc;
  DeviceGuard guard_(cfg->device);
  const size_t N = cfg->num_envs;
  // pack the model(s) into the lane-interleaved float layout the quad kernel stages in shared memory
  const size_t nm = (size_t)n_models;
  std::vector<float> packed(nm * JBP_SIZE);
  for (size_t i = 0; i < nm; i++) jb_pack_model<float, JbModel>(models_host[i], packed.data() + i * JBP_SIZE);
  DevBuf pdev, idev, full;                   // released on every early return; release() hands ownership on
  CUDA_TRY(cudaMalloc(&pdev.p, packed.size() * sizeof(float)));
  CUDA_TRY(cudaMemcpy(pdev.p, packed.data(), packed.size() * sizeof(float), cudaMemcpyHostToDevice));
  if (model_index_host) {
    // K distinct models + per-env index: expand on the device (the host never holds num_envs models)
    CUDA_TRY(cudaMalloc(&idev.p, N * sizeof(int)));
    CUDA_TRY(cudaMemcpy(idev.p, model_index_host, N * sizeof(int), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&full.p, N * JBP_SIZE * sizeof(float)));
    jb_expand_models_kernel<<<(unsigned)N, 128>>>((float*)full.p, (const float*)pdev.p, (const int*)idev.p, (int)N);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    return sim_create_common(out, cfg, task_ids_host, (float*)full.release(), true);
  }
  return sim_create_common(out, cfg, task_ids_host, (float*)pdev.release(), n_models != 1);
}

int jb_sim_create_packed(JbSim** out, const JbSimConfig* cfg, const float* packed_dev, int n_models,
                         const int32_t* task_ids_host) {
  if (!out || !cfg || !packed_dev) return fail("jb_sim_create_packed: null argument");
  if (n_models != 1 && n_models != cfg->num_envs) return fail("jb_sim_create_packed: n_models must be 1 or num_envs");
  if (int rc = check_config(cfg)) return rc;
  DeviceGuard guard_(cfg->device);
  DevBuf pdev;
  const size_t bytes = (size_t)n_models * JBP_SIZE * sizeof(float);
  CUDA_TRY(cudaMalloc(&pdev.p, bytes));
  CUDA_TRY(cudaMemcpy(pdev.p, packed_dev, bytes, cudaMemcpyDeviceToDevice));
  return sim_create_common(out, cfg, task_ids_host, (float*)pdev.release(), n_models != 1);
}

void jb_sim_destroy(JbSim* s) {
  if (!s) return;
  JB_GUARD(s);
  JbDev& d = s->d;
  cudaFree(d.qpos); cudaFree(d.qvel); cudaFree(d.warm); cudaFree(d.target); cudaFree(d.rev); cudaFree(d.task);
  cudaFree(d.step_count); cudaFree(d.episode); cudaFree(d.reset_next); cudaFree(d.counters);
  cudaFree(d.load); cudaFree(s->perm_buf); cudaFree(s->sort_scratch);
  if (d.packed) cudaFree((void*)d.packed);
  if (s->rep_geom) cudaFree(s->rep_geom);
  if (s->rep_dist) cudaFree(s->rep_dist);
  cudaFreeHost(s->h_act); cudaFreeHost(s->h_obs); cudaFreeHost(s->h_rew); cudaFreeHost(s->h_st);
  cudaFree(s->s_act); cudaFree(s->s_obs); cudaFree(s->s_rew); cudaFree(s->s_st);
  delete s;
}

int jb_sim_reset(JbSim* s, const uint8_t* mask_dev, float* obs_dev, void* stream) {
  if (!s) return fail("jb_sim_reset: null sim");
  JB_GUARD(s);
  cudaStream_t st = (cudaStream_t)stream;
  JbLaunchArgs a = {JB_MODE_RESET, mask_dev, nullptr, obs_dev, nullptr, nullptr, nullptr, nullptr, 0, 1, 1};
  CUDA_TRY(launch(s, a, st));
  return 0;
}

int jb_sim_step(JbSim* s, const float* actions_dev, float* obs_dev, float* reward_dev, int32_t* step_type_dev, void* stream) {
  if (!s || !actions_dev) return fail("jb_sim_step: null argument");
  JB_GUARD(s);
  cudaStream_t st = (cudaStream_t)stream;
  JbLaunchArgs a = {JB_MODE_STEP, nullptr, actions_dev, obs_dev, reward_dev, step_type_dev, nullptr, nullptr, 0, 0, 0};
  const int n_sub = s->d.n_sub, seg = s->auto_sort ? s->seg_sub : n_sub;
  const size_t N = s->d.N;
  // Slot order: segment i reads perm buffer i & 1 and the sort that follows it writes the other one; every control step
  // starts from buffer 0 (an odd number of segments ends with a device copy 1 -> 0), so the sequence of launches and
  // pointers is the same for every step -- a step can be captured in a CUDA graph and replayed.
  int cur = 0;
  for (int k0 = 0; k0 < n_sub; k0 += seg) {
    a.sub_count = n_sub - k0 < seg ? n_sub - k0 : seg;
    a.first = k0 == 0;
    a.last = k0 + a.sub_count >= n_sub;
    if (s->auto_sort) s->d.perm = s->perm_buf + (size_t)cur * N;
    CUDA_TRY(launch(s, a, st));
    if (s->auto_sort) {
      CUDA_TRY(jb_sort_by_load(s->d.load, s->d.N, a.sub_count, s->perm_buf + (size_t)(cur ^ 1) * N, s->sort_scratch, st));
      s->launches += 3;
      cur ^= 1;
    }
  }
  if (s->auto_sort) {
    if (cur == 1) CUDA_TRY(cudaMemcpyAsync(s->perm_buf, s->perm_buf + N, N * sizeof(int), cudaMemcpyDeviceToDevice, st));
    s->d.perm = s->perm_buf;
  }
  // the period is counted on the host: a CUDA graph captured from this call replays the monitor on every step or on none,
  // whichever the captured step did (it only feeds counters[5..7], never the simulation)
  if (s->monitor && (s->steps++ % JB_MONITOR_PERIOD) == 0) {
    CUDA_TRY(jb_launch_self_clearance(s->d, nullptr, st));
    s->launches++;
  }
  return 0;
}

// staging of the host-buffer entry point, allocated on its first use (170 MB at 1 Mi envs that device-resident callers
// never need)
static int ensure_host_staging(JbSim* s) {
  if (s->s_act) return 0;
  const size_t N = s->d.N, os = s->cfg.obs_stride;
  CUDA_TRY(cudaMallocHost(&s->h_act, N * sizeof(float)));
  CUDA_TRY(cudaMallocHost(&s->h_obs, N * os * sizeof(float)));
  CUDA_TRY(cudaMallocHost(&s->h_rew, N * sizeof(float)));
  CUDA_TRY(cudaMallocHost(&s->h_st, N * sizeof(int32_t)));
  CUDA_TRY(cudaMalloc(&s->s_obs, N * os * sizeof(float)));
  CUDA_TRY(cudaMalloc(&s->s_rew, N * sizeof(float)));
  CUDA_TRY(cudaMalloc(&s->s_st, N * sizeof(int32_t)));
  CUDA_TRY(cudaMalloc(&s->s_act, N * sizeof(float)));
  return 0;
}

// page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory) host memory can be the DMA end point itself
static bool is_pinned(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}

int jb_sim_step_host(JbSim* s, const float* actions_host, float* obs_host, float* reward_host, int32_t* step_type_host, void* stream) {
  if (!s || !actions_host) return fail("jb_sim_step_host: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  JB_GUARD(s);
  if (ensure_host_staging(s)) return 1;
  const size_t N = s->d.N, os = s->cfg.obs_stride;
  // pageable caller buffers go through the simulator's pinned staging; pinned ones are copied to / from directly
  const float* a_src = actions_host;
  if (!is_pinned(actions_host)) { memcpy(s->h_act, actions_host, N * sizeof(float)); a_src = s->h_act; }
  float* o_dst = obs_host && is_pinned(obs_host) ? obs_host : s->h_obs;
  float* r_dst = reward_host && is_pinned(reward_host) ? reward_host : s->h_rew;
  int32_t* t_dst = step_type_host && is_pinned(step_type_host) ? step_type_host : s->h_st;
  CUDA_TRY(cudaMemcpyAsync(s->s_act, a_src, N * sizeof(float), cudaMemcpyHostToDevice, st));
  if (jb_sim_step(s, s->s_act, s->s_obs, s->s_rew, s->s_st, stream)) return 1;
  if (obs_host) CUDA_TRY(cudaMemcpyAsync(o_dst, s->s_obs, N * os * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (reward_host) CUDA_TRY(cudaMemcpyAsync(r_dst, s->s_rew, N * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (step_type_host) CUDA_TRY(cudaMemcpyAsync(t_dst, s->s_st, N * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (obs_host && o_dst != obs_host) memcpy(obs_host, s->h_obs, N * os * sizeof(float));
  if (reward_host && r_dst != reward_host) memcpy(reward_host, s->h_rew, N * sizeof(float));
  if (step_type_host && t_dst != step_type_host) memcpy(step_type_host, s->h_st, N * sizeof(int32_t));
  return 0;
}

int jb_sim_observe(JbSim* s, float* obs_dev, float* reward_dev, void* stream) {
  if (!s || (!obs_dev && !reward_dev)) return fail("jb_sim_observe: null argument");
  JB_GUARD(s);
  s->launches++;
  jb_observe_kernel<<<grid_of(s->d.N), JB_BLOCK, 0, (cudaStream_t)stream>>>(s->d, obs_dev, reward_dev);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int jb_sim_heuristic_actions(JbSim* s, const float* obs_dev, float* actions_dev, void* stream) {
  if (!s || !obs_dev || !actions_dev) return fail("jb_sim_heuristic_actions: null argument");
  JB_GUARD(s);
  s->launches++;
  jb_heuristic_kernel<<<grid_of(s->d.N), JB_BLOCK, 0, (cudaStream_t)stream>>>(s->d, obs_dev, actions_dev);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int jb_sim_get_state(JbSim* s, float* qpos_dev, float* qvel_dev, float* warm_dev, void* stream) {
  if (!s) return fail("jb_sim_get_state: null sim");
  JB_GUARD(s);
  jb_get_state_kernel<<<grid_of(s->d.N), JB_BLOCK, 0, (cudaStream_t)stream>>>(s->d, qpos_dev, qvel_dev, warm_dev);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
int jb_sim_set_state(JbSim* s, const float* qpos_dev, const float* qvel_dev, const float* warm_dev, void* stream) {
  if (!s) return fail("jb_sim_set_state: null sim");
  JB_GUARD(s);
  jb_set_state_kernel<<<grid_of(s->d.N), JB_BLOCK, 0, (cudaStream_t)stream>>>(s->d, qpos_dev, qvel_dev, warm_dev);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
int jb_sim_get_target(JbSim* s, float* target_dev, void* stream) {
  if (!s || !target_dev) return fail("jb_sim_get_target: null argument");
  JB_GUARD(s);
  jb_target_kernel<<<grid_of(s->d.N), JB_BLOCK, 0, (cudaStream_t)stream>>>(s->d, target_dev, 0);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
int jb_sim_set_target(JbSim* s, const float* target_dev, void* stream) {
  if (!s || !target_dev) return fail("jb_sim_set_target: null argument");
  JB_GUARD(s);
  jb_target_kernel<<<grid_of(s->d.N), JB_BLOCK, 0, (cudaStream_t)stream>>>(s->d, (float*)target_dev, 1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
int jb_sim_get_step_count(JbSim* s, int32_t* count_dev, void* stream) {
  if (!s || !count_dev) return fail("jb_sim_get_step_count: null argument");
  JB_GUARD(s);
  jb_count_kernel<<<grid_of(s->d.N), JB_BLOCK, 0, (cudaStream_t)stream>>>(s->d, count_dev, 0);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
int jb_sim_set_step_count(JbSim* s, const int32_t* count_dev, void* stream) {
  if (!s || !count_dev) return fail("jb_sim_set_step_count: null argument");
  JB_GUARD(s);
  jb_count_kernel<<<grid_of(s->d.N), JB_BLOCK, 0, (cudaStream_t)stream>>>(s->d, (int*)count_dev, 1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
// ---- exact checkpoint / resume: the per-env arrays of JbDev that a step reads and writes, back to back in one blob
struct SnapPart { void* p; size_t bytes; };
static int snap_parts(JbSim* s, SnapPart* out) {
  const size_t N = s->d.N;
  JbDev& d = s->d;
  const SnapPart parts[] = {{d.qpos, 16 * N * sizeof(float)}, {d.qvel, 15 * N * sizeof(float)}, {d.warm, 15 * N * sizeof(float)},
                            {d.target, 4 * N * sizeof(float)}, {d.rev, N * sizeof(int)}, {d.step_count, N * sizeof(int)},
                            {d.episode, N * sizeof(int)}, {d.reset_next, N}};
  const int n = (int)(sizeof(parts) / sizeof(parts[0]));
  for (int i = 0; i < n; i++) out[i] = parts[i];
  return n;
}
size_t jb_sim_snapshot_size(const JbSim* s) {
  if (!s) return 0;
  SnapPart parts[8];
  const int n = snap_parts(const_cast<JbSim*>(s), parts);
  size_t total = 0;
  for (int i = 0; i < n; i++) total += (parts[i].bytes + 15) & ~(size_t)15;
  return total;
}
static int snap_copy(JbSim* s, char* blob, bool restore, cudaStream_t st) {
  SnapPart parts[8];
  const int n = snap_parts(s, parts);
  size_t off = 0;
  for (int i = 0; i < n; i++) {
    if (restore) CUDA_TRY(cudaMemcpyAsync(parts[i].p, blob + off, parts[i].bytes, cudaMemcpyDeviceToDevice, st));
    else CUDA_TRY(cudaMemcpyAsync(blob + off, parts[i].p, parts[i].bytes, cudaMemcpyDeviceToDevice, st));
    off += (parts[i].bytes + 15) & ~(size_t)15;
  }
  return 0;
}
int jb_sim_snapshot(JbSim* s, void* blob_dev, void* stream) {
  if (!s || !blob_dev) return fail("jb_sim_snapshot: null argument");
  JB_GUARD(s);
  return snap_copy(s, (char*)blob_dev, false, (cudaStream_t)stream);
}
int jb_sim_restore(JbSim* s, const void* blob_dev, void* stream) {
  if (!s || !blob_dev) return fail("jb_sim_restore: null argument");
  JB_GUARD(s);
  return snap_copy(s, (char*)const_cast<void*>(blob_dev), true, (cudaStream_t)stream);
}
int jb_sim_get_contacts(JbSim* s, int32_t* ncon_dev, int32_t* geom_dev, float* dist_dev, int max_con, void* stream) {
  if (!s || !ncon_dev || !geom_dev || !dist_dev || max_con <= 0) return fail("jb_sim_get_contacts: bad argument");
  JB_GUARD(s);
  cudaStream_t st = (cudaStream_t)stream;
  if (!s->rep_geom) {
    CUDA_TRY(cudaMalloc(&s->rep_geom, (size_t)s->d.N * JB_REP * sizeof(int)));
    CUDA_TRY(cudaMalloc(&s->rep_dist, (size_t)s->d.N * JB_REP * sizeof(float)));
  }
  JbLaunchArgs a = {JB_MODE_CONTACTS, nullptr, nullptr, nullptr, nullptr, nullptr, s->rep_geom, s->rep_dist, 0, 1, 1};
  CUDA_TRY(launch(s, a, st));
  jb_compact_contacts_kernel<<<grid_of(s->d.N), JB_BLOCK, 0, st>>>(s->d.N, s->rep_geom, s->rep_dist, ncon_dev, geom_dev,
                                                                    dist_dev, max_con);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
// profiling builds (-DJBQ_PHASE_TIMERS): per-phase clock64 sums of lane 0 of every warp, slots 9..15
int jb_sim_get_phase_cycles(JbSim* s, unsigned long long* out8_host) {
  if (!s || !out8_host) return fail("jb_sim_get_phase_cycles: null argument");
  JB_GUARD(s);
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(out8_host, s->d.counters + 8, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return 0;
}

__global__ void jb_copy_load_kernel(const int4* src, int4* dst, int N) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < N) dst[e] = src[e];
}

int jb_sim_get_load(JbSim* s, int32_t* load_dev, void* stream) {
  if (!s || !load_dev) return fail("jb_sim_get_load: null argument");
  JB_GUARD(s);
  jb_copy_load_kernel<<<grid_of(s->d.N), JB_BLOCK, 0, (cudaStream_t)stream>>>(s->d.load, reinterpret_cast<int4*>(load_dev), s->d.N);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int jb_sim_set_slot_order(JbSim* s, int mode, const int32_t* perm_dev, void* stream) {
  if (!s) return fail("jb_sim_set_slot_order: null sim");
  JB_GUARD(s);
  if (mode == JB_ORDER_AUTO) {             // also for batches below JB_SORT_MIN_ENVS (tests)
    if (!s->auto_sort && !s->d.perm) {
      jb_iota_kernel<<<grid_of(s->d.N), JB_BLOCK, 0, (cudaStream_t)stream>>>(s->perm_buf, s->d.N);
      CUDA_TRY(cudaGetLastError());
    }
    s->auto_sort = 1;
    s->d.perm = s->perm_buf;
    return 0;
  }
  s->auto_sort = 0;
  if (mode == JB_ORDER_IDENTITY) { s->d.perm = nullptr; return 0; }
  if (mode != JB_ORDER_EXPLICIT || !perm_dev) return fail("jb_sim_set_slot_order: bad mode / null permutation");
  CUDA_TRY(cudaMemcpyAsync(s->perm_buf, perm_dev, (size_t)s->d.N * sizeof(int), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  s->d.perm = s->perm_buf;
  return 0;
}

int jb_sim_get_slot_order(JbSim* s, int32_t* perm_dev, void* stream) {
  if (!s || !perm_dev) return fail("jb_sim_get_slot_order: null argument");
  JB_GUARD(s);
  if (s->d.perm) {
    CUDA_TRY(cudaMemcpyAsync(perm_dev, s->d.perm, (size_t)s->d.N * sizeof(int), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  } else {
    jb_iota_kernel<<<grid_of(s->d.N), JB_BLOCK, 0, (cudaStream_t)stream>>>(perm_dev, s->d.N);
    CUDA_TRY(cudaGetLastError());
  }
  return 0;
}

int jb_sim_self_clearance(JbSim* s, float* clearance_dev, double* out3_host, void* stream) {
  if (!s) return fail("jb_sim_self_clearance: null sim");
  JB_GUARD(s);
  if (clearance_dev) {
    CUDA_TRY(jb_launch_self_clearance(s->d, clearance_dev, (cudaStream_t)stream));
    s->launches++;
  }
  if (out3_host) {
    unsigned long long c[3];
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(c, s->d.counters + 5, sizeof(c), cudaMemcpyDeviceToHost));
    out3_host[0] = (double)c[0];
    out3_host[1] = (double)c[1];
    out3_host[2] = c[0] ? (double)(long long)c[2] * 1e-9 : 0.0;
  }
  return 0;
}

int jb_sim_get_counters(JbSim* s, unsigned long long* out5_host) {
  if (!s || !out5_host) return fail("jb_sim_get_counters: null argument");
  JB_GUARD(s);
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(out5_host, s->d.counters, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return 0;
}

int jb_measure_fp32_peak(int device, double* tflops_out) {
  if (!tflops_out) return fail("jb_measure_fp32_peak: null argument");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 1 << 16;
  float* buf = nullptr;
  CUDA_TRY(cudaMalloc(&buf, (size_t)threads * blocks * sizeof(float)));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    CUDA_TRY(cudaEventRecord(e0));
    jb_fma_peak_kernel<<<blocks, threads>>>(buf, iters, 0.999f, 1e-3f);
    CUDA_TRY(cudaEventRecord(e1));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    double tf = 2.0 * 16.0 * (double)iters * threads * blocks / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf);
  *tflops_out = best;
  return 0;
}

}  // extern "C"
