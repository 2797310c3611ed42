// The fused kernel (sm_100a): one QUAD of lanes per environment runs a whole control step -- n_sub_steps physics
// sub-steps + reward + observation + episode bookkeeping -- without touching HBM in between.  CTA = up to 384 threads =
// 96 envs (the launcher picks the size); the packed model sits in shared memory (one copy per CTA for the shared model, one per env for augmented
// models).  See jb_quad.cuh for the per-quad algorithm and DESIGN.md for the roofline model.
#include "jb_sim_common.cuh"
#include <stdlib.h>

// This file is compiled twice (jitterbug_b200/build.py): as is for the shared-model kernel (384-thread CTAs, 168
// registers), and with -DJB_PER_ENV_BUILD for the per-env-model kernel: there every env's 1.8 KB model is staged in shared
// memory, which comes out of the L1 that caches the local-memory contact slots and spills, so that build trades CTA
// size for registers (256-thread bound: ~222 registers, no spills) and runs 256-thread CTAs.
#ifdef JB_PER_ENV_BUILD
#undef JB_BLOCK
#ifndef JB_PERENV_BLOCK
#define JB_PERENV_BLOCK 256
#endif
#define JB_BLOCK JB_PERENV_BLOCK
#ifndef JBQ_SPREAD
#define JBQ_SPREAD 0          /* spreading a heavy lane's contacts over its quad paid before the load sort (+3 %); with heavy
                                 envs packed into their own CTAs it costs every other CTA: 9.3 -> 10.2 M steps/s without it */
#endif
#define JB_SIM_KERNEL jb_sim_kernel_perenv
#define JB_LAUNCH_SIM jb_launch_sim_perenv
#define JB_CONFIGURE jb_configure_kernels_perenv
#else
#ifndef JBQ_SPREAD
#define JBQ_SPREAD 0
#endif
#define JB_SIM_KERNEL jb_sim_kernel
#define JB_LAUNCH_SIM jb_launch_sim
#define JB_CONFIGURE jb_configure_kernels
#endif
#include "jb_quad.cuh"
#include "jb_task.cuh"

#ifndef JB_SYNC_SUBSTEP
#define JB_SYNC_SUBSTEP 0       /* extra CTA barrier at the top of every sub-step: redundant next to the post-solver one (+2 %
                                   without it since the load sort) */
#endif
#ifndef JB_BLOCK_NEWTON
#define JB_BLOCK_NEWTON 0
#endif
#ifndef JB_PHASE_SYNC
#define JB_PHASE_SYNC 0x42      /* bit p: barrier at phase boundary p (1..6; 8..11 inside the Newton loop need JB_BLOCK_NEWTON).
                                   0x42 = after the leg phase and after the solver: two whole-CTA barriers per sub-step keep
                                   the warps in the same code region; measured ahead of all six (0x7e, round 1's setting)
                                   by 2 % once the load sort is on (profiles/README.md) */
#endif
#ifndef JB_MIN_BLOCKS
#define JB_MIN_BLOCKS 1
#endif

// Phase barriers synchronise JB_SYNC_GROUPS independent groups of warps (group = warp index mod JB_SYNC_GROUPS, i.e. with
// 4 groups the warps that share an SM sub-partition); 1 = the whole CTA.
#ifndef JB_SYNC_GROUPS
#define JB_SYNC_GROUPS 1
#endif
__device__ __forceinline__ void jb_group_sync() {
#if JB_SYNC_GROUPS <= 1
  __syncthreads();
#else
  const int w = threadIdx.x >> 5;
  asm volatile("bar.sync %0, %1;" ::"r"(1 + (w % JB_SYNC_GROUPS)), "r"((int)(blockDim.x / JB_SYNC_GROUPS)) : "memory");
#endif
}

// Quad sums as xor-butterflies over lanes {l, l^1, l^2, l^3}.  The kernel keeps control flow warp-uniform around every
// call (see jb_quad.cuh), so all 32 lanes execute each shuffle together and the full mask is legal.
struct QuadCommGPU {
  __device__ __forceinline__ float sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
  }
  __device__ __forceinline__ void sum2(float& a, float& b) { a = sum(a); b = sum(b); }
  __device__ __forceinline__ void sumN(float* v, int n) {
#pragma unroll
    for (int i = 0; i < 28; i++) if (i < n) v[i] = sum(v[i]);
  }
  __device__ __forceinline__ int isum(int v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
  }
  // value held by lane `src` (0..3) of the caller's quad
  __device__ __forceinline__ float get(float v, int src) { return __shfl_sync(0xffffffffu, v, (threadIdx.x & 28) | src); }
  __device__ __forceinline__ int iget(int v, int src) { return __shfl_sync(0xffffffffu, v, (threadIdx.x & 28) | src); }
  __device__ __forceinline__ bool warp_any(bool p) { return __any_sync(0xffffffffu, p) != 0; }
  __device__ __forceinline__ bool loop_any(bool p) {
#if JB_BLOCK_NEWTON
    return __syncthreads_or(p) != 0;     // CTA-uniform Newton trip count: allows barriers inside the loop
#else
    return __any_sync(0xffffffffu, p) != 0;
#endif
  }
  // CTA barrier at the phase boundaries of the sub-step: the CTA's warps then execute the same few-KB code region at
  // the same time and share its instruction-cache lines (the whole sub-step is several times the I-cache)
  __device__ __forceinline__ void phase_sync(int p) {
#if JB_PHASE_SYNC
    if ((JB_PHASE_SYNC >> p) & 1) jb_group_sync();
#endif
  }
#ifdef JBQ_PHASE_TIMERS
  long long t_last; long long acc[8];
  __device__ __forceinline__ void tick(int p) { long long t = clock64(); if (p > 0) acc[p] += t - t_last; t_last = t; }
#endif
};

__device__ __forceinline__ void load_state(const JbDev& d, int e, int lane, QState<float>& s) {
  const int N = d.N;
  s.pos[0] = d.qpos[0 * N + e]; s.pos[1] = d.qpos[1 * N + e]; s.pos[2] = d.qpos[2 * N + e];
#pragma unroll
  for (int i = 0; i < 4; i++) s.quat[i] = d.qpos[(3 + i) * N + e];
  s.qa = d.qpos[(7 + 2 * lane) * N + e]; s.qb = d.qpos[(8 + 2 * lane) * N + e];
  s.qm = d.qpos[15 * N + e];
  s.rev = d.rev[e];
#pragma unroll
  for (int i = 0; i < 3; i++) { s.vlin[i] = d.qvel[i * N + e]; s.w[i] = d.qvel[(3 + i) * N + e]; }
  s.va = d.qvel[(6 + 2 * lane) * N + e]; s.vb = d.qvel[(7 + 2 * lane) * N + e];
  s.vm = d.qvel[14 * N + e];
#pragma unroll
  for (int i = 0; i < 6; i++) s.wr[i] = d.warm[i * N + e];
  s.wla = d.warm[(6 + 2 * lane) * N + e]; s.wlb = d.warm[(7 + 2 * lane) * N + e];
  s.wmo = d.warm[14 * N + e];
}

__device__ __forceinline__ void store_state(const JbDev& d, int e, int lane, const QState<float>& s) {
  const int N = d.N;
  d.qpos[(7 + 2 * lane) * N + e] = s.qa; d.qpos[(8 + 2 * lane) * N + e] = s.qb;
  d.qvel[(6 + 2 * lane) * N + e] = s.va; d.qvel[(7 + 2 * lane) * N + e] = s.vb;
  d.warm[(6 + 2 * lane) * N + e] = s.wla; d.warm[(7 + 2 * lane) * N + e] = s.wlb;
  // replicated components: spread the stores over the four lanes
  if (lane == 0) {
    d.qpos[0 * N + e] = s.pos[0]; d.qpos[1 * N + e] = s.pos[1]; d.qpos[2 * N + e] = s.pos[2];
#pragma unroll
    for (int i = 0; i < 4; i++) d.qpos[(3 + i) * N + e] = s.quat[i];
  } else if (lane == 1) {
#pragma unroll
    for (int i = 0; i < 3; i++) { d.qvel[i * N + e] = s.vlin[i]; d.qvel[(3 + i) * N + e] = s.w[i]; }
  } else if (lane == 2) {
#pragma unroll
    for (int i = 0; i < 6; i++) d.warm[i * N + e] = s.wr[i];
  } else {
    d.qpos[15 * N + e] = s.qm; d.rev[e] = s.rev; d.qvel[14 * N + e] = s.vm; d.warm[14 * N + e] = s.wmo;
  }
}

__device__ __forceinline__ bool state_bad(const QState<float>& s) {
  float acc = fabsf(s.qa) + fabsf(s.qb) + fabsf(s.va) + fabsf(s.vb) + fabsf(s.qm) + fabsf(s.vm);
#pragma unroll
  for (int i = 0; i < 3; i++) acc += fabsf(s.pos[i]) + fabsf(s.vlin[i]) + fabsf(s.w[i]);
#pragma unroll
  for (int i = 0; i < 4; i++) acc += fabsf(s.quat[i]);
  return !(acc < 1e10f);      // NaN or huge (mj_checkPos / mj_checkVel threshold)
}

__device__ __forceinline__ void do_reset(const float* g, const JbDev& d, int e, int lane, QState<float>& s,
                                         JbTarget<float>& tg, int task, bool live, unsigned char reset_flag) {
  int ep = d.episode[e];
  jb_reset_env<float>(g, s, tg, task, d.random_pose, d.seed, d.env_id_offset + (unsigned)e, (unsigned)ep);
  s.qa = s.qb = s.va = s.vb = 0.f;
#pragma unroll
  for (int i = 0; i < 6; i++) s.wr[i] = 0.f;
  s.wla = s.wlb = s.wmo = 0.f;
  __syncwarp(__activemask());                   // all four lanes have read episode[e] before lane 0 bumps it
  if (lane == 0 && live) {
    d.episode[e] = ep + 1;
    d.step_count[e] = 0;
    d.reset_next[e] = reset_flag;
    const int N = d.N;
    d.target[0 * N + e] = tg.x; d.target[1 * N + e] = tg.y; d.target[2 * N + e] = tg.qw; d.target[3 * N + e] = tg.qz;
  }
}

__device__ __forceinline__ void write_obs(const float* g, const JbDev& d, int e, const QState<float>& s, int task,
                                          const JbTarget<float>& tg, float* obs_out, float* reward_out) {
  float obs[JB_OBS_MAX], rew;
  int n = jb_obs_reward<float>(g, s, task, tg, obs, &rew);
  if (obs_out) {
    float* o = obs_out + (size_t)e * d.obs_stride;
#pragma unroll
    for (int i = 0; i < JB_OBS_MAX; i++)
      if (i < d.obs_stride) o[i] = i < n ? obs[i] : 0.f;
  }
  if (reward_out) reward_out[e] = rew;
}

extern __shared__ float jb_smem[];

#ifdef JB_MAXNREG          /* tuning: explicit register budget instead of the one ptxas derives from the CTA size */
__global__ void __maxnreg__(JB_MAXNREG)
#else
__global__ void __launch_bounds__(JB_BLOCK, JB_MIN_BLOCKS)
#endif
JB_SIM_KERNEL(const __grid_constant__ JbDev d, const __grid_constant__ JbLaunchArgs a) {
  const int tid = threadIdx.x;
  const int lane = tid & 3;
  const int epb = blockDim.x >> 2;            // envs per CTA (the CTA size is chosen at launch, <= JB_BLOCK)
  const int e_raw = blockIdx.x * epb + (tid >> 2);
  // slot -> env: the quads of the grid work on the envs in the order of d.perm (load-sorted, see jb_sort.cu), so that
  // the heavy envs share warps and CTAs; padding quads past the end of the batch and envs not selected by a reset mask
  // shadow a valid env and skip all stores (every thread runs the whole body: quad shuffles and warp votes need full warps)
  const int slot = min(e_raw, d.N - 1);
  const int e = d.perm ? d.perm[slot] : slot;
  // ---- packed model -> shared memory
  if (d.per_env) {
    // 16-byte copies; each env's model starts JBP_ENV_STRIDE floats after the previous one (see jb_sim_common.cuh)
    const int e0 = blockIdx.x * epb;
    const int nenv = min(epb, d.N - e0);
    float4* dst = reinterpret_cast<float4*>(jb_smem);
    for (int i = tid; i < nenv * (JBP_SIZE / 4); i += blockDim.x) {
      const int q = i / (JBP_SIZE / 4), off = i - q * (JBP_SIZE / 4);
      const int env = d.perm ? d.perm[e0 + q] : e0 + q;
      dst[q * (JBP_ENV_STRIDE / 4) + off] = reinterpret_cast<const float4*>(d.packed + (size_t)env * JBP_SIZE)[off];
    }
  } else {
    for (int i = tid; i < JBP_SIZE; i += blockDim.x) jb_smem[i] = d.packed[i];
  }
  __syncthreads();
  const float* P = d.per_env ? jb_smem + (tid >> 2) * JBP_ENV_STRIDE : jb_smem;
  QuadCommGPU cm;
#ifdef JBQ_PHASE_TIMERS
  for (int i = 0; i < 8; i++) cm.acc[i] = 0;
  cm.t_last = clock64();
#endif
  const int N = d.N;
  const int task = d.task[e];
  QState<float> s;
  JbTarget<float> tg;
  const unsigned char rflag = d.reset_next[e];
  const bool reset_now = a.mode == JB_MODE_RESET ? (!a.mask || a.mask[e]) : (a.mode == JB_MODE_STEP && a.first && rflag == 1);
  // an env that was reset by an earlier segment of this control step sits the remaining segments out
  const bool sit_out = a.mode == JB_MODE_STEP && !a.first && rflag == 2;
  const bool live = e_raw < d.N && !(a.mode == JB_MODE_RESET && !reset_now);
  // a masked reset usually selects few envs (episode boundaries of a VecEnv loop): CTAs without one leave at once
  if (a.mode == JB_MODE_RESET && a.mask && !__syncthreads_or(live ? 1 : 0)) return;
  int iters = 0, ncon_sum = 0, overflow = 0, nsub = 0, lane_load = 0;
  float ctrl = 0.f;
  if (reset_now) {
    do_reset(P, d, e, lane, s, tg, task, live, a.last ? 0 : 2);
  } else {
    load_state(d, e, lane, s);
    tg.x = d.target[0 * N + e]; tg.y = d.target[1 * N + e]; tg.qw = d.target[2 * N + e]; tg.qz = d.target[3 * N + e];
    if (a.mode == JB_MODE_STEP && !sit_out) { ctrl = a.actions[e]; nsub = a.sub_count; }
  }
  // a reset / contact report runs one forward-only pass (mj_forward) for the warm start / the contact list.
  // Envs of one warp can be in different modes only at episode boundaries; the trip count is made warp-uniform.
  const bool fwd_only = nsub == 0;
  QReport<float> rep = {a.rep_geom + (size_t)e * 4 * JBQ_CAP, a.rep_dist + (size_t)e * 4 * JBQ_CAP, 1};
  const bool want_rep = a.mode == JB_MODE_CONTACTS && live;
  if (want_rep) {
#pragma unroll
    for (int i = 0; i < JBQ_CAP; i++) rep.geom[lane * JBQ_CAP + i] = -1;
  }
  int trips = fwd_only ? 1 : nsub;
  trips = __reduce_max_sync(0xffffffffu, trips);
  for (int k = 0; k < trips; k++) {
#if JB_SYNC_SUBSTEP
    jb_group_sync();            // keeps the CTA's warps in the same code region: instruction-cache lines are shared
#endif
    // an env that is resetting inside a stepping warp (episode boundary) just repeats zero-length sub-steps: they
    // leave the state untouched and recompute the same warm start
    jbq_substep<float, QuadCommGPU>(cm, P, lane, s, ctrl, fwd_only, &iters, &ncon_sum, &overflow, want_rep ? &rep : nullptr,
                                    &lane_load);
  }
  if (a.mode == JB_MODE_CONTACTS) return;
  // load record of these sub-steps (jb_sort.cu keys the next launch's slot order on it): the fullest lane's figures
  int ld_n = lane_load & 0xfff, ld_3 = (lane_load >> 12) & 0xff, ld_2 = (lane_load >> 20) & 0xff;
  int ld_last = (int)((unsigned)lane_load >> 28);
#pragma unroll
  for (int m = 1; m <= 2; m <<= 1) {
    ld_n = max(ld_n, __shfl_xor_sync(0xffffffffu, ld_n, m));
    ld_3 = max(ld_3, __shfl_xor_sync(0xffffffffu, ld_3, m));
    ld_2 = max(ld_2, __shfl_xor_sync(0xffffffffu, ld_2, m));
    ld_last = max(ld_last, __shfl_xor_sync(0xffffffffu, ld_last, m));
  }
  int bad = 0;
  if (a.last) {                 // grid-uniform
    if (!fwd_only) {            // mj_step1 of the last sub-step normalises qpos quaternions in place
      float qn = rsqrtf(s.quat[0] * s.quat[0] + s.quat[1] * s.quat[1] + s.quat[2] * s.quat[2] + s.quat[3] * s.quat[3]);
      s.quat[0] *= qn; s.quat[1] *= qn; s.quat[2] *= qn; s.quat[3] *= qn;
    }
    // mj_checkPos / mj_checkVel: MuJoCo resets the data (dm_control raises PhysicsError); here the env restarts
    bad = cm.isum(!fwd_only && state_bad(s) ? 1 : 0);
    if (bad) do_reset(P, d, e, lane, s, tg, task, live, 0);
  }
  if (!live) return;
  if (fwd_only) {
    if (reset_now) {
      store_state(d, e, lane, s);
      if (lane == 0) {
        write_obs(P, d, e, s, task, tg, a.obs, nullptr);
        if (a.reward) a.reward[e] = 0.f;
        if (a.step_type) a.step_type[e] = JB_STEP_FIRST;
      }
    } else if (a.last && lane == 0) {
      d.reset_next[e] = 0;      // sat out the control step it was reset in: back to stepping
    }
    return;
  }
  if (overflow) atomicAdd(&d.counters[3], (unsigned long long)overflow);     // rare: any lane that ran out of slots
  store_state(d, e, lane, s);
  if (lane == 0) {
    atomicAdd(&d.counters[1], (unsigned long long)iters);
    atomicAdd(&d.counters[2], (unsigned long long)ncon_sum);
    if (e == 0) atomicAdd(&d.counters[0], (unsigned long long)a.sub_count * (unsigned long long)N);
    if (d.load) d.load[e] = make_int4(ld_n, ld_3, ld_2, iters | (ld_last << 16));
    if (a.last) {
      int sc = bad ? 0 : d.step_count[e] + 1;
      d.step_count[e] = sc;
      int st = JB_STEP_MID;
      if (d.step_limit > 0 && sc >= d.step_limit) { st = JB_STEP_LAST; d.reset_next[e] = 1; }
      write_obs(P, d, e, s, task, tg, a.obs, a.reward);
      // a diverged env has just been restarted: its caller sees the episode boundary (FIRST, reward 0, the observation
      // of the fresh episode) -- dm_control raises PhysicsError here and the caller resets
      if (bad) { st = JB_STEP_FIRST; if (a.reward) a.reward[e] = 0.f; }
      if (a.step_type) a.step_type[e] = st;
      if (bad) atomicAdd(&d.counters[4], 1ull);
    }
  }
#ifdef JBQ_PHASE_TIMERS
  if ((tid & 31) == 0) for (int i = 1; i < 8; i++) atomicAdd(&d.counters[8 + i], (unsigned long long)cm.acc[i]);
#endif
}

// CTA size: JB_BLOCK threads (one CTA per SM) when the batch fills the GPU; for small batches (e.g. the 4096-env parity
// configuration) smaller CTAs spread the envs over more SMs.  JBSIM_BLOCK overrides (tuning experiments).
static int jb_pick_block(int N, int per_env) {
  static int sms = 0, forced = -1;
  if (forced < 0) {
    const char* e = getenv("JBSIM_BLOCK");
    forced = e ? atoi(e) : 0;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  if (forced >= 32 && forced <= JB_BLOCK && forced % 32 == 0) return forced;
  // minimise waves x time-per-CTA with the measured per-CTA step times (ms per control step of one resident CTA).
  // Shared model: 128 threads 0.92, 256 threads 1.10, 384 threads 1.28 -- large batches want the 384-thread CTA (lowest
  // cost per env), small ones (e.g. the 4096-env parity configuration) smaller CTAs that spread over more SMs.
  // Per-env models (the JB_PER_ENV_BUILD instantiation): 128 threads 0.78, 192 threads 1.05, 256 threads 1.31.
  const int cand[2][3] = {{384, 256, 128}, {JB_BLOCK, 192, 128}};
  const float tcta[2][3] = {{1.28f, 1.10f, 0.92f}, {1.31f * JB_BLOCK / 256, 1.05f, 0.78f}};
  const int m = per_env ? 1 : 0;
  int threads = 128;
  float best = 1e30f;
  for (int i = 0; i < 3; i++) {
    if (cand[m][i] > JB_BLOCK) continue;
    const int epb = cand[m][i] / 4;
    const int blocks = (N + epb - 1) / epb;
    const float cost = (float)((blocks + sms - 1) / sms) * tcta[m][i];
    if (cost < best * 0.999f) { best = cost; threads = cand[m][i]; }
  }
  return threads;
}

cudaError_t JB_LAUNCH_SIM(const JbDev& d, const JbLaunchArgs& a, cudaStream_t st) {
  const int threads = jb_pick_block(d.N, d.per_env);
  const int epb = threads / 4;
  const size_t smem = (d.per_env ? (size_t)epb * JBP_ENV_STRIDE : (size_t)JBP_SIZE) * sizeof(float);
  const int blocks = (d.N + epb - 1) / epb;
  JB_SIM_KERNEL<<<blocks, threads, smem, st>>>(d, a);
  return cudaGetLastError();
}

cudaError_t JB_CONFIGURE(void) {
  // per-env models stage JB_ENVS_PER_BLOCK packed models per CTA; opt in to as much dynamic shared memory as that needs
  // (CTAs larger than 496 threads exceed the 227 KB limit in per-env mode and are rejected at launch)
  size_t need = (size_t)JB_ENVS_PER_BLOCK * JBP_ENV_STRIDE * sizeof(float);
  if (need > 227 * 1024) need = 227 * 1024;
#ifdef JB_CARVEOUT_L1
  cudaFuncSetAttribute(JB_SIM_KERNEL, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
#endif
  return cudaFuncSetAttribute(JB_SIM_KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need);
}
