// Load-sorted slot order for the step kernel (sm_100a).
//
// The step kernel advances a warp -- and, through its phase barriers, a CTA -- at the pace of its heaviest env: the
// per-contact loops run as many trips as the fullest lane of the warp has contacts, and a warp whose quad needs a second
// Newton trip keeps the other eleven warps of the CTA at the next barrier.  Heavy envs (a leg standing on the rim of its
// lower cylinder: 3-4 contacts in one lane) are few (~5 %) and persistent from one control step to the next, so the
// envs are handed to the grid's quads in the order of a small integer key computed from the PREVIOUS control step's
// load record (JbDev::load, written by the step kernel): light envs share warps and CTAs with light envs, the heavy
// ones are packed together at the end of the grid.  Results per env do not depend on the order (every quad's arithmetic
// is its own; the warp-uniform loops only add zero-length steps), so this is purely a scheduling decision.
//
// The sort is a stable counting sort of the env ids (keys in [0, JB_SORT_BUCKETS)): within a bucket envs stay in
// ascending env id, which keeps the SoA state accesses of neighbouring quads in the same 32-byte sectors.  Three small
// launches per control step (histogram, scan of JB_SORT_BUCKETS x n_tiles counts, scatter), ~10 us at 1 Mi envs next to
// an ~80 ms step.
#include "jb_sim_common.cuh"
#include <stdlib.h>

#define JB_SORT_THREADS 256
#define JB_SORT_ROUNDS 8
#define JB_SORT_TILE (JB_SORT_THREADS * JB_SORT_ROUNDS)

// key of an env from its load record (x: contacts summed over the sub-steps in the fullest lane, y / z: sub-steps in
// which some lane held >= 3 / >= 2 contacts, w: Newton trips of those sub-steps + 65536 * contacts in the fullest lane at
// the latest sub-step).  `variant`: tuning (JBSIM_SORT_KEY).
__device__ __forceinline__ int jb_load_key(int4 ld, int n_sub, int variant) {
  const int trips = ld.w & 0xffff, last = (ld.w >> 16) & 15;
  const int extra = max(trips - n_sub, 0);                   // Newton trips beyond one per sub-step
  if (variant == 1) return min(last, 4) * 3 + min(extra, 2);
  if (variant == 5) return min(last, 3) * 4 + min(extra, 3);
  if (variant == 6) return min(extra, 3) * 4 + min(last, 3);
  if (variant == 2) return min((ld.x * 4 + n_sub / 2) / n_sub, 15);
  if (variant == 3) return min((ld.x * 2 + n_sub / 2) / n_sub, 6) * 2 + (extra > 0 ? 1 : 0);
  if (variant == 4) {
    const int h = max(min(last, 4), ld.y > 0 ? 3 : (ld.z > 0 ? 2 : (ld.x > 0 ? 1 : 0)));
    return h * 3 + min(extra, 2);
  }
  if (ld.x == 0) return 0;                                   // airborne for all of them
  if (ld.z == 0) return extra == 0 ? 1 : 2;                  // never more than one contact per lane
  if (ld.y == 0) return ld.z <= n_sub / 4 ? 3 : (extra <= 2 ? 4 : 5);
  if (ld.y <= n_sub / 4) return 6;
  return ld.y <= (3 * n_sub) / 4 ? 7 : 8;
}

__global__ void __launch_bounds__(JB_SORT_THREADS)
jb_sort_hist_kernel(const int4* __restrict__ load, int N, int n_sub, int n_tiles, int* __restrict__ hist, int desc, int variant) {
  __shared__ int cnt[JB_SORT_BUCKETS];
  if (threadIdx.x < JB_SORT_BUCKETS) cnt[threadIdx.x] = 0;
  __syncthreads();
  const int base = blockIdx.x * JB_SORT_TILE;
#pragma unroll
  for (int j = 0; j < JB_SORT_ROUNDS; j++) {
    const int e = base + j * JB_SORT_THREADS + threadIdx.x;
    const int key = e < N ? (desc ? JB_SORT_BUCKETS - 1 - jb_load_key(load[e], n_sub, variant) : jb_load_key(load[e], n_sub, variant)) : -1;
    // one shared-memory atomic per (warp, key) instead of one per thread
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    if (key >= 0 && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&cnt[key], __popc(peers));
  }
  __syncthreads();
  if (threadIdx.x < JB_SORT_BUCKETS) hist[threadIdx.x * n_tiles + blockIdx.x] = cnt[threadIdx.x];
}

// exclusive scan of hist[JB_SORT_BUCKETS * n_tiles] (bucket-major) in place, one CTA
__global__ void __launch_bounds__(1024) jb_sort_scan_kernel(int* __restrict__ hist, int n) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < n ? hist[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) warp_tot[w] = x;
    __syncthreads();
    if (w == 0) {
      int t = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
      warp_tot[lane] = t;                                   // inclusive over warps
    }
    __syncthreads();
    const int carry = carry_s;
    const int incl = x + (w > 0 ? warp_tot[w - 1] : 0) + carry;
    if (i < n) hist[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = incl;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(JB_SORT_THREADS)
jb_sort_scatter_kernel(const int4* __restrict__ load, int N, int n_sub, int n_tiles, const int* __restrict__ offs,
                       int* __restrict__ perm, int desc, int variant) {
  __shared__ int run[JB_SORT_BUCKETS];                          // next free position of each bucket for this tile
  __shared__ int cnt[JB_SORT_THREADS / 32][JB_SORT_BUCKETS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x < JB_SORT_BUCKETS) run[threadIdx.x] = offs[threadIdx.x * n_tiles + blockIdx.x];
  const int base = blockIdx.x * JB_SORT_TILE;
  for (int j = 0; j < JB_SORT_ROUNDS; j++) {
    if (threadIdx.x < JB_SORT_BUCKETS) {
#pragma unroll
      for (int k = 0; k < JB_SORT_THREADS / 32; k++) cnt[k][threadIdx.x] = 0;
    }
    __syncthreads();
    const int e = base + j * JB_SORT_THREADS + threadIdx.x;
    const int key = e < N ? (desc ? JB_SORT_BUCKETS - 1 - jb_load_key(load[e], n_sub, variant) : jb_load_key(load[e], n_sub, variant)) : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    if (key >= 0 && rank == 0) cnt[w][key] = __popc(peers);
    __syncthreads();
    if (key >= 0) {
      int pos = run[key] + rank;
      for (int k = 0; k < w; k++) pos += cnt[k][key];
      perm[pos] = e;
    }
    __syncthreads();
    if (threadIdx.x < JB_SORT_BUCKETS) {
      int t = 0;
#pragma unroll
      for (int k = 0; k < JB_SORT_THREADS / 32; k++) t += cnt[k][threadIdx.x];
      run[threadIdx.x] += t;
    }
  }
}

int jb_sort_scratch_ints(int N) { return JB_SORT_BUCKETS * ((N + JB_SORT_TILE - 1) / JB_SORT_TILE); }

// perm_out[slot] = env, envs in ascending (key, env id); scratch: jb_sort_scratch_ints(N) ints.  3 launches on `st`.
cudaError_t jb_sort_by_load(const int4* load, int N, int n_sub, int* perm_out, int* scratch, cudaStream_t st) {
  const int n_tiles = (N + JB_SORT_TILE - 1) / JB_SORT_TILE;
  static int desc = -1;                      // heaviest first: the grid's last wave is made of light CTAs (JBSIM_SORT_ASC=1: tuning)
  static int variant = 1;                    // measured best of the five keys (profiles/README.md, round 2)
  if (desc < 0) {
    const char* e = getenv("JBSIM_SORT_ASC"); desc = (e && atoi(e) == 1) ? 0 : 1;
    if ((e = getenv("JBSIM_SORT_KEY"))) variant = atoi(e);
  }
  jb_sort_hist_kernel<<<n_tiles, JB_SORT_THREADS, 0, st>>>(load, N, n_sub, n_tiles, scratch, desc, variant);
  jb_sort_scan_kernel<<<1, 1024, 0, st>>>(scratch, JB_SORT_BUCKETS * n_tiles);
  jb_sort_scatter_kernel<<<n_tiles, JB_SORT_THREADS, 0, st>>>(load, N, n_sub, n_tiles, scratch, perm_out, desc, variant);
  return cudaGetLastError();
}
