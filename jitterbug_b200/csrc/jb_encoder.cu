// Fused observation-encoder hook: obs[N, D] -> (normalise) -> small dense network on the first k features -> latent,
// remaining features passed through.  One launch per control step on the simulator's stream, thread per env,
// network parameters staged in shared memory; HBM-bound (reads D, writes latent + D - k floats per env).
// Follows the reference's Jitterbug.encode_obs (jitterbug_dmc/jitterbug.py:927-993) and the two encoder families it
// feeds: the tanh (denoising) auto-encoder's hidden layer (benchmarks/denoising_autoencoder.py:66-98, :195-205, :218-232)
// and the VAE encoder (benchmarks/VAE.py:13-46, :134-146).
#include <stdint.h>
#include <string>

#include "../../include/jbsim.h"
#include "jb_task.cuh"

int jb_set_error(const std::string& msg);   // jb_api.cu

#define JB_ENC_BLOCK 256

struct JbEncArgs {
  JbEncoderDesc d;
  const float* params;     // per ensemble member: for each layer W[in][out] row-major then b[out]
  const float* extrema;    // [obs_dim][2] (min, max) or NULL
  const float* obs;
  float* out;
  int n, obs_stride, out_stride, n_params;
  uint64_t seed, counter;
};

__device__ __forceinline__ float jb_enc_act(float x, int a) {
  return a == JB_ACT_TANH ? tanhf(x) : (a == JB_ACT_RELU ? fmaxf(x, 0.f) : x);
}

// standard normal from two Philox words (Box-Muller); u1 in (0, 1]
__device__ __forceinline__ float jb_enc_normal(uint32_t a, uint32_t b) {
  float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f), u2 = jb_u01(b);
  return sqrtf(-2.0f * logf(u1)) * cosf(6.28318530717958647692f * u2);
}

__global__ void __launch_bounds__(JB_ENC_BLOCK) jb_encode_kernel(const __grid_constant__ JbEncArgs a) {
  extern __shared__ float sm[];
  float* P = sm;                      // n_models * n_params
  float* X = sm + a.d.n_models * a.n_params;   // extrema, 2 * obs_dim
  for (int i = threadIdx.x; i < a.d.n_models * a.n_params; i += blockDim.x) P[i] = a.params[i];
  if (a.extrema)
    for (int i = threadIdx.x; i < 2 * a.d.obs_dim; i += blockDim.x) X[i] = a.extrema[i];
  __syncthreads();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.n) return;
  const JbEncoderDesc& d = a.d;
  float x[JB_ENC_MAX_WIDTH], y[JB_ENC_MAX_WIDTH], z[JB_ENC_MAX_WIDTH];
  const float* o = a.obs + (size_t)e * a.obs_stride;
  float* out = a.out + (size_t)e * a.out_stride;
  const int latent = d.dims[d.n_layers];
  for (int i = 0; i < d.obs_dim; i++) {
    float v = o[i];
    if (d.normalize == JB_ENC_NORM_PM1) v = -1.0f + 2.0f * (v - X[2 * i]) / (X[2 * i + 1] - X[2 * i]);
    else if (d.normalize == JB_ENC_NORM_01) v = (v - X[2 * i]) / (X[2 * i + 1] - X[2 * i]);
    if (i < d.k_encode) x[i] = v;
    else out[latent + i - d.k_encode] = v;                  // untouched features ride along (jitterbug.py:962-981)
  }
  for (int j = 0; j < latent; j++) z[j] = 0.f;
  for (int m = 0; m < d.n_models; m++) {
    const float* W = P + m * a.n_params;
    float in[JB_ENC_MAX_WIDTH];        // activations of the current layer; x is kept for the next ensemble member
    for (int i = 0; i < d.k_encode; i++) in[i] = x[i];
    for (int l = 0; l < d.n_layers; l++) {
      const int ni = d.dims[l], no = d.dims[l + 1];
      const bool heads = d.vae && l == d.n_layers - 1;      // (mean | std) heads share the input (VAE.py:29-30, :46)
      const int cols = heads ? 2 * no : no;
      const float* b = W + ni * cols;
      for (int j = 0; j < cols; j++) {
        float s = b[j];
        for (int i = 0; i < ni; i++) s = fmaf(in[i], W[i * cols + j], s);
        y[j] = heads ? s : jb_enc_act(s, d.act[l]);
      }
      if (heads) {
        // z = mean + std * eps with the std head used as is (VAE.py:143-145)
        for (int j = 0; j < no; j += 2) {
          JbPhilox r = jb_philox((uint32_t)e + a.d.env_id_offset, (uint32_t)a.counter, (uint32_t)(a.counter >> 32) ^ (uint32_t)(j / 2 + 64 * m),
                                 0x454E43u, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
          in[j] = y[j] + y[no + j] * jb_enc_normal(r.v[0], r.v[1]);
          if (j + 1 < no) in[j + 1] = y[j + 1] + y[no + j + 1] * jb_enc_normal(r.v[2], r.v[3]);
        }
      } else {
        for (int j = 0; j < no; j++) in[j] = y[j];
      }
      W = b + cols;
    }
    for (int j = 0; j < latent; j++) z[j] += in[j];
  }
  const float inv = 1.0f / (float)d.n_models;               // ensemble mean (jitterbug.py:983-991)
  for (int j = 0; j < latent; j++) out[j] = z[j] * inv;
}

extern "C" int jb_encoder_num_params(const JbEncoderDesc* d) {
  if (!d || d->n_layers < 1 || d->n_layers > JB_ENC_MAX_LAYERS) return -1;
  int n = 0;
  for (int l = 0; l < d->n_layers; l++) {
    int cols = d->dims[l + 1] * ((d->vae && l == d->n_layers - 1) ? 2 : 1);
    n += (d->dims[l] + 1) * cols;
  }
  return n;
}

extern "C" int jb_encoder_out_dim(const JbEncoderDesc* d) {
  if (!d || d->n_layers < 1 || d->n_layers > JB_ENC_MAX_LAYERS) return -1;
  return d->dims[d->n_layers] + d->obs_dim - d->k_encode;
}

extern "C" int jb_encode_obs(const JbEncoderDesc* d, const float* params_dev, const float* extrema_dev,
                             const float* obs_dev, int n, int obs_stride, float* out_dev, int out_stride,
                             uint64_t seed, uint64_t counter, void* stream) {
  if (!d || !params_dev || !obs_dev || !out_dev) return jb_set_error("jb_encode_obs: null argument");
  if (d->n_layers < 1 || d->n_layers > JB_ENC_MAX_LAYERS) return jb_set_error("jb_encode_obs: n_layers out of range");
  if (d->obs_dim < 1 || d->obs_dim > obs_stride || d->k_encode < 1 || d->k_encode > d->obs_dim)
    return jb_set_error("jb_encode_obs: need 1 <= k_encode <= obs_dim <= obs_stride");
  if (d->dims[0] != d->k_encode) return jb_set_error("jb_encode_obs: dims[0] must equal k_encode");
  for (int l = 0; l <= d->n_layers; l++) {
    int w = d->dims[l] * ((d->vae && l == d->n_layers) ? 2 : 1);
    if (d->dims[l] < 1 || w > JB_ENC_MAX_WIDTH) return jb_set_error("jb_encode_obs: layer width out of range");
  }
  for (int l = 0; l < d->n_layers; l++)
    if (d->act[l] < JB_ACT_LINEAR || d->act[l] > JB_ACT_RELU) return jb_set_error("jb_encode_obs: unknown activation");
  if (d->normalize != JB_ENC_NORM_NONE && !extrema_dev) return jb_set_error("jb_encode_obs: normalisation needs extrema");
  if (d->normalize < JB_ENC_NORM_NONE || d->normalize > JB_ENC_NORM_01) return jb_set_error("jb_encode_obs: unknown normalisation");
  if (d->n_models < 1) return jb_set_error("jb_encode_obs: n_models must be >= 1");
  if (out_stride < jb_encoder_out_dim(d)) return jb_set_error("jb_encode_obs: out_stride too small");
  if (n <= 0) return 0;
  JbEncArgs a;
  a.d = *d; a.params = params_dev; a.extrema = extrema_dev; a.obs = obs_dev; a.out = out_dev;
  a.n = n; a.obs_stride = obs_stride; a.out_stride = out_stride; a.n_params = jb_encoder_num_params(d);
  a.seed = seed; a.counter = counter;
  size_t smem = sizeof(float) * ((size_t)a.n_params * d->n_models + 2 * d->obs_dim);
  if (smem > 200 * 1024) return jb_set_error("jb_encode_obs: parameters exceed shared memory");
  cudaError_t e = cudaSuccess;
  if (smem > 48 * 1024) e = cudaFuncSetAttribute(jb_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) {
    jb_encode_kernel<<<(n + JB_ENC_BLOCK - 1) / JB_ENC_BLOCK, JB_ENC_BLOCK, smem, (cudaStream_t)stream>>>(a);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) return jb_set_error(std::string("jb_encode_obs: ") + cudaGetErrorString(e));
  return 0;
}
