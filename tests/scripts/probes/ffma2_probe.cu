// FP32 issue-rate probe: scalar FFMA vs packed FFMA2 (fma.rn.f32x2), 8 independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, int iters, float s) {
  float2 a[8], b = make_float2(s, s * 1.0001f), c = make_float2(0.5f, 0.25f);
  for (int i = 0; i < 8; i++) a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }
      else if (MODE == 1) a[i] = __ffma2_rn(a[i], b, c);
      else { a[i] = __ffma2_rn(a[i], b, c); a[i].x = __int_as_float(__float_as_int(a[i].x) ^ (it & 1)); }   // FFMA2 + 1 ALU op
    }
  }
  float r = 0;
  for (int i = 0; i < 8; i++) r += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, float* d) {
  int iters = 20000, blocks = 148 * 4;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, 512>>>(d, 100, 0.999f);
  cudaEventRecord(e0); k<MODE><<<blocks, 512>>>(d, iters, 0.999f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fma = (double)blocks * 512 * iters * 16;
  printf("%s: %.3f ms  %.1f TFLOP/s (2 flops per lane-FMA)\n", name, ms, 2 * fma / ms * 1e-9);
}
int main() { float* d; cudaMalloc(&d, 148 * 4 * 512 * 4); run<0>("scalar FFMA x2 ", d); run<1>("packed FFMA2    ", d); run<2>("FFMA2 + LOP3    ", d); return 0; }
