// Does the packed fp32x2 FMA of sm_100 (FFMA2) relieve an issue-bound fp32 kernel?  Three kernels, same flops:
//   scalar : 8 independent FFMA chains per thread
//   packed : 4 independent FFMA2 chains per thread (two fp32 FMAs per instruction)
//   mixed  : like the solver loop -- per 127 arithmetic instructions 51 non-arithmetic ones (integer adds / selects)
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float *out, int iters, float a, float b, int salt) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3f + i;
  int junk = salt;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int r = 0; r < 16; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);
    } else {
#pragma unroll
      for (int r = 0; r < 16; ++r)
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          float2 v = __ffma2_rn(make_float2(x[i], x[i + 1]), make_float2(a, a), make_float2(b, b));
          x[i] = v.x;
          x[i + 1] = v.y;
        }
    }
    if (MODE >= 2) {  // 51 / 127 * 128 = 51 integer instructions per 128 fp ops
#pragma unroll
      for (int r = 0; r < 51; ++r) junk = (junk ^ (junk >> 3)) + r;
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + junk;
}

template <int MODE>
void run(const char *name) {
  float *out;
  const int blocks = 148 * 8, threads = 256, iters = 4096;
  cudaMalloc(&out, blocks * threads * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(out, iters, 0.999f, 0.001f, 1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, iters, 0.999f, 0.001f, 1);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fma = (double)blocks * threads * iters * 128;
  printf("%-28s %.3f ms  %.2f TFMA/s  (%.1f fp32 FMA / clk / SM at 1.965 GHz)\n", name, ms, fma / ms * 1e-9,
         fma / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
}

int main() {
  run<0>("scalar FFMA");
  run<1>("packed FFMA2");
  run<2>("scalar FFMA + 40% integer");
  run<3>("packed FFMA2 + 40% integer");
  return 0;
}
