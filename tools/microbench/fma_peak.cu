// Measured arithmetic / issue peaks of the GPU this runs on, as ONE JSON line (the roofline denominators bench.py uses
// for the kernels that are bound by the FMA pipe or by instruction issue; MEASURED_PEAKS.json only holds the HBM copy
// and the bf16 GEMM peaks).  Every figure is device-timed with CUDA events over a second launch (the first one warms up):
//   ffma_tflops   scalar fp32 FFMA, 8 independent chains per thread        (2 flop per FMA)
//   ffma2_tflops  packed fp32 FFMA2 (two FMAs per instruction), 4 chains
//   dfma_tflops   fp64 DFMA, 8 chains
//   issue_ginst   warp instructions per second of the scalar FFMA kernel: the FMA pipe accepts one warp instruction per
//                 scheduler per clock, which is also the issue ceiling (4 schedulers x SMs x clock)
// usage: fma_peak [device]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float *out, int iters, float a, float b, double da, double db, int salt) {
  float x[8];
  double d[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = threadIdx.x * 1e-3f + i;
    d[i] = threadIdx.x * 1e-3 + i;
  }
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int r = 0; r < 16; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);
    } else if (MODE == 1) {
#pragma unroll
      for (int r = 0; r < 16; ++r)
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          const float2 v = __ffma2_rn(make_float2(x[i], x[i + 1]), make_float2(a, a), make_float2(b, b));
          x[i] = v.x;
          x[i + 1] = v.y;
        }
    } else if (MODE == 2) {
#pragma unroll
      for (int r = 0; r < 16; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = fma(d[i], da, db);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i] + (float)d[i] + (float)salt;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static double run_ms(int sms, int iters) {
  float *out;
  const int blocks = sms * 8, threads = 256;
  cudaMalloc(&out, (size_t)blocks * threads * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(out, iters, 0.999f, 0.001f, 0.999, 0.001, 1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, 0.999f, 0.001f, 0.999, 0.001, 1);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaFree(out);
  return best;
}

int main(int argc, char **argv) {
  const int dev = argc > 1 ? atoi(argv[1]) : 0;
  if (cudaSetDevice(dev) != cudaSuccess) {
    printf("{\"error\": \"no CUDA device %d\"}\n", dev);
    return 1;
  }
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, dev);
  const int sms = pr.multiProcessorCount;
  const int iters = 2048;
  const double threads = (double)sms * 8 * 256;
  const double ms0 = run_ms<0>(sms, iters), ms1 = run_ms<1>(sms, iters), ms2 = run_ms<2>(sms, iters / 8);
  const double fma32 = threads * iters * 128, fma64 = threads * (iters / 8) * 128;
  const double inst0 = threads / 32 * iters * 128;  // warp instructions of the scalar FFMA kernel
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"ffma_tflops\": %.2f, \"ffma2_tflops\": %.2f, \"dfma_tflops\": %.3f, "
         "\"issue_ginst\": %.1f, \"how\": \"tools/microbench/fma_peak.cu: independent FMA chains, %d x 256 threads, CUDA "
         "events, best of 3\"}\n",
         pr.name, sms, 2 * fma32 / ms0 * 1e-9, 2 * fma32 / ms1 * 1e-9, 2 * fma64 / ms2 * 1e-9, inst0 / ms0 * 1e-6,
         sms * 8);
  return 0;
}
