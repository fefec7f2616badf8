// Micro-benchmark 4: write-out WITHOUT shared-memory staging -- every thread owns one environment row and appends to
// it with 16-byte stores (4 fp32 leaf samples per store, 2 observation pairs per store); the L2 merges the pieces of
// a 128-byte line before it is evicted.  Rows have the reference pitch (N+1 = 2001 elements, so only 4-byte aligned):
// a lane flushes its 4-entry register ring when its own element index reaches a 16-byte boundary.
#include <cstdio>
#include <cuda_runtime.h>

template <int WORK>
__global__ void __launch_bounds__(128) direct(float *l0, float *l1, float *obs, long n_save, int n_steps, long B) {
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float *r0 = l0 + b * n_save, *r1 = l1 + b * n_save, *ro = obs + b * n_save * 2;
  const int ph4 = (int)((b * n_save) & 3), ph2 = (int)((b * n_save) & 1);
  float a0 = 0, a1 = 0, a2 = 0, a3 = 0, c0 = 0, c1 = 0, c2 = 0, c3 = 0, o0 = 0, o1 = 0, o2 = 0, o3 = 0;
  float x = (float)b * 1e-6f, v = 0.5f;
  for (int s = 0; s < n_steps; ++s) {
#pragma unroll
    for (int w = 0; w < WORK; ++w) {  // stand-in for the solver arithmetic
      const float dv = -100.f * x - v;
      x = fmaf(1e-4f, v, x);
      v = fmaf(1e-4f, dv, v);
    }
    const int j = (ph4 + s) & 3;
    a0 = j == 0 ? x : a0; a1 = j == 1 ? x : a1; a2 = j == 2 ? x : a2; a3 = j == 3 ? x : a3;
    c0 = j == 0 ? v : c0; c1 = j == 1 ? v : c1; c2 = j == 2 ? v : c2; c3 = j == 3 ? v : c3;
    if (j == 3) {
      if (s >= 3) {
        *reinterpret_cast<float4 *>(r0 + s - 3) = make_float4(a0, a1, a2, a3);
        *reinterpret_cast<float4 *>(r1 + s - 3) = make_float4(c0, c1, c2, c3);
      } else {
        for (int i = 0; i <= s; ++i) { r0[i] = x; r1[i] = v; }
      }
    }
    const int j2 = (ph2 + s) & 1;
    if (j2 == 0) { o0 = x * 0.1f; o1 = v * 0.1f; } else { o2 = x * 0.1f; o3 = v * 0.1f; }
    if (j2 == 1) {
      if (s >= 1) *reinterpret_cast<float4 *>(ro + 2 * (s - 1)) = make_float4(o0, o1, o2, o3);
      else { ro[0] = o2; ro[1] = o3; }
    }
  }
  // tails
  const int j = (ph4 + n_steps - 1) & 3;
  if (j != 3) for (int i = 0; i <= j; ++i) { r0[n_steps - 1 - i] = x; r1[n_steps - 1 - i] = v; }
  if (((ph2 + n_steps - 1) & 1) == 0) { ro[2 * (n_steps - 1)] = o0; ro[2 * (n_steps - 1) + 1] = o1; }
}

template <int WORK>
void run(float *l0, float *l1, float *obs, long B, long n_save, int n_steps, int occ) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int smem = 200 * 1024 / occ;
  cudaFuncSetAttribute(direct<WORK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int it = 0; it < 2; ++it) {
    if (it == 1) cudaEventRecord(e0);
    direct<WORK><<<(int)(B / 128), 128, smem>>>(l0, l1, obs, n_save, n_steps, B);
  }
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double steps = (double)B * n_steps;
  printf("direct 16B stores WORK=%2d ctas/SM=%2d: %.2f ms  %.3e env-steps/s  output %.0f GB/s  err=%s\n", WORK, occ, ms,
         steps / ms * 1e3, steps * 16 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const long B = 1 << 20, n_save = 2001;
  const int n_steps = 2000;
  float *l0, *l1, *obs;
  cudaMalloc(&l0, B * n_save * 4);
  cudaMalloc(&l1, B * n_save * 4);
  cudaMalloc(&obs, B * n_save * 8);
  for (int occ : {4, 8, 12, 16}) run<1>(l0, l1, obs, B, n_save, n_steps, occ);
  for (int occ : {4, 8, 12, 16}) run<30>(l0, l1, obs, B, n_save, n_steps, occ);
  return 0;
}
