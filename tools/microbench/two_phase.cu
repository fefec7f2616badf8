// Micro-benchmark 3: two-phase trajectory write-out through an L2-resident scratch.
//   phase A (per 32-step tile): every thread (= env) appends its 2 leaf values per step to a CTA-private scratch laid
//                               out [slot][leaf][env] -> 512-byte coalesced runs
//   phase B (every G tiles)   : the CTA reads its scratch back (coalesced, hopefully L2 hits), transposes through smem
//                               and writes final rows: leaves in runs of G*32*4 bytes, obs in runs of G*32*8 bytes
// compared with mode 0 = direct 32-slot tiles (128-byte leaf segments, 256-byte obs segments) from smem-like staging.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int E = 128, K = 32;

__global__ void __launch_bounds__(E) direct(float *l0, float *l1, float *obs, long n_save, int n_tiles) {
  __shared__ float st[K * 2][E + 1];
  const int e = threadIdx.x, lane = e & 31, warp = e >> 5;
  const long blk0 = (long)blockIdx.x * E;
  for (int t = 0; t < n_tiles; ++t) {
    for (int k = 0; k < K; ++k) { st[2 * k][e] = k + t; st[2 * k + 1][e] = k - t; }
    __syncthreads();
    for (int r = warp; r < E; r += 4) {
      const long row = (blk0 + r) * n_save + (long)t * K;
      l0[row + lane] = st[2 * lane][r];
      l1[row + lane] = st[2 * lane + 1][r];
      obs[row * 2 + lane] = st[lane][r] * 0.5f;
      obs[row * 2 + 32 + lane] = st[32 + lane][r] * 0.5f;
    }
    __syncthreads();
  }
}

template <int G>
__global__ void __launch_bounds__(E) two_phase(float *l0, float *l1, float *obs, float *scratch, long n_save, int n_tiles) {
  __shared__ float st[32][G * K + 1];  // one warp-block of 32 envs x (G*K) slots for one leaf at a time
  const int e = threadIdx.x, lane = e & 31, warp = e >> 5;
  const long blk0 = (long)blockIdx.x * E;
  float *my = scratch + (size_t)blockIdx.x * (G * K * 2 * E);  // [slot][leaf][env]
  for (int t0 = 0; t0 < n_tiles; t0 += G) {
    // phase A: G tiles of K steps, coalesced scratch writes
    for (int s = 0; s < G * K; ++s) {
      my[(s * 2 + 0) * E + e] = s + t0;
      my[(s * 2 + 1) * E + e] = s - t0;
    }
    __syncthreads();
    // phase B: per leaf, per 32-env block: scratch -> smem -> rows
    for (int leaf = 0; leaf < 2; ++leaf) {
      float *dstleaf = leaf ? l1 : l0;
      for (int wb = 0; wb < 4; ++wb) {  // 32-env blocks
        // all 4 warps cooperate: warp w loads slots w, w+4, ...
        for (int s = warp; s < G * K; s += 4) st[lane][s] = my[(s * 2 + leaf) * E + wb * 32 + lane];
        __syncthreads();
        for (int r = warp; r < 32; r += 4) {
          const long row = (blk0 + wb * 32 + r) * n_save + (long)t0 * K;
          for (int s = lane; s < G * K; s += 32) {
            const float v = st[r][s];
            dstleaf[row + s] = v;
            obs[(row + s) * 2 + leaf] = v * 0.5f;   // NB: strided by 2 -> written as two half-used passes (pessimistic)
          }
        }
        __syncthreads();
      }
    }
  }
}

template <int G>
void run2(float *l0, float *l1, float *obs, float *scratch, long B, long n_save, int n_tiles) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int it = 0; it < 2; ++it) {
    if (it == 1) cudaEventRecord(e0);
    two_phase<G><<<(int)(B / E), E>>>(l0, l1, obs, scratch, n_save, n_tiles);
  }
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double steps = (double)B * n_tiles * K;
  printf("two-phase G=%d (leaf runs %4d B): %.2f ms  %.3e env-steps/s  final-output %.0f GB/s  err=%s\n", G, G * K * 4, ms,
         steps / ms * 1e3, steps * 16 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const long B = 1 << 19, n_save = 2001;
  const int n_tiles = 60;  // 1920 steps
  float *l0, *l1, *obs, *scratch;
  cudaMalloc(&l0, B * n_save * 4);
  cudaMalloc(&l1, B * n_save * 4);
  cudaMalloc(&obs, B * n_save * 8);
  cudaMalloc(&scratch, (size_t)(B / E) * 4 * K * 2 * E * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int it = 0; it < 2; ++it) {
    if (it == 1) cudaEventRecord(e0);
    direct<<<(int)(B / E), E>>>(l0, l1, obs, n_save, n_tiles);
  }
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double steps = (double)B * n_tiles * K;
  printf("direct (leaf runs 128 B)          : %.2f ms  %.3e env-steps/s  final-output %.0f GB/s\n", ms, steps / ms * 1e3,
         steps * 16 / ms / 1e6);
  run2<1>(l0, l1, obs, scratch, B, n_save, n_tiles);
  run2<2>(l0, l1, obs, scratch, B, n_save, n_tiles);
  run2<4>(l0, l1, obs, scratch, B, n_save, n_tiles);
  return 0;
}
