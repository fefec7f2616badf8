// Micro-benchmark: achievable HBM write/read bandwidth for the sim_ahead access pattern -- every CTA owns 128 rows of
// a (B, N) float array and walks along the rows in tiles of L floats; each warp-wide access covers one contiguous
// L*4-byte (<=128 B per instruction) row segment, rows are `pitch` floats apart.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void write_rows(float *out, long pitch, int n_tiles, int L, int rows_per_cta) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const long row0 = (long)blockIdx.x * rows_per_cta;
  for (int t = 0; t < n_tiles; ++t) {
    for (int r = warp; r < rows_per_cta; r += nw) {
      float *dst = out + (row0 + r) * pitch + (long)t * L;
      for (int k = lane; k < L; k += 32) dst[k] = (float)(t + k);
    }
    __syncthreads();
  }
}
__global__ void read_rows(const float *in, float *sink, long pitch, int n_tiles, int L, int rows_per_cta) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const long row0 = (long)blockIdx.x * rows_per_cta;
  float acc = 0.f;
  for (int t = 0; t < n_tiles; ++t) {
    for (int r = warp; r < rows_per_cta; r += nw) {
      const float *src = in + (row0 + r) * pitch + (long)t * L;
      for (int k = lane; k < L; k += 32) acc += __ldg(src + k);
    }
  }
  if (acc == 1234.5f) sink[0] = acc;
}

int main() {
  const long B = 1 << 20;
  const int Ntot = 2048;
  float *buf, *sink;
  cudaMalloc(&sink, 4);
  for (long pitch : {2048L, 2001L}) {
    cudaMalloc(&buf, B * pitch * 4);
    for (int L : {16, 32, 64, 128, 256, 512}) {
      const int n_tiles = (int)((pitch < Ntot ? pitch : Ntot) / L);
      for (int mode = 0; mode < 2; ++mode) {
        for (int occ : {4, 8}) {
          cudaEvent_t e0, e1;
          cudaEventCreate(&e0);
          cudaEventCreate(&e1);
          const int rows_per_cta = 128;
          const int grid = (int)(B / rows_per_cta);
          // occupancy is steered with dynamic smem: 227 KB / occ per CTA
          const int smem = 200 * 1024 / occ;
          cudaFuncSetAttribute(write_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
          cudaFuncSetAttribute(read_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
          for (int it = 0; it < 2; ++it) {
            if (it == 1) cudaEventRecord(e0);
            if (mode == 0) write_rows<<<grid, 128, smem>>>(buf, pitch, n_tiles, L, rows_per_cta);
            else read_rows<<<grid, 128, smem>>>(buf, sink, pitch, n_tiles, L, rows_per_cta);
          }
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          const double bytes = (double)B * n_tiles * L * 4;
          printf("%s pitch=%ld L=%3d floats (%4d B) ctas/SM=%d : %.2f ms  %.0f GB/s\n", mode ? "read " : "write", pitch, L, L * 4,
                 occ, ms, bytes / ms / 1e6);
        }
      }
    }
    cudaFree(buf);
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
