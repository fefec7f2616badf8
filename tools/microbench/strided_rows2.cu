// Micro-benchmark 2: misaligned rows (pitch 2001 floats) written with row segments whose boundaries are shifted per row
// so that they fall on 32-byte sector (align=8 floats) or 128-byte line (align=32 floats) boundaries of the global
// address space; align=1 is the plain tile-aligned pattern.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void write_rows(float *out, long pitch, int n_tiles, int L, int rows_per_cta, int align) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const long row0 = (long)blockIdx.x * rows_per_cta;
  for (int t = 0; t < n_tiles; ++t) {
    for (int r = warp; r < rows_per_cta; r += nw) {
      const long base = (row0 + r) * pitch;
      const int shift = (int)(base % align);  // elements by which this row's windows are moved back
      long lo = (long)t * L - shift, hi = (long)(t + 1) * L - shift;
      if (t == 0) lo = 0;
      if (t == n_tiles - 1) hi = (long)n_tiles * L;
      float *dst = out + base;
      for (long k = lo + lane; k < hi; k += 32) dst[k] = (float)k;
    }
    __syncthreads();
  }
}

int main() {
  const long B = 1 << 20;
  float *buf;
  for (long pitch : {2001L, 2008L, 2016L}) {
    cudaMalloc(&buf, B * pitch * 4);
    for (int L : {32, 64, 128}) {
      const int n_tiles = (int)(2000 / L);
      for (int align : {1, 8, 32}) {
        for (int occ : {4, 8}) {
          cudaEvent_t e0, e1;
          cudaEventCreate(&e0);
          cudaEventCreate(&e1);
          const int rows_per_cta = 128;
          const int grid = (int)(B / rows_per_cta);
          const int smem = 200 * 1024 / occ;
          cudaFuncSetAttribute(write_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
          for (int it = 0; it < 2; ++it) {
            if (it == 1) cudaEventRecord(e0);
            write_rows<<<grid, 128, smem>>>(buf, pitch, n_tiles, L, rows_per_cta, align);
          }
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          const double bytes = (double)B * n_tiles * L * 4;
          printf("write pitch=%ld L=%3d (%4d B) align=%2d ctas/SM=%d : %.2f ms  %.0f GB/s\n", pitch, L, L * 4, align, occ, ms,
                 bytes / ms / 1e6);
        }
      }
    }
    cudaFree(buf);
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
