// Micro-benchmark 5: tensor memory (TMEM) as the per-thread staging buffer of the trajectory write-out.
// Every thread (= environment, = one TMEM lane) appends its samples to its own TMEM lane with tcgen05.st, column =
// value * K + slot.  Flush (per warp, no CTA barrier): tcgen05.ld 32 columns (32 slots of one value) -> registers ->
// per-warp smem transpose buffer [slot][env] -> warp-wide stores with lanes along time.  K = 64 slots of 2 values use
// 128 TMEM columns per CTA (4 CTAs/SM fill the 512 columns) while shared memory only holds the 4.2 KB/warp transpose
// buffer, so row segments are twice as long as with 32-slot shared-memory tiles at the same occupancy.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int E = 128;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int K>
__global__ void __launch_bounds__(E) tmem_writeout(float *l0, float *l1, float *obs, long n_save, int n_tiles) {
  __shared__ uint32_t tmem_base_s;
  __shared__ float tb[4][2][32][33];
  const int e = threadIdx.x, lane = e & 31, warp = e >> 5;
  constexpr int COLS = 2 * K;  // power of two for K = 16, 32, 64, 128
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tbase = tmem_base_s + ((uint32_t)(warp * 32) << 16);
  const long blk0 = (long)blockIdx.x * E;
  float x = e * 1e-3f, v = 0.25f;
  for (int t = 0; t < n_tiles; ++t) {
    // "compute": K saves of 2 values, staged in this thread's TMEM lane
    for (int k = 0; k < K; ++k) {
      x = fmaf(1e-4f, v, x);
      v = fmaf(-1e-2f, x, v);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tbase + k), "r"(__float_as_uint(x)));
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tbase + K + k), "r"(__float_as_uint(v)));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;");
    // flush: this warp's 32 environments, 32 slots at a time, both values transposed before the row loop so that
    // the observation pair of a slot goes out as one 8-byte store (256 contiguous bytes per warp-wide store)
    for (int c0 = 0; c0 < K; c0 += 32) {
#pragma unroll
      for (int val = 0; val < 2; ++val) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
            "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(tbase + val * K + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
        for (int j = 0; j < 32; ++j) tb[warp][val][j][lane] = __uint_as_float(r[j]);
      }
      __syncwarp();
      const long col0 = (long)t * K + c0;
#pragma unroll 4
      for (int rr = 0; rr < 32; ++rr) {
        const float y0 = tb[warp][0][lane][rr], y1 = tb[warp][1][lane][rr];
        const long row = (blk0 + warp * 32 + rr) * n_save + col0 + lane;
        l0[row] = y0;
        l1[row] = y1;
        *reinterpret_cast<float2 *>(obs + row * 2) = make_float2(y0 * 0.5f, y1 * 0.5f);
      }
      __syncwarp();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "n"(COLS));
}

template <int K>
void run(float *l0, float *l1, float *obs, long B, long n_save) {
  const int n_tiles = 1920 / K;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int it = 0; it < 2; ++it) {
    if (it == 1) cudaEventRecord(e0);
    tmem_writeout<K><<<(int)(B / E), E>>>(l0, l1, obs, n_save, n_tiles);
  }
  cudaEventRecord(e1);
  cudaError_t err = cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double steps = (double)B * n_tiles * K;
  printf("TMEM staging K=%3d (leaf runs %4d B, %3d columns/CTA): %.2f ms  %.3e env-steps/s  output %.0f GB/s  err=%s\n", K,
         K * 4, 2 * K, ms, steps / ms * 1e3, steps * 16 / ms / 1e6, cudaGetErrorString(err));
}

int main() {
  const long B = 1 << 19, n_save = 2001;
  float *l0, *l1, *obs;
  cudaMalloc(&l0, B * n_save * 4);
  cudaMalloc(&l1, B * n_save * 4);
  cudaMalloc(&obs, B * n_save * 8);
  cudaMemset(l0, 0, B * n_save * 4);
  run<32>(l0, l1, obs, B, n_save);
  run<64>(l0, l1, obs, B, n_save);
  run<128>(l0, l1, obs, B, n_save);
  // spot check: row 1000, a few samples must be finite and increasing-ish
  float h[8];
  cudaMemcpy(h, l0 + 1000 * n_save + 60, sizeof(h), cudaMemcpyDeviceToHost);
  printf("samples: %g %g %g %g %g %g %g %g\n", h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
  printf("final err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
