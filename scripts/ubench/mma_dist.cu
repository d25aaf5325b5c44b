// Microbenchmark: cost of a tcgen05.mma (SS, K-major SW128, M = 128, K = 16) as a function of the DISTANCE between the A and B tiles in
// shared memory and of N (diagnostics only).  One CTA per SM; one thread issues L MMAs back to back (4 k-slices of a 64-column tile).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../bpmult_b200/csrc/tc_common.cuh"
void bpm_set_error(const char*, ...) {}

__global__ void __launch_bounds__(128, 1) mma_dist(int N, int L, int aoff, int boff, int bmn, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tptr;
  const uint32_t bar_a = smem_u32(&bar[0]);
  if (threadIdx.x == 0) { mbar_init(bar_a, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tptr), 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tptr;
  if (threadIdx.x < 32) {
    const uint64_t dA = umma_desc(base + aoff, 16, 1024, BPM_SWZ_128B);
    const uint64_t dB = bmn ? umma_desc(base + boff, 16384, 1024, BPM_SWZ_128B) : umma_desc(base + boff, 16, 1024, BPM_SWZ_128B);
    const uint32_t idesc = umma_idesc_bf16(128, N, 0, bmn);
    long long t0 = clock64();
    if (elect_one()) {
      for (int it = 0; it < L / 8; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const uint64_t ka = (uint64_t)(((u & 3) * 32) >> 4), kb = bmn ? (uint64_t)(((u & 3) * 2048) >> 4) : ka;
          umma_bf16(tmem, dA + ka, dB + kb, idesc, (it > 0) ? 1u : 0u);
        }
      }
      umma_commit(bar_a);
    }
    __syncwarp();
    mbar_wait(bar_a, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t2 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long* out; cudaMalloc(&out, 32);
  cudaFuncSetAttribute(mma_dist, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  const int L = 4096;
  for (int bmn : {0, 1})
    for (int N : {64, 128})
      for (int aoff : {0, 98304})
        for (int boff : {16384, 32768, 49152, 65536, 81920, 98304, 114688, 131072, 163840, 196608}) {
          if (boff == aoff) continue;
          mma_dist<<<148, 128, 226 * 1024>>>(N, L, aoff, boff, bmn, out);
          cudaError_t e = cudaDeviceSynchronize();
          long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
          printf("B %s N=%3d  A at %3d KB  B at %3d KB: %6.1f clk/MMA %s\n", bmn ? "MN-major" : "K-major ", N, aoff / 1024, boff / 1024, (double)h / L,
                 e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
  return 0;
}
