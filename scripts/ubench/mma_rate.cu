// Microbenchmark: tcgen05.mma issue / execution rate as a function of N and operand source (diagnostics only).
// One CTA per SM; one thread issues L MMAs (K = 16 each) back to back, commits, and waits; cycles are measured with clock64.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../bpmult_b200/csrc/tc_common.cuh"
void bpm_set_error(const char*, ...) {}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__global__ void __launch_bounds__(128, 1) mma_rate(int N, int L, int mode, int kstep_bytes, int nacc, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const uint32_t bar_a = smem_u32(&bar);
  if (threadIdx.x == 0) { mbar_init(bar_a, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tptr), 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tptr;
  if (threadIdx.x < 32) {
    // mode 0: A smem K-major SW128, B smem K-major SW128 (GEMM);  1: A from TMEM, B smem K-major;  2: A smem, B MN-major SW64 (attention acc)
    // 3: A TMEM, B MN-major SW64;  4: A, B K-major SW64 (attention S^T)
    const uint64_t dA = (mode == 4 ? umma_desc(base, 16, 512, BPM_SWZ_64B) : umma_desc(base, 16, 1024, BPM_SWZ_128B));
    const uint32_t boff = 65536;
    const uint64_t dB = (mode == 2 || mode == 3) ? umma_desc(base + boff, 512, 512, BPM_SWZ_64B)
                        : (mode == 4 ? umma_desc(base + boff, 16, 512, BPM_SWZ_64B) : umma_desc(base + boff, 16, 1024, BPM_SWZ_128B));
    const uint32_t idesc = umma_idesc_bf16(128, N, 0, (mode == 2 || mode == 3) ? 1 : 0);
    const bool ts = (mode == 1 || mode == 3);
    long long t0 = clock64();
    if (elect_one()) {
      for (int it = 0; it < L / 8; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const uint64_t ko = (uint64_t)(((u & 3) * kstep_bytes) >> 4);
          // nacc independent accumulators (compile-time pattern: u % 4 -> up to 4)
          const uint32_t d = tmem + (uint32_t)((nacc == 1 ? 0 : (nacc == 2 ? (u & 1) : (u & 3))) * N);
          if (ts) umma_bf16_ts(d, tmem + 384 + (u & 3) * 8, dB + ko, idesc, (it > 0) ? 1u : 0u);
          else umma_bf16(d, dA + ko, dB + ko, idesc, (it > 0) ? 1u : 0u);
        }
      }
    }
    __syncwarp();
    long long t1 = clock64();
    if (elect_one()) umma_commit(bar_a);
    __syncwarp();
    mbar_wait(bar_a, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long* out; cudaMalloc(&out, 16);
  cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int sms = 148;
  const char* names[] = {"SS Kmaj SW128 (GEMM)", "TS, B Kmaj SW128", "SS, B MNmaj SW64 (attn acc)", "TS, B MNmaj SW64", "SS Kmaj SW64 (attn S^T)"};
  for (int mode : {0, 1, 2, 3})
    for (int N : {32, 64, 128, 256}) {
      if (mode >= 2 && N > 64) continue;
      for (int nacc : {1, 2, 4}) {
        if (nacc * N > 384) continue;
        int L = 256;
        mma_rate<<<sms, 128, 200 * 1024>>>(N, L, mode, mode >= 2 ? 1024 : 32, nacc, out);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        printf("%-24s N=%3d nacc=%d: issue %5.1f clk/MMA  complete %5.1f clk/MMA  %s\n", names[mode], N, nacc, (double)h[0] / L, (double)h[1] / L,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
    }
  return 0;
}
