// Microbenchmark: tcgen05.mma issue / execution rate as a function of N and operand source (diagnostics only).
// One CTA per SM; one thread issues L MMAs (K = 16 each) back to back, commits, and waits; cycles are measured with clock64.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../bpmult_b200/csrc/tc_common.cuh"
void bpm_set_error(const char*, ...) {}

__global__ void __launch_bounds__(128, 1) mma_rate(int N, int L, int mode, int kstep_bytes, int nacc, int batch, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tptr;
  const uint32_t bar_a = smem_u32(&bar[0]);
  if (threadIdx.x == 0) { mbar_init(bar_a, 1); mbar_init(bar_a + 8, 1); mbar_fence_init(); }
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
    unsigned long long g0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    long long t0 = clock64();
    if (batch == 8) {
      if (elect_one()) {
        for (int it = 0; it < L / 8; it++) {
#pragma unroll
          for (int u = 0; u < 8; u++) {
            const uint64_t ko = (uint64_t)(((u & 3) * kstep_bytes) >> 4);
            if (ts) umma_bf16_ts(tmem, tmem + 384 + (u & 3) * 8, dB + ko, idesc, (it > 0) ? 1u : 0u);
            else umma_bf16(tmem, dA + ko, dB + ko, idesc, (it > 0) ? 1u : 0u);
          }
        }
      }
    } else {
      // GEMM-like issue loop: per k-block one elect_one block with 4 MMAs whose descriptors are re-formed from a stage address,
      // followed by __syncwarp (batch == 4), optionally with a commit per k-block (batch == 5)
      int st = 0;
      for (int it = 0; it < L / 4; it++) {
        const uint32_t sa = base + st * 49152;
        uint64_t da = (dA & ~0x3FFFull) | (uint64_t)((sa & 0x3FFFFu) >> 4), db = (dB & ~0x3FFFull) | (uint64_t)(((sa + 16384) & 0x3FFFFu) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; k++) {
            umma_bf16(tmem, da, db, idesc, (uint32_t)(it > 0) | (uint32_t)k);
            da += 2; db += 2;
          }
          if (batch == 5) umma_commit(bar_a + 8);
        }
        __syncwarp();
        if (++st == 3) st = 0;
      }
    }
    __syncwarp();
    long long t1 = clock64();
    if (elect_one()) umma_commit(bar_a);
    __syncwarp();
    mbar_wait(bar_a, 0);
    long long t2 = clock64();
    unsigned long long g1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; out[2] = (long long)(g1 - g0); }
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long* out; cudaMalloc(&out, 32);
  cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int sms = 148;
  const char* names[] = {"SS Kmaj SW128 (GEMM)", "TS, B Kmaj SW128", "SS, B MNmaj SW64 (attn acc)", "TS, B MNmaj SW64", "SS Kmaj SW64 (attn S^T)"};
  for (int batch : {8, 4, 5})
    for (int N : {192, 256}) {
      int L = 8192, mode = 0, nacc = 1;
      mma_rate<<<sms, 128, 200 * 1024>>>(N, L, mode, 32, nacc, batch, out);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[3]; cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost);
      printf("batch %d N=%3d: issue %5.1f clk/MMA  complete %5.1f clk/MMA  %.2f GHz  %s\n", batch, N, (double)h[0] / L, (double)h[1] / L,
             (double)h[1] / (double)h[2], e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
