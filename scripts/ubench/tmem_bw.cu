// Microbenchmark: tcgen05.ld / tcgen05.st throughput per SM vs number of warps (diagnostics only).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../bpmult_b200/csrc/tc_common.cuh"
void bpm_set_error(const char*, ...) {}

__global__ void __launch_bounds__(512, 1) tmem_bw(int iters, int mode, long long* out, float* sink) {
  __shared__ uint32_t tptr;
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tptr), 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tptr;
  const int warp = threadIdx.x >> 5;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t col0 = (uint32_t)((warp >> 2) * 64) & 511u;
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    float v[32];
    if (mode == 0) {            // 32 columns per load, wait after each
      tmem_ld32(tmem + lane_off + col0 + (i & 1) * 32, v);
      tmem_ld_wait();
      acc += v[0] + v[31];
    } else if (mode == 1) {     // two loads in flight
      float w[32];
      tmem_ld32(tmem + lane_off + col0, v);
      tmem_ld32(tmem + lane_off + col0 + 32, w);
      tmem_ld_wait();
      acc += v[0] + w[31];
    } else {                    // stores of 16 columns
      uint32_t r[16];
#pragma unroll
      for (int u = 0; u < 16; u++) r[u] = i + u;
      tmem_st16(tmem + lane_off + col0 + (i & 3) * 16, r);
      tmem_st_wait();
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long* out; cudaMalloc(&out, 16);
  float* sink; cudaMalloc(&sink, 16);
  const int iters = 2000;
  for (int mode = 0; mode < 3; mode++)
    for (int warps : {1, 4, 8, 16}) {
      tmem_bw<<<148, warps * 32>>>(iters, mode, out, sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
      double bytes = (double)iters * warps * 32 * (mode == 0 ? 32 : (mode == 1 ? 64 : 16)) * 4;
      printf("mode %d (%s) warps %2d: %8lld clk  -> %6.1f B/clk/SM  (%5.1f clk per warp-instr)  %s\n", mode,
             mode == 0 ? "ld x32 serial" : (mode == 1 ? "ld 2 x32 in flight" : "st x16"), warps, h, bytes / h,
             (double)h / (iters * (mode == 1 ? 2 : 1)), e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
