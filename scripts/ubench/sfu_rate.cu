// Micro-benchmark: sustained ex2 rate of the softmax inner loop pattern (scale-subtract, ex2, row sum, bf16 pack) per SM,
// as a function of resident warps and of what surrounds the MUFU instructions.  One CTA per SM, clock64 timing.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sfu_rate sfu_rate.cu && ./sfu_rate
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void k(float* out, uint32_t* outp, long long* clk, int iters, float scale, float mneg) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; i++) v[i] = (float)(threadIdx.x + i) * 1e-3f;
  float r0 = 0, r1 = 0, r2 = 0, r3 = 0, mx = -1e30f;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    float s[32];
#pragma unroll
    for (int i = 0; i < 32; i++) s[i] = v[i] + (float)it;        // stand-in for the tcgen05.ld (1 FADD / element, all modes)
    if (MODE >= 3) {                                              // pass 1: row max
#pragma unroll
      for (int i = 0; i < 32; i++) mx = fmaxf(mx, s[i]);
    }
    if (MODE >= 1) {
#pragma unroll
      for (int i = 0; i < 32; i++) s[i] = fmaf(s[i], scale, mneg);
    }
#pragma unroll
    for (int i = 0; i < 32; i++) s[i] = ex2f(s[i]);
    if (MODE >= 1) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) { r0 += s[i]; r1 += s[i + 1]; r2 += s[i + 2]; r3 += s[i + 3]; }
    } else {
      r0 += s[0] + s[31];
    }
    if (MODE >= 2) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) { __nv_bfloat162 p = __floats2bfloat162_rn(s[i], s[i + 1]); acc ^= *(uint32_t*)&p; }
    } else {
#pragma unroll
      for (int i = 1; i < 31; i++) acc ^= __float_as_uint(s[i]) & (uint32_t)it;   // keep the values alive cheaply (LOP3)
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = r0 + r1 + r2 + r3 + mx;
  outp[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(int warps, const char* what) {
  float* out; uint32_t* outp; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&outp, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
  const int iters = 2000;
  k<MODE><<<148, warps * 32>>>(out, outp, clk, iters, 1.44f, -3.f);
  cudaDeviceSynchronize();
  k<MODE><<<148, warps * 32>>>(out, outp, clk, iters, 1.44f, -3.f);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  double c = (double)h[0];
  printf("%-44s warps/SM %2d: %7.1f clk per 32x32-warp chunk/SMSP-normalised,  %5.2f ex2/clk/SM\n", what, warps, c / iters,
         (double)iters * 32 * 32 * warps / c);
  cudaFree(out); cudaFree(outp); cudaFree(clk);
}

int main() {
  for (int w : {4, 8, 16}) {
    run<0>(w, "ex2 only (+1 FADD, +1 LOP3 / element)");
    run<1>(w, "+ scale-subtract FFMA, row-sum FADD");
    run<2>(w, "+ bf16 pack");
    run<3>(w, "+ row max (FMNMX)");
  }
  cudaError_t e = cudaGetLastError();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
