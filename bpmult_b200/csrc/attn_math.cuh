// Math helpers shared by the tensor-core attention kernels (attn_tc.cu: head dim 32; attn_tc128.cu: head dim 64 / 128).
#pragma once
#include "tc_common.cuh"

#define LOG2E_F 1.4426950408889634f
#define LN2_F 0.6931471805599453f

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// packed fp32 pairs (sm_100 FFMA2 / FADD2): one issue slot for two lanes of work
__device__ __forceinline__ void ffma2(float& a0, float& a1, float b, float c) {
  asm("{.reg .b64 x, y, z;\n mov.b64 x, {%0, %1};\n mov.b64 y, {%2, %2};\n mov.b64 z, {%3, %3};\n fma.rn.f32x2 x, x, y, z;\n mov.b64 {%0, %1}, x;}"
      : "+f"(a0), "+f"(a1) : "f"(b), "f"(c));
}
__device__ __forceinline__ void fadd2(float& a0, float& a1, float b0, float b1) {
  asm("{.reg .b64 x, y;\n mov.b64 x, {%0, %1};\n mov.b64 y, {%2, %3};\n add.rn.f32x2 x, x, y;\n mov.b64 {%0, %1}, x;}"
      : "+f"(a0), "+f"(a1) : "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fmul2(float& a0, float& a1, float b) {
  asm("{.reg .b64 x, y;\n mov.b64 x, {%0, %1};\n mov.b64 y, {%2, %2};\n mul.rn.f32x2 x, x, y;\n mov.b64 {%0, %1}, x;}" : "+f"(a0), "+f"(a1) : "f"(b));
}
__device__ __forceinline__ void fmul2v(float& a0, float& a1, float b0, float b1) {
  asm("{.reg .b64 x, y;\n mov.b64 x, {%0, %1};\n mov.b64 y, {%2, %3};\n mul.rn.f32x2 x, x, y;\n mov.b64 {%0, %1}, x;}" : "+f"(a0), "+f"(a1) : "f"(b0), "f"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *(uint32_t*)&v;
}

