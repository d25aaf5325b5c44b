// Shared device/host helpers for the bpmult_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/bpmult_b200.h"

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------- host-side error plumbing
void bpm_set_error(const char* fmt, ...);
#define BPM_REQUIRE(cond, ...)                      \
  do {                                              \
    if (!(cond)) {                                  \
      bpm_set_error(__VA_ARGS__);                   \
      return BPM_EINVAL;                            \
    }                                               \
  } while (0)
#define BPM_CHECK_LAUNCH(name)                                               \
  do {                                                                       \
    cudaError_t e__ = cudaPeekAtLastError();                                 \
    if (e__ != cudaSuccess) {                                                \
      bpm_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      (void)cudaGetLastError();                                              \
      return BPM_ELAUNCH;                                                    \
    }                                                                        \
  } while (0)

static inline int bpm_cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
int bpm_num_sms();

// ---------------------------------------------------------------- dtype helpers
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float ld_as_f(const void* p, int dtype, int64_t i) {
  return dtype == BPM_BF16 ? __bfloat162float(((const bf16*)p)[i]) : ((const float*)p)[i];
}
__device__ __forceinline__ void st_from_f(void* p, int dtype, int64_t i, float v) {
  if (dtype == BPM_BF16) ((bf16*)p)[i] = __float2bfloat16_rn(v);
  else ((float*)p)[i] = v;
}

// 8-element vector access (row pitches are multiples of 8 elements and 16 B aligned by construction)
template <typename T> struct Vec8;
template <> struct Vec8<float> {
  float v[8];
  __device__ __forceinline__ void load(const float* p) {
    float4 a = *(const float4*)p, b = *(const float4*)(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    *(float4*)p = make_float4(v[0], v[1], v[2], v[3]);
    *(float4*)(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <> struct Vec8<bf16> {
  float v[8];
  __device__ __forceinline__ void load(const bf16* p) {
    uint4 r = *(const uint4*)p;
    const __nv_bfloat162* h = (const __nv_bfloat162*)&r;
#pragma unroll
    for (int i = 0; i < 4; i++) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  __device__ __forceinline__ void store(bf16* p) const {
    uint4 r;
    __nv_bfloat162* h = (__nv_bfloat162*)&r;
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *(uint4*)p = r;
  }
};

// ---------------------------------------------------------------- Philox4x32-10 dropout
// One Philox call yields 4 x 32 bit = 8 half-words = the keep decisions of 8 consecutive elements:
//   element e uses call counter e >> 3, half-word e & 7 (word (e & 7) >> 1, low half first); keep <=> hw >= round(p * 65536).
struct DropCtx {
  uint32_t k0, k1;      // key = seed
  uint32_t s0, s1;      // site (counter words 2,3)
  uint32_t thresh;      // drop if half-word < thresh (0 .. 65536)
  float inv_keep;
  bool on;
};

__device__ __forceinline__ DropCtx make_drop(const bpm_dropout_t& d) {
  DropCtx c;
  c.on = d.p > 0.f;
  uint64_t seed = d.seed_ptr ? *d.seed_ptr : d.seed;
  c.k0 = (uint32_t)seed; c.k1 = (uint32_t)(seed >> 32);
  c.s0 = (uint32_t)d.site; c.s1 = (uint32_t)(d.site >> 32);
  c.thresh = d.p >= 1.f ? 65536u : (uint32_t)((double)d.p * 65536.0 + 0.5);
  c.inv_keep = d.p < 1.f ? 1.f / (1.f - d.p) : 0.f;
  return c;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  return make_uint4(c0, c1, c2, c3);
}

// random half-words for elements 8*g .. 8*g+7
__device__ __forceinline__ uint4 drop_rand8(const DropCtx& c, uint64_t g) {
  return philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), c.s0, c.s1, c.k0, c.k1);
}
// keep decisions of 8 consecutive elements as a bit mask (bit i = element 8*g + i is kept)
__device__ __forceinline__ uint32_t drop_keep8(const DropCtx& c, uint64_t g) {
  uint4 r = drop_rand8(c, g);
  uint32_t m = 0;
  m |= ((r.x & 0xFFFFu) >= c.thresh) ? 1u : 0u;   m |= ((r.x >> 16) >= c.thresh) ? 2u : 0u;
  m |= ((r.y & 0xFFFFu) >= c.thresh) ? 4u : 0u;   m |= ((r.y >> 16) >= c.thresh) ? 8u : 0u;
  m |= ((r.z & 0xFFFFu) >= c.thresh) ? 16u : 0u;  m |= ((r.z >> 16) >= c.thresh) ? 32u : 0u;
  m |= ((r.w & 0xFFFFu) >= c.thresh) ? 64u : 0u;  m |= ((r.w >> 16) >= c.thresh) ? 128u : 0u;
  return m;
}
// multiplier (0 or 1/(1-p)) for a single element index e
__device__ __forceinline__ float drop_mult1(const DropCtx& c, uint64_t e) {
  if (!c.on) return 1.f;
  uint32_t m = drop_keep8(c, e >> 3);
  return ((m >> (uint32_t)(e & 7)) & 1u) ? c.inv_keep : 0.f;
}
// multipliers for 8 consecutive elements starting at e (e % 8 == 0)
__device__ __forceinline__ void drop_mult8(const DropCtx& c, uint64_t e, float* m) {
  if (!c.on) {
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = 1.f;
    return;
  }
  uint32_t k = drop_keep8(c, e >> 3);
#pragma unroll
  for (int i = 0; i < 8; i++) m[i] = ((k >> i) & 1u) ? c.inv_keep : 0.f;
}

// ---------------------------------------------------------------- warp / block reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
