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
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-(function, device) setting: remembered per device, so a second GPU used
// from the same process gets it too (api.cu)
int bpm_func_smem(const void* func, int bytes, const char* what);
int bpm_debug_get(int slot);   // diagnostic knobs (api.cu)
void* bpm_debug_get_ptr();     // optional device buffer for kernel event traces

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// A training step is ~3000 short kernels, each of which fills the GPU and depends on its predecessor.  Launched with the
// programmatic-stream-serialization attribute a kernel may START while its predecessor's last CTAs are still running: its CTAs take
// the SMs the predecessor frees and run their prologue (barrier init, TMEM allocation, descriptor prefetch) there, then block in
// pdl_wait() until the predecessor has completed and flushed its memory.  Every kernel calls pdl_trigger() first (lets ITS successor
// be scheduled) and pdl_wait() before its first access to global memory.  Without the attribute both are no-ops.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
int bpm_pdl_enabled();          // 0 when BPM_NO_PDL=1 (api.cu)

template <typename... KArgs, typename... Args>
static inline cudaError_t bpm_launch_cluster(int cluster, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                             Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (bpm_pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    n++;
  }
  if (cluster > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = cluster; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
    n++;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
static inline cudaError_t bpm_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  return bpm_launch_cluster(1, kernel, grid, block, smem, stream, static_cast<Args&&>(args)...);
}

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

// ---------------------------------------------------------------- counter-based dropout RNG
// Dropout needs one cheap, reproducible decision per element, regenerated identically in the backward pass -- not a
// cryptographic stream.  Philox4x32-10 costs ~15 instructions per element and made every GEMM epilogue with dropout RNG-bound
// (profiles/r01: fc1 100 us vs 19 us roofline), so decisions come from a multiply-xorshift-multiply hash of the element-pair
// counter (4 integer instructions per 32 random bits):
//   pair counter c = e >> 1;   x = lo32(c) * 0x7feb352d + (k0 ^ hi32(c) * 0x85ebca77);   x ^= x >> 16;   x = x * 0x846ca68b + k1
//   element e uses the low (e even) / high (e odd) 16 bits;  keep <=> half-word >= round(p * 65536);  kept values are scaled by 1/(1-p).
//   k0 / k1 are derived from (seed, site) with a full avalanche mix once per kernel.
struct DropCtx {
  uint32_t k0, k1;      // stream key (seed, site)
  uint32_t thresh;      // drop if half-word < thresh (0 .. 65536)
  uint32_t t16;         // thresh << 16 (saturated): "x >= t16" tests the high half-word, "(x << 16) >= t16" the low one
  float inv_keep;
  bool on;
  bool all;             // p >= 1: everything is dropped
};

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du;
  x ^= x >> 15; x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

__device__ __forceinline__ DropCtx make_drop(const bpm_dropout_t& d) {
  DropCtx c;
  c.on = d.p > 0.f;
  uint64_t seed = d.seed_ptr ? *d.seed_ptr : d.seed;
  c.k0 = mix32((uint32_t)seed ^ mix32((uint32_t)d.site + 0x9E3779B9u));
  c.k1 = mix32((uint32_t)(seed >> 32) ^ mix32((uint32_t)(d.site >> 32) + 0x85EBCA6Bu) ^ 0xC2B2AE35u);
  c.thresh = d.p >= 1.f ? 65536u : (uint32_t)((double)d.p * 65536.0 + 0.5);
  c.all = c.thresh >= 65536u;
  c.t16 = c.all ? 0xFFFFFFFFu : (c.thresh << 16);
  c.inv_keep = d.p < 1.f ? 1.f / (1.f - d.p) : 0.f;
  return c;
}

// 32 random bits for the element pair (2c, 2c+1)
__device__ __forceinline__ uint32_t drop_rand_pair(const DropCtx& c, uint64_t pair) {
  uint32_t x = (uint32_t)pair * 0x7feb352du + (c.k0 ^ ((uint32_t)(pair >> 32) * 0x85ebca77u));
  x ^= x >> 16;
  return x * 0x846ca68bu + c.k1;
}
// keep decisions of the pair's two elements from its 32 random bits
__device__ __forceinline__ bool drop_keep_lo(const DropCtx& c, uint32_t x) { return (x << 16) >= c.t16 && !c.all; }
__device__ __forceinline__ bool drop_keep_hi(const DropCtx& c, uint32_t x) { return x >= c.t16 && !c.all; }
// keep decisions of 8 consecutive elements 8*g .. 8*g+7 as a bit mask (bit i = element 8*g + i is kept)
__device__ __forceinline__ uint32_t drop_keep8(const DropCtx& c, uint64_t g) {
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const uint32_t r = drop_rand_pair(c, (g << 2) + i);
    m |= (drop_keep_lo(c, r) ? 1u : 0u) << (2 * i);
    m |= (drop_keep_hi(c, r) ? 1u : 0u) << (2 * i + 1);
  }
  return m;
}
// multiplier (0 or 1/(1-p)) for a single element index e
__device__ __forceinline__ float drop_mult1(const DropCtx& c, uint64_t e) {
  if (!c.on) return 1.f;
  const uint32_t r = drop_rand_pair(c, e >> 1);
  return ((e & 1) ? drop_keep_hi(c, r) : drop_keep_lo(c, r)) ? c.inv_keep : 0.f;
}
// multipliers for 8 consecutive elements starting at e (e % 8 == 0)
__device__ __forceinline__ void drop_mult8(const DropCtx& c, uint64_t e, float* m) {
  if (!c.on) {
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = 1.f;
    return;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const uint32_t r = drop_rand_pair(c, (e >> 1) + i);
    m[2 * i] = drop_keep_lo(c, r) ? c.inv_keep : 0.f;
    m[2 * i + 1] = drop_keep_hi(c, r) ? c.inv_keep : 0.f;
  }
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
