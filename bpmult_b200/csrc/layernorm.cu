// LayerNorm forward / backward over the padded row layout (sm_100a).  HBM-bound: one warp per row, the row lives
// in registers (16-byte vector loads), statistics over the D real columns only, pad columns written as zero.
// Reference: nn.LayerNorm(embed_dim) eps 1e-5, models/transformer.py:197-202,227-229.
// Algorithmic bytes: rows*D*(e_in + e_out) fwd; rows*D*(e_dy + e_x + 4 [+4 accumulate read]) bwd.
#include "bpm_common.cuh"
#include "tc_common.cuh"

#define LN_WARPS 4

// Pad columns of x, gamma and beta are ZERO by contract (include/bpmult_b200.h), so the statistics need no per-element mask:
//   sum over Dp == sum over D,  sum (x - mean)^2 over D == sum over Dp - (Dp - D) * mean^2,  and y_pad = (0 - mean) * rstd * 0 + 0 = 0.
template <typename TI, typename TO, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_fwd_kernel(const TI* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               int rows, int D, int Dp, float eps, TO* __restrict__ y, float* __restrict__ mean_out,
                                                               float* __restrict__ rstd_out) {
  pdl_trigger();
  pdl_wait();
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int nvec = Dp >> 3;
  float invD = 1.f / (float)D;
  float npad = (float)(Dp - D);
  Vec8<float> g[NV], b[NV];
#pragma unroll
  for (int i = 0; i < NV; i++) {
    int c = lane + 32 * i;
    if (c < nvec) { g[i].load(gamma + c * 8); b[i].load(beta + c * 8); }
  }
  for (int row = blockIdx.x * LN_WARPS + warp; row < rows; row += gridDim.x * LN_WARPS) {
    const TI* xr = x + (int64_t)row * Dp;
    Vec8<TI> v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; i++) {
      int c = lane + 32 * i;
      if (c < nvec) {
        v[i].load(xr + c * 8);
        s += ((v[i].v[0] + v[i].v[1]) + (v[i].v[2] + v[i].v[3])) + ((v[i].v[4] + v[i].v[5]) + (v[i].v[6] + v[i].v[7]));
      }
    }
    float mean = warp_sum(s) * invD;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; i++) {
      int c = lane + 32 * i;
      if (c < nvec) {
#pragma unroll
        for (int j = 0; j < 8; j++) { float d = v[i].v[j] - mean; v[i].v[j] = d; q = fmaf(d, d, q); }
      }
    }
    float rstd = rsqrtf((warp_sum(q) - npad * mean * mean) * invD + eps);
    if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
    TO* yr = y + (int64_t)row * Dp;
#pragma unroll
    for (int i = 0; i < NV; i++) {
      int c = lane + 32 * i;
      if (c < nvec) {
        Vec8<TO> o;
#pragma unroll
        for (int j = 0; j < 8; j++) o.v[j] = fmaf(v[i].v[j] * rstd, g[i].v[j], b[i].v[j]);
        o.store(yr + c * 8);
      }
    }
  }
}

template <typename T> struct Vec4;
template <> struct Vec4<float> {
  float v[4];
  __device__ __forceinline__ void load(const float* p) { const float4 a = *(const float4*)p; v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; }
  __device__ __forceinline__ void store(float* p) const { *(float4*)p = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct Vec4<bf16> {
  float v[4];
  __device__ __forceinline__ void load(const bf16* p) {
    const uint2 r = *(const uint2*)p;
    const float2 a = __bfloat1622float2(*(const __nv_bfloat162*)&r.x), b = __bfloat1622float2(*(const __nv_bfloat162*)&r.y);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  __device__ __forceinline__ void store(bf16* p) const {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 r;
    r.x = *(const uint32_t*)&a; r.y = *(const uint32_t*)&b;
    *(uint2*)p = r;
  }
};
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// forward with half a warp per row (two rows per warp, 4-element vectors): for Dp = 64 * n every lane is busy and a warp has two rows
// of loads in flight (the warp-per-row kernel is latency-bound: ncu long-scoreboard 50 %, and fills 40 of 64 lane slots at Dp = 320)
template <typename TI, typename TO, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_fwd_hw_kernel(const TI* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  int rows, int D, int Dp, float eps, TO* __restrict__ y, float* __restrict__ mean_out,
                                                                  float* __restrict__ rstd_out) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = lane >> 4, hl = lane & 15;
  const float invD = 1.f / (float)D, npad = (float)(Dp - D);
  Vec4<float> g[NV], b[NV];
#pragma unroll
  for (int i = 0; i < NV; i++) { g[i].load(gamma + (hl + 16 * i) * 4); b[i].load(beta + (hl + 16 * i) * 4); }
  for (int64_t pr = blockIdx.x * LN_WARPS + warp; 2 * pr < rows; pr += gridDim.x * LN_WARPS) {
    const int64_t row = 2 * pr + half;
    const bool valid = row < rows;
    const TI* xr = x + row * Dp;
    Vec4<TI> v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; i++) {
      if (valid) v[i].load(xr + (hl + 16 * i) * 4);
      else v[i].v[0] = v[i].v[1] = v[i].v[2] = v[i].v[3] = 0.f;
      s += (v[i].v[0] + v[i].v[1]) + (v[i].v[2] + v[i].v[3]);
    }
    const float mean = half_warp_sum(s) * invD;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; i++) {
#pragma unroll
      for (int j = 0; j < 4; j++) { const float d = v[i].v[j] - mean; v[i].v[j] = d; q = fmaf(d, d, q); }
    }
    const float rstd = rsqrtf((half_warp_sum(q) - npad * mean * mean) * invD + eps);
    if (!valid) continue;
    if (hl == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
    TO* yr = y + row * Dp;
#pragma unroll
    for (int i = 0; i < NV; i++) {
      Vec4<TO> o;
#pragma unroll
      for (int j = 0; j < 4; j++) o.v[j] = fmaf(v[i].v[j] * rstd, g[i].v[j], b[i].v[j]);
      o.store(yr + (hl + 16 * i) * 4);
    }
  }
}

// backward: persistent grid, each warp walks rows; dgamma/dbeta partials stay in registers until the end.
// Optional fused epilogue: cast_out (bf16 / fp32, pitch Dp) = dropmask * dx_new, i.e. the GEMM operand the NEXT backward block would
// otherwise produce with a separate pass over dx (bpm_cast_drop); element index of the mask = row * Dp + column.
template <typename TG, typename TX, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32, 4) ln_bwd_kernel(const TG* __restrict__ dy, const TX* __restrict__ x, const float* __restrict__ mean_in,
                                                               const float* __restrict__ rstd_in, const float* __restrict__ gamma, int rows, int D,
                                                               int Dp, float* __restrict__ dx, int accumulate, float* __restrict__ dgamma,
                                                               float* __restrict__ dbeta, void* __restrict__ cast_out, int cast_dtype,
                                                               bpm_dropout_t cast_drop) {
  extern __shared__ float sm[];  // [LN_WARPS][2][Dp]
  pdl_trigger();
  pdl_wait();
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int nvec = Dp >> 3;
  float invD = 1.f / (float)D;
  float ag[NV][8], ab[NV][8];
  Vec8<float> gm[NV];
#pragma unroll
  for (int i = 0; i < NV; i++) {
    int c = lane + 32 * i;
#pragma unroll
    for (int j = 0; j < 8; j++) { ag[i][j] = 0.f; ab[i][j] = 0.f; gm[i].v[j] = 0.f; }
    if (c < nvec) gm[i].load(gamma + c * 8);
  }
  const DropCtx dc = make_drop(cast_drop);
  for (int row = blockIdx.x * LN_WARPS + warp; row < rows; row += gridDim.x * LN_WARPS) {
    const TG* gr = dy + (int64_t)row * Dp;
    const TX* xr = x + (int64_t)row * Dp;
    float* dr = dx + (int64_t)row * Dp;
    float mean = mean_in[row], rstd = rstd_in[row];
    Vec8<TG> g[NV];
    Vec8<TX> xv[NV];
    Vec8<float> o[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; i++) {                          // every load of the row is issued before anything is consumed
      int c = lane + 32 * i;
      if (c < nvec) {
        g[i].load(gr + c * 8);
        xv[i].load(xr + c * 8);
        if (accumulate) o[i].load(dr + c * 8);
        else {
#pragma unroll
          for (int j = 0; j < 8; j++) o[i].v[j] = 0.f;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NV; i++) {
      int c = lane + 32 * i;
      if (c < nvec) {
        // dy pads are zero (they come from GEMMs against zero-padded weights) and gamma pads are zero, so only x-hat needs care:
        // x-hat_pad = -mean * rstd is finite and always multiplied by a zero gradient.
#pragma unroll
        for (int j = 0; j < 8; j++) {
          float xh = (xv[i].v[j] - mean) * rstd;
          float gy = g[i].v[j];
          float gh = gy * gm[i].v[j];
          xv[i].v[j] = xh; g[i].v[j] = gh;
          s1 += gh; s2 = fmaf(gh, xh, s2);
          ag[i][j] = fmaf(gy, xh, ag[i][j]); ab[i][j] += gy;
        }
      }
    }
    s1 = warp_sum(s1) * invD;
    s2 = warp_sum(s2) * invD;
#pragma unroll
    for (int i = 0; i < NV; i++) {
      int c = lane + 32 * i;
      if (c < nvec) {
        if (c * 8 + 8 <= D) {
#pragma unroll
          for (int j = 0; j < 8; j++) o[i].v[j] += rstd * (g[i].v[j] - s1 - xv[i].v[j] * s2);
        } else {                                              // the chunk that straddles D (and pure pad chunks): keep pads at zero
#pragma unroll
          for (int j = 0; j < 8; j++) o[i].v[j] += (c * 8 + j < D) ? rstd * (g[i].v[j] - s1 - xv[i].v[j] * s2) : 0.f;
        }
        o[i].store(dr + c * 8);
        if (cast_out != nullptr) {
          float m[8];
          drop_mult8(dc, (uint64_t)row * (uint64_t)Dp + (uint64_t)(c * 8), m);
          if (cast_dtype == BPM_BF16) {
            Vec8<bf16> t;
#pragma unroll
            for (int j = 0; j < 8; j++) t.v[j] = o[i].v[j] * m[j];
            t.store((bf16*)cast_out + (int64_t)row * Dp + c * 8);
          } else {
            Vec8<float> t;
#pragma unroll
            for (int j = 0; j < 8; j++) t.v[j] = o[i].v[j] * m[j];
            t.store((float*)cast_out + (int64_t)row * Dp + c * 8);
          }
        }
      }
    }
  }
  // block-level reduction of the parameter gradients, then one atomic per column per block
  float* sg = sm + (size_t)warp * 2 * Dp;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    int c = lane + 32 * i;
    if (c < nvec) {
#pragma unroll
      for (int j = 0; j < 8; j++) { sg[c * 8 + j] = ag[i][j]; sg[Dp + c * 8 + j] = ab[i][j]; }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * Dp; c += blockDim.x) {
    int col = c % Dp;
    if (col >= D) continue;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < LN_WARPS; w++) s += sm[(size_t)w * 2 * Dp + c];
    atomicAdd((c < Dp ? dgamma : dbeta) + col, s);
  }
}

#define LN_PF 3                                  // rows in flight per warp in the prefetching backward
template <typename TG, typename TX, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32, 4) ln_bwd_pf_kernel(const TG* __restrict__ dy, const TX* __restrict__ x, const float* __restrict__ mean_in,
                                                               const float* __restrict__ rstd_in, const float* __restrict__ gamma, int rows, int D,
                                                               int Dp, float* __restrict__ dx, int accumulate, float* __restrict__ dgamma,
                                                               float* __restrict__ dbeta, void* __restrict__ cast_out, int cast_dtype,
                                                               bpm_dropout_t cast_drop) {
  // Same math as ln_bwd_kernel; the rows are PREFETCHED: lane 0 of every warp keeps LN_PF bulk copies (dy, x, dx rows) in flight
  // into a per-warp shared-memory ring, so the memory latency of row k + LN_PF overlaps the arithmetic of row k without holding the
  // rows in registers (the register version waits on its own loads half of the time: ncu long-scoreboard 49 %).
  extern __shared__ float sm[];  // [LN_WARPS][2][Dp] column sums | rings [LN_WARPS][LN_PF][dy row | x row | dx row] | barriers
  pdl_trigger();
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t row_g = (uint32_t)Dp * sizeof(TG), row_x = (uint32_t)Dp * sizeof(TX), row_d = (uint32_t)Dp * 4u;
  const uint32_t stb = row_g + row_x + row_d;
  uint8_t* const ring_gen = (uint8_t*)sm + (size_t)LN_WARPS * 2 * Dp * 4 + (size_t)warp * LN_PF * stb;
  const uint32_t ring_s = smem_u32(ring_gen);
  const uint32_t bars = smem_u32((uint8_t*)sm + (size_t)LN_WARPS * 2 * Dp * 4 + (size_t)LN_WARPS * LN_PF * stb) + (uint32_t)warp * LN_PF * 8u;
  if (lane == 0) {
    for (int s2 = 0; s2 < LN_PF; s2++) mbar_init(bars + 8u * s2, 1);
    mbar_fence_init();
  }
  __syncwarp();
  pdl_wait();
  const int row0 = blockIdx.x * LN_WARPS + warp, rstride = gridDim.x * LN_WARPS;
  auto issue = [&](int k) {                       // lane 0: row k of this warp -> stage k % LN_PF
    const int64_t row = row0 + (int64_t)k * rstride;
    if (row >= rows) return;
    const int s2 = k % LN_PF;
    const uint32_t dst = ring_s + (uint32_t)s2 * stb, bar = bars + 8u * s2;
    mbar_expect_tx(bar, row_g + row_x + (accumulate ? row_d : 0u));
    bulk_load_1d(dst, dy + row * Dp, row_g, bar);
    bulk_load_1d(dst + row_g, x + row * Dp, row_x, bar);
    if (accumulate) bulk_load_1d(dst + row_g + row_x, dx + row * Dp, row_d, bar);
  };
  if (lane == 0) {
    for (int k = 0; k < LN_PF; k++) issue(k);
  }
  int kk = 0;
  int nvec = Dp >> 3;
  float invD = 1.f / (float)D;
  float ag[NV][8], ab[NV][8];
  Vec8<float> gm[NV];
#pragma unroll
  for (int i = 0; i < NV; i++) {
    int c = lane + 32 * i;
#pragma unroll
    for (int j = 0; j < 8; j++) { ag[i][j] = 0.f; ab[i][j] = 0.f; gm[i].v[j] = 0.f; }
    if (c < nvec) gm[i].load(gamma + c * 8);
  }
  const DropCtx dc = make_drop(cast_drop);
  for (int row = row0; row < rows; row += rstride, kk++) {
    const int st = kk % LN_PF;
    const TG* gr = (const TG*)(ring_gen + (size_t)st * stb);
    const TX* xr = (const TX*)(ring_gen + (size_t)st * stb + row_g);
    const float* dri = (const float*)(ring_gen + (size_t)st * stb + row_g + row_x);
    float* dr = dx + (int64_t)row * Dp;
    float mean = mean_in[row], rstd = rstd_in[row];
    mbar_wait(bars + 8u * st, (uint32_t)(kk / LN_PF) & 1u);
    Vec8<TG> g[NV];
    Vec8<TX> xv[NV];
    Vec8<float> o[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; i++) {                          // every load of the row is issued before anything is consumed
      int c = lane + 32 * i;
      if (c < nvec) {
        g[i].load(gr + c * 8);
        xv[i].load(xr + c * 8);
        if (accumulate) o[i].load(dri + c * 8);
        else {
#pragma unroll
          for (int j = 0; j < 8; j++) o[i].v[j] = 0.f;
        }
      }
    }
    __syncwarp();                                           // every lane holds its part of the row: the stage may be refilled
    if (lane == 0) issue(kk + LN_PF);
#pragma unroll
    for (int i = 0; i < NV; i++) {
      int c = lane + 32 * i;
      if (c < nvec) {
        // dy pads are zero (they come from GEMMs against zero-padded weights) and gamma pads are zero, so only x-hat needs care:
        // x-hat_pad = -mean * rstd is finite and always multiplied by a zero gradient.
#pragma unroll
        for (int j = 0; j < 8; j++) {
          float xh = (xv[i].v[j] - mean) * rstd;
          float gy = g[i].v[j];
          float gh = gy * gm[i].v[j];
          xv[i].v[j] = xh; g[i].v[j] = gh;
          s1 += gh; s2 = fmaf(gh, xh, s2);
          ag[i][j] = fmaf(gy, xh, ag[i][j]); ab[i][j] += gy;
        }
      }
    }
    s1 = warp_sum(s1) * invD;
    s2 = warp_sum(s2) * invD;
#pragma unroll
    for (int i = 0; i < NV; i++) {
      int c = lane + 32 * i;
      if (c < nvec) {
        if (c * 8 + 8 <= D) {
#pragma unroll
          for (int j = 0; j < 8; j++) o[i].v[j] += rstd * (g[i].v[j] - s1 - xv[i].v[j] * s2);
        } else {                                              // the chunk that straddles D (and pure pad chunks): keep pads at zero
#pragma unroll
          for (int j = 0; j < 8; j++) o[i].v[j] += (c * 8 + j < D) ? rstd * (g[i].v[j] - s1 - xv[i].v[j] * s2) : 0.f;
        }
        o[i].store(dr + c * 8);
        if (cast_out != nullptr) {
          float m[8];
          drop_mult8(dc, (uint64_t)row * (uint64_t)Dp + (uint64_t)(c * 8), m);
          if (cast_dtype == BPM_BF16) {
            Vec8<bf16> t;
#pragma unroll
            for (int j = 0; j < 8; j++) t.v[j] = o[i].v[j] * m[j];
            t.store((bf16*)cast_out + (int64_t)row * Dp + c * 8);
          } else {
            Vec8<float> t;
#pragma unroll
            for (int j = 0; j < 8; j++) t.v[j] = o[i].v[j] * m[j];
            t.store((float*)cast_out + (int64_t)row * Dp + c * 8);
          }
        }
      }
    }
  }
  // block-level reduction of the parameter gradients, then one atomic per column per block
  float* sg = sm + (size_t)warp * 2 * Dp;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    int c = lane + 32 * i;
    if (c < nvec) {
#pragma unroll
      for (int j = 0; j < 8; j++) { sg[c * 8 + j] = ag[i][j]; sg[Dp + c * 8 + j] = ab[i][j]; }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * Dp; c += blockDim.x) {
    int col = c % Dp;
    if (col >= D) continue;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < LN_WARPS; w++) s += sm[(size_t)w * 2 * Dp + c];
    atomicAdd((c < Dp ? dgamma : dbeta) + col, s);
  }
}

// ---------------------------------------------------------------- backward, half a warp per row
// With one warp per row and 8-element vectors a row of Dp = 320 occupies 40 of the 64 lane slots of its two passes (62 %), and the kernel
// is bound by issue slots (ncu: 618 warp instructions per row, 62 % of the issue slots busy, 37 us against 26 us of memory time).  Here a
// warp takes TWO consecutive rows, 16 lanes each, 4-element vectors: Dp = 64 * NV fills every lane (320 = 16 x 5 x 4).  Same prefetch
// scheme (one bulk copy brings both rows of a stage: consecutive rows are contiguous), same arithmetic and summation order per row.
#define LN_HPF 2                                 // stages (row pairs) in flight per warp
template <typename TG, typename TX, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32, 3) ln_bwd_hw_kernel(const TG* __restrict__ dy, const TX* __restrict__ x, const float* __restrict__ mean_in,
                                                               const float* __restrict__ rstd_in, const float* __restrict__ gamma, int rows, int D,
                                                               int Dp, float* __restrict__ dx, int accumulate, float* __restrict__ dgamma,
                                                               float* __restrict__ dbeta, void* __restrict__ cast_out, int cast_dtype,
                                                               bpm_dropout_t cast_drop) {
  extern __shared__ float sm[];  // [LN_WARPS][2][Dp] column sums | rings [LN_WARPS][LN_HPF][2 dy rows | 2 x rows | 2 dx rows] | barriers
  pdl_trigger();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = lane >> 4, hl = lane & 15;
  const uint32_t row_g = (uint32_t)Dp * sizeof(TG), row_x = (uint32_t)Dp * sizeof(TX), row_d = (uint32_t)Dp * 4u;
  const uint32_t stb = 2u * (row_g + row_x + row_d);
  uint8_t* const ring_gen = (uint8_t*)sm + (size_t)LN_WARPS * 2 * Dp * 4 + (size_t)warp * LN_HPF * stb;
  const uint32_t ring_s = smem_u32(ring_gen);
  const uint32_t bars = smem_u32((uint8_t*)sm + (size_t)LN_WARPS * 2 * Dp * 4 + (size_t)LN_WARPS * LN_HPF * stb) + (uint32_t)warp * LN_HPF * 8u;
  if (lane == 0) {
    for (int s2 = 0; s2 < LN_HPF; s2++) mbar_init(bars + 8u * s2, 1);
    mbar_fence_init();
  }
  __syncwarp();
  pdl_wait();
  const int pair0 = blockIdx.x * LN_WARPS + warp, pstride = gridDim.x * LN_WARPS;            // in row PAIRS
  auto issue = [&](int k) {                       // lane 0: row pair k of this warp -> stage k % LN_HPF
    const int64_t row = 2 * (pair0 + (int64_t)k * pstride);
    if (row >= rows) return;
    const uint32_t n = rows - row >= 2 ? 2u : 1u;
    const int s2 = k % LN_HPF;
    const uint32_t dst = ring_s + (uint32_t)s2 * stb, bar = bars + 8u * s2;
    mbar_expect_tx(bar, n * (row_g + row_x + (accumulate ? row_d : 0u)));
    bulk_load_1d(dst, dy + row * Dp, n * row_g, bar);
    bulk_load_1d(dst + 2u * row_g, x + row * Dp, n * row_x, bar);
    if (accumulate) bulk_load_1d(dst + 2u * (row_g + row_x), dx + row * Dp, n * row_d, bar);
  };
  if (lane == 0) {
    for (int k = 0; k < LN_HPF; k++) issue(k);
  }
  const float invD = 1.f / (float)D;
  float ag[NV][4], ab[NV][4];
  Vec4<float> gm[NV];
#pragma unroll
  for (int i = 0; i < NV; i++) {
#pragma unroll
    for (int j = 0; j < 4; j++) { ag[i][j] = 0.f; ab[i][j] = 0.f; }
    gm[i].load(gamma + (hl + 16 * i) * 4);
  }
  const DropCtx dc = make_drop(cast_drop);
  int kk = 0;
  for (int64_t row = 2 * (int64_t)pair0; row < rows; row += 2 * (int64_t)pstride, kk++) {
    const int st = kk % LN_HPF;
    const int64_t my = row + half;
    const bool valid = my < rows;
    uint8_t* const stage = ring_gen + (size_t)st * stb;
    const TG* gr = (const TG*)(stage + (size_t)half * row_g);
    const TX* xr = (const TX*)(stage + 2u * row_g + (size_t)half * row_x);
    const float* dri = (const float*)(stage + 2u * (row_g + row_x) + (size_t)half * row_d);
    const float mean = valid ? mean_in[my] : 0.f, rstd = valid ? rstd_in[my] : 0.f;
    mbar_wait(bars + 8u * st, (uint32_t)(kk / LN_HPF) & 1u);
    Vec4<TG> g[NV];
    Vec4<TX> xv[NV];
    Vec4<float> o[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) {                            // every load of the row is issued before anything is consumed
      const int c = hl + 16 * i;
      if (valid) {
        g[i].load(gr + c * 4);
        xv[i].load(xr + c * 4);
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++) { g[i].v[j] = 0.f; xv[i].v[j] = 0.f; }
      }
      if (valid && accumulate) o[i].load(dri + c * 4);
      else {
#pragma unroll
        for (int j = 0; j < 4; j++) o[i].v[j] = 0.f;
      }
    }
    __syncwarp();                                             // both rows are in registers: the stage may be refilled
    if (lane == 0) issue(kk + LN_HPF);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; i++) {
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const float xh = (xv[i].v[j] - mean) * rstd;
        const float gy = g[i].v[j];
        const float gh = gy * gm[i].v[j];
        xv[i].v[j] = xh; g[i].v[j] = gh;
        s1 += gh; s2 = fmaf(gh, xh, s2);
        ag[i][j] = fmaf(gy, xh, ag[i][j]); ab[i][j] += gy;
      }
    }
    s1 = half_warp_sum(s1) * invD;
    s2 = half_warp_sum(s2) * invD;
    if (!valid) continue;
    float* dr = dx + my * Dp;
#pragma unroll
    for (int i = 0; i < NV; i++) {
      const int c = hl + 16 * i;
      if (c * 4 + 4 <= D) {
#pragma unroll
        for (int j = 0; j < 4; j++) o[i].v[j] += rstd * (g[i].v[j] - s1 - xv[i].v[j] * s2);
      } else {                                                // the chunk that straddles D (and pure pad chunks): keep pads at zero
#pragma unroll
        for (int j = 0; j < 4; j++) o[i].v[j] += (c * 4 + j < D) ? rstd * (g[i].v[j] - s1 - xv[i].v[j] * s2) : 0.f;
      }
      o[i].store(dr + c * 4);
      if (cast_out != nullptr) {
        float m[4];
        const uint64_t e = (uint64_t)my * (uint64_t)Dp + (uint64_t)(c * 4);
        if (dc.on) {
#pragma unroll
          for (int q = 0; q < 2; q++) {
            const uint32_t r = drop_rand_pair(dc, (e >> 1) + q);
            m[2 * q] = drop_keep_lo(dc, r) ? dc.inv_keep : 0.f;
            m[2 * q + 1] = drop_keep_hi(dc, r) ? dc.inv_keep : 0.f;
          }
        } else {
          m[0] = m[1] = m[2] = m[3] = 1.f;
        }
        if (cast_dtype == BPM_BF16) {
          Vec4<bf16> t;
#pragma unroll
          for (int j = 0; j < 4; j++) t.v[j] = o[i].v[j] * m[j];
          t.store((bf16*)cast_out + my * Dp + c * 4);
        } else {
          Vec4<float> t;
#pragma unroll
          for (int j = 0; j < 4; j++) t.v[j] = o[i].v[j] * m[j];
          t.store((float*)cast_out + my * Dp + c * 4);
        }
      }
    }
  }
  // the two halves of a warp hold the same columns: fold them, then the block-level reduction and one atomic per column per block
#pragma unroll
  for (int i = 0; i < NV; i++) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      ag[i][j] += __shfl_xor_sync(0xffffffffu, ag[i][j], 16);
      ab[i][j] += __shfl_xor_sync(0xffffffffu, ab[i][j], 16);
    }
  }
  float* sg = sm + (size_t)warp * 2 * Dp;
  if (half == 0) {
#pragma unroll
    for (int i = 0; i < NV; i++) {
      const int c = hl + 16 * i;
#pragma unroll
      for (int j = 0; j < 4; j++) { sg[c * 4 + j] = ag[i][j]; sg[Dp + c * 4 + j] = ab[i][j]; }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * Dp; c += blockDim.x) {
    int col = c % Dp;
    if (col >= D) continue;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < LN_WARPS; w++) s += sm[(size_t)w * 2 * Dp + c];
    atomicAdd((c < Dp ? dgamma : dbeta) + col, s);
  }
}

template <typename TI, typename TO>
static int ln_fwd_launch(const void* x, const float* gamma, const float* beta, int rows, int D, int Dp, float eps, void* y, float* mean, float* rstd,
                         cudaStream_t s) {
  int nv = bpm_cdiv(Dp / 8, 32);
  int grid = min(bpm_cdiv(rows, LN_WARPS), bpm_num_sms() * 16);
  if (Dp % 64 == 0 && Dp / 64 <= 6 && Dp % 256 != 0 && !(bpm_debug_get(0) & 2048)) {
    const int hgrid = min(bpm_cdiv(bpm_cdiv(rows, 2), LN_WARPS), bpm_num_sms() * 16);
#define LNFH(NV) (void)bpm_launch(ln_fwd_hw_kernel<TI, TO, NV>, dim3(hgrid), dim3(LN_WARPS * 32), 0, s, (const TI*)x, gamma, beta, rows, D, Dp, eps, (TO*)y, mean, rstd)
    switch (Dp / 64) {
      case 1: LNFH(1); break;
      case 2: LNFH(2); break;
      case 3: LNFH(3); break;
      case 5: LNFH(5); break;
      default: LNFH(6); break;
    }
#undef LNFH
    return BPM_OK;
  }
#define LNF(NV) (void)bpm_launch(ln_fwd_kernel<TI, TO, NV>, dim3(grid), dim3(LN_WARPS * 32), 0, s, (const TI*)x, gamma, beta, rows, D, Dp, eps, (TO*)y, mean, rstd)
  switch (nv) {
    case 1: LNF(1); break;
    case 2: LNF(2); break;
    case 3: LNF(3); break;
    case 4: LNF(4); break;
    default: bpm_set_error("layernorm: Dp %d > 1024 unsupported", Dp); return BPM_EINVAL;
  }
#undef LNF
  return BPM_OK;
}

extern "C" int bpm_layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, int rows, int D, int Dp, float eps, void* y,
                                 int y_dtype, float* mean, float* rstd, void* stream) {
  BPM_REQUIRE(x && gamma && beta && y && mean && rstd && rows > 0 && D > 0 && Dp >= D && Dp % 8 == 0, "layernorm_fwd: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  if (x_dtype == BPM_F32 && y_dtype == BPM_F32) rc = ln_fwd_launch<float, float>(x, gamma, beta, rows, D, Dp, eps, y, mean, rstd, s);
  else if (x_dtype == BPM_F32 && y_dtype == BPM_BF16) rc = ln_fwd_launch<float, bf16>(x, gamma, beta, rows, D, Dp, eps, y, mean, rstd, s);
  else if (x_dtype == BPM_BF16 && y_dtype == BPM_BF16) rc = ln_fwd_launch<bf16, bf16>(x, gamma, beta, rows, D, Dp, eps, y, mean, rstd, s);
  else if (x_dtype == BPM_BF16 && y_dtype == BPM_F32) rc = ln_fwd_launch<bf16, float>(x, gamma, beta, rows, D, Dp, eps, y, mean, rstd, s);
  else BPM_REQUIRE(false, "layernorm_fwd: bad dtype");
  if (rc) return rc;
  BPM_CHECK_LAUNCH("layernorm_fwd");
  return BPM_OK;
}

template <typename TG, typename TX>
static int ln_bwd_launch(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, int rows, int D, int Dp, float* dx,
                         int accumulate, float* dgamma, float* dbeta, void* cast_out, int cast_dtype, bpm_dropout_t cast_drop, cudaStream_t s) {
  int nv = bpm_cdiv(Dp / 8, 32);
  int grid = min(bpm_cdiv(rows, LN_WARPS), bpm_num_sms() * 4);
  size_t smem = (size_t)LN_WARPS * 2 * Dp * sizeof(float);
  const size_t hring = (size_t)LN_WARPS * LN_HPF * 2 * Dp * (sizeof(TG) + sizeof(TX) + 4) + LN_WARPS * LN_HPF * 8;
  if (Dp % 64 == 0 && Dp / 64 <= 6 && Dp % 256 != 0 && smem + hring <= 100 * 1024 && !(bpm_debug_get(0) & 2048)) {
    // half a warp per row: every lane busy (the full-warp mapping fills Dp / 256 of its last pass)
    const int hgrid = min(bpm_cdiv(bpm_cdiv(rows, 2), LN_WARPS), bpm_num_sms() * 3);
#define LNH(NV) \
  do { \
    if (int rc = bpm_func_smem((const void*)ln_bwd_hw_kernel<TG, TX, NV>, 100 * 1024, "layernorm_bwd")) return rc; \
    (void)bpm_launch(ln_bwd_hw_kernel<TG, TX, NV>, dim3(hgrid), dim3(LN_WARPS * 32), smem + hring, s, (const TG*)dy, (const TX*)x, mean, rstd, gamma, rows, D, \
                     Dp, dx, accumulate, dgamma, dbeta, cast_out, cast_dtype, cast_drop); \
  } while (0)
    switch (Dp / 64) {
      case 1: LNH(1); break;
      case 2: LNH(2); break;
      case 3: LNH(3); break;
      case 5: LNH(5); break;
      default: LNH(6); break;
    }
#undef LNH
    return BPM_OK;
  }
  const size_t ring = (size_t)LN_WARPS * LN_PF * Dp * (sizeof(TG) + sizeof(TX) + 4) + LN_WARPS * LN_PF * 8;
  const bool pf = smem + ring <= 52 * 1024 && !(bpm_debug_get(0) & 1024);          // 4 CTAs per SM must still fit
#define LNB(NV) \
  do { \
    if (pf) { \
      if (int rc = bpm_func_smem((const void*)ln_bwd_pf_kernel<TG, TX, NV>, 56 * 1024, "layernorm_bwd")) return rc; \
      (void)bpm_launch(ln_bwd_pf_kernel<TG, TX, NV>, dim3(grid), dim3(LN_WARPS * 32), smem + ring, s, (const TG*)dy, (const TX*)x, mean, rstd, gamma, rows, D, \
                       Dp, dx, accumulate, dgamma, dbeta, cast_out, cast_dtype, cast_drop); \
    } else { \
      (void)bpm_launch(ln_bwd_kernel<TG, TX, NV>, dim3(grid), dim3(LN_WARPS * 32), smem, s, (const TG*)dy, (const TX*)x, mean, rstd, gamma, rows, D, Dp, dx, \
                       accumulate, dgamma, dbeta, cast_out, cast_dtype, cast_drop); \
    } \
  } while (0)
  switch (nv) {
    case 1: LNB(1); break;
    case 2: LNB(2); break;
    case 3: LNB(3); break;
    case 4: LNB(4); break;
    default: bpm_set_error("layernorm: Dp %d > 1024 unsupported", Dp); return BPM_EINVAL;
  }
#undef LNB
  return BPM_OK;
}

extern "C" int bpm_layernorm_bwd_cast(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* mean, const float* rstd,
                                      const float* gamma, int rows, int D, int Dp, float* dx, int accumulate, float* dgamma, float* dbeta,
                                      void* cast_out, int cast_dtype, bpm_dropout_t cast_drop, void* stream) {
  BPM_REQUIRE(dy && x && mean && rstd && gamma && dx && dgamma && dbeta && rows > 0 && D > 0 && Dp >= D && Dp % 8 == 0, "layernorm_bwd: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
#define LNB_ARGS dy, x, mean, rstd, gamma, rows, D, Dp, dx, accumulate, dgamma, dbeta, cast_out, cast_dtype, cast_drop, s
  if (dy_dtype == BPM_F32 && x_dtype == BPM_F32) rc = ln_bwd_launch<float, float>(LNB_ARGS);
  else if (dy_dtype == BPM_BF16 && x_dtype == BPM_F32) rc = ln_bwd_launch<bf16, float>(LNB_ARGS);
  else if (dy_dtype == BPM_BF16 && x_dtype == BPM_BF16) rc = ln_bwd_launch<bf16, bf16>(LNB_ARGS);
  else if (dy_dtype == BPM_F32 && x_dtype == BPM_BF16) rc = ln_bwd_launch<float, bf16>(LNB_ARGS);
  else BPM_REQUIRE(false, "layernorm_bwd: bad dtype");
#undef LNB_ARGS
  if (rc) return rc;
  BPM_CHECK_LAUNCH("layernorm_bwd");
  return BPM_OK;
}

extern "C" int bpm_layernorm_bwd(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* mean, const float* rstd, const float* gamma,
                                 int rows, int D, int Dp, float* dx, int accumulate, float* dgamma, float* dbeta, void* stream) {
  bpm_dropout_t none = {0, nullptr, 0, 0.f};
  return bpm_layernorm_bwd_cast(dy, dy_dtype, x, x_dtype, mean, rstd, gamma, rows, D, Dp, dx, accumulate, dgamma, dbeta, nullptr, 0, none, stream);
}
