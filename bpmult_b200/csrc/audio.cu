// AudioEncoder of the 4-modality model (reference models/mmtr.py:93-108: Conv1d(96, 96, k=128, stride=2) x 2 + AdaptiveAvgPool1d(200)),
// SURVEY 8 f2.  A strided k = 128 convolution is an implicit GEMM with K = 96 * 128 = 12288: the contraction itself runs on the tcgen05
// GEMM of gemm_tc.cu (forward, weight gradient, input gradient); this file holds the HBM-bound glue around it, on time-major rows
// [B*T, C] (channels contiguous):
//   im2col    col[(b*Tout + t), k*C + c] = x[(b*Tin + stride*t + k), c]      -- a (row, tap) pair is ONE contiguous run of C elements
//   col2im    dx[(b*Tin + u), c] = sum over taps k with (u - k) % stride == 0 of dcol[(b*Tout + (u-k)/stride), k*C + c]   (a gather, no atomics)
//   weights   W[co][ci][k] (reference layout, fp32) <-> Wp[co][k*C + ci]    (so that K runs tap-major like the im2col rows)
//   adaptive average pooling over time and its backward
#include "bpm_common.cuh"

template <typename T>
__global__ void im2col_kernel(const T* __restrict__ x, int B, int Tin, int C, int ldx, int KW, int stride, T* __restrict__ col, int Tout) {
  pdl_trigger();
  pdl_wait();
  // one thread per 8 channels of a (row, tap): C % 8 == 0
  const int c8 = C / 8;
  const int64_t n = (int64_t)B * Tout * KW * c8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int cc = (int)(i % c8);
    const int64_t rk = i / c8;
    const int k = (int)(rk % KW);
    const int64_t row = rk / KW;
    const int t = (int)(row % Tout), b = (int)(row / Tout);
    const T* src = x + ((int64_t)b * Tin + (int64_t)stride * t + k) * ldx + cc * 8;
    T* dst = col + row * ((int64_t)KW * C) + (int64_t)k * C + cc * 8;
    if (sizeof(T) == 2) *(uint4*)dst = *(const uint4*)src;
    else { *(float4*)dst = *(const float4*)src; *((float4*)dst + 1) = *((const float4*)src + 1); }
  }
}

template <typename T>
__global__ void col2im_kernel(const T* __restrict__ dcol, int B, int Tin, int C, int KW, int stride, int Tout, float* __restrict__ dx, int lddx) {
  pdl_trigger();
  pdl_wait();
  const int c8 = C / 8;
  const int64_t n = (int64_t)B * Tin * c8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int cc = (int)(i % c8);
    const int64_t bu = i / c8;
    const int u = (int)(bu % Tin), b = (int)(bu / Tin);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    // taps k = u - stride*t with 0 <= t < Tout, 0 <= k < KW
    const int t_lo = max(0, (u - (KW - 1) + stride - 1) / stride), t_hi = min(Tout - 1, u / stride);
    for (int t = t_lo; t <= t_hi; t++) {
      const int k = u - stride * t;
      Vec8<T> v;
      v.load(dcol + ((int64_t)b * Tout + t) * ((int64_t)KW * C) + (int64_t)k * C + cc * 8);
#pragma unroll
      for (int j = 0; j < 8; j++) acc[j] += v.v[j];
    }
    float* dst = dx + ((int64_t)b * Tin + u) * lddx + cc * 8;
    *(float4*)dst = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *((float4*)dst + 1) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

// mode 0: Wp[co, k*Cin + ci] = W[co, ci, k] (dst dtype);  mode 1: gW[co, ci, k] (+)= gWp[co, k*Cin + ci] (fp32 both)
__global__ void conv_weight_kernel(const float* __restrict__ src, void* __restrict__ dst, int Cout, int Cin, int KW, int dst_dtype, int mode, int accumulate) {
  pdl_trigger();
  pdl_wait();
  const int64_t n = (int64_t)Cout * Cin * KW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (mode == 0) {                                   // i indexes the packed destination [co][k][ci]
      const int ci = (int)(i % Cin), k = (int)((i / Cin) % KW), co = (int)(i / ((int64_t)Cin * KW));
      st_from_f(dst, dst_dtype, i, src[((int64_t)co * Cin + ci) * KW + k]);
    } else {                                           // i indexes the reference-layout destination [co][ci][k]
      const int k = (int)(i % KW), ci = (int)((i / KW) % Cin), co = (int)(i / ((int64_t)Cin * KW));
      const float g = src[((int64_t)co * KW + k) * Cin + ci];
      float* d = (float*)dst;
      d[i] = accumulate ? d[i] + g : g;
    }
  }
}

// y[(b*Tp + i), c] = mean of x[(b*T + t), c] over t in [floor(i*T/Tp), ceil((i+1)*T/Tp))   (torch AdaptiveAvgPool1d window rule)
template <typename TX, typename TY>
__global__ void adaptive_pool_fwd_kernel(const TX* __restrict__ x, int B, int T, int C, int ldx, int Tp, TY* __restrict__ y, int ldy) {
  pdl_trigger();
  pdl_wait();
  const int64_t n = (int64_t)B * Tp * C;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const int64_t bi = idx / C;
    const int i = (int)(bi % Tp), b = (int)(bi / Tp);
    const int t0 = (int)(((int64_t)i * T) / Tp), t1 = (int)((((int64_t)(i + 1)) * T + Tp - 1) / Tp);
    float acc = 0.f;
    for (int t = t0; t < t1; t++) acc += to_f(x[((int64_t)b * T + t) * ldx + c]);
    y[((int64_t)b * Tp + i) * ldy + c] = from_f<TY>(acc / (float)(t1 - t0));
  }
}

// dx[(b*T + t), c] = sum over the windows i that contain t of dy[(b*Tp + i), c] / len_i
__global__ void adaptive_pool_bwd_kernel(const float* __restrict__ dy, int B, int T, int C, int Tp, int lddy, float* __restrict__ dx, int lddx) {
  pdl_trigger();
  pdl_wait();
  const int64_t n = (int64_t)B * T * C;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const int64_t bt = idx / C;
    const int t = (int)(bt % T), b = (int)(bt / T);
    // windows containing t: i with floor(i*T/Tp) <= t < ceil((i+1)*T/Tp); a superset range, then the exact test
    int i_lo = (int)(((int64_t)t * Tp) / T) - (Tp + T - 1) / T - 1, i_hi = (int)((((int64_t)(t + 1)) * Tp + T - 1) / T);
    i_lo = max(i_lo, 0); i_hi = min(i_hi, Tp - 1);
    float acc = 0.f;
    for (int i = i_lo; i <= i_hi; i++) {
      const int t0 = (int)(((int64_t)i * T) / Tp), t1 = (int)((((int64_t)(i + 1)) * T + Tp - 1) / Tp);
      if (t >= t0 && t < t1) acc += dy[((int64_t)b * Tp + i) * lddy + c] / (float)(t1 - t0);
    }
    dx[((int64_t)b * T + t) * lddx + c] = acc;
  }
}

static int grid_for(int64_t n) {
  int64_t blocks = (n + 255) / 256, cap = (int64_t)bpm_num_sms() * 16;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

extern "C" int bpm_conv1d_im2col(const void* x, int dtype, int B, int Tin, int C, int ldx, int KW, int stride, void* col, int Tout, void* stream) {
  BPM_REQUIRE(x && col && B > 0 && C % 8 == 0 && ldx % 8 == 0 && KW > 0 && stride > 0 && Tout == (Tin - KW) / stride + 1 && Tout > 0, "conv1d_im2col: bad shape");
  const int64_t n = (int64_t)B * Tout * KW * (C / 8);
  cudaError_t e;
  if (dtype == BPM_BF16) e = bpm_launch(im2col_kernel<bf16>, dim3(grid_for(n)), dim3(256), 0, (cudaStream_t)stream, (const bf16*)x, B, Tin, C, ldx, KW, stride, (bf16*)col, Tout);
  else e = bpm_launch(im2col_kernel<float>, dim3(grid_for(n)), dim3(256), 0, (cudaStream_t)stream, (const float*)x, B, Tin, C, ldx, KW, stride, (float*)col, Tout);
  if (e != cudaSuccess) { bpm_set_error("conv1d_im2col: launch failed: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return BPM_ELAUNCH; }
  return BPM_OK;
}

extern "C" int bpm_conv1d_col2im(const void* dcol, int dtype, int B, int Tin, int C, int KW, int stride, int Tout, float* dx, int lddx, void* stream) {
  BPM_REQUIRE(dcol && dx && B > 0 && C % 8 == 0 && lddx % 4 == 0 && Tout == (Tin - KW) / stride + 1, "conv1d_col2im: bad shape");
  const int64_t n = (int64_t)B * Tin * (C / 8);
  cudaError_t e;
  if (dtype == BPM_BF16) e = bpm_launch(col2im_kernel<bf16>, dim3(grid_for(n)), dim3(256), 0, (cudaStream_t)stream, (const bf16*)dcol, B, Tin, C, KW, stride, Tout, dx, lddx);
  else e = bpm_launch(col2im_kernel<float>, dim3(grid_for(n)), dim3(256), 0, (cudaStream_t)stream, (const float*)dcol, B, Tin, C, KW, stride, Tout, dx, lddx);
  if (e != cudaSuccess) { bpm_set_error("conv1d_col2im: launch failed: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return BPM_ELAUNCH; }
  return BPM_OK;
}

extern "C" int bpm_conv1d_pack_weight(const float* W, int Cout, int Cin, int KW, void* Wp, int dst_dtype, void* stream) {
  BPM_REQUIRE(W && Wp && Cout > 0 && Cin > 0 && KW > 0, "conv1d_pack_weight: bad args");
  cudaError_t e = bpm_launch(conv_weight_kernel, dim3(grid_for((int64_t)Cout * Cin * KW)), dim3(256), 0, (cudaStream_t)stream, W, Wp, Cout, Cin, KW, dst_dtype, 0, 0);
  if (e != cudaSuccess) { bpm_set_error("conv1d_pack_weight: launch failed: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return BPM_ELAUNCH; }
  return BPM_OK;
}

extern "C" int bpm_conv1d_unpack_wgrad(const float* gWp, int Cout, int Cin, int KW, float* gW, int accumulate, void* stream) {
  BPM_REQUIRE(gWp && gW && Cout > 0 && Cin > 0 && KW > 0, "conv1d_unpack_wgrad: bad args");
  cudaError_t e = bpm_launch(conv_weight_kernel, dim3(grid_for((int64_t)Cout * Cin * KW)), dim3(256), 0, (cudaStream_t)stream, gWp, (void*)gW, Cout, Cin, KW, BPM_F32, 1, accumulate);
  if (e != cudaSuccess) { bpm_set_error("conv1d_unpack_wgrad: launch failed: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return BPM_ELAUNCH; }
  return BPM_OK;
}

extern "C" int bpm_adaptive_pool_fwd(const void* x, int x_dtype, int B, int T, int C, int ldx, int Tp, void* y, int y_dtype, int ldy, void* stream) {
  BPM_REQUIRE(x && y && B > 0 && T > 0 && C > 0 && Tp > 0, "adaptive_pool_fwd: bad args");
  const int g = grid_for((int64_t)B * Tp * C);
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e;
  if (x_dtype == BPM_BF16 && y_dtype == BPM_BF16) e = bpm_launch(adaptive_pool_fwd_kernel<bf16, bf16>, dim3(g), dim3(256), 0, s, (const bf16*)x, B, T, C, ldx, Tp, (bf16*)y, ldy);
  else if (x_dtype == BPM_BF16) e = bpm_launch(adaptive_pool_fwd_kernel<bf16, float>, dim3(g), dim3(256), 0, s, (const bf16*)x, B, T, C, ldx, Tp, (float*)y, ldy);
  else if (y_dtype == BPM_BF16) e = bpm_launch(adaptive_pool_fwd_kernel<float, bf16>, dim3(g), dim3(256), 0, s, (const float*)x, B, T, C, ldx, Tp, (bf16*)y, ldy);
  else e = bpm_launch(adaptive_pool_fwd_kernel<float, float>, dim3(g), dim3(256), 0, s, (const float*)x, B, T, C, ldx, Tp, (float*)y, ldy);
  if (e != cudaSuccess) { bpm_set_error("adaptive_pool_fwd: launch failed: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return BPM_ELAUNCH; }
  return BPM_OK;
}

extern "C" int bpm_adaptive_pool_bwd(const float* dy, int B, int T, int C, int Tp, int lddy, float* dx, int lddx, void* stream) {
  BPM_REQUIRE(dy && dx && B > 0 && T > 0 && C > 0 && Tp > 0, "adaptive_pool_bwd: bad args");
  cudaError_t e = bpm_launch(adaptive_pool_bwd_kernel, dim3(grid_for((int64_t)B * T * C)), dim3(256), 0, (cudaStream_t)stream, dy, B, T, C, Tp, lddy, dx, lddx);
  if (e != cudaSuccess) { bpm_set_error("adaptive_pool_bwd: launch failed: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return BPM_ELAUNCH; }
  return BPM_OK;
}
