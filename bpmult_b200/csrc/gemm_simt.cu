// Exact-fp32 FFMA GEMM ("precision mode", fp32 storage) with the shared fused epilogue.  Also used for the
// [B, D]-row head of the model (tiny M) where tensor-core tiles would be >90% padding.
// 64x64x16 tiles, 256 threads, 4x4 micro-tile per thread, arbitrary M/N/K/pitches, all four transpose combinations.
#include "gemm_epilogue.cuh"

#define SG_BM 64
#define SG_BN 64
#define SG_BK 16

template <typename T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(int ta, int tb, int M, int N, int K, const T* __restrict__ A, int lda, const T* __restrict__ B,
                                                        int ldb, EpiParams ep, int k_per_split) {
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Bs[SG_BK][SG_BN + 4];
  int t = threadIdx.x;
  int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  int kbeg = blockIdx.z * k_per_split, kend = min(K, kbeg + k_per_split);
  int tm = t >> 4, tn = t & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += SG_BK) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int m, k;
      if (ta == 0) { k = t & 15; m = (t >> 4) + 16 * i; }
      else { m = t & 63; k = (t >> 6) + 4 * i; }
      int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < kend) v = to_f<T>(ta == 0 ? A[(int64_t)gm * lda + gk] : A[(int64_t)gk * lda + gm]);
      As[k][m] = v;
      int n;
      if (tb == 0) { k = t & 15; n = (t >> 4) + 16 * i; }
      else { n = t & 63; k = (t >> 6) + 4 * i; }
      int gn = n0 + n; gk = k0 + k;
      v = 0.f;
      if (gn < N && gk < kend) v = to_f<T>(tb == 0 ? B[(int64_t)gn * ldb + gk] : B[(int64_t)gk * ldb + gn]);
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SG_BK; k++) {
      float4 a = *(const float4*)&As[k][tm * 4];
      float4 b = *(const float4*)&Bs[k][tn * 4];
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  DropCtx dc = make_drop(ep.drop);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int m = m0 + tm * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int n = n0 + tn * 4 + j;
      if (n < N) epi_store1(ep, dc, m, n, acc[i][j]);
    }
  }
}

// Skinny problems (M <= 64 rows: the [B, D] head of the model -- final GMU, residual MLP, out_layer and their input gradients).  The 64x64
// tiles above give such a problem N/64 CTAs (5 of 148 SMs for N = 320) and a K loop of 20..80 barrier-separated steps: ~30 us a launch,
// 0.8 ms of the cfg-2 step over the head's 27 GEMMs.  Here a CTA owns 8 output columns and ALL rows, K is split 4 ways inside the CTA
// (thread = (row, k-slice)) and reduced through shared memory: N/8 CTAs, 64-deep K steps.
#define SK_BN 8
#define SK_BK 64
template <int TB>
__global__ void __launch_bounds__(256) gemm_skinny_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                                          EpiParams ep) {
  __shared__ float As[SK_BK][64 + 1];        // [k][m]
  __shared__ float Bs[SK_BN][SK_BK + 1];     // [n][k]
  __shared__ float red[3][64][SK_BN];
  const int t = threadIdx.x, m = t & 63, ks = t >> 6;
  const int n0 = blockIdx.x * SK_BN;
  float acc[SK_BN];
#pragma unroll
  for (int j = 0; j < SK_BN; j++) acc[j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += SK_BK) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const int idx = t + 256 * i, k = idx & 63, mm = idx >> 6;
      As[k][mm] = (mm < M && k0 + k < K) ? A[(int64_t)mm * lda + k0 + k] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const int idx = t + 256 * i;
      int k, n;
      if (TB == 0) { k = idx & 63; n = idx >> 6; }
      else { n = idx & 7; k = idx >> 3; }
      float v = 0.f;
      if (n0 + n < N && k0 + k < K) v = TB == 0 ? B[(int64_t)(n0 + n) * ldb + k0 + k] : B[(int64_t)(k0 + k) * ldb + n0 + n];
      Bs[n][k] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; kk++) {
      const int k = ks * 16 + kk;
      const float a = As[k][m];
#pragma unroll
      for (int j = 0; j < SK_BN; j++) acc[j] = fmaf(a, Bs[j][k], acc[j]);
    }
    __syncthreads();
  }
  if (ks > 0) {
#pragma unroll
    for (int j = 0; j < SK_BN; j++) red[ks - 1][m][j] = acc[j];
  }
  __syncthreads();
  if (ks == 0 && m < M) {
    const DropCtx dc = make_drop(ep.drop);
#pragma unroll
    for (int j = 0; j < SK_BN; j++) {
      const float v = acc[j] + red[0][m][j] + red[1][m][j] + red[2][m][j];
      if (n0 + j < N) epi_store1(ep, dc, m, n0 + j, v);
    }
  }
}

int bpm_gemm_simt(const bpm_gemm_t* g, cudaStream_t stream) {
  EpiParams ep = make_epi(g);
  if (g->ab_dtype == BPM_F32 && g->ta == 0 && g->M <= 64 && g->K >= 64) {
    dim3 grid(bpm_cdiv(g->N, SK_BN));
    if (g->tb == 0) gemm_skinny_kernel<0><<<grid, 256, 0, stream>>>(g->M, g->N, g->K, (const float*)g->A, g->lda, (const float*)g->B, g->ldb, ep);
    else gemm_skinny_kernel<1><<<grid, 256, 0, stream>>>(g->M, g->N, g->K, (const float*)g->A, g->lda, (const float*)g->B, g->ldb, ep);
    BPM_CHECK_LAUNCH("gemm_skinny");
    return BPM_OK;
  }
  int gx = bpm_cdiv(g->N, SG_BN), gy = bpm_cdiv(g->M, SG_BM);
  int split = 1;
  if (g->accumulate) {
    split = g->split_k > 0 ? g->split_k : max(1, min(bpm_cdiv(g->K, 256), (2 * bpm_num_sms()) / max(1, gx * gy)));
  }
  int kps = bpm_cdiv(bpm_cdiv(g->K, split), SG_BK) * SG_BK;
  split = bpm_cdiv(g->K, kps);
  dim3 grid(gx, gy, split);
  if (g->ab_dtype == BPM_BF16)
    gemm_simt_kernel<bf16><<<grid, 256, 0, stream>>>(g->ta, g->tb, g->M, g->N, g->K, (const bf16*)g->A, g->lda, (const bf16*)g->B, g->ldb, ep, kps);
  else
    gemm_simt_kernel<float><<<grid, 256, 0, stream>>>(g->ta, g->tb, g->M, g->N, g->K, (const float*)g->A, g->lda, (const float*)g->B, g->ldb, ep, kps);
  BPM_CHECK_LAUNCH("gemm_simt");
  return BPM_OK;
}
