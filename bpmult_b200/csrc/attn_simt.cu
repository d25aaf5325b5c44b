// Exact-fp32 crossmodal attention ("precision mode", fp32 or bf16 storage, fp32 math): one warp per query row
// (forward, dQ) or per key row (dK/dV).  Never materialises the [B*H, T, S] score tensor in HBM
// (reference: models/multihead_attention.py:110-126 does, in fp32).  Also serves head dims the tensor-core kernel
// does not cover.  Mask: key j visible to query i iff j <= i + mask_off (models/transformer.py:209-216).
#include "bpm_common.cuh"

#define AS_WARPS 4

template <typename T>
__device__ __forceinline__ float dot_row(const float* __restrict__ a_s, const T* __restrict__ row, int dhp) {
  float s = 0.f;
  for (int d = 0; d < dhp; d += 8) {
    Vec8<T> kv; kv.load(row + d);
#pragma unroll
    for (int j = 0; j < 8; j++) s = fmaf(a_s[d + j], kv.v[j], s);
  }
  return s;
}

template <typename T>
__global__ void __launch_bounds__(AS_WARPS * 32) attn_fwd_simt(bpm_attn_t a, const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                                                               T* __restrict__ out, float* __restrict__ lse) {
  extern __shared__ float sm[];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int bh = blockIdx.x, b = bh / a.H, h = bh % a.H;
  int i = blockIdx.y * AS_WARPS + warp;
  if (i >= a.T) return;
  float* sc = sm + (size_t)warp * (a.S + a.dhp);
  float* qs = sc + a.S;
  int pitch = a.H * a.dhp, pkv = a.ld_kv ? a.ld_kv : pitch;
  const T* qrow = q + ((int64_t)b * a.T + i) * pitch + h * a.dhp;
  for (int d = lane; d < a.dhp; d += 32) qs[d] = to_f<T>(qrow[d]);
  __syncwarp();
  int jmax = a.mask_off >= 0 ? min(a.S - 1, i + a.mask_off) : a.S - 1;
  const T* kb = k + (int64_t)b * a.S * pkv + h * a.dhp;
  const T* vb = v + (int64_t)b * a.S * pkv + h * a.dhp;
  float m = -INFINITY;
  for (int j = lane; j <= jmax; j += 32) {
    float s = (a.key_pad && a.key_pad[(int64_t)b * a.S + j]) ? -INFINITY : dot_row<T>(qs, kb + (int64_t)j * pkv, a.dhp);
    sc[j] = s;
    m = fmaxf(m, s);
  }
  m = warp_max(m);
  float l = 0.f;
  for (int j = lane; j <= jmax; j += 32) { float p = expf(sc[j] - m); sc[j] = p; l += p; }
  l = warp_sum(l);
  float inv_l = 1.f / l;
  DropCtx dc = make_drop(a.drop);
  uint64_t ebase = ((uint64_t)bh * a.T + i) * (uint64_t)a.S;
  for (int j = lane; j <= jmax; j += 32) sc[j] = sc[j] * inv_l * drop_mult1(dc, ebase + j);
  __syncwarp();
  if (lane == 0) lse[(int64_t)bh * a.T + i] = m + logf(l);
  T* orow = out + ((int64_t)b * a.T + i) * pitch + h * a.dhp;
  for (int d = lane; d < a.dhp; d += 32) {
    float acc = 0.f;
    for (int j = 0; j <= jmax; j++) acc = fmaf(sc[j], to_f<T>(vb[(int64_t)j * pkv + d]), acc);
    orow[d] = from_f<T>(acc);
  }
}

// dQ (and delta = rowsum(dO * O)) : warp per query row
template <typename T>
__global__ void __launch_bounds__(AS_WARPS * 32) attn_bwd_dq_simt(bpm_attn_t a, const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                                                                  const T* __restrict__ out, const T* __restrict__ dout, const float* __restrict__ lse,
                                                                  float* __restrict__ delta, T* __restrict__ dq, float dq_scale) {
  extern __shared__ float sm[];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int bh = blockIdx.x, b = bh / a.H, h = bh % a.H;
  int i = blockIdx.y * AS_WARPS + warp;
  if (i >= a.T) return;
  float* sc = sm + (size_t)warp * (a.S + 2 * a.dhp);
  float* qs = sc + a.S;
  float* gs = qs + a.dhp;
  int pitch = a.H * a.dhp, pkv = a.ld_kv ? a.ld_kv : pitch;
  int64_t ro = ((int64_t)b * a.T + i) * pitch + h * a.dhp;
  float dl = 0.f;
  for (int d = lane; d < a.dhp; d += 32) {
    qs[d] = to_f<T>(q[ro + d]);
    float g = to_f<T>(dout[ro + d]);
    gs[d] = g;
    dl += g * to_f<T>(out[ro + d]);
  }
  dl = warp_sum(dl);
  __syncwarp();
  float L = lse[(int64_t)bh * a.T + i];
  if (lane == 0) {
    delta[(int64_t)bh * a.T + i] = dl;
    delta[(int64_t)a.B * a.H * a.T + (int64_t)bh * a.T + i] = L * 1.4426950408889634f;      // second half of the workspace: lse * log2e
  }
  int jmax = a.mask_off >= 0 ? min(a.S - 1, i + a.mask_off) : a.S - 1;
  const T* kb = k + (int64_t)b * a.S * pkv + h * a.dhp;
  const T* vb = v + (int64_t)b * a.S * pkv + h * a.dhp;
  DropCtx dc = make_drop(a.drop);
  uint64_t ebase = ((uint64_t)bh * a.T + i) * (uint64_t)a.S;
  for (int j = lane; j <= jmax; j += 32) {
    float ds = 0.f;
    if (!(a.key_pad && a.key_pad[(int64_t)b * a.S + j])) {
      float s = dot_row<T>(qs, kb + (int64_t)j * pkv, a.dhp);
      float p = expf(s - L);
      float dpt = dot_row<T>(gs, vb + (int64_t)j * pkv, a.dhp);
      ds = p * (dpt * drop_mult1(dc, ebase + j) - dl);
    }
    sc[j] = ds;
  }
  __syncwarp();
  for (int d = lane; d < a.dhp; d += 32) {
    float acc = 0.f;
    for (int j = 0; j <= jmax; j++) acc = fmaf(sc[j], to_f<T>(kb[(int64_t)j * pkv + d]), acc);
    dq[ro + d] = from_f<T>(acc * dq_scale);
  }
}

// dK, dV : warp per key row
template <typename T>
__global__ void __launch_bounds__(AS_WARPS * 32) attn_bwd_dkv_simt(bpm_attn_t a, const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                                                                   const T* __restrict__ dout, const float* __restrict__ lse,
                                                                   const float* __restrict__ delta, T* __restrict__ dk, T* __restrict__ dv) {
  extern __shared__ float sm[];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int bh = blockIdx.x, b = bh / a.H, h = bh % a.H;
  int j = blockIdx.y * AS_WARPS + warp;
  if (j >= a.S) return;
  float* pt = sm + (size_t)warp * (2 * a.T + 2 * a.dhp);
  float* ds = pt + a.T;
  float* ks = ds + a.T;
  float* vs = ks + a.dhp;
  int pitch = a.H * a.dhp, pkv = a.ld_kv ? a.ld_kv : pitch, pdkv = a.ld_dkv ? a.ld_dkv : pitch;
  int64_t ko = ((int64_t)b * a.S + j) * pkv + h * a.dhp, kdo = ((int64_t)b * a.S + j) * pdkv + h * a.dhp;
  for (int d = lane; d < a.dhp; d += 32) { ks[d] = to_f<T>(k[ko + d]); vs[d] = to_f<T>(v[ko + d]); }
  __syncwarp();
  bool padded = a.key_pad && a.key_pad[(int64_t)b * a.S + j];
  int imin = a.mask_off >= 0 ? max(0, j - a.mask_off) : 0;
  if (padded) imin = a.T;
  const T* qb = q + (int64_t)b * a.T * pitch + h * a.dhp;
  const T* gb = dout + (int64_t)b * a.T * pitch + h * a.dhp;
  DropCtx dc = make_drop(a.drop);
  for (int i = imin + lane; i < a.T; i += 32) {
    float s = dot_row<T>(ks, qb + (int64_t)i * pitch, a.dhp);
    float p = expf(s - lse[(int64_t)bh * a.T + i]);
    float dpt = dot_row<T>(vs, gb + (int64_t)i * pitch, a.dhp);
    float mult = drop_mult1(dc, ((uint64_t)bh * a.T + i) * (uint64_t)a.S + j);
    pt[i] = p * mult;
    ds[i] = p * (dpt * mult - delta[(int64_t)bh * a.T + i]);
  }
  __syncwarp();
  for (int d = lane; d < a.dhp; d += 32) {
    float av = 0.f, ak = 0.f;
    for (int i = imin; i < a.T; i++) {
      av = fmaf(pt[i], to_f<T>(gb[(int64_t)i * pitch + d]), av);
      ak = fmaf(ds[i], to_f<T>(qb[(int64_t)i * pitch + d]), ak);
    }
    dv[kdo + d] = from_f<T>(av);
    dk[kdo + d] = from_f<T>(ak);
  }
}

// head-averaged probabilities (after dropout, like the reference's second return value)
template <typename T>
__global__ void attn_weights_simt(bpm_attn_t a, const T* __restrict__ q, const T* __restrict__ k, const float* __restrict__ lse, float* __restrict__ w) {
  int64_t n = (int64_t)a.B * a.T * a.S;
  int pitch = a.H * a.dhp;
  DropCtx dc = make_drop(a.drop);
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    int j = (int)(e % a.S), i = (int)((e / a.S) % a.T), b = (int)(e / ((int64_t)a.S * a.T));
    float acc = 0.f;
    bool vis = (a.mask_off < 0 || j <= i + a.mask_off) && !(a.key_pad && a.key_pad[(int64_t)b * a.S + j]);
    if (vis) {
      for (int h = 0; h < a.H; h++) {
        const T* qr = q + ((int64_t)b * a.T + i) * pitch + h * a.dhp;
        const T* kr = k + ((int64_t)b * a.S + j) * (a.ld_kv ? a.ld_kv : pitch) + h * a.dhp;
        float s = 0.f;
        for (int d = 0; d < a.dhp; d++) s = fmaf(to_f<T>(qr[d]), to_f<T>(kr[d]), s);
        int64_t bh = (int64_t)b * a.H + h;
        acc += expf(s - lse[bh * a.T + i]) * drop_mult1(dc, ((uint64_t)bh * a.T + i) * (uint64_t)a.S + j);
      }
    }
    w[e] = acc / (float)a.H;
  }
}

static int check_attn(const bpm_attn_t* a) {
  BPM_REQUIRE(a && a->B > 0 && a->T > 0 && a->S > 0 && a->H > 0 && a->dhp >= a->dh && a->dhp % 8 == 0 && a->dhp <= 256, "xattn: bad shape");
  return BPM_OK;
}

template <typename K>
static int set_smem(K kern, size_t bytes) {
  if (bytes > 48 * 1024) {
    BPM_REQUIRE(bytes <= 227 * 1024, "xattn(simt): sequence too long for the fp32 precision-mode kernel (%zu B smem)", bytes);
    // the requirement grows with S: always raise the limit to the maximum once per (kernel, device)
    if (int rc = bpm_func_smem((const void*)kern, 227 * 1024, "xattn(simt)")) return rc;
  }
  return BPM_OK;
}

int bpm_xattn_fwd_simt(const bpm_attn_t* a, const void* q, const void* k, const void* v, void* out, float* lse, cudaStream_t s) {
  int rc = check_attn(a);
  if (rc) return rc;
  dim3 grid(a->B * a->H, bpm_cdiv(a->T, AS_WARPS));
  size_t smem = (size_t)AS_WARPS * (a->S + a->dhp) * sizeof(float);
  if (a->dtype == BPM_BF16) {
    if ((rc = set_smem(attn_fwd_simt<bf16>, smem))) return rc;
    attn_fwd_simt<bf16><<<grid, AS_WARPS * 32, smem, s>>>(*a, (const bf16*)q, (const bf16*)k, (const bf16*)v, (bf16*)out, lse);
  } else {
    if ((rc = set_smem(attn_fwd_simt<float>, smem))) return rc;
    attn_fwd_simt<float><<<grid, AS_WARPS * 32, smem, s>>>(*a, (const float*)q, (const float*)k, (const float*)v, (float*)out, lse);
  }
  BPM_CHECK_LAUNCH("xattn_fwd_simt");
  return BPM_OK;
}

int bpm_xattn_bwd_simt(const bpm_attn_t* a, const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse,
                       float* delta, void* dq, float dq_scale, void* dk, void* dv, cudaStream_t s) {
  int rc = check_attn(a);
  if (rc) return rc;
  dim3 g1(a->B * a->H, bpm_cdiv(a->T, AS_WARPS)), g2(a->B * a->H, bpm_cdiv(a->S, AS_WARPS));
  size_t sm1 = (size_t)AS_WARPS * (a->S + 2 * a->dhp) * sizeof(float), sm2 = (size_t)AS_WARPS * (2 * a->T + 2 * a->dhp) * sizeof(float);
  if (a->dtype == BPM_BF16) {
    if ((rc = set_smem(attn_bwd_dq_simt<bf16>, sm1))) return rc;
    if ((rc = set_smem(attn_bwd_dkv_simt<bf16>, sm2))) return rc;
    attn_bwd_dq_simt<bf16><<<g1, AS_WARPS * 32, sm1, s>>>(*a, (const bf16*)q, (const bf16*)k, (const bf16*)v, (const bf16*)out, (const bf16*)dout, lse, delta,
                                                          (bf16*)dq, dq_scale);
    attn_bwd_dkv_simt<bf16><<<g2, AS_WARPS * 32, sm2, s>>>(*a, (const bf16*)q, (const bf16*)k, (const bf16*)v, (const bf16*)dout, lse, delta, (bf16*)dk,
                                                           (bf16*)dv);
  } else {
    if ((rc = set_smem(attn_bwd_dq_simt<float>, sm1))) return rc;
    if ((rc = set_smem(attn_bwd_dkv_simt<float>, sm2))) return rc;
    attn_bwd_dq_simt<float><<<g1, AS_WARPS * 32, sm1, s>>>(*a, (const float*)q, (const float*)k, (const float*)v, (const float*)out, (const float*)dout, lse,
                                                           delta, (float*)dq, dq_scale);
    attn_bwd_dkv_simt<float><<<g2, AS_WARPS * 32, sm2, s>>>(*a, (const float*)q, (const float*)k, (const float*)v, (const float*)dout, lse, delta,
                                                            (float*)dk, (float*)dv);
  }
  BPM_CHECK_LAUNCH("xattn_bwd_simt");
  return BPM_OK;
}

extern "C" int bpm_xattn_weights(const bpm_attn_t* a, const void* q, const void* k, const float* lse, float* w, void* stream) {
  int rc = check_attn(a);
  if (rc) return rc;
  BPM_REQUIRE(q && k && lse && w, "xattn_weights: null pointer");
  int64_t n = (int64_t)a->B * a->T * a->S;
  int grid = (int)((n + 255) / 256 < (int64_t)bpm_num_sms() * 16 ? (n + 255) / 256 : (int64_t)bpm_num_sms() * 16);
  if (a->dtype == BPM_BF16) attn_weights_simt<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>(*a, (const bf16*)q, (const bf16*)k, lse, w);
  else attn_weights_simt<float><<<grid, 256, 0, (cudaStream_t)stream>>>(*a, (const float*)q, (const float*)k, lse, w);
  BPM_CHECK_LAUNCH("xattn_weights");
  return BPM_OK;
}
