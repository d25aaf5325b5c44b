// Fused GEMM epilogue shared by the tcgen05 (bf16) and FFMA (fp32) GEMM kernels.
//   v = (acc + bias[n]) * alpha; relu; dropout; gate (relu-backward mask from a saved activation); + residual;
//   store or atomically accumulate.   See bpm_gemm_t in include/bpmult_b200.h.
#pragma once
#include "bpm_common.cuh"

struct EpiParams {
  void* C; int ldc; int c_dtype;
  const float* bias; float alpha; int act;
  bpm_dropout_t drop;
  const void* gate; int ldg; int gate_dtype; float gate_scale;
  const void* residual; int ldr; int res_dtype;
  int accumulate;
  int M, N;
};

static inline EpiParams make_epi(const bpm_gemm_t* g) {
  EpiParams e;
  e.C = g->C; e.ldc = g->ldc; e.c_dtype = g->c_dtype;
  e.bias = g->bias; e.alpha = g->alpha; e.act = g->act;
  e.drop = g->drop;
  e.gate = g->gate; e.ldg = g->ldg; e.gate_dtype = g->gate_dtype; e.gate_scale = g->gate_scale;
  e.residual = g->residual; e.ldr = g->ldr; e.res_dtype = g->res_dtype;
  e.accumulate = g->accumulate;
  e.M = g->M; e.N = g->N;
  return e;
}

// scalar epilogue for one element (used by the FFMA kernel and by ragged edges)
__device__ __forceinline__ void epi_store1(const EpiParams& p, const DropCtx& dc, int m, int n, float acc) {
  float v = acc;
  if (p.bias) v += p.bias[n];
  v *= p.alpha;
  if (p.act == 1) v = fmaxf(v, 0.f);
  if (dc.on) v *= drop_mult1(dc, (uint64_t)m * (uint64_t)p.ldc + (uint64_t)n);
  if (p.gate) v = ld_as_f(p.gate, p.gate_dtype, (int64_t)m * p.ldg + n) > 0.f ? v * p.gate_scale : 0.f;
  if (p.residual) v += ld_as_f(p.residual, p.res_dtype, (int64_t)m * p.ldr + n);
  if (p.accumulate) atomicAdd((float*)p.C + (int64_t)m * p.ldc + n, v);
  else st_from_f(p.C, p.c_dtype, (int64_t)m * p.ldc + n, v);
}

// 8 consecutive columns n0..n0+7 of row m (n0 % 8 == 0, all in range, pitches multiples of 8): vector loads/stores
__device__ __forceinline__ void epi_store8(const EpiParams& p, const DropCtx& dc, int m, int n0, float* acc) {
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; j++) v[j] = acc[j];
  if (p.bias) {
    Vec8<float> b; b.load(p.bias + n0);
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] += b.v[j];
  }
#pragma unroll
  for (int j = 0; j < 8; j++) v[j] *= p.alpha;
  if (p.act == 1) {
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = fmaxf(v[j], 0.f);
  }
  if (dc.on) {
    float mlt[8];
    drop_mult8(dc, (uint64_t)m * (uint64_t)p.ldc + (uint64_t)n0, mlt);
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] *= mlt[j];
  }
  if (p.gate) {
    float gt[8];
    if (p.gate_dtype == BPM_BF16) { Vec8<bf16> t; t.load((const bf16*)p.gate + (int64_t)m * p.ldg + n0);
#pragma unroll
      for (int j = 0; j < 8; j++) gt[j] = t.v[j]; }
    else { Vec8<float> t; t.load((const float*)p.gate + (int64_t)m * p.ldg + n0);
#pragma unroll
      for (int j = 0; j < 8; j++) gt[j] = t.v[j]; }
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = gt[j] > 0.f ? v[j] * p.gate_scale : 0.f;
  }
  if (p.residual) {
    if (p.res_dtype == BPM_BF16) { Vec8<bf16> t; t.load((const bf16*)p.residual + (int64_t)m * p.ldr + n0);
#pragma unroll
      for (int j = 0; j < 8; j++) v[j] += t.v[j]; }
    else { Vec8<float> t; t.load((const float*)p.residual + (int64_t)m * p.ldr + n0);
#pragma unroll
      for (int j = 0; j < 8; j++) v[j] += t.v[j]; }
  }
  if (p.accumulate) {
    float* c = (float*)p.C + (int64_t)m * p.ldc + n0;
#pragma unroll
    for (int j = 0; j < 8; j++) atomicAdd(c + j, v[j]);
  } else if (p.c_dtype == BPM_BF16) {
    Vec8<bf16> o;
#pragma unroll
    for (int j = 0; j < 8; j++) o.v[j] = v[j];
    o.store((bf16*)p.C + (int64_t)m * p.ldc + n0);
  } else {
    Vec8<float> o;
#pragma unroll
    for (int j = 0; j < 8; j++) o.v[j] = v[j];
    o.store((float*)p.C + (int64_t)m * p.ldc + n0);
  }
}
