// C-ABI dispatch: picks the tensor-core (bf16) or exact-fp32 kernel for GEMM and attention.
// There is no CPU path: every entry point launches CUDA kernels on the caller's stream.
#include <stdlib.h>
#include <mutex>
#include <set>
#include <utility>
#include "bpm_common.cuh"

int bpm_func_smem(const void* func, int bytes, const char* what) {
  static std::mutex mu;
  static std::set<std::pair<const void*, int>> done;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); bpm_set_error("%s: cudaGetDevice failed", what); return BPM_ELAUNCH; }
  std::lock_guard<std::mutex> lk(mu);
  const auto key = std::make_pair(func, dev);
  if (done.count(key)) return BPM_OK;
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) { bpm_set_error("%s: cudaFuncSetAttribute(%d B): %s", what, bytes, cudaGetErrorString(e)); (void)cudaGetLastError(); return BPM_ELAUNCH; }
  done.insert(key);
  return BPM_OK;
}

int bpm_gemm_simt(const bpm_gemm_t* g, cudaStream_t stream);
int bpm_gemm_tc(const bpm_gemm_t* g, cudaStream_t stream);
int bpm_xattn_fwd_simt(const bpm_attn_t* a, const void* q, const void* k, const void* v, void* out, float* lse, cudaStream_t s);
int bpm_xattn_bwd_simt(const bpm_attn_t* a, const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse,
                       float* delta, void* dq, float dq_scale, void* dk, void* dv, cudaStream_t s);
int bpm_xattn_fwd_tc(const bpm_attn_t* a, const void* q, const void* k, const void* v, void* out, float* lse, cudaStream_t s);
int bpm_xattn_bwd_tc(const bpm_attn_t* a, const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse,
                     float* delta, void* dq, float dq_scale, void* dk, void* dv, cudaStream_t s);
int bpm_xattn_tc_supported(const bpm_attn_t* a);
// head dim 64 / 128 (attn_tc128.cu)
int bpm_xattn128_supported(const bpm_attn_t* a, int backward);
int64_t bpm_xattn128_ws_floats(const bpm_attn_t* a);
int bpm_xattn_fwd_tc128(const bpm_attn_t* a, const void* q, const void* k, const void* v, void* out, float* lse, cudaStream_t s);
int bpm_xattn_bwd_tc128(const bpm_attn_t* a, const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse,
                        float* ws, void* dq, float dq_scale, void* dk, void* dv, cudaStream_t s);
extern "C" int bpm_colsum(const void* X, int dtype, int M, int N, int ld, float* out, void* stream);

// diagnostic knobs for kernel bring-up / profiling (scripts/ only; the product never sets them): slot 0 = GEMM, 1 = attention
static int g_dbg[4] = {0, 0, 0, 0};
int bpm_debug_get(int slot) { return g_dbg[slot & 3]; }
extern "C" int bpm_debug_set(int slot, int value) { g_dbg[slot & 3] = value; return BPM_OK; }
static void* g_dbg_ptr = nullptr;
void* bpm_debug_get_ptr() { return g_dbg_ptr; }
extern "C" int bpm_debug_set_ptr(void* p) { g_dbg_ptr = p; return BPM_OK; }

int bpm_pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("BPM_NO_PDL"); v = (e && e[0] == '1') ? 0 : 1; }
  return v;
}

// BPM_DEBUG_FFMA=1 routes bf16 problems through the FFMA kernels too (kernel bring-up / bisecting only).
static int debug_ffma() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("BPM_DEBUG_FFMA"); v = (e && e[0] == '1') ? 1 : 0; }
  return v;
}

extern "C" int bpm_gemm(const bpm_gemm_t* g, void* stream) {
  BPM_REQUIRE(g && g->A && g->B && g->C && g->M > 0 && g->N > 0 && g->K > 0, "gemm: bad args");
  BPM_REQUIRE(g->ab_dtype == BPM_F32 || g->ab_dtype == BPM_BF16, "gemm: bad dtype");
  BPM_REQUIRE(!g->accumulate || g->c_dtype == BPM_F32, "gemm: accumulate needs fp32 C");
  // tiny-M problems (the [B, D] head) are >90% tile padding on a 128-row MMA: they stay on the FFMA kernel
  BPM_REQUIRE(!g->colsum_out || g->ta == 1, "gemm: colsum_out needs ta = 1");
  if (g->ab_dtype == BPM_F32 || debug_ffma()) {
    if (g->colsum_out) {      // op(A) = A^T with A stored [K, M]: row sums of op(A) = column sums of the stored matrix
      int rc = bpm_colsum(g->A, g->ab_dtype, g->K, g->M, g->lda, g->colsum_out, stream);
      if (rc) return rc;
    }
    return bpm_gemm_simt(g, (cudaStream_t)stream);
  }
  return bpm_gemm_tc(g, (cudaStream_t)stream);
}

extern "C" int bpm_xattn_fwd(const bpm_attn_t* a, const void* q, const void* k, const void* v, void* out, float* lse, void* stream) {
  BPM_REQUIRE(a && q && k && v && out && lse, "xattn_fwd: null pointer");
  if (a->dtype == BPM_BF16 && !debug_ffma()) {
    if (bpm_xattn_tc_supported(a)) return bpm_xattn_fwd_tc(a, q, k, v, out, lse, (cudaStream_t)stream);
    if (bpm_xattn128_supported(a, 0) && !(bpm_debug_get(1) & 65536)) return bpm_xattn_fwd_tc128(a, q, k, v, out, lse, (cudaStream_t)stream);
  }
  return bpm_xattn_fwd_simt(a, q, k, v, out, lse, (cudaStream_t)stream);
}

extern "C" int bpm_xattn_bwd(const bpm_attn_t* a, const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse,
                             float* delta, void* dq, float dq_scale, void* dk, void* dv, void* stream) {
  BPM_REQUIRE(a && q && k && v && out && dout && lse && delta && dq && dk && dv, "xattn_bwd: null pointer");
  if (a->dtype == BPM_BF16 && !debug_ffma()) {
    if (bpm_xattn_tc_supported(a)) return bpm_xattn_bwd_tc(a, q, k, v, out, dout, lse, delta, dq, dq_scale, dk, dv, (cudaStream_t)stream);
    if (bpm_xattn128_supported(a, 1) && !(bpm_debug_get(1) & 131072))
      return bpm_xattn_bwd_tc128(a, q, k, v, out, dout, lse, delta, dq, dq_scale, dk, dv, (cudaStream_t)stream);
  }
  return bpm_xattn_bwd_simt(a, q, k, v, out, dout, lse, delta, dq, dq_scale, dk, dv, (cudaStream_t)stream);
}

extern "C" int64_t bpm_xattn_bwd_workspace(const bpm_attn_t* a) {
  if (!a) return 0;
  const int64_t base = 2 * (int64_t)a->B * a->H * a->T;
  if (a->dtype == BPM_BF16 && !debug_ffma() && !bpm_xattn_tc_supported(a) && bpm_xattn128_supported(a, 1)) return bpm_xattn128_ws_floats(a);
  return base;
}
