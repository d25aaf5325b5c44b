// Crossmodal flash attention on tcgen05 tensor cores (sm_100a), head dim 25 padded to 32, bf16 in / fp32 accumulate.
// Reference semantics: models/multihead_attention.py:95-127 (scores, fp32 softmax, dropout, PV) with the mask of
// models/transformer.py:209-216 evaluated from indices (key j visible to query i iff j <= i + mask_off).
//
// Forward, one CTA per (128-query tile, batch*head):
//   warp 0      TMA producer: Q tile once, K/V tiles (64 keys) through a 3-stage ring          [SWIZZLE_64B boxes, 64-byte rows]
//   warp 1      TMEM allocator + single-thread MMA issuer:  S_j = Q K_j^T  (M128 N64 K32)  and  O_j = P_j V_j  (M128 N32 K64)
//   warps 2-5   softmax, one thread per query row (= TMEM lane): tcgen05.ld S_j -> online softmax in registers (exp2, running
//               max / sum) -> Philox dropout -> P_j as bf16 into shared memory in the canonical K-major SWIZZLE_128B layout ->
//               tcgen05.ld O_j (32 columns) and accumulate O in registers with the usual rescale.
//   S is double-buffered in TMEM and P in shared memory, so the MMA of tile j+1 overlaps the softmax of tile j; two CTAs are
//   co-resident per SM (192 TMEM columns, ~66 KB smem each).  With dh = 32 the kernel is bound by exponentials (128x64 ex2 per
//   tile on the 16/clk/SM MUFU vs 64+64 MMA cycles), see DESIGN.md.
//   V is consumed directly from its [keys, dh] tile as an MN-major B operand (no transpose).  Fully masked KV tiles are skipped.
#include "tc_common.cuh"

#define AT_BM 128
#define AT_BN 64
#define AT_DH 32
#define AT_KV_STAGES 3
#define AT_THREADS 192
#define LOG2E_F 1.4426950408889634f
#define LN2_F 0.6931471805599453f

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *(uint32_t*)&v;
}

struct AttnFwdSmem {
  // offsets from the 1024-aligned base
  static constexpr int Q = 0;                                   // 128 x 64 B
  static constexpr int K = Q + AT_BM * 64;                      // stages x 64 x 64 B
  static constexpr int V = K + AT_KV_STAGES * AT_BN * 64;
  static constexpr int P = V + AT_KV_STAGES * AT_BN * 64;       // 2 x (128 rows x 128 B)
  static constexpr int BAR = P + 2 * AT_BM * 128;
  static constexpr int NBAR = 1 + 2 * AT_KV_STAGES + 2 + 2 + 2 + 2;
  static constexpr int TOTAL = BAR + 8 * NBAR + 16;
};

__global__ void __launch_bounds__(AT_THREADS, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   bf16* __restrict__ out, float* __restrict__ lse, int B, int T, int S, int H, int mask_off, bpm_dropout_t drop) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + AttnFwdSmem::BAR;
  const uint32_t q_full = bar0;
  auto kv_full = [&](int s) { return bar0 + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar0 + 8u * (1 + AT_KV_STAGES + s); };
  auto s_full = [&](int i) { return bar0 + 8u * (1 + 2 * AT_KV_STAGES + i); };
  auto s_empty = [&](int i) { return bar0 + 8u * (3 + 2 * AT_KV_STAGES + i); };
  auto p_full = [&](int i) { return bar0 + 8u * (5 + 2 * AT_KV_STAGES + i); };
  auto o_full = [&](int i) { return bar0 + 8u * (7 + 2 * AT_KV_STAGES + i); };
  const uint32_t tmem_ptr_addr = bar0 + 8u * AttnFwdSmem::NBAR;
  volatile uint32_t* tmem_ptr_gen = (volatile uint32_t*)(base_gen + AttnFwdSmem::BAR + 8 * AttnFwdSmem::NBAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = gridDim.x - 1 - blockIdx.x;          // heavy (late) query tiles first
  const int q0 = qt * AT_BM;
  const int bh = blockIdx.y, b = bh / H, h = bh % H;
  // KV tiles this query tile can see
  int jmax = S - 1;
  if (mask_off >= 0) jmax = min(jmax, q0 + AT_BM - 1 + mask_off);
  const int n_tiles = jmax / AT_BN + 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < AT_KV_STAGES; s++) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    for (int i = 0; i < 2; i++) { mbar_init(s_full(i), 1); mbar_init(s_empty(i), 4); mbar_init(p_full(i), 4); mbar_init(o_full(i), 1); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_addr, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_gen;
  const uint32_t tS[2] = {tmem + 0, tmem + 64};
  const uint32_t tO[2] = {tmem + 128, tmem + 160};

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, AT_BM * 64);
      tma_load_3d(base + AttnFwdSmem::Q, &tmQ, q_full, h * AT_DH, q0, b);
      for (int j = 0; j < n_tiles; j++) {
        int s = j % AT_KV_STAGES;
        mbar_wait(kv_empty(s), ((uint32_t)(j / AT_KV_STAGES) & 1u) ^ 1u);
        mbar_expect_tx(kv_full(s), 2 * AT_BN * 64);
        tma_load_3d(base + AttnFwdSmem::K + s * AT_BN * 64, &tmK, kv_full(s), h * AT_DH, j * AT_BN, b);
        tma_load_3d(base + AttnFwdSmem::V + s * AT_BN * 64, &tmV, kv_full(s), h * AT_DH, j * AT_BN, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(AT_BM, AT_BN, 0, 0);       // S = Q K^T : both operands K-major
      const uint32_t idesc_o = umma_idesc_bf16(AT_BM, AT_DH, 0, 1);       // O = P V   : V is MN-major (dh contiguous)
      mbar_wait(q_full, 0);
      auto issue_s = [&](int j) {
        int s = j % AT_KV_STAGES, i = j & 1;
        mbar_wait(kv_full(s), (uint32_t)(j / AT_KV_STAGES) & 1u);
        mbar_wait(s_empty(i), ((uint32_t)(j >> 1) & 1u) ^ 1u);
        tc_fence_after();
        uint32_t qa = base + AttnFwdSmem::Q, ka = base + AttnFwdSmem::K + s * AT_BN * 64;
#pragma unroll
        for (int k = 0; k < AT_DH / 16; k++)
          umma_bf16(tS[i], umma_desc(qa + k * 32, 16, 512, BPM_SWZ_64B), umma_desc(ka + k * 32, 16, 512, BPM_SWZ_64B), idesc_s, k > 0);
        umma_commit(s_full(i));
      };
      auto issue_pv = [&](int j) {
        int s = j % AT_KV_STAGES, i = j & 1;
        mbar_wait(p_full(i), (uint32_t)(j >> 1) & 1u);
        tc_fence_after();
        uint32_t pa = base + AttnFwdSmem::P + i * AT_BM * 128, va = base + AttnFwdSmem::V + s * AT_BN * 64;
#pragma unroll
        for (int k = 0; k < AT_BN / 16; k++)
          umma_bf16(tO[i], umma_desc(pa + k * 32, 16, 1024, BPM_SWZ_128B), umma_desc(va + k * 16 * 64, 512, 512, BPM_SWZ_64B), idesc_o, k > 0);
        umma_commit(o_full(i));        // O_j ready (and P buffer i reusable)
        umma_commit(kv_empty(s));      // K/V stage s free
      };
      issue_s(0);
      for (int j = 0; j < n_tiles; j++) {
        if (j + 1 < n_tiles) issue_s(j + 1);
        issue_pv(j);
      }
    }
  } else {
    // ===================== softmax warps: one thread per query row =====================
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int qi = q0 + r;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    DropCtx dc = make_drop(drop);
    const uint64_t ebase = ((uint64_t)bh * T + (uint64_t)min(qi, T - 1)) * (uint64_t)S;
    float m = -INFINITY, l = 0.f;
    float oacc[AT_DH];
#pragma unroll
    for (int d = 0; d < AT_DH; d++) oacc[d] = 0.f;
    const int row_lim = (mask_off >= 0) ? min(qi + mask_off, S - 1) : S - 1;     // last visible key of this row
    uint8_t* prow[2];
    prow[0] = base_gen + AttnFwdSmem::P + (r >> 3) * 1024 + (r & 7) * 128;
    prow[1] = prow[0] + AT_BM * 128;

    for (int j = 0; j < n_tiles; j++) {
      const int i = j & 1;
      float sv[AT_BN];
      mbar_wait(s_full(i), (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      tmem_ld32(tS[i] + lane_off, sv);
      tmem_ld32(tS[i] + lane_off + 32, sv + 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty(i));
      // ---- online softmax (log2 domain)
      const int k0 = j * AT_BN;
      if (k0 + AT_BN - 1 > row_lim) {
#pragma unroll
        for (int c = 0; c < AT_BN; c++) sv[c] = (k0 + c <= row_lim) ? sv[c] : -INFINITY;
      }
      float mx = sv[0];
#pragma unroll
      for (int c = 1; c < AT_BN; c++) mx = fmaxf(mx, sv[c]);
      const float m_new = fmaxf(m, mx * LOG2E_F);
      const float alpha = ex2f(m - m_new);
      float rs = 0.f;
#pragma unroll
      for (int c = 0; c < AT_BN; c++) { sv[c] = ex2f(fmaf(sv[c], LOG2E_F, -m_new)); rs += sv[c]; }
      l = fmaf(l, alpha, rs);
      m = m_new;
      if (dc.on) {
        // Philox words are shared by 4 consecutive element indices e = ebase + key
#pragma unroll
        for (int c = 0; c < AT_BN; c++) {
          uint64_t e = ebase + (uint64_t)(k0 + c);
          sv[c] *= drop_mult1(dc, e);
        }
      }
      // ---- P_j -> shared memory (bf16, K-major SWIZZLE_128B: 16-byte chunk u of row r lands at chunk u ^ (r & 7))
      uint8_t* pr = prow[i];
#pragma unroll
      for (int u = 0; u < AT_BN / 8; u++) {
        uint4 w;
        w.x = pack_bf16x2(sv[u * 8 + 0], sv[u * 8 + 1]); w.y = pack_bf16x2(sv[u * 8 + 2], sv[u * 8 + 3]);
        w.z = pack_bf16x2(sv[u * 8 + 4], sv[u * 8 + 5]); w.w = pack_bf16x2(sv[u * 8 + 6], sv[u * 8 + 7]);
        *(uint4*)(pr + ((u ^ (r & 7)) << 4)) = w;
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(i));
      // ---- fold the previous tile's P V into the register accumulator, then rescale to the new max
      if (j > 0) {
        const int ip = (j - 1) & 1;
        float ov[AT_DH];
        mbar_wait(o_full(ip), (uint32_t)((j - 1) >> 1) & 1u);
        tc_fence_after();
        tmem_ld32(tO[ip] + lane_off, ov);
        tmem_ld_wait();
        tc_fence_before();
#pragma unroll
        for (int d = 0; d < AT_DH; d++) oacc[d] = (oacc[d] + ov[d]) * alpha;
      }
    }
    {
      const int ip = (n_tiles - 1) & 1;
      float ov[AT_DH];
      mbar_wait(o_full(ip), (uint32_t)((n_tiles - 1) >> 1) & 1u);
      tc_fence_after();
      tmem_ld32(tO[ip] + lane_off, ov);
      tmem_ld_wait();
      tc_fence_before();
      const float inv_l = 1.f / l;
#pragma unroll
      for (int d = 0; d < AT_DH; d++) oacc[d] = (oacc[d] + ov[d]) * inv_l;
    }
    if (qi < T) {
      bf16* orow = out + ((int64_t)b * T + qi) * (int64_t)(H * AT_DH) + h * AT_DH;
#pragma unroll
      for (int u = 0; u < AT_DH / 8; u++) {
        uint4 w;
        w.x = pack_bf16x2(oacc[u * 8 + 0], oacc[u * 8 + 1]); w.y = pack_bf16x2(oacc[u * 8 + 2], oacc[u * 8 + 3]);
        w.z = pack_bf16x2(oacc[u * 8 + 4], oacc[u * 8 + 5]); w.w = pack_bf16x2(oacc[u * 8 + 6], oacc[u * 8 + 7]);
        *(uint4*)(orow + u * 8) = w;
      }
      lse[(int64_t)bh * T + qi] = (m + log2f(l)) * LN2_F;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

// ---------------------------------------------------------------- host
static int make_qkv_map(CUtensorMap* m, const void* p, int B, int rows, int HP, int box_rows) {
  uint64_t dims[3] = {(uint64_t)HP, (uint64_t)rows, (uint64_t)B};
  uint64_t str[2] = {(uint64_t)HP * 2, (uint64_t)rows * HP * 2};
  uint32_t box[3] = {AT_DH, (uint32_t)box_rows, 1};
  return bpm_make_tmap_bf16(m, p, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B);
}

int bpm_xattn_tc_supported(const bpm_attn_t* a) {
  return a->dtype == BPM_BF16 && a->dhp == AT_DH && a->key_pad == nullptr;
}

int bpm_xattn_fwd_tc(const bpm_attn_t* a, const void* q, const void* k, const void* v, void* out, float* lse, cudaStream_t stream) {
  BPM_REQUIRE(((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) % 16 == 0, "xattn_fwd: pointers must be 16-byte aligned");
  const int HP = a->H * a->dhp;
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_qkv_map(&tq, q, a->B, a->T, HP, AT_BM))) return rc;
  if ((rc = make_qkv_map(&tk, k, a->B, a->S, HP, AT_BN))) return rc;
  if ((rc = make_qkv_map(&tv, v, a->B, a->S, HP, AT_BN))) return rc;
  size_t smem = AttnFwdSmem::TOTAL + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { bpm_set_error("xattn_fwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return BPM_ELAUNCH; }
    attr_set = true;
  }
  dim3 grid(bpm_cdiv(a->T, AT_BM), a->B * a->H);
  attn_fwd_tc_kernel<<<grid, AT_THREADS, smem, stream>>>(tq, tk, tv, (bf16*)out, lse, a->B, a->T, a->S, a->H, a->mask_off, a->drop);
  BPM_CHECK_LAUNCH("xattn_fwd_tc");
  return BPM_OK;
}

int bpm_xattn_bwd_simt(const bpm_attn_t* a, const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse,
                       float* delta, void* dq, float dq_scale, void* dk, void* dv, cudaStream_t s);
int bpm_xattn_bwd_tc(const bpm_attn_t* a, const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse,
                     float* delta, void* dq, float dq_scale, void* dk, void* dv, cudaStream_t s) {
  // the tensor-core backward lands next; until then the fp32-math kernel (bf16 storage) serves this entry point
  return bpm_xattn_bwd_simt(a, q, k, v, out, dout, lse, delta, dq, dq_scale, dk, dv, s);
}
