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
                   bf16* __restrict__ out, float* __restrict__ lse, int B, int T, int S, int H, int mask_off, bpm_dropout_t drop,
                   uint32_t* __restrict__ drop_bits) {
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
      // 4-way trees: a 64-long dependent FMNMX / FADD chain would leave the warp latency-bound
      float mx4[4] = {sv[0], sv[1], sv[2], sv[3]};
#pragma unroll
      for (int c = 4; c < AT_BN; c += 4) {
        mx4[0] = fmaxf(mx4[0], sv[c]); mx4[1] = fmaxf(mx4[1], sv[c + 1]); mx4[2] = fmaxf(mx4[2], sv[c + 2]); mx4[3] = fmaxf(mx4[3], sv[c + 3]);
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float m_new = fmaxf(m, mx * LOG2E_F);
      const float alpha = ex2f(m - m_new);
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < AT_BN; c += 4) {
#pragma unroll
        for (int e = 0; e < 4; e++) { sv[c + e] = ex2f(fmaf(sv[c + e], LOG2E_F, -m_new)); rs4[e] += sv[c + e]; }
      }
      l = fmaf(l, alpha, (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]));
      m = m_new;
      if (dc.on) {
        // keep decisions: one Philox call per 8 consecutive keys (element index e = ebase + key)
        uint32_t kb[2] = {0u, 0u};
        const uint64_t e0 = ebase + (uint64_t)k0;
        if ((e0 & 7) == 0) {
#pragma unroll
          for (int u = 0; u < AT_BN / 8; u++) kb[u >> 2] |= drop_keep8(dc, (e0 >> 3) + u) << ((u & 3) * 8);
        } else {
#pragma unroll
          for (int c = 0; c < AT_BN; c++) kb[c >> 5] |= (drop_mult1(dc, e0 + c) != 0.f ? 1u : 0u) << (c & 31);
        }
#pragma unroll
        for (int c = 0; c < AT_BN; c++) sv[c] = ((kb[c >> 5] >> (c & 31)) & 1u) ? sv[c] * dc.inv_keep : 0.f;
        if (drop_bits != nullptr && qi < T) {
          uint32_t* wrow = drop_bits + ((int64_t)bh * T + qi) * (int64_t)((S + 31) >> 5) + (k0 >> 5);
          wrow[0] = kb[0];
          if (k0 + 32 < S) wrow[1] = kb[1];
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
  attn_fwd_tc_kernel<<<grid, AT_THREADS, smem, stream>>>(tq, tk, tv, (bf16*)out, lse, a->B, a->T, a->S, a->H, a->mask_off, a->drop, a->drop_bits);
  BPM_CHECK_LAUNCH("xattn_fwd_tc");
  return BPM_OK;
}


// =====================================================================================================================
// Backward.  One CTA per (batch, head) for T <= 512: the dQ accumulators of all (<= 4) query tiles stay resident in TMEM for
// the whole kernel and dK / dV of the current 128-key tile are accumulated in TMEM over the inner query-tile loop, so NO
// atomics and no second pass are needed.  The score tile is computed TRANSPOSED (rows = keys, columns = queries):
//     S^T  = K_j Q_i^T          dP^T = V_j dO_i^T                               (M128 N128 K32, both operands K-major)
//     P^T  = exp2(S^T log2e - lse_q),   P~^T = P^T * dropmask,   dS^T = P^T * (dP^T * dropmask - delta_q)   [8 compute warps]
//     dV_j += P~^T dO_i         dK_j += dS^T Q_i      (A = the bf16 tile just written to smem, K-major; B = dO_i / Q_i MN-major)
//     dQ_i += dS K_j            (A = the SAME dS^T smem tile read as an MN-major operand; B = K_j MN-major)
// No row reductions are needed in the backward (lse and delta are per query = per column), so the 128x128 tile is split
// between two warps per TMEM lane quarter (64 columns each).  TMEM: S^T 128 + dP^T 128 + dK 32 + dV 32 + dQ 4x32 = 448 cols.
// =====================================================================================================================
#define AB_THREADS 320
#define AB_MAXQT 4

struct AttnBwdSmem {
  static constexpr int KV = 0;                                  // 2 stages x (K 8 KB + V 8 KB)
  static constexpr int QD_STAGE = 2 * 128 * 64 + 2048;          // Q 8 KB + dO 8 KB + dropout keep bits (128 queries x 4 words)
  static constexpr int QD = KV + 2 * 2 * 128 * 64;              // 2 stages
  static constexpr int PT = QD + 2 * QD_STAGE;                  // 2 chunk tiles x 16 KB
  static constexpr int DST = PT + 2 * 128 * 128;
  static constexpr int LSE = DST + 2 * 128 * 128;               // 512 floats (pre-multiplied by log2e; +inf beyond T)
  static constexpr int DEL = LSE + 512 * 4;
  static constexpr int BAR = DEL + 512 * 4;
  static constexpr int NBAR = 15;
  static constexpr int TOTAL = BAR + 8 * NBAR + 16;
};

__global__ void attn_delta_kernel(const bf16* __restrict__ out, const bf16* __restrict__ dout, float* __restrict__ delta, int B, int T, int H) {
  // delta[b, h, t] = sum_d dO * O   (one thread per (row, head): 2 x 64 B)
  int64_t n = (int64_t)B * T * H;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    int h = (int)(idx % H);
    int64_t row = idx / H;
    int t = (int)(row % T), b = (int)(row / T);
    const bf16* o = out + row * (H * AT_DH) + h * AT_DH;
    const bf16* g = dout + row * (H * AT_DH) + h * AT_DH;
    float acc = 0.f;
#pragma unroll
    for (int u = 0; u < AT_DH / 8; u++) {
      Vec8<bf16> a, c; a.load(o + u * 8); c.load(g + u * 8);
#pragma unroll
      for (int j = 0; j < 8; j++) acc = fmaf(a.v[j], c.v[j], acc);
    }
    delta[((int64_t)b * H + h) * T + t] = acc;
  }
}

__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const __grid_constant__ CUtensorMap tmdO, const __grid_constant__ CUtensorMap tmBits, const int bits_tma,
                   const float* __restrict__ lse, const float* __restrict__ delta, bf16* __restrict__ dq,
                   bf16* __restrict__ dk, bf16* __restrict__ dv, float dq_scale, int B, int T, int S, int H, int mask_off, bpm_dropout_t drop,
                   const uint32_t* __restrict__ drop_bits) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + AttnBwdSmem::BAR;
  auto kv_full = [&](int s) { return bar0 + 8u * s; };
  auto kv_empty = [&](int s) { return bar0 + 8u * (2 + s); };
  auto q_full = [&](int s) { return bar0 + 8u * (4 + s); };
  auto q_empty = [&](int s) { return bar0 + 8u * (6 + s); };
  const uint32_t st_full = bar0 + 8u * 8, st_free = bar0 + 8u * 9, pt_full = bar0 + 8u * 10, pair_done = bar0 + 8u * 11;
  const uint32_t dkv_full = bar0 + 8u * 12, dkv_free = bar0 + 8u * 13, dq_full = bar0 + 8u * 14;
  const uint32_t tmem_ptr_addr = bar0 + 8u * AttnBwdSmem::NBAR;
  volatile uint32_t* tmem_ptr_gen = (volatile uint32_t*)(base_gen + AttnBwdSmem::BAR + 8 * AttnBwdSmem::NBAR);
  float* lse_s = (float*)(base_gen + AttnBwdSmem::LSE);
  float* del_s = (float*)(base_gen + AttnBwdSmem::DEL);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  const int nq = (T + 127) / 128, nkv = (S + 127) / 128;
  // first query tile that can see key tile j (visible iff key <= q + off)
  auto i_min_of = [&](int j) { int qlo = j * 128 - mask_off; return (mask_off < 0 || qlo <= 0) ? 0 : qlo / 128; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
    for (int s = 0; s < 2; s++) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); mbar_init(q_full(s), 1); mbar_init(q_empty(s), 1); }
    mbar_init(st_full, 1); mbar_init(st_free, 8); mbar_init(pt_full, 8); mbar_init(pair_done, 1);
    mbar_init(dkv_full, 1); mbar_init(dkv_free, 8); mbar_init(dq_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_addr, 512);
  for (int t = threadIdx.x; t < 512; t += AB_THREADS) {
    lse_s[t] = t < T ? lse[(int64_t)bh * T + t] * LOG2E_F : INFINITY;      // exp2(x - inf) = 0 for the padded query columns
    del_s[t] = t < T ? delta[(int64_t)bh * T + t] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_gen;
  const uint32_t tST = tmem, tDPT = tmem + 128, tDK = tmem + 256, tDV = tmem + 288, tDQ = tmem + 320;

  if (warp == 0) {
    if (lane == 0) {
      int pc = 0, jc = 0;
      for (int j = 0; j < nkv; j++) {
        int imin = i_min_of(j);
        if (imin >= nq) continue;
        int ks = jc & 1;
        mbar_wait(kv_empty(ks), ((uint32_t)(jc >> 1) & 1u) ^ 1u);
        mbar_expect_tx(kv_full(ks), 2 * 128 * 64);
        tma_load_3d(base + AttnBwdSmem::KV + ks * 16384, &tmK, kv_full(ks), h * AT_DH, j * 128, b);
        tma_load_3d(base + AttnBwdSmem::KV + ks * 16384 + 8192, &tmV, kv_full(ks), h * AT_DH, j * 128, b);
        jc++;
        for (int i = imin; i < nq; i++, pc++) {
          int qs = pc & 1;
          mbar_wait(q_empty(qs), ((uint32_t)(pc >> 1) & 1u) ^ 1u);
          mbar_expect_tx(q_full(qs), 2 * 128 * 64 + (bits_tma ? 2048 : 0));
          tma_load_3d(base + AttnBwdSmem::QD + qs * AttnBwdSmem::QD_STAGE, &tmQ, q_full(qs), h * AT_DH, i * 128, b);
          tma_load_3d(base + AttnBwdSmem::QD + qs * AttnBwdSmem::QD_STAGE + 8192, &tmdO, q_full(qs), h * AT_DH, i * 128, b);
          if (bits_tma) tma_load_2d(base + AttnBwdSmem::QD + qs * AttnBwdSmem::QD_STAGE + 16384, &tmBits, q_full(qs), j * 4, bh * T + i * 128);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t id_st = umma_idesc_bf16(128, 128, 0, 0);      // S^T, dP^T
      const uint32_t id_kv = umma_idesc_bf16(128, AT_DH, 0, 1);    // dV, dK : A K-major (smem tile), B MN-major
      const uint32_t id_dq = umma_idesc_bf16(128, AT_DH, 1, 1);    // dQ     : A = dS^T read MN-major, B = K_j MN-major
      // pair iterator over (key tile j, query tile i); jc counts the non-empty key tiles (K/V ring index)
      auto next_pair = [&](int& j, int& i, int& jc) {
        i++;
        if (i >= nq) {
          do { j++; } while (j < nkv && i_min_of(j) >= nq);
          if (j < nkv) { i = i_min_of(j); jc++; }
        }
      };
      // S^T / dP^T of pair pc are issued one pair AHEAD of the accumulate MMAs, so the compute warps never wait for them
      auto issue_st = [&](int j, int i, int pc, int jc) {
        const int ks = jc & 1, qs = pc & 1;
        if (i == i_min_of(j)) mbar_wait(kv_full(ks), (uint32_t)(jc >> 1) & 1u);
        const uint32_t ka = base + AttnBwdSmem::KV + ks * 16384, va = ka + 8192;
        const uint32_t qa = base + AttnBwdSmem::QD + qs * AttnBwdSmem::QD_STAGE, ga = qa + 8192;
        mbar_wait(q_full(qs), (uint32_t)(pc >> 1) & 1u);
        mbar_wait(st_free, ((uint32_t)pc & 1u) ^ 1u);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 2; k++) {
          umma_bf16(tST, umma_desc(ka + k * 32, 16, 512, BPM_SWZ_64B), umma_desc(qa + k * 32, 16, 512, BPM_SWZ_64B), id_st, k > 0);
          umma_bf16(tDPT, umma_desc(va + k * 32, 16, 512, BPM_SWZ_64B), umma_desc(ga + k * 32, 16, 512, BPM_SWZ_64B), id_st, k > 0);
        }
        umma_commit(st_full);
      };
      int j = 0, jc = 0;
      while (j < nkv && i_min_of(j) >= nq) j++;
      if (j < nkv) {
        int i = i_min_of(j), pc = 0;
        issue_st(j, i, 0, 0);
        while (j < nkv) {
          int nj = j, ni = i, njc = jc;
          next_pair(nj, ni, njc);
          if (nj < nkv) issue_st(nj, ni, pc + 1, njc);
          const int ks = jc & 1, qs = pc & 1, imin = i_min_of(j);
          const uint32_t ka = base + AttnBwdSmem::KV + ks * 16384;
          const uint32_t qa = base + AttnBwdSmem::QD + qs * AttnBwdSmem::QD_STAGE, ga = qa + 8192;
          mbar_wait(pt_full, (uint32_t)pc & 1u);
          if (i == imin) mbar_wait(dkv_free, ((uint32_t)jc & 1u) ^ 1u);        // previous key tile's dK/dV drained from TMEM
          tc_fence_after();
          const uint32_t pa = base + AttnBwdSmem::PT, da = base + AttnBwdSmem::DST;
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const uint32_t acc = (i > imin || k > 0) ? 1u : 0u;
            const uint32_t aoff = (k >> 2) * 16384 + (k & 3) * 32;
            umma_bf16(tDV, umma_desc(pa + aoff, 16, 1024, BPM_SWZ_128B), umma_desc(ga + k * 1024, 512, 512, BPM_SWZ_64B), id_kv, acc);
            umma_bf16(tDK, umma_desc(da + aoff, 16, 1024, BPM_SWZ_128B), umma_desc(qa + k * 1024, 512, 512, BPM_SWZ_64B), id_kv, acc);
          }
#pragma unroll
          for (int k = 0; k < 8; k++)
            umma_bf16(tDQ + 32 * i, umma_desc(da + k * 2048, 16384, 1024, BPM_SWZ_128B), umma_desc(ka + k * 1024, 512, 512, BPM_SWZ_64B), id_dq,
                      (j > 0 || k > 0) ? 1u : 0u);
          umma_commit(pair_done);
          umma_commit(q_empty(qs));
          if (i == nq - 1) { umma_commit(dkv_full); umma_commit(kv_empty(ks)); }
          j = nj; i = ni; jc = njc; pc++;
        }
      }
      umma_commit(dq_full);
    }
  } else {
    // ===================== compute warps =====================
    const int cw = warp - 2;
    const int quarter = warp & 3, half = cw >> 2;
    const int r = quarter * 32 + lane;                      // key row inside the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    DropCtx dc = make_drop(drop);
    const int HP = H * AT_DH;
    const int W = (S + 31) >> 5;
    uint8_t* const prow = base_gen + AttnBwdSmem::PT + half * 16384 + (r >> 3) * 1024 + (r & 7) * 128;
    uint8_t* const drow = base_gen + AttnBwdSmem::DST + half * 16384 + (r >> 3) * 1024 + (r & 7) * 128;
    int pc = 0, jc = 0;
    for (int j = 0; j < nkv; j++) {
      const int key = j * 128 + r;
      int imin = i_min_of(j);
      if (imin >= nq) {                                     // no query sees this key tile: dK = dV = 0
        if (key < S) {
          bf16* dst = (half == 0 ? dk : dv) + ((int64_t)b * S + key) * HP + h * AT_DH;
#pragma unroll
          for (int u = 0; u < AT_DH / 8; u++) *(uint4*)(dst + u * 8) = make_uint4(0, 0, 0, 0);
        }
        continue;
      }
      for (int i = imin; i < nq; i++, pc++) {
        const int q0 = i * 128;
        const bool diag = (mask_off >= 0) && (j * 128 + 127 > q0 + mask_off);
        const int cmin = key - mask_off - q0;               // columns < cmin are masked for this key row (diag tiles only)
        // keep bits written by the forward: word (query, 32-key group).  Fast path: the TMA producer staged the 128 x 4-word tile of
        // this pair next to Q / dO (one broadcast LDS per column); otherwise each lane fetches the words of 2 of its 64 columns.
        const uint32_t* bits_s = (const uint32_t*)(base_gen + AttnBwdSmem::QD + (pc & 1) * AttnBwdSmem::QD_STAGE + 16384) + quarter;
        uint32_t mw[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
        if (dc.on && drop_bits != nullptr && !bits_tma) {
#pragma unroll
          for (int h2 = 0; h2 < 2; h2++) {
            const int qq = q0 + half * 64 + h2 * 32 + lane;
            mw[h2] = (qq < T && j * 4 + quarter < W) ? drop_bits[((int64_t)bh * T + qq) * W + j * 4 + quarter] : 0u;
          }
        }
        mbar_wait(st_full, (uint32_t)pc & 1u);        // (S^T was computed from this pair's Q stage, so its keep-bit tile has landed too)
        tc_fence_after();
#pragma unroll
        for (int ch = 0; ch < 2; ch++) {
          const int c0 = half * 64 + ch * 32;
          float sv[32], dpv[32];
          tmem_ld32(tST + lane_off + c0, sv);
          tmem_ld32(tDPT + lane_off + c0, dpv);
          tmem_ld_wait();
          if (ch == 1) {                                    // all of this warp's S^T / dP^T columns are in registers
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(st_free);
          }
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            const float4 l4 = *(const float4*)(lse_s + q0 + c0 + c);
            const float4 d4 = *(const float4*)(del_s + q0 + c0 + c);
            const float ls[4] = {l4.x, l4.y, l4.z, l4.w}, dl[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
              const int col = c0 + c + e;
              float p = ex2f(fmaf(sv[c + e], LOG2E_F, -ls[e]));
              if (key >= S || (diag && col < cmin)) p = 0.f;
              float mult = 1.f;
              if (dc.on) {
                if (bits_tma) {
                  mult = ((bits_s[col * 4] >> lane) & 1u) ? dc.inv_keep : 0.f;
                } else if (drop_bits != nullptr) {
                  const uint32_t wq = __shfl_sync(0xffffffffu, mw[ch], c + e);
                  mult = ((wq >> lane) & 1u) ? dc.inv_keep : 0.f;
                } else {
                  mult = drop_mult1(dc, ((uint64_t)bh * T + (uint64_t)min(q0 + col, T - 1)) * (uint64_t)S + (uint64_t)min(key, S - 1));
                }
              }
              sv[c + e] = p * mult;                                     // P~^T
              dpv[c + e] = p * fmaf(dpv[c + e], mult, -dl[e]);          // dS^T
            }
          }
          // the previous pair's accumulate MMAs must have finished reading the P^T / dS^T tiles before they are overwritten
          if (ch == 0) mbar_wait(pair_done, ((uint32_t)pc & 1u) ^ 1u);
          // this thread owns row r, columns [64*half + 32*ch, +32) = chunk tile `half`, 16-byte units 4*ch .. 4*ch+3
#pragma unroll
          for (int uu = 0; uu < 4; uu++) {
            const int u = ch * 4 + uu;
            *(uint4*)(prow + ((u ^ (r & 7)) << 4)) = make_uint4(pack_bf16x2(sv[uu * 8], sv[uu * 8 + 1]), pack_bf16x2(sv[uu * 8 + 2], sv[uu * 8 + 3]),
                                                               pack_bf16x2(sv[uu * 8 + 4], sv[uu * 8 + 5]), pack_bf16x2(sv[uu * 8 + 6], sv[uu * 8 + 7]));
            *(uint4*)(drow + ((u ^ (r & 7)) << 4)) = make_uint4(pack_bf16x2(dpv[uu * 8], dpv[uu * 8 + 1]), pack_bf16x2(dpv[uu * 8 + 2], dpv[uu * 8 + 3]),
                                                               pack_bf16x2(dpv[uu * 8 + 4], dpv[uu * 8 + 5]), pack_bf16x2(dpv[uu * 8 + 6], dpv[uu * 8 + 7]));
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(pt_full);
      }
      // ---- dK (half 0) / dV (half 1) of key tile j
      mbar_wait(dkv_full, (uint32_t)jc & 1u);
      tc_fence_after();
      {
        float acc[AT_DH];
        tmem_ld32((half == 0 ? tDK : tDV) + lane_off, acc);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dkv_free);
        if (key < S) {
          bf16* dst = (half == 0 ? dk : dv) + ((int64_t)b * S + key) * HP + h * AT_DH;
#pragma unroll
          for (int u = 0; u < AT_DH / 8; u++)
            *(uint4*)(dst + u * 8) = make_uint4(pack_bf16x2(acc[u * 8], acc[u * 8 + 1]), pack_bf16x2(acc[u * 8 + 2], acc[u * 8 + 3]),
                                                pack_bf16x2(acc[u * 8 + 4], acc[u * 8 + 5]), pack_bf16x2(acc[u * 8 + 6], acc[u * 8 + 7]));
        }
      }
      jc++;
    }
    // ---- dQ of every query tile (tile i is drained by the warps of half i & 1)
    mbar_wait(dq_full, 0);
    tc_fence_after();
    for (int i = half; i < nq; i += 2) {
      float acc[AT_DH];
      tmem_ld32(tDQ + 32 * i + lane_off, acc);
      tmem_ld_wait();
      const int qi = i * 128 + r;
      if (qi < T) {
        bf16* dst = dq + ((int64_t)b * T + qi) * HP + h * AT_DH;
#pragma unroll
        for (int u = 0; u < AT_DH / 8; u++)
          *(uint4*)(dst + u * 8) = make_uint4(pack_bf16x2(acc[u * 8] * dq_scale, acc[u * 8 + 1] * dq_scale), pack_bf16x2(acc[u * 8 + 2] * dq_scale, acc[u * 8 + 3] * dq_scale),
                                              pack_bf16x2(acc[u * 8 + 4] * dq_scale, acc[u * 8 + 5] * dq_scale), pack_bf16x2(acc[u * 8 + 6] * dq_scale, acc[u * 8 + 7] * dq_scale));
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

int bpm_xattn_bwd_simt(const bpm_attn_t* a, const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse,
                       float* delta, void* dq, float dq_scale, void* dk, void* dv, cudaStream_t s);

int bpm_xattn_bwd_tc(const bpm_attn_t* a, const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse,
                     float* delta, void* dq, float dq_scale, void* dk, void* dv, cudaStream_t stream) {
  if (a->T > 128 * AB_MAXQT)      // dQ accumulators of all query tiles must fit in TMEM; longer targets use the fp32-math kernel
    return bpm_xattn_bwd_simt(a, q, k, v, out, dout, lse, delta, dq, dq_scale, dk, dv, stream);
  BPM_REQUIRE(((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out | (uintptr_t)dout | (uintptr_t)dq | (uintptr_t)dk | (uintptr_t)dv) % 16 == 0,
              "xattn_bwd: pointers must be 16-byte aligned");
  const int HP = a->H * a->dhp;
  {
    int64_t n = (int64_t)a->B * a->T * a->H;
    int grid = (int)((n + 255) / 256 < (int64_t)bpm_num_sms() * 8 ? (n + 255) / 256 : (int64_t)bpm_num_sms() * 8);
    attn_delta_kernel<<<grid, 256, 0, stream>>>((const bf16*)out, (const bf16*)dout, delta, a->B, a->T, a->H);
    BPM_CHECK_LAUNCH("xattn_delta");
  }
  CUtensorMap tq, tk, tv, tg;
  int rc;
  if ((rc = make_qkv_map(&tq, q, a->B, a->T, HP, 128))) return rc;
  if ((rc = make_qkv_map(&tk, k, a->B, a->S, HP, 128))) return rc;
  if ((rc = make_qkv_map(&tv, v, a->B, a->S, HP, 128))) return rc;
  if ((rc = make_qkv_map(&tg, dout, a->B, a->T, HP, 128))) return rc;
  // dropout keep bits as a [B*H*T, W] uint32 tensor, box {4 words, 128 queries}; needs a 16-byte pitch (S % 128 == 0)
  CUtensorMap tb = tq;
  const int W = (a->S + 31) / 32;
  const int bits_tma = (a->drop.p > 0.f && a->drop_bits != nullptr && a->S % 128 == 0) ? 1 : 0;
  if (bits_tma) {
    bpm_encode_tiled_fn enc = bpm_get_encode_tiled();
    cuuint64_t gd[2] = {(cuuint64_t)W, (cuuint64_t)a->B * a->H * a->T}, gs[1] = {(cuuint64_t)W * 4};
    cuuint32_t bx[2] = {4, 128}, es[2] = {1, 1};
    CUresult r = enc(&tb, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void*)a->drop_bits, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    BPM_REQUIRE(r == CUDA_SUCCESS, "xattn_bwd: tensor map for the dropout bits failed (%d)", (int)r);
  }
  size_t smem = AttnBwdSmem::TOTAL + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { bpm_set_error("xattn_bwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return BPM_ELAUNCH; }
    attr_set = true;
  }
  attn_bwd_tc_kernel<<<a->B * a->H, AB_THREADS, smem, stream>>>(tq, tk, tv, tg, tb, bits_tma, lse, delta, (bf16*)dq, (bf16*)dk, (bf16*)dv, dq_scale, a->B, a->T, a->S,
                                                                  a->H, a->mask_off, a->drop, a->drop_bits);
  BPM_CHECK_LAUNCH("xattn_bwd_tc");
  return BPM_OK;
}
