// Crossmodal flash attention on tcgen05 tensor cores (sm_100a), head dim 25 padded to 32, bf16 in / fp32 accumulate.
// Reference semantics: models/multihead_attention.py:95-127 (scores, fp32 softmax, dropout, PV) with the mask of
// models/transformer.py:209-216 evaluated from indices (key j visible to query i iff j <= i + mask_off).
//
// Forward.  Persistent: one CTA per SM, two independent GROUPS per CTA; each group walks its own list of (batch*head, 128-query
// tile) items (heavy = late query tiles first, snake order over the groups) and owns its own warps, shared-memory rings and
// 256 TMEM columns.  While one group's softmax warps wait for the tensor core, the other group's keep the MUFU / FMA pipes busy.
//   per group:  1 TMA producer warp  (Q double-buffered, K/V tiles of 128 keys through a 3-stage ring; SWIZZLE_64B boxes)
//               1 MMA issuer warp    S_j = Q K_j^T (M128 N128 K32)  and  O_j = P_j V_j (M128 N32 K128, A = P_j FROM TMEM)
//               4 softmax warps      one thread per query row (= TMEM lane): tcgen05.ld S_j twice (row max, then exp2 / row sum),
//                                    P_j back to TMEM as packed bf16 (tcgen05.st), O_{j-1} folded into registers with the rescale.
//   Converged issuer warps + elect.sync, descriptors advanced by adds (see tc_common.cuh: elect_one).
//   With dh = 32 the kernel is bound by exponentials (128x128 ex2 per tile on the 16/clk/SM MUFU = 1024 clk vs ~410 MMA clk).
//   V is consumed directly from its [keys, dh] tile as an MN-major B operand (no transpose).  Fully masked KV tiles are skipped.
#include "tc_common.cuh"
#include "attn_math.cuh"

#define AT_BM 128
#define AT_BN 128
#define AT_DH 32
#define AT_KV_STAGES 4
#define AF_GROUPS 2
#define AF_THREADS(SW) (32 * AF_GROUPS * (2 + (SW)))     // per group: 1 producer, 1 MMA issuer, SW softmax warps (4, or 8 = two per TMEM lane quarter)
struct AttnFwdSmem {
  // per group, offsets from the group's 1024-aligned base
  static constexpr int Q = 0;                                   // 2 x (128 x 64 B)
  static constexpr int K = Q + 2 * AT_BM * 64;                  // stages x 128 x 64 B
  static constexpr int V = K + AT_KV_STAGES * AT_BN * 64;
  static constexpr int GROUP = V + AT_KV_STAGES * AT_BN * 64;   // 64 KB
  static constexpr int STG = AF_GROUPS * GROUP;                 // output staging: one 32-row x 64-byte slice per (group, lane quarter) (TMA store)
  static constexpr int XCH = STG + AF_GROUPS * 4 * 2048;        // SW = 8: row-max exchange [2 tile parities][2 halves][128] + row sums [128], per group
  static constexpr int XCH_G = (2 * 2 * AT_BM + AT_BM) * 4;
  static constexpr int BAR = XCH + AF_GROUPS * XCH_G;
  static constexpr int NBAR_G = 4 + 2 * AT_KV_STAGES + 2 + 1 + 2;   // q_full[2], q_free[2], kv_full/empty, s_full, s_free, p_full, o_full[2]
  static constexpr int TOTAL = BAR + 8 * AF_GROUPS * NBAR_G + 16;
};

// causal mask of one 32-column chunk of a score row: column c is visible iff c <= lim (lim = last visible key - first key of the chunk)
__device__ __forceinline__ void fwd_mask32(float* sv, int lim) {
#pragma unroll
  for (int c = 0; c < 32; c++) sv[c] = (c <= lim) ? sv[c] : -INFINITY;
}
// ... and the chunk's key-padding bits (bit c set = key c of the chunk is padded)
__device__ __forceinline__ void fwd_mask32_pad(float* sv, int lim, uint32_t padm) {
#pragma unroll
  for (int c = 0; c < 32; c++) sv[c] = (c <= lim && !((padm >> c) & 1u)) ? sv[c] : -INFINITY;
}

// ONES (no dropout, head dim <= 26 of 32): the MMA warp writes ones into two padding columns of every V tile in shared memory, so
// column AF_PAD0 of the P V accumulator is the row sum of P -- rescaled with the other columns, summed by the tensor core from the
// same bf16 probabilities that weight V -- and the softmax warps need no add per probability.
#define AF_PAD0 26
// SW = 8 (with ONES): two softmax warps per TMEM lane quarter, 64 of the tile's 128 key columns and 16 of the 32 output columns
// each; the pair shares the row maximum through shared memory (one 64-thread named barrier per tile).  The kernel is bound by the
// serial latency of a tile's softmax, not by a pipe (35 % issue slots, 41 % MUFU with SW = 4): halving the columns per warp shortens it.
// KP: key-padding mask (uint8 [B, S], 1 = padded key; north star (2) -- the reference has none, multihead_attention.py:52): a separate
// instantiation, so the hot ones carry no code for it.
template <bool DROP, bool ONES, int SW, bool KP = false>
__global__ void __launch_bounds__(AF_THREADS(SW), 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const __grid_constant__ CUtensorMap tmO, float* __restrict__ lse, int B, int T, int S, int H, int mask_off, bpm_dropout_t drop,
                   uint32_t* __restrict__ drop_bits, const uint8_t* __restrict__ key_pad = nullptr) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_gen = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  static_assert(SW == 4 || (SW == 8 && ONES && !DROP), "8 softmax warps per group: only the light no-dropout / ones-column math fits 96 registers");
  constexpr int GW = 2 + SW;
  const int grp = warp / GW, gw = warp % GW;                              // group, role inside the group (0 producer, 1 MMA, 2.. softmax)
  const uint32_t gbase = base + grp * AttnFwdSmem::GROUP;
  const uint32_t bar0 = base + AttnFwdSmem::BAR + 8u * grp * AttnFwdSmem::NBAR_G;
  auto q_full = [&](int s) { return bar0 + 8u * s; };
  auto q_free = [&](int s) { return bar0 + 8u * (2 + s); };
  auto kv_full = [&](int s) { return bar0 + 8u * (4 + s); };
  auto kv_empty = [&](int s) { return bar0 + 8u * (4 + AT_KV_STAGES + s); };
  const uint32_t s_full = bar0 + 8u * (4 + 2 * AT_KV_STAGES), s_free = s_full + 8u, p_full = s_full + 16u;
  auto o_full = [&](int i) { return s_full + 24u + 8u * i; };
  const uint32_t tmem_ptr_addr = base + AttnFwdSmem::BAR + 8u * AF_GROUPS * AttnFwdSmem::NBAR_G;
  volatile uint32_t* tmem_ptr_gen = (volatile uint32_t*)(base_gen + AttnFwdSmem::BAR + 8 * AF_GROUPS * AttnFwdSmem::NBAR_G);

  const int nbh = B * H, nq = (T + AT_BM - 1) / AT_BM;
  const int n_items = nbh * nq;
  const int G = AF_GROUPS * gridDim.x, gid = blockIdx.x * AF_GROUPS + grp;
  // r-th item of this group: heavy (late) query tiles first, snake order over the groups so that the loads even out
  auto item_of = [&](int r, int& bh, int& qt) {
    const int idx = r * G + ((r & 1) ? (G - 1 - gid) : gid);
    if (idx >= n_items) return false;
    qt = nq - 1 - idx / nbh;
    bh = idx % nbh;
    return true;
  };
  auto tiles_of = [&](int qt) {
    int jmax = S - 1;
    if (mask_off >= 0) jmax = min(jmax, qt * AT_BM + AT_BM - 1 + mask_off);
    return jmax / AT_BN + 1;
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    for (int g = 0; g < AF_GROUPS; g++) {
      const uint32_t b0 = base + AttnFwdSmem::BAR + 8u * g * AttnFwdSmem::NBAR_G;
      for (int s = 0; s < 2; s++) { mbar_init(b0 + 8u * s, 1); mbar_init(b0 + 8u * (2 + s), 1); }
      for (int s = 0; s < AT_KV_STAGES; s++) { mbar_init(b0 + 8u * (4 + s), 1); mbar_init(b0 + 8u * (4 + AT_KV_STAGES + s), 1); }
      const uint32_t sf = b0 + 8u * (4 + 2 * AT_KV_STAGES);
      mbar_init(sf, 1); mbar_init(sf + 8u, SW); mbar_init(sf + 16u, SW); mbar_init(sf + 24u, 1); mbar_init(sf + 32u, 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_addr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem = *tmem_ptr_gen + (uint32_t)(grp * 256);
  const uint32_t tS = tmem, tP = tmem + 128, tO0 = tmem + 192;            // S 128 | P 64 (bf16 pairs) | O 2 x 32
  auto adr = [](uint32_t a) { return (uint64_t)((a & 0x3FFFFu) >> 4); };

  if (gw == 0) {
    // ===================== TMA producer (converged warp, one elected lane issues) =====================
    int kc = 0;                                                          // running K/V tile counter (ring position)
    for (int r = 0;; r++) {
      int bh, qt;
      if (!item_of(r, bh, qt)) break;
      const int b = bh / H, h = bh % H, qs = r & 1;
      mbar_wait(q_free(qs), ((uint32_t)(r >> 1) & 1u) ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(q_full(qs), AT_BM * 64);
        tma_load_3d(gbase + AttnFwdSmem::Q + qs * AT_BM * 64, &tmQ, q_full(qs), h * AT_DH, qt * AT_BM, b);
      }
      __syncwarp();
      const int nt = tiles_of(qt);
      for (int j = 0; j < nt; j++, kc++) {
        const int s = kc % AT_KV_STAGES;
        mbar_wait(kv_empty(s), ((uint32_t)(kc / AT_KV_STAGES) & 1u) ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(kv_full(s), 2 * AT_BN * 64);
          tma_load_3d(gbase + AttnFwdSmem::K + s * AT_BN * 64, &tmK, kv_full(s), h * AT_DH, j * AT_BN, b);
          tma_load_3d(gbase + AttnFwdSmem::V + s * AT_BN * 64, &tmV, kv_full(s), h * AT_DH, j * AT_BN, b);
        }
        __syncwarp();
      }
    }
  } else if (gw == 1) {
    // ===================== MMA issuer (converged warp, one elected lane issues) =====================
    const uint32_t idesc_s = umma_idesc_bf16(AT_BM, AT_BN, 0, 0);       // S = Q K^T : both operands K-major
    const uint32_t idesc_o = umma_idesc_bf16(AT_BM, AT_DH, 0, 1);       // O = P V   : A from TMEM, V is MN-major (dh contiguous)
    const uint64_t d_k64 = umma_desc(0, 16, 512, BPM_SWZ_64B);          // K-major 64-byte rows; +32 B per k16
    const uint64_t d_mn64 = umma_desc(0, 512, 512, BPM_SWZ_64B);        // V as MN-major B operand; +1024 B (16 keys) per k16
    int kc = 0, tc = 0;                                                  // K/V ring position; running tile counter (S / P / O phases)
    for (int r = 0;; r++) {
      int bh, qt;
      if (!item_of(r, bh, qt)) break;
      const int qs = r & 1, nt = tiles_of(qt);
      mbar_wait(q_full(qs), (uint32_t)(r >> 1) & 1u);
      const uint64_t dq = d_k64 | adr(gbase + AttnFwdSmem::Q + qs * AT_BM * 64);
      auto issue_s = [&](int kcj, int tcj, bool last) {
        const int s = kcj % AT_KV_STAGES;
        mbar_wait(kv_full(s), (uint32_t)(kcj / AT_KV_STAGES) & 1u);
        if (ONES) {                                                        // V[:, AF_PAD0 .. +1] = 1 (16-byte unit 3 of the swizzled 64-byte row)
          uint8_t* vt = base_gen + (gbase - base) + AttnFwdSmem::V + s * AT_BN * 64;
#pragma unroll
          for (int rr = 0; rr < 4; rr++) {
            const int r = rr * 32 + lane;
            *(uint32_t*)(vt + r * 64 + ((3 ^ ((r >> 1) & 3)) << 4) + 4) = 0x3F803F80u;
          }
          fence_async_smem();
          __syncwarp();
        }
        mbar_wait(s_free, ((uint32_t)tcj & 1u) ^ 1u);                      // the softmax warps have read the previous S tile
        tc_fence_after();
        if (elect_one()) {
          const uint64_t dk = d_k64 | adr(gbase + AttnFwdSmem::K + s * AT_BN * 64);
#pragma unroll
          for (int k = 0; k < AT_DH / 16; k++) umma_bf16(tS, dq + 2 * k, dk + 2 * k, idesc_s, (uint32_t)k);
          umma_commit(s_full);
          if (last) umma_commit(q_free(qs));                               // this item's Q tile has been read for the last time
        }
        __syncwarp();
      };
      issue_s(kc, tc, nt == 1);
      for (int j = 0; j < nt; j++, kc++, tc++) {
        if (j + 1 < nt) issue_s(kc + 1, tc + 1, j + 2 == nt);
        const int s = kc % AT_KV_STAGES;
        mbar_wait(p_full, (uint32_t)tc & 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t dv = d_mn64 | adr(gbase + AttnFwdSmem::V + s * AT_BN * 64);
#pragma unroll
          for (int k = 0; k < AT_BN / 16; k++) umma_bf16_ts(tO0 + 32 * (tc & 1), tP + 8 * k, dv + 64 * k, idesc_o, (uint32_t)k);
          umma_commit(o_full(tc & 1));      // O_j ready, P consumed
          umma_commit(kv_empty(s));         // K/V stage free
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== softmax warps: one thread per query row =====================
    const int quarter = warp & 3;                                          // TMEM lane quarter this warp may access
    const int half = (gw - 2) >> 2;                                        // SW = 8: key-column half / output-column half of this warp
    constexpr int NCH = SW == 8 ? 2 : 4;                                   // 32-key chunks of a tile per warp
    constexpr int OC = SW == 8 ? 16 : 32;                                  // output columns per warp
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int rr = quarter * 32 + lane;
    float* const xmax = (float*)(base_gen + AttnFwdSmem::XCH + grp * AttnFwdSmem::XCH_G);
    float* const xsum = xmax + 2 * 2 * AT_BM;
    const int pair_bar = 1 + grp * 4 + quarter;                            // named barrier of this (group, quarter) pair of warps
    auto pair_sync = [&]() { if (SW == 8) asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory"); };
    const DropCtx dc = make_drop(drop);
    const int W = (S + 31) >> 5;
    int tc = 0;
    for (int r = 0;; r++) {
      int bh, qt;
      if (!item_of(r, bh, qt)) break;
      const int b = bh / H, h = bh % H, q0 = qt * AT_BM, qi = q0 + rr, nt = tiles_of(qt);
      const int row_lim = (mask_off >= 0) ? min(qi + mask_off, S - 1) : S - 1;     // last visible key of this row
      // warp-uniform visibility of a 32-key chunk: keys <= vis_all are visible to all 32 rows of this warp, keys > vis_any to none.
      // Fully masked chunks (a suffix of the tile) cost no loads and no exponentials, fully visible ones no mask arithmetic.
      const int vis_all = (mask_off >= 0) ? min(q0 + quarter * 32 + mask_off, S - 1) : S - 1;
      const int vis_any = (mask_off >= 0) ? min(q0 + quarter * 32 + 31 + mask_off, S - 1) : S - 1;
      const uint64_t ebase = ((uint64_t)bh * T + (uint64_t)min(qi, T - 1)) * (uint64_t)S;
      float m = -INFINITY, l = 0.f;
      float oacc[OC];
#pragma unroll
      for (int d = 0; d < OC; d++) oacc[d] = 0.f;
      for (int j = 0; j < nt; j++, tc++) {
        const int k0 = j * AT_BN;
        const int nvis_t = max(0, min(AT_BN / 32, (vis_any - k0 + 32) >> 5));  // chunks of this tile with at least one visible key
        const int c_lo = SW == 8 ? 2 * half : 0;                               // this warp's chunks: [c_lo, c_lo + NCH)
        const int nvis = max(0, min(NCH, nvis_t - c_lo));                      // ... of which the first nvis have visible keys
        mbar_wait(s_full, (uint32_t)tc & 1u);
        tc_fence_after();
        // ---- pass 1: row maximum
        float mx = -INFINITY;
#pragma unroll 1
        for (int ci = 0; ci < nvis; ci++) {
          const int c = (c_lo + ci) * 32;
          float sv[32];
          tmem_ld32(tS + lane_off + c, sv);
          tmem_ld_wait();
          if (KP) {
            const int key = k0 + c + lane;
            const uint32_t padm = __ballot_sync(0xffffffffu, key < S && key_pad[(int64_t)b * S + key] != 0);
            if (k0 + c + 31 > vis_all || padm) fwd_mask32_pad(sv, row_lim - (k0 + c), padm);
          } else if (k0 + c + 31 > vis_all) fwd_mask32(sv, row_lim - (k0 + c));
          float m4[4] = {sv[0], sv[1], sv[2], sv[3]};
#pragma unroll
          for (int e = 4; e < 32; e += 4) {
            m4[0] = fmaxf(m4[0], sv[e]); m4[1] = fmaxf(m4[1], sv[e + 1]); m4[2] = fmaxf(m4[2], sv[e + 2]); m4[3] = fmaxf(m4[3], sv[e + 3]);
          }
          mx = fmaxf(mx, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
        }
        if (SW == 8) {                                                     // row maximum over both column halves
          float* xm = xmax + (tc & 1) * 2 * AT_BM;
          xm[half * AT_BM + rr] = mx;
          pair_sync();
          mx = fmaxf(mx, xm[(half ^ 1) * AT_BM + rr]);
        }
        const float m_new = fmaxf(m, mx * LOG2E_F);
        const float alpha = (KP && m_new == -INFINITY) ? 1.f : ex2f(m - m_new);
        // P (single TMEM buffer) is free once the previous tile's PV has been issued AND completed; its result O_{j-1} is then ready too
        if (tc > 0) { mbar_wait(o_full((tc - 1) & 1), (uint32_t)((tc - 1) >> 1) & 1u); tc_fence_after(); }
        // ---- pass 2: P = exp2(S log2e - m_new), row sum, dropout, packed bf16 back to TMEM
        float rs0 = 0.f, rs1 = 0.f;
        if (nvis == 0) {                                                   // nothing of this tile is visible to this warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_free);
        }
#pragma unroll 1
        for (int ci = 0; ci < NCH; ci++) {
          const int c = (c_lo + ci) * 32;
          uint32_t pk[16];
          if (ci >= nvis) {
#pragma unroll
            for (int u = 0; u < 16; u++) pk[u] = 0u;
            tmem_st16(tP + lane_off + (uint32_t)(c >> 1), pk);
            continue;
          }
          float sv[32];
          tmem_ld32(tS + lane_off + c, sv);
          tmem_ld_wait();
          if (ci + 1 == nvis) {                                            // the S tile is in registers / consumed: the next Q K^T may start
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_free);
          }
          if (KP) {
            const int key = k0 + c + lane;
            const uint32_t padm = __ballot_sync(0xffffffffu, key < S && key_pad[(int64_t)b * S + key] != 0);
            if (k0 + c + 31 > vis_all || padm) fwd_mask32_pad(sv, row_lim - (k0 + c), padm);
          } else if (k0 + c + 31 > vis_all) fwd_mask32(sv, row_lim - (k0 + c));
          float r4[4] = {0.f, 0.f, 0.f, 0.f};
          const float neg_m = (KP && m_new == -INFINITY) ? 0.f : -m_new;     // (a row whose keys so far are all padded: exp2(-inf - 0) = 0)
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            ffma2(sv[e], sv[e + 1], LOG2E_F, neg_m);
            ffma2(sv[e + 2], sv[e + 3], LOG2E_F, neg_m);
#pragma unroll
            for (int u = 0; u < 4; u++) sv[e + u] = ex2f(sv[e + u]);
            if (!ONES) {
              fadd2(r4[0], r4[1], sv[e], sv[e + 1]);
              fadd2(r4[2], r4[3], sv[e + 2], sv[e + 3]);
            }
          }
          rs0 += r4[0] + r4[2];
          rs1 += r4[1] + r4[3];
          if (DROP) {
            // keep decisions of keys k0+c .. k0+c+31 (element index e = ebase + key): one 32-bit word per (query, 32-key group).
            // The 1/(1-p) scale is applied once to the output row (every kept probability of the row shares it).
            uint32_t kb = 0u;
            const uint64_t e0 = ebase + (uint64_t)(k0 + c);
            if ((e0 & 1) == 0) {
#pragma unroll
              for (int u = 0; u < 16; u++) {
                const uint32_t x = drop_rand_pair(dc, (e0 >> 1) + u);
                const bool kl = drop_keep_lo(dc, x), kh = drop_keep_hi(dc, x);
                sv[2 * u] = kl ? sv[2 * u] : 0.f;
                sv[2 * u + 1] = kh ? sv[2 * u + 1] : 0.f;
                kb |= (kl ? 1u : 0u) << (2 * u) | (kh ? 1u : 0u) << (2 * u + 1);
              }
            } else {
#pragma unroll
              for (int u = 0; u < 32; u++) {
                const bool kp = drop_mult1(dc, e0 + u) != 0.f;
                sv[u] = kp ? sv[u] : 0.f;
                kb |= (kp ? 1u : 0u) << u;
              }
            }
            if (drop_bits != nullptr && qi < T && k0 + c < S) drop_bits[((int64_t)bh * T + qi) * (int64_t)W + ((k0 + c) >> 5)] = kb;
          }
#pragma unroll
          for (int u = 0; u < 16; u++) pk[u] = pack_bf16x2(sv[2 * u], sv[2 * u + 1]);
          tmem_st16(tP + lane_off + (uint32_t)(c >> 1), pk);
        }
        const float rs = rs0 + rs1;
        l = fmaf(l, alpha, rs);
        m = m_new;
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        // ---- fold the previous tile's P V into the register accumulator, then rescale to the new max
        if (j > 0) {
          float ov[OC];
          if (SW == 8) tmem_ld16(tO0 + 32 * ((tc - 1) & 1) + 16 * half + lane_off, ov);
          else tmem_ld32(tO0 + 32 * ((tc - 1) & 1) + lane_off, ov);
          tmem_ld_wait();
#pragma unroll
          for (int d = 0; d < OC; d++) oacc[d] = (oacc[d] + ov[d]) * alpha;
        }
      }
      {
        float ov[OC];
        mbar_wait(o_full((tc - 1) & 1), (uint32_t)((tc - 1) >> 1) & 1u);
        tc_fence_after();
        if (SW == 8) tmem_ld16(tO0 + 32 * ((tc - 1) & 1) + 16 * half + lane_off, ov);
        else tmem_ld32(tO0 + 32 * ((tc - 1) & 1) + lane_off, ov);
        tmem_ld_wait();
        tc_fence_before();
        // the 32 output rows of this lane quarter leave through shared memory and one TMA store (a thread owns a 64-byte row: direct
        // stores would be 32 half-used sectors per instruction); rows beyond T are clipped by the tensor map
        uint8_t* const stg = base_gen + AttnFwdSmem::STG + (grp * 4 + quarter) * 2048;
        const bool issuer = SW == 4 || half == 0;
        if (ONES) {                                                        // the ones column of V: row sum of the bf16 probabilities
          constexpr int PC = AF_PAD0 % OC;                                 // (column 26 sits in the second half's 16 columns)
          if (SW == 4 || half == 1) {
            l = oacc[PC] + ov[PC];
            oacc[PC] = oacc[PC + 1] = 0.f;
            ov[PC] = ov[PC + 1] = 0.f;
            if (SW == 8) xsum[rr] = l;
          }
        }
        if (issuer) {
          if (elect_one()) bulk_wait_read<0>();                            // the previous item's store has read the slice
          __syncwarp();
        }
        pair_sync();                                                       // SW = 8: row sum visible to the first half; staging slice free
        if (SW == 8 && half == 0) l = xsum[rr];
        const float inv_l = (DROP ? dc.inv_keep : 1.f) / l;
#pragma unroll
        for (int d = 0; d < OC; d++) oacc[d] = (oacc[d] + ov[d]) * inv_l;
#pragma unroll
        for (int u = 0; u < OC / 8; u++) {
          uint4 w;
          w.x = pack_bf16x2(oacc[u * 8 + 0], oacc[u * 8 + 1]); w.y = pack_bf16x2(oacc[u * 8 + 2], oacc[u * 8 + 3]);
          w.z = pack_bf16x2(oacc[u * 8 + 4], oacc[u * 8 + 5]); w.w = pack_bf16x2(oacc[u * 8 + 6], oacc[u * 8 + 7]);
          const int uu = (SW == 8 ? 2 * half : 0) + u;                     // 16-byte unit of the 64-byte row
          *(uint4*)(stg + lane * 64 + ((uu ^ ((lane >> 1) & 3)) << 4)) = w;
        }
        fence_async_smem();
        pair_sync();                                                       // SW = 8: both halves of the rows are staged
        __syncwarp();
        if (issuer) {
          if (elect_one()) {
            tma_store_3d(&tmO, base + AttnFwdSmem::STG + (grp * 4 + quarter) * 2048, h * AT_DH, q0 + quarter * 32, b);
            bulk_commit();
          }
          __syncwarp();
          if (qi < T) lse[(int64_t)bh * T + qi] = (m + log2f(l)) * LN2_F;
        }
      }
    }
    if (elect_one()) bulk_wait_read<0>();                                 // the staging slice must outlive the last TMA store's read
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(*tmem_ptr_gen, 512);
}

// ---------------------------------------------------------------- host
static int make_qkv_map(CUtensorMap* m, const void* p, int B, int rows, int HP, int box_rows, int pitch = 0, int batch_rows = 0) {
  if (pitch == 0) pitch = HP;                         // row pitch in elements (k / v may be column slices of a wider buffer)
  if (batch_rows == 0) batch_rows = rows;             // rows between two batch entries (a query chunk is a row range of every batch entry)
  uint64_t dims[3] = {(uint64_t)HP, (uint64_t)rows, (uint64_t)B};
  uint64_t str[2] = {(uint64_t)pitch * 2, (uint64_t)batch_rows * pitch * 2};
  uint32_t box[3] = {AT_DH, (uint32_t)box_rows, 1};
  return bpm_make_tmap_bf16(m, p, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B);
}

int bpm_xattn_tc_supported(const bpm_attn_t* a) {
  return a->dtype == BPM_BF16 && a->dhp == AT_DH;
}

int bpm_xattn_fwd_tc(const bpm_attn_t* a, const void* q, const void* k, const void* v, void* out, float* lse, cudaStream_t stream) {
  BPM_REQUIRE(((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) % 16 == 0, "xattn_fwd: pointers must be 16-byte aligned");
  const int HP = a->H * a->dhp;
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_qkv_map(&tq, q, a->B, a->T, HP, AT_BM))) return rc;
  BPM_REQUIRE(a->ld_kv % 8 == 0, "xattn_fwd: ld_kv must be a multiple of 8 elements");
  if ((rc = make_qkv_map(&tk, k, a->B, a->S, HP, AT_BN, a->ld_kv))) return rc;
  if ((rc = make_qkv_map(&tv, v, a->B, a->S, HP, AT_BN, a->ld_kv))) return rc;
  size_t smem = AttnFwdSmem::TOTAL + 1024;
  // 0: no dropout, row sums through a ones column of V (needs two free padding columns);  1: dropout;  2: no dropout, no ones column
  const bool kp = a->key_pad != nullptr;
  const int dm = a->drop.p > 0.f ? 1 : ((!kp && a->dh <= AF_PAD0 && !(bpm_debug_get(1) & 8192)) ? 0 : 2);
  const int sw = (dm == 0 && !(bpm_debug_get(1) & 32768)) ? 8 : 4;
  auto kern = kp ? (dm == 1 ? attn_fwd_tc_kernel<true, false, 4, true> : attn_fwd_tc_kernel<false, false, 4, true>)
                 : (dm == 1 ? attn_fwd_tc_kernel<true, false, 4>
                            : (dm == 0 ? (sw == 8 ? attn_fwd_tc_kernel<false, true, 8> : attn_fwd_tc_kernel<false, true, 4>) : attn_fwd_tc_kernel<false, false, 4>));
  CUtensorMap to;
  if ((rc = make_qkv_map(&to, out, a->B, a->T, HP, 32))) return rc;                      // output: {32 columns, 32 rows} store boxes
  if (int rc2 = bpm_func_smem((const void*)kern, (int)smem, "xattn_fwd_tc")) return rc2;
  const int n_items = a->B * a->H * bpm_cdiv(a->T, AT_BM);
  const int ctas = min(bpm_num_sms(), bpm_cdiv(n_items, AF_GROUPS));
  cudaError_t le = bpm_launch(kern, dim3(ctas), dim3(AF_THREADS(sw)), smem, stream, tq, tk, tv, to, lse, a->B, a->T, a->S, a->H,
                              a->mask_off, a->drop, a->drop_bits, a->key_pad);
  if (le != cudaSuccess) { bpm_set_error("xattn_fwd_tc: launch failed: %s", cudaGetErrorString(le)); (void)cudaGetLastError(); return BPM_ELAUNCH; }
  return BPM_OK;
}


// =====================================================================================================================
// Backward.  Persistent: one CTA per SM walks (batch, head) pairs (T <= 512).  For one (b, h) the dQ accumulators of all (<= 4)
// query tiles stay resident in TMEM and dK / dV of the current 128-key tile are accumulated in TMEM over the inner query-tile
// loop, so NO atomics and no second pass are needed.  The score tile is computed TRANSPOSED (rows = keys, columns = queries):
//     S^T  = K_j Q_i^T          dP^T = V_j dO_i^T                               (M128 N128 K32, both operands K-major)
//     P^T  = exp2(S^T log2e - lse_q log2e),   P~^T = P^T * dropmask,   dS^T = P^T * (dP^T * dropmask - delta_q)   [8 compute warps]
//     dV_j += P~^T dO_i         dK_j += dS^T Q_i      (A = the bf16 tile just written to smem, K-major; B = dO_i / Q_i MN-major)
//     dQ_i += dS K_j            (A = the SAME dS^T smem tile read as an MN-major operand; B = K_j MN-major)
// No row reductions are needed in the backward (lse and delta are per query = per column), so the 128x128 tile is split
// between two warps per TMEM lane quarter (64 columns each).
//   warp 0    TMA producer: per (b,h) the lse*log2e / delta vectors (bulk copies), K/V tiles (2 stages), Q/dO(/keep-bit) tiles (4 stages)
//   warp 1    issues S^T / dP^T (one thread);  warp 2 issues the accumulate MMAs dV, dK, dQ (one thread).  Two issuers, because a
//             tcgen05.mma issue blocks while the tensor-core queue is full: each stream's barrier waits hide behind the other's
//             MMAs.  Every shared-memory descriptor is formed once and advanced by adds.
//   warps 3.. compute (AB_CW / 4 per TMEM lane quarter; 16-column sub-chunks).  P~^T goes back to TENSOR MEMORY as packed bf16 (tcgen05.st) and feeds dV as the A operand from TMEM:
//             with N = dh = 32 an MMA whose A comes from shared memory is bound by the 4 KB A read (~80 clk), not by its math.
//             dS^T is needed in both orientations (dK and dQ), so it is staged in shared memory (double-buffered).
// TMEM: S^T 128 + dP^T 128 + P~^T 64 (bf16 pairs) + dK 32 + dV 32 + dQ 4x32 = 512 columns.
// The pre-kernel computes delta = rowsum(dO * O) and lse * log2e into the [2, B, H, T] workspace.
// =====================================================================================================================
// CW compute warps (template parameter: 8, or 16 for the light folded math): CW/4 per TMEM lane quarter, 128/(CW/4) query columns each
#define AB_THREADS(CW) (96 + 32 * (CW))
#define AB_MAXQT 4
#define AB_QD_STAGES 4
// optional per-role event trace of CTA 0 (bpm_debug_set_ptr; scripts/trace_attn.py): entry = (clock64 << 8) | event id.
// Compiled in only with `make EXTRA=-DBPM_ATTN_TRACE` (the checks cost ~5 % of the kernel when they sit in the pair loop).
#define AB_TRACE_N 4096
#ifdef BPM_ATTN_TRACE
#define TRACE(role, id)                                                                 \
  do {                                                                                  \
    if (trace != nullptr && blockIdx.x == 0 && tr_n < AB_TRACE_N) {                     \
      trace[(role) * AB_TRACE_N + tr_n] = ((unsigned long long)clock64() << 8) | (id);  \
      tr_n++;                                                                           \
    }                                                                                   \
  } while (0)
#else
#define TRACE(role, id) do { (void)tr_n; } while (0)
#endif

struct AttnBwdSmem {
  static constexpr int KV = 0;                                  // 2 stages x (K 8 KB + V 8 KB)
  static constexpr int QD_STAGE = 2 * 128 * 64 + 2048;          // Q 8 KB + dO 8 KB + dropout keep bits (128 queries x 4 words)
  static constexpr int QD = KV + 2 * 2 * 128 * 64;
  static constexpr int DST = QD + AB_QD_STAGES * QD_STAGE;      // 2 buffers x (2 chunk tiles x 16 KB)
  static constexpr int LD = DST + 2 * 2 * 128 * 128;            // 2 buffers x (512 floats lse*log2e (+inf beyond T) + 512 floats delta)
  static constexpr int STG = LD + 2 * 4096;                     // per compute warp: 2 staging slices of 32 rows x 64 B for the dQ / dK / dV TMA stores
  static constexpr int BAR = STG + 8 * 2 * 2048;                // 32 KB: 2 slices per warp with 8 compute warps, 1 with 16
  static constexpr int NBAR = 4 + 2 * AB_QD_STAGES + 16;
  static constexpr int TOTAL = BAR + 8 * NBAR + 16;
};
static_assert(AttnBwdSmem::DST % 1024 == 0 && AttnBwdSmem::STG % 1024 == 0, "swizzled tiles need 1024-byte alignment");
static_assert(AttnBwdSmem::TOTAL + 1024 <= 227 * 1024, "shared memory budget");

#define AD_TT 32                                  // rows (time steps) per block of the delta kernel
__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16* __restrict__ out, const bf16* __restrict__ dout, const float* __restrict__ lse,
                                                         float* __restrict__ ws, int B, int T, int H) {
  // ws[0][b, h, t] = delta = sum_d dO * O;   ws[1][b, h, t] = lse * log2e.  A block takes AD_TT consecutive rows of one sample: the dot
  // products run head-fastest (a warp reads 2 KB contiguous of O and dO), land in shared memory as [h][t] and leave time-fastest, so
  // that both sides are coalesced (thread-per-(row, head) wrote 4-byte words T * 4 bytes apart: 14.8 us for 50 MB).
  __shared__ float d_s[32 * (AD_TT + 1)];
  pdl_trigger();
  pdl_wait();
  const int tiles = (T + AD_TT - 1) / AD_TT;
  const int b = blockIdx.x / tiles, t0 = (blockIdx.x % tiles) * AD_TT;
  const int nt = min(AD_TT, T - t0), HP = H * AT_DH;
  const int64_t n = (int64_t)B * T * H;
  for (int i = threadIdx.x; i < nt * H; i += blockDim.x) {
    const int tl = i / H, h = i - tl * H;
    const int64_t row = (int64_t)b * T + t0 + tl;
    const bf16* o = out + row * HP + h * AT_DH;
    const bf16* g = dout + row * HP + h * AT_DH;
    float acc = 0.f;
#pragma unroll
    for (int u = 0; u < AT_DH / 8; u++) {
      Vec8<bf16> a, c; a.load(o + u * 8); c.load(g + u * 8);
#pragma unroll
      for (int j = 0; j < 8; j++) acc = fmaf(a.v[j], c.v[j], acc);
    }
    d_s[h * (AD_TT + 1) + tl] = acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < AD_TT * H; i += blockDim.x) {
    const int h = i / AD_TT, tl = i - h * AD_TT;
    if (tl < nt) {
      const int64_t o_idx = ((int64_t)b * H + h) * T + t0 + tl;
      ws[o_idx] = d_s[h * (AD_TT + 1) + tl];
      ws[n + o_idx] = lse[o_idx] * LOG2E_F;
    }
  }
}

// One NC-column chunk of a (128 keys x 128 queries) pair for key row `key`: sv (S^T) -> P~^T, dpv (dP^T) -> dS^T, in place.
//   MASKED: columns [0, cmin) of the chunk are invisible to this key row (cmin >= NC: all of them, e.g. a key beyond S).
//   DROP: 0 no dropout, 1 keep bits staged in shared memory by TMA (word (query, 32-key group), bit = key lane),
//         2 keep bits fetched from global memory into mw (one word per lane = query), 3 regenerate the decisions.
//   MASKED: the tile touches the mask diagonal or the end of the key sequence.
//   FOLD (no dropout only): -lse and -delta already sit inside the MMA results (two padding columns of the head dimension carry them
//         as bf16 hi + lo parts against ones in K / V, see the S^T issuer), so sv = S^T - lse and dpv = dP^T - delta on entry:
//         per element one multiply, one ex2, one multiply, and no shared-memory operands.
template <int DROP, bool MASKED, int NC, int FOLD>
__device__ __forceinline__ void bwd_chunk_math(float* sv, float* dpv, const float* __restrict__ lse_c, const float* __restrict__ del_c, int c_lo, int cmin,
                                               bool key_oob, bool diag, const DropCtx& dc, const uint32_t* __restrict__ bits_c, uint32_t mw, int lane,
                                               uint64_t e_row, int q_lo, int T, int S) {
  if (FOLD == 1) {
#pragma unroll
    for (int c = 0; c < NC; c += 2) {
      float t0 = sv[c], t1 = sv[c + 1];
      fmul2(t0, t1, LOG2E_F);
      float p0 = ex2f(t0), p1 = ex2f(t1);
      if (MASKED) { p0 = (c < cmin) ? 0.f : p0; p1 = (c + 1 < cmin) ? 0.f : p1; }
      sv[c] = p0; sv[c + 1] = p1;
      fmul2v(dpv[c], dpv[c + 1], p0, p1);
    }
    return;
  }
#pragma unroll
  for (int c = 0; c < NC; c += 4) {
    const float4 l4 = FOLD == 2 ? make_float4(0.f, 0.f, 0.f, 0.f) : *(const float4*)(lse_c + c);
    const float4 d4 = *(const float4*)(del_c + c);
    const float ls[4] = {l4.x, l4.y, l4.z, l4.w}, dl[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
    for (int e = 0; e < 4; e++) {
      float p = FOLD == 2 ? ex2f(sv[c + e] * LOG2E_F) : ex2f(fmaf(sv[c + e], LOG2E_F, -ls[e]));     // FOLD 2: -lse is inside S^T already
      if (MASKED) p = (c + e < cmin) ? 0.f : p;                    // cmin: first visible column of this key row, relative to the chunk
      if (DROP == 0) {
        sv[c + e] = p;
        dpv[c + e] = p * (dpv[c + e] - dl[e]);
      } else {
        float mult;
        if (DROP == 1) mult = ((bits_c[(c + e) * 4] >> lane) & 1u) ? dc.inv_keep : 0.f;
        else if (DROP == 2) mult = ((__shfl_sync(0xffffffffu, mw, c + e) >> lane) & 1u) ? dc.inv_keep : 0.f;
        else mult = drop_mult1(dc, ((uint64_t)min(q_lo + c + e, T - 1)) * (uint64_t)S + e_row);
        sv[c + e] = p * mult;                                     // P~^T
        dpv[c + e] = p * fmaf(dpv[c + e], mult, -dl[e]);          // dS^T
      }
    }
  }
}

// DM = 0: instantiation without any dropout code (two thirds of the calls of the training step: 20% faster than the general one,
// whose hot loop with all four dropout modes does not fit the instruction cache as well); DM = 1: dropout mode resolved at run time.
// (A third instantiation with only the TMA-staged keep bits was measured 15% SLOWER than the general one and is not kept.)
// FOLD (with DM = 0, head dim <= 26 of 32): the S^T issuer warp patches two padding columns of every operand tile in shared memory
// before it issues the MMAs -- Q rows get (-lse) and dO rows get (-delta), each as bf16 hi + lo, K and V rows get ones -- so the
// tensor core delivers S^T - lse and dP^T - delta directly.  The garbage this leaves in the same two columns of dQ / dK / dV is
// zeroed when the accumulators are drained.
#define AB_PAD0 26                       // patched head-dim columns (AB_PAD0, AB_PAD0 + 1): 4-byte aligned inside the 64-byte row
template <int DM, bool FOLD, int CW>
__global__ void __launch_bounds__(AB_THREADS(CW), 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const __grid_constant__ CUtensorMap tmdO, const __grid_constant__ CUtensorMap tmBits, const __grid_constant__ CUtensorMap tmDQ,
                   const __grid_constant__ CUtensorMap tmDK, const __grid_constant__ CUtensorMap tmDV, const int bits_tma,
                   const float* __restrict__ ws, bf16* __restrict__ dq, bf16* __restrict__ dk, bf16* __restrict__ dv, float dq_scale, int B, int T,
                   int S, int H, int mask_off, bpm_dropout_t drop, const uint32_t* __restrict__ drop_bits, const int ld_dkv, const int dbg,
                   unsigned long long* __restrict__ trace, const uint8_t* __restrict__ key_pad, const int q_base, const int T_full,
                   const int kv_add) {
  // Query chunking (T_full > 512): the host launches this kernel once per chunk of <= 512 queries.  T is the chunk length (tiling, TMEM),
  // q_base its first query and T_full the whole sequence (lse / delta rows, dropout element indices); the tensor maps of Q / dO / dQ
  // already start at the chunk, mask_off already includes q_base.  Chunks after the first ADD their dK / dV to what is there (kv_add:
  // TMA reduce-add instead of store).
  int tr_n = 0;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + AttnBwdSmem::BAR;
  auto kv_full = [&](int s) { return bar0 + 8u * s; };
  auto kv_empty = [&](int s) { return bar0 + 8u * (2 + s); };
  auto q_full = [&](int s) { return bar0 + 8u * (4 + s); };
  auto q_empty = [&](int s) { return bar0 + 8u * (4 + AB_QD_STAGES + s); };
  constexpr int B1 = 4 + 2 * AB_QD_STAGES;
  const uint32_t st_full = bar0 + 8u * B1, st_free = bar0 + 8u * (B1 + 1);
  const uint32_t pv_free = bar0 + 8u * (B1 + 2);                 // the dV MMAs have read P~^T from TMEM
  const uint32_t dkv_full = bar0 + 8u * (B1 + 3);
  auto pt_full = [&](int s) { return bar0 + 8u * (B1 + 4 + s); };   // P~^T (TMEM) and dS^T (smem buffer s) of a pair are written
  auto ds_free = [&](int s) { return bar0 + 8u * (B1 + 6 + s); };   // the dK / dQ MMAs have read dS^T buffer s
  const uint32_t dkv_free = bar0 + 8u * (B1 + 8);
  const uint32_t dq_full = bar0 + 8u * (B1 + 9), dq_free = bar0 + 8u * (B1 + 10);
  auto ld_full = [&](int s) { return bar0 + 8u * (B1 + 11 + s); };
  auto ld_empty = [&](int s) { return bar0 + 8u * (B1 + 13 + s); };
  static_assert(B1 + 15 <= AttnBwdSmem::NBAR, "barrier count");
  const uint32_t tmem_ptr_addr = bar0 + 8u * AttnBwdSmem::NBAR;
  volatile uint32_t* tmem_ptr_gen = (volatile uint32_t*)(base_gen + AttnBwdSmem::BAR + 8 * AttnBwdSmem::NBAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nbh = B * H;
  const int64_t nrow = (int64_t)nbh * T_full;
  const int nq = (T + 127) / 128, nkv = (S + 127) / 128;
  // first query tile that can see key tile j (visible iff key <= q + off)
  auto i_min_of = [&](int j) { int qlo = j * 128 - mask_off; return (mask_off < 0 || qlo <= 0) ? 0 : qlo / 128; };
  // pair iterator over (bh, key tile j, query tile i); jc counts the non-empty key tiles (K/V ring index)
  struct Pair { int bh, j, i, jc; };
  auto first_j = [&](int j) { while (j < nkv && i_min_of(j) >= nq) j++; return j; };
  auto next_pair = [&](Pair& p) {
    p.i++;
    if (p.i >= nq) {
      p.j = first_j(p.j + 1);
      p.jc++;
      if (p.j >= nkv) { p.bh += gridDim.x; p.j = first_j(0); }
      p.i = i_min_of(p.j < nkv ? p.j : 0);
    }
  };
  auto first_pair = [&]() { Pair p; p.bh = blockIdx.x; p.j = first_j(0); p.jc = 0; p.i = i_min_of(p.j < nkv ? p.j : 0); if (p.j >= nkv) p.bh = nbh; return p; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
    for (int s = 0; s < 2; s++) {
      mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1);
      mbar_init(pt_full(s), CW); mbar_init(ds_free(s), 1);
      mbar_init(ld_full(s), 1); mbar_init(ld_empty(s), CW);
    }
    for (int s = 0; s < AB_QD_STAGES; s++) { mbar_init(q_full(s), 1); mbar_init(q_empty(s), 1); }
    mbar_init(st_full, 1); mbar_init(st_free, CW); mbar_init(pv_free, 1);
    mbar_init(dkv_full, 1); mbar_init(dkv_free, 8);              // drained by 8 warps: (lane quarter) x (dK | dV)
    mbar_init(dq_full, 1); mbar_init(dq_free, CW);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_addr, 512);
  // query columns beyond T: lse = +inf (exp2(x - inf) = 0), delta = 0; the bulk copies only ever write the first T entries
  for (int t = threadIdx.x; t < 2 * 1024; t += AB_THREADS(CW)) {
    float* ld = (float*)(base_gen + AttnBwdSmem::LD);
    ld[t] = ((t & 1023) < 512) ? INFINITY : 0.f;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem = *tmem_ptr_gen;
  const uint32_t tST = tmem, tDPT = tmem + 128, tPT = tmem + 256, tDK = tmem + 320, tDV = tmem + 352, tDQ = tmem + 384;
  auto adr = [](uint32_t a) { return (uint64_t)((a & 0x3FFFFu) >> 4); };

  if (warp == 0) {
    {   // converged warp; one elected lane issues
      int pc = 0, jc = 0, bc = 0;
      for (int bh = blockIdx.x; bh < nbh; bh += gridDim.x, bc++) {
        const int b = bh / H, h = bh % H;
        {
          const int s = bc & 1;
          mbar_wait(ld_empty(s), ((uint32_t)(bc >> 1) & 1u) ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(ld_full(s), 2u * (uint32_t)T * 4u);
            bulk_load_1d(base + AttnBwdSmem::LD + s * 4096, ws + nrow + (int64_t)bh * T_full + q_base, (uint32_t)T * 4u, ld_full(s));
            bulk_load_1d(base + AttnBwdSmem::LD + s * 4096 + 2048, ws + (int64_t)bh * T_full + q_base, (uint32_t)T * 4u, ld_full(s));
          }
          __syncwarp();
        }
        for (int j = 0; j < nkv; j++) {
          const int imin = i_min_of(j);
          if (imin >= nq) continue;
          const int ks = jc & 1;
          mbar_wait(kv_empty(ks), ((uint32_t)(jc >> 1) & 1u) ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(kv_full(ks), 2 * 128 * 64);
            tma_load_3d(base + AttnBwdSmem::KV + ks * 16384, &tmK, kv_full(ks), h * AT_DH, j * 128, b);
            tma_load_3d(base + AttnBwdSmem::KV + ks * 16384 + 8192, &tmV, kv_full(ks), h * AT_DH, j * 128, b);
          }
          __syncwarp();
          jc++;
          for (int i = imin; i < nq; i++, pc++) {
            const int qs = pc % AB_QD_STAGES;
            mbar_wait(q_empty(qs), ((uint32_t)(pc / AB_QD_STAGES) & 1u) ^ 1u);
            if (elect_one()) {
              mbar_expect_tx(q_full(qs), 2 * 128 * 64 + (bits_tma ? 2048 : 0));
              tma_load_3d(base + AttnBwdSmem::QD + qs * AttnBwdSmem::QD_STAGE, &tmQ, q_full(qs), h * AT_DH, i * 128, b);
              tma_load_3d(base + AttnBwdSmem::QD + qs * AttnBwdSmem::QD_STAGE + 8192, &tmdO, q_full(qs), h * AT_DH, i * 128, b);
              if (bits_tma) tma_load_2d(base + AttnBwdSmem::QD + qs * AttnBwdSmem::QD_STAGE + 16384, &tmBits, q_full(qs), j * 4, bh * T_full + q_base + i * 128);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== S^T = K_j Q_i^T and dP^T = V_j dO_i^T (one pair ahead of the compute warps) =====================
    {   // converged warp; one elected lane issues
      const uint32_t id_st = umma_idesc_bf16(128, 128, 0, 0);
      const uint64_t d_k64 = umma_desc(0, 16, 512, BPM_SWZ_64B);          // K-major 64-byte rows; +32 B per k16
      int pc = 0, bcs = -1, cur_bh = -1;
      // patched word of tile row r: 16-byte unit 3 of the 64-byte row (SWIZZLE_64B: unit ^ ((r >> 1) & 3)), bytes 4..7 = columns 26, 27
      auto pad_word = [&](int tile_off, int r) { return (uint32_t*)(base_gen + tile_off + r * 64 + ((3 ^ ((r >> 1) & 3)) << 4) + 4); };
      auto hi_lo = [](float x) {
        const __nv_bfloat16 hi = __float2bfloat16_rn(x);
        const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
        return (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
      };
      for (Pair p = first_pair(); p.bh < nbh; next_pair(p), pc++) {
        const int ks = p.jc & 1, qs = pc % AB_QD_STAGES;
        TRACE(1, 10);
        if (FOLD && p.bh != cur_bh) {                          // this (b,h)'s lse / delta rows (the compute warps release the buffer)
          cur_bh = p.bh; bcs++;
          mbar_wait(ld_full(bcs & 1), (uint32_t)(bcs >> 1) & 1u);
        }
        if (p.i == i_min_of(p.j)) {
          mbar_wait(kv_full(ks), (uint32_t)(p.jc >> 1) & 1u);
          if (FOLD) {
            const int kt = AttnBwdSmem::KV + ks * 16384;
#pragma unroll
            for (int rr = 0; rr < 4; rr++) {
              *pad_word(kt, rr * 32 + lane) = 0x3F803F80u;             // K: (1, 1)
              *pad_word(kt + 8192, rr * 32 + lane) = 0x3F803F80u;      // V: (1, 1)
            }
          }
        }
        const uint32_t ka = base + AttnBwdSmem::KV + ks * 16384, va = ka + 8192;
        const uint32_t qa = base + AttnBwdSmem::QD + qs * AttnBwdSmem::QD_STAGE, ga = qa + 8192;
        mbar_wait(q_full(qs), (uint32_t)(pc / AB_QD_STAGES) & 1u);
        if (FOLD) {
          const float* lse_s = (const float*)(base_gen + AttnBwdSmem::LD + (bcs & 1) * 4096);
          const float* del_s = lse_s + 512;
          const int qt = AttnBwdSmem::QD + qs * AttnBwdSmem::QD_STAGE;
#pragma unroll
          for (int rr = 0; rr < 4; rr++) {
            const int row = rr * 32 + lane, qi = p.i * 128 + row;
            const float l2 = lse_s[qi];                                 // lse * log2e, +inf beyond T
            const bool live = qi < T && l2 < 1e30f;
            *pad_word(qt, row) = live ? hi_lo(-l2 * LN2_F) : 0x0000C6EAu;   // (-30000, 0): exp2 underflows to 0 for dead query rows
            if (DM == 0) *pad_word(qt + 8192, row) = live ? hi_lo(-del_s[qi]) : 0u;     // (with dropout, delta stays outside: dS = P (mask dP - delta))
          }
          fence_async_smem();                                          // generic-proxy writes -> visible to the MMAs' operand reads
          __syncwarp();
        }
        TRACE(1, 11);
        mbar_wait(st_free, ((uint32_t)pc & 1u) ^ 1u);
        TRACE(1, 12);
        tc_fence_after();
        const uint64_t dka = d_k64 | adr(ka), dva = d_k64 | adr(va), dqa = d_k64 | adr(qa), dga = d_k64 | adr(ga);
        if (elect_one()) {
          if (!(dbg & 8)) {
#pragma unroll
            for (int k = 0; k < 2; k++) umma_bf16(tST, dka + 2 * k, dqa + 2 * k, id_st, k > 0);
#pragma unroll
            for (int k = 0; k < 2; k++) umma_bf16(tDPT, dva + 2 * k, dga + 2 * k, id_st, k > 0);
          }
          umma_commit(st_full);
        }
        __syncwarp();
        TRACE(1, 13);
      }
    }
  } else if (warp == 2) {
    // ===================== dV_j += P~^T dO_i (A from TMEM),  dK_j += dS^T Q_i,  dQ_i += dS K_j =====================
    {   // converged warp; one elected lane issues
      const uint32_t id_kv = umma_idesc_bf16(128, AT_DH, 0, 1);    // dV, dK : A K-major (TMEM / smem tile), B MN-major
      const uint32_t id_dq = umma_idesc_bf16(128, AT_DH, 1, 1);    // dQ     : A = dS^T read MN-major, B = K_j MN-major
      const uint64_t d_mn64 = umma_desc(0, 512, 512, BPM_SWZ_64B);        // Q / dO / K tiles as MN-major B operands; +1024 B per k16
      const uint64_t d_k128 = umma_desc(0, 16, 1024, BPM_SWZ_128B);       // dS^T chunk tiles, K-major
      const uint64_t d_mn128 = umma_desc(0, 16384, 1024, BPM_SWZ_128B);   // dS^T read MN-major (LBO = chunk tile pitch); +2048 B per k16
      int pc = 0, bc = 0;
      for (Pair cur = first_pair(); cur.bh < nbh; pc++) {
        Pair nx = cur;
        next_pair(nx);
        const int ks = cur.jc & 1, qs = pc % AB_QD_STAGES, pb = pc & 1, imin = i_min_of(cur.j);
        const bool first_of_bh = cur.j == first_j(0) && cur.i == imin;
        const uint32_t ka = base + AttnBwdSmem::KV + ks * 16384;
        const uint32_t qa = base + AttnBwdSmem::QD + qs * AttnBwdSmem::QD_STAGE, ga = qa + 8192;
        const uint32_t da = base + AttnBwdSmem::DST + pb * 32768;
        TRACE(2, 14);
        mbar_wait(pt_full(pb), (uint32_t)(pc >> 1) & 1u);
        if (cur.i == imin) mbar_wait(dkv_free, ((uint32_t)cur.jc & 1u) ^ 1u);                // dK / dV of the previous key tile have been drained
        if (first_of_bh) mbar_wait(dq_free, ((uint32_t)bc & 1u) ^ 1u);                        // previous (b,h)'s dQ has been drained
        tc_fence_after();
        TRACE(2, 15);
        const uint64_t dda = d_k128 | adr(da), ddm = d_mn128 | adr(da);
        const uint64_t dgb = d_mn64 | adr(ga), dqb = d_mn64 | adr(qa), dkb = d_mn64 | adr(ka);
        const uint32_t t_dq = tDQ + 32 * cur.i;
        if (elect_one()) {
          const uint32_t acc_kv = cur.i > imin ? 1u : 0u, acc_q = cur.j > 0 ? 1u : 0u;
          if (!(dbg & 4)) {
#pragma unroll
            for (int k = 0; k < 8; k++) umma_bf16_ts(tDV, tPT + 8 * k, dgb + 64 * k, id_kv, acc_kv | (uint32_t)k);
          }
          umma_commit(pv_free);
          if (!(dbg & 4)) {
#pragma unroll
            for (int k = 0; k < 8; k++)
              umma_bf16(tDK, dda + (uint64_t)(((k >> 2) * 16384 + (k & 3) * 32) >> 4), dqb + 64 * k, id_kv, acc_kv | (uint32_t)k);
#pragma unroll
            for (int k = 0; k < 8; k++) umma_bf16(t_dq, ddm + 128 * k, dkb + 64 * k, id_dq, acc_q | (uint32_t)k);
          }
          umma_commit(ds_free(pb));
          umma_commit(q_empty(qs));
          if (cur.i == nq - 1) { umma_commit(dkv_full); umma_commit(kv_empty(ks)); }
          if (nx.bh != cur.bh) umma_commit(dq_full);
        }
        __syncwarp();
        if (nx.bh != cur.bh) bc++;
        TRACE(2, 16);
        cur = nx;
      }
    }
  } else {
    // ===================== compute warps =====================
    constexpr int NCG = CW / 4;                             // column groups (warps per lane quarter)
    constexpr int NCOL = 128 / NCG;                         // query columns per warp
    const int cw = warp - 3;
    const int quarter = warp & 3, colq = cw >> 2;          // TMEM lane quarter; column group [NCOL*colq, +NCOL) of the pair tile
    const int r = quarter * 32 + lane;                      // key row inside the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int c0 = colq * NCOL;
    const DropCtx dc = make_drop(drop);
    const int drop_mode = DM == 0 ? 0 : (!dc.on ? 0 : (bits_tma ? 1 : (drop_bits != nullptr ? 2 : 3)));
    const int HP = H * AT_DH;
    const int W = (S + 31) >> 5;
    // dS^T row r lives at (r >> 3) * 1024 + (r & 7) * 128 inside each 64-column chunk tile (16 KB), 16-byte unit u at u ^ (r & 7)
    const uint32_t row_off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
    int pc = 0, jc = 0, bc = 0;
    // dK / dV of a finished key tile are drained while the NEXT pair is being computed (no stall on the last accumulate MMAs):
    // dK (all 32 columns) is drained by the first column group of each lane quarter, dV by group NCG/2; the others only compute
    constexpr int DCOL = 32;
    const bool drain_dv = colq >= NCG / 2;
    const bool drains_kv = (colq % (NCG / 2)) == 0;
    const int dcol0 = 0;
    int pend_jc = -1, pend_b = 0, pend_h = 0, pend_key = 0;
    // dQ of a finished (b,h) is drained after the first pair of the NEXT (b,h) has been computed, for the same reason
    int pend_q_bc = -1, pend_q_b = 0, pend_q_h = 0;
    // accumulator rows leave through shared memory and the TMA unit: a thread owns one 64-byte row, so direct stores would be 32
    // half-used sectors per instruction (measured: 3 k clk per (b,h) for dQ, 1 k per key tile for dK / dV); staged and stored as
    // {32 columns, 32 rows} boxes they are fully coalesced, asynchronous, and rows beyond T / S are clipped by the tensor map
    constexpr int NSL = 16 / CW;                              // staging slices per warp (2 KB each)
    uint8_t* const stg_gen = base_gen + AttnBwdSmem::STG + cw * NSL * 2048;
    const uint32_t stg_s = base + AttnBwdSmem::STG + cw * NSL * 2048;
    int stg_n = 0;                                            // staging slices used so far (slice = stg_n % NSL)
    auto stage_store = [&](const float* acc, float scale, const CUtensorMap* map, int c0, int row0, int bidx, bool add = false) {
      const int sl = stg_n % NSL;
      stg_n++;
      if (elect_one()) bulk_wait_read<NSL - 1>();             // the store that last used this slice has read it
      __syncwarp();
#pragma unroll
      for (int u = 0; u < 4; u++)
        *(uint4*)(stg_gen + sl * 2048 + lane * 64 + ((u ^ ((lane >> 1) & 3)) << 4)) =
            make_uint4(pack_bf16x2(acc[u * 8] * scale, acc[u * 8 + 1] * scale), pack_bf16x2(acc[u * 8 + 2] * scale, acc[u * 8 + 3] * scale),
                       pack_bf16x2(acc[u * 8 + 4] * scale, acc[u * 8 + 5] * scale), pack_bf16x2(acc[u * 8 + 6] * scale, acc[u * 8 + 7] * scale));
      fence_async_smem();
      __syncwarp();
      if (elect_one()) {
        if (add) tma_reduce_add_3d(map, stg_s + sl * 2048, c0, row0, bidx);
        else tma_store_3d(map, stg_s + sl * 2048, c0, row0, bidx);
        bulk_commit();
      }
      __syncwarp();
    };
    auto drain_dq = [&]() {
      mbar_wait(dq_full, (uint32_t)pend_q_bc & 1u);
      tc_fence_after();
      TRACE(3 + (colq & 1), 35);
      for (int i = colq; i < nq; i += NCG) {                  // query tile i is drained by column group i % NCG
        float acc[AT_DH];
        tmem_ld32(tDQ + 32 * i + lane_off, acc);
        tmem_ld_wait();
        if (FOLD) acc[AB_PAD0] = acc[AB_PAD0 + 1] = 0.f;
        stage_store(acc, dq_scale, &tmDQ, pend_q_h * AT_DH, i * 128 + quarter * 32, pend_q_b);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free);
      pend_q_bc = -1;
    };
    auto drain_dkv = [&]() {
      if (!drains_kv) { pend_jc = -1; return; }
      mbar_wait(dkv_full, (uint32_t)pend_jc & 1u);
      tc_fence_after();
      TRACE(3 + (colq & 1), 36);
      float acc[DCOL];
      tmem_ld32((drain_dv ? tDV : tDK) + lane_off, acc);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dkv_free);
      if (FOLD) acc[AB_PAD0] = acc[AB_PAD0 + 1] = 0.f;
      stage_store(acc, 1.f, drain_dv ? &tmDV : &tmDK, pend_h * AT_DH, pend_key - lane, pend_b, kv_add != 0);
      pend_jc = -1;
    };
    for (int bh = blockIdx.x; bh < nbh; bh += gridDim.x, bc++) {
      const int b = bh / H, h = bh % H;
      const float* lse_s = (const float*)(base_gen + AttnBwdSmem::LD + (bc & 1) * 4096);
      const float* del_s = lse_s + 512;
      mbar_wait(ld_full(bc & 1), (uint32_t)(bc >> 1) & 1u);
      for (int j = 0; j < nkv; j++) {
        const int key = j * 128 + r;
        const int imin = i_min_of(j);
        if (imin >= nq) {                                     // no query sees this key tile: dK = dV = 0
          if (key < S && drains_kv && !kv_add) {
            bf16* dst = (drain_dv ? dv : dk) + ((int64_t)b * S + key) * ld_dkv + h * AT_DH + dcol0;
#pragma unroll
            for (int u = 0; u < DCOL / 8; u++) *(uint4*)(dst + u * 8) = make_uint4(0, 0, 0, 0);
          }
          continue;
        }
        // a padded key behaves like a key beyond S: probability 0, dS = 0, so dK = dV = 0 for its row
        const bool key_oob = key >= S || (key_pad != nullptr && key_pad[(int64_t)b * S + key] != 0);
        const bool any_pad = key_pad != nullptr && __any_sync(0xffffffffu, key_oob);
        // element index of (query q, key) in the [B*H, T, S] probability tensor = (bh*T + q)*S + key
        const uint64_t e_row = (uint64_t)bh * (uint64_t)T_full * (uint64_t)S + (uint64_t)min(key, S - 1);
        for (int i = imin; i < nq; i++, pc++) {
          const int q0 = i * 128;
          const int pb = pc & 1, qs = pc % AB_QD_STAGES;
          const bool diag = (mask_off >= 0) && (j * 128 + 127 > q0 + mask_off);
          const bool masked = diag || (j * 128 + 127 >= S) || any_pad;
          const int cmin = key - mask_off - q0;               // columns < cmin are masked for this key row (diag tiles only)
          // warp-uniform view: 16-query sub-chunks starting at or beyond cmin_hi (the last key row's cmin) need no mask arithmetic
          const int cmin_hi = j * 128 + quarter * 32 + 31 - mask_off - q0;
          const bool warp_oob = j * 128 + quarter * 32 + 31 >= S || any_pad;
          uint8_t* const dtile = base_gen + AttnBwdSmem::DST + pb * 32768 + row_off;
          // keep bits written by the forward: word (query, 32-key group).  Fast path: the TMA producer staged the 128 x 4-word tile of
          // this pair next to Q / dO (one broadcast LDS per column); otherwise each lane fetches the words of its columns (32 per word).
          const uint32_t* bits_s = (const uint32_t*)(base_gen + AttnBwdSmem::QD + qs * AttnBwdSmem::QD_STAGE + 16384) + quarter;
          uint32_t mw[NCOL / 32];
#pragma unroll
          for (int g = 0; g < NCOL / 32; g++) {
            mw[g] = 0xFFFFFFFFu;
            if (drop_mode == 2) {
              const int qq = q0 + c0 + g * 32 + lane;
              mw[g] = (qq < T && j * 4 + quarter < W) ? drop_bits[((int64_t)bh * T_full + q_base + qq) * W + j * 4 + quarter] : 0u;
            }
          }
          TRACE(3 + (colq & 1), 20);
          mbar_wait(st_full, (uint32_t)pc & 1u);        // (S^T was computed from this pair's Q stage, so its keep-bit tile has landed too)
          tc_fence_after();
          TRACE(3 + (colq & 1), 21);
#pragma unroll 1
          for (int ch = 0; ch < NCOL / 32; ch++) {            // (not unrolled: the hot loop has to stay inside the instruction cache)
          uint32_t pk[16];                                    // P~^T of 32 queries, packed bf16 pairs
#pragma unroll
          for (int s2 = 0; s2 < 2; s2++) {                    // 16-column sub-chunks keep the live register set small
            const int sc = ch * 2 + s2;
            const int cs = c0 + sc * 16;
            float sv[16], dpv[16];
            if (!(dbg & 16)) {
              tmem_ld16(tST + lane_off + cs, sv);
              tmem_ld16(tDPT + lane_off + cs, dpv);
              tmem_ld_wait();
            }
            if (sc == NCOL / 16 - 1) {
              tc_fence_before();                              // all of this warp's S^T / dP^T columns are in registers
              __syncwarp();
              if (lane == 0) mbar_arrive(st_free);
            }
            if (!(dbg & 1)) {
              const float* lse_c = lse_s + q0 + cs;
              const float* del_c = del_s + q0 + cs;
              const uint32_t mwc = NCOL / 32 == 1 ? mw[0] : (ch == 0 ? mw[0] : mw[NCOL / 32 - 1]);
              const uint32_t mws = s2 == 0 ? mwc : __shfl_down_sync(0xffffffffu, mwc, 16);   // lane l: word of query cs + l
              const bool need_mask = masked && (warp_oob || (diag && cs < cmin_hi));
              const int cm = key_oob ? 0x40000000 : (diag ? cmin - cs : -0x40000000);
#define BWD_MATH(DROP)                                                                                                                        \
  do {                                                                                                                                        \
    if (need_mask) bwd_chunk_math<DROP, true, 16, FOLD ? (DROP == 0 ? 1 : 2) : 0>(sv, dpv, lse_c, del_c, cs, cm, key_oob, diag, dc, bits_s + cs * 4, mws, lane, e_row, q_base + q0 + cs, T_full, S); \
    else bwd_chunk_math<DROP, false, 16, FOLD ? (DROP == 0 ? 1 : 2) : 0>(sv, dpv, lse_c, del_c, cs, cmin, key_oob, diag, dc, bits_s + cs * 4, mws, lane, e_row, q_base + q0 + cs, T_full, S);       \
  } while (0)
              if (drop_mode == 0) BWD_MATH(0);
              else if (drop_mode == 1) BWD_MATH(1);
              else if (drop_mode == 2) BWD_MATH(2);
              else BWD_MATH(3);
#undef BWD_MATH
            }
            if (sc == 0) {
              TRACE(3 + (colq & 1), 23);
              // the dV MMAs of the previous pair have read P~^T.  Their commit also covers every earlier MMA of that issuer, i.e. the
              // dK / dQ MMAs that read this dS^T buffer two pairs ago.
              mbar_wait(pv_free, ((uint32_t)pc & 1u) ^ 1u);
              tc_fence_after();
              TRACE(3 + (colq & 1), 24);
            }
            if (!(dbg & 2)) {
#pragma unroll
              for (int u = 0; u < 8; u++) pk[s2 * 8 + u] = pack_bf16x2(sv[2 * u], sv[2 * u + 1]);
              // P~^T: row r, queries [cs - 16, cs + 16) -> 16 packed columns of the dV A operand in TMEM
              if (s2 == 1) tmem_st16(tPT + lane_off + (uint32_t)((cs - 16) >> 1), pk);
              // dS^T: columns [cs, cs + 16) = chunk tile cs / 64, 16-byte units (cs % 64) / 8, +1
#pragma unroll
              for (int uu = 0; uu < 2; uu++) {
                const int u = ((cs & 63) >> 3) + uu;
                *(uint4*)(dtile + (cs >> 6) * 16384 + ((u ^ (r & 7)) << 4)) =
                    make_uint4(pack_bf16x2(dpv[uu * 8], dpv[uu * 8 + 1]), pack_bf16x2(dpv[uu * 8 + 2], dpv[uu * 8 + 3]),
                               pack_bf16x2(dpv[uu * 8 + 4], dpv[uu * 8 + 5]), pack_bf16x2(dpv[uu * 8 + 6], dpv[uu * 8 + 7]));
              }
            }
          }
          }
          TRACE(3 + (colq & 1), 30);
          if (pend_jc >= 0) drain_dkv();                      // previous key tile's dK / dV: its MMAs finished while this pair was computed
          TRACE(3 + (colq & 1), 32);
          if (pend_q_bc >= 0) drain_dq();                     // previous (b,h)'s dQ
          TRACE(3 + (colq & 1), 33);
          tmem_st_wait();
          tc_fence_before();
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(pt_full(pb));
          TRACE(3 + (colq & 1), 31);
        }
        pend_jc = jc; pend_b = b; pend_h = h; pend_key = key;
        jc++;
      }
      // ---- end of this (b,h): release the lse / delta buffer; the last key tile's dK / dV and dQ are drained one pair later
      __syncwarp();
      if (lane == 0) mbar_arrive(ld_empty(bc & 1));
      pend_q_bc = bc; pend_q_b = b; pend_q_h = h;
      TRACE(3 + (colq & 1), 40);
    }
    if (pend_q_bc >= 0) {                                     // after the last (b,h): dq_full also covers the last key tile's dK / dV
      mbar_wait(dq_full, (uint32_t)pend_q_bc & 1u);
      tc_fence_after();
      if (pend_jc >= 0) drain_dkv();
      drain_dq();
    }
    if (elect_one()) bulk_wait_read<0>();                     // the staging slices must outlive the last TMA stores' reads
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

int bpm_xattn_bwd_simt(const bpm_attn_t* a, const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse,
                       float* delta, void* dq, float dq_scale, void* dk, void* dv, cudaStream_t s);

int bpm_xattn_bwd_tc(const bpm_attn_t* a, const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse,
                     float* delta, void* dq, float dq_scale, void* dk, void* dv, cudaStream_t stream) {
  // The per-(b,h) lse / delta vectors travel as 16-byte bulk copies (T % 4); other shapes use the fp32-math kernel.  The dQ accumulators
  // of a (b,h) live in TMEM (4 query tiles): longer sequences run as chunks of 512 queries, one launch each, every chunk producing its
  // own dQ rows and ADDING its share of dK / dV (TMA reduce-add in the storage type) to the first chunk's.
  if (a->T % 4 != 0) return bpm_xattn_bwd_simt(a, q, k, v, out, dout, lse, delta, dq, dq_scale, dk, dv, stream);
  BPM_REQUIRE(((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out | (uintptr_t)dout | (uintptr_t)dq | (uintptr_t)dk | (uintptr_t)dv |
               (uintptr_t)delta) % 16 == 0, "xattn_bwd: pointers must be 16-byte aligned");
  const int HP = a->H * a->dhp;
  {
    BPM_REQUIRE(a->H <= 32, "xattn_bwd: more than 32 heads");
    const int grid = a->B * ((a->T + AD_TT - 1) / AD_TT);
    cudaError_t le = bpm_launch(attn_delta_kernel, dim3(grid), dim3(256), 0, stream, (const bf16*)out, (const bf16*)dout, lse, delta, a->B, a->T, a->H);
    if (le != cudaSuccess) { bpm_set_error("xattn_delta: launch failed: %s", cudaGetErrorString(le)); (void)cudaGetLastError(); return BPM_ELAUNCH; }
  }
  BPM_REQUIRE(a->ld_kv % 8 == 0 && a->ld_dkv % 8 == 0, "xattn_bwd: ld_kv / ld_dkv must be multiples of 8 elements");
  CUtensorMap tk, tv, tdk, tdv;
  int rc;
  if ((rc = make_qkv_map(&tdk, dk, a->B, a->S, HP, 32, a->ld_dkv))) return rc;            // outputs: {32 columns, 32 rows} store boxes
  if ((rc = make_qkv_map(&tdv, dv, a->B, a->S, HP, 32, a->ld_dkv))) return rc;
  if ((rc = make_qkv_map(&tk, k, a->B, a->S, HP, 128, a->ld_kv))) return rc;
  if ((rc = make_qkv_map(&tv, v, a->B, a->S, HP, 128, a->ld_kv))) return rc;
  // dropout keep bits as a [B*H*T, W] uint32 tensor, box {4 words, 128 queries}; needs a 16-byte pitch (S % 128 == 0)
  CUtensorMap tb = tk;
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
  // 0: no dropout, -lse / -delta folded into the MMAs (needs two free padding columns);  1: dropout;  2: no dropout, no fold
  const int dm = a->drop.p > 0.f ? 1 : ((a->dh <= AB_PAD0 && !(bpm_debug_get(1) & 2048)) ? 0 : 2);
  const int cw = (dm == 0 && !(bpm_debug_get(1) & 4096)) ? 16 : 8;
  const bool fold1 = dm == 1 && a->dh <= AB_PAD0 && !(bpm_debug_get(1) & 2048);     // dropout: lse still folds into S^T
  auto kern = dm == 0 ? (cw == 16 ? attn_bwd_tc_kernel<0, true, 16> : attn_bwd_tc_kernel<0, true, 8>)
                      : (dm == 1 ? (fold1 ? attn_bwd_tc_kernel<1, true, 8> : attn_bwd_tc_kernel<1, false, 8>) : attn_bwd_tc_kernel<0, false, 8>);
  if (int rc2 = bpm_func_smem((const void*)kern, (int)smem, "xattn_bwd_tc")) return rc2;
  const int ctas = min(a->B * a->H, bpm_num_sms());
  const int TC = 128 * AB_MAXQT;
  for (int q_base = 0; q_base < a->T; q_base += TC) {
    const int Tc = min(TC, a->T - q_base);
    const int64_t row_off = (int64_t)q_base * HP;                                         // element offset of the chunk inside a batch entry
    CUtensorMap tq, tg, tdq;
    if ((rc = make_qkv_map(&tq, (const bf16*)q + row_off, a->B, Tc, HP, 128, 0, a->T))) return rc;
    if ((rc = make_qkv_map(&tg, (const bf16*)dout + row_off, a->B, Tc, HP, 128, 0, a->T))) return rc;
    if ((rc = make_qkv_map(&tdq, (bf16*)dq + row_off, a->B, Tc, HP, 32, 0, a->T))) return rc;
    const int off = a->mask_off >= 0 ? a->mask_off + q_base : a->mask_off;
    cudaError_t le = bpm_launch(kern, dim3(ctas), dim3(AB_THREADS(cw)), smem, stream, tq, tk, tv, tg, tb, tdq, tdk, tdv, bits_tma, delta, (bf16*)dq, (bf16*)dk,
                                (bf16*)dv, dq_scale, a->B, Tc, a->S, a->H, off, a->drop, a->drop_bits, a->ld_dkv ? a->ld_dkv : HP, bpm_debug_get(1),
                                (unsigned long long*)bpm_debug_get_ptr(), a->key_pad, q_base, a->T, q_base > 0 ? 1 : 0);
    if (le != cudaSuccess) { bpm_set_error("xattn_bwd_tc: launch failed: %s", cudaGetErrorString(le)); (void)cudaGetLastError(); return BPM_ELAUNCH; }
  }
  return BPM_OK;
}
