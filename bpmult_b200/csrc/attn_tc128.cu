// Crossmodal flash attention on tcgen05 tensor cores (sm_100a) for head dims 64 and 128 (the 4-modality model: hidden 768 / 6 heads,
// reference README.md:30,36; semantics models/multihead_attention.py:95-127 with the mask of models/transformer.py:209-216 evaluated
// from indices, plus the optional key-padding mask of the north star).  bf16 in, fp32 accumulate.
//
// At head dim 128 the two MMAs of a score tile cost as much tensor time as its exponentials cost MUFU time (1 : 1; at head dim 32 it is
// 1 : 4, attn_tc.cu), and an output row no longer fits in registers, so the design differs from attn_tc.cu:
//
// Forward.  Persistent, one CTA per SM.  An item = (batch*head, PAIR of consecutive 128-query tiles); both query tiles share every
// K / V tile the CTA loads (half the operand traffic per FLOP) and ping-pong on the tensor core: while the softmax warps of tile A work
// on S_A, the tensor core runs P_B V and the next Q_B K^T.
//   warp 0      TMA producer: Q_A, Q_B (per item), K_j / V_j through one ring of 4 tile slots (32 KB each, SWIZZLE_128B sub-tiles of 64 columns)
//   warp 1      MMA issuer:   S_g = Q_g K_j^T (M128 N128 K=DH) and O_g += P_g V_j (M128 N=DH K128, A = P_g FROM TMEM, V as MN-major B)
//   warps 2-5   softmax of tile A, warps 6-9 softmax of tile B: one thread per query row (TMEM lane)
//   TMEM (512 columns): S_A 128 | S_B 128 | O_A DH | O_B DH.  P_g (packed bf16, 64 columns) ALIASES S_g: a thread overwrites score
//   columns it has already consumed (chunk c of P lands in columns [16c, 16c+16), inside what chunks <= c of S occupied), and the next
//   Q K^T into S_g is ordered behind the P V that reads P_g because one thread issues both (tcgen05.mma executes in issue order).
//   O accumulates in TMEM over the key tiles.  The running maximum is applied LAZILY: a row keeps exponentiating against the reference
//   maximum it already used until the true maximum has grown by more than 2^8; only then the row's O and row sum are rescaled (by the
//   owning softmax thread: tcgen05.ld / multiply / tcgen05.st).  exp2 arguments stay <= 8, probabilities <= 256: exact in fp32 / bf16.
//
// Backward.  Persistent; an item = (batch*head, 128-key tile j).  dK_j and dV_j (128 x 128 fp32 each) accumulate in TMEM over the query
// loop, which walks HALF tiles of 64 queries so that the remaining 256 columns hold two generations of S^T / dP^T:
//     S^T  = K_j Q_h^T,   dP^T = V_j dO_h^T                      (M128 N64 K128; rows = keys, columns = queries)
//     P^T  = exp2(S^T log2e - lse),  P~^T = P^T * dropmask,  dS^T = P^T (dP^T * dropmask - delta)         [8 compute warps]
//     dV_j += P~^T dO_h   (A = P~^T FROM TMEM, written over the S^T columns its own thread has consumed)
//     dK_j += dS^T Q_h    (A = dS^T from shared memory, K-major)
//     dQ_h^T = K_j^T dS_h (M = 128 head dims, N = 64 queries, K = 128 keys: A = K_j read MN-major, B = the same dS^T tile read MN-major)
//   dQ_h^T lands in the columns dP^T occupied, is drained TRANSPOSED (thread = head dim) into [query][dim] fp32 staging and leaves with a
//   TMA reduce-add into an fp32 accumulator in global memory (every key tile contributes to every query row); a small kernel scales
//   and casts it to bf16 afterwards.  No atomics on dK / dV, no second pass, any T and S.
#include "tc_common.cuh"
#include "attn_math.cuh"

#define A8_THREADS 320                  // forward: producer, issuer, 2 x 4 softmax warps
#define A8_NSTG 4                       // K / V ring slots

template <int DH> struct A8Fwd {
  static constexpr int NSUB = DH / 64;
  static constexpr int SUBT = 128 * 128;                 // one sub-tile: 128 rows x 64 bf16 columns (128-byte rows, SWIZZLE_128B)
  static constexpr int TILE = NSUB * SUBT;               // 128 rows x DH
  static constexpr int Q = 0;                            // 2 query tiles
  static constexpr int KV = Q + 2 * TILE;                // ring of A8_NSTG tiles: K_0, V_0, K_1, V_1, ...
  static constexpr int STG = KV + A8_NSTG * TILE;        // output staging: per softmax warp 32 rows x 128 B
  static constexpr int BAR = STG + 8 * 4096;
  static constexpr int NBAR = 4 + 2 * A8_NSTG + 8;       // q_full[2] q_free[2] kv_full[] kv_empty[] s_full[2] p_full[2] o_done[2] o_free[2]
  static constexpr int TOTAL = BAR + 8 * NBAR + 16;
};
static_assert(A8Fwd<128>::TOTAL + 1024 <= 227 * 1024, "forward shared memory budget");

// mask of one 32-column chunk of a score row: column c is visible iff c <= lim and its key is not padded
__device__ __forceinline__ void a8_mask32(float* sv, int lim, uint32_t padm) {
#pragma unroll
  for (int c = 0; c < 32; c++) sv[c] = (c <= lim && !((padm >> c) & 1u)) ? sv[c] : -INFINITY;
}

template <int DH, bool DROP>
__global__ void __launch_bounds__(A8_THREADS, 1)
attn128_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const __grid_constant__ CUtensorMap tmO, float* __restrict__ lse, int B, int T, int S, int H, int mask_off,
                   const uint8_t* __restrict__ key_pad, bpm_dropout_t drop, uint32_t* __restrict__ drop_bits) {
  using L = A8Fwd<DH>;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_gen = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = base + L::BAR;
  auto q_full = [&](int g) { return bar0 + 8u * g; };
  auto q_free = [&](int g) { return bar0 + 8u * (2 + g); };
  auto kv_full = [&](int s) { return bar0 + 8u * (4 + s); };
  auto kv_empty = [&](int s) { return bar0 + 8u * (4 + A8_NSTG + s); };
  constexpr int B1 = 4 + 2 * A8_NSTG;
  auto s_full = [&](int g) { return bar0 + 8u * (B1 + g); };
  auto p_full = [&](int g) { return bar0 + 8u * (B1 + 2 + g); };
  auto o_done = [&](int g) { return bar0 + 8u * (B1 + 4 + g); };
  auto o_free = [&](int g) { return bar0 + 8u * (B1 + 6 + g); };
  const uint32_t tmem_ptr_addr = bar0 + 8u * L::NBAR;
  volatile uint32_t* tmem_ptr_gen = (volatile uint32_t*)(base_gen + L::BAR + 8 * L::NBAR);

  const int nbh = B * H, nq = (T + 127) / 128, nqp = (nq + 1) / 2;
  const int n_items = nbh * nqp;
  const int G = gridDim.x, gid = blockIdx.x;
  // r-th item of this CTA: heavy (late) query pairs first, snake order over the CTAs so that the loads even out
  auto item_of = [&](int r, int& bh, int& qp) {
    const int idx = r * G + ((r & 1) ? (G - 1 - gid) : gid);
    if (idx >= n_items) return false;
    qp = nqp - 1 - idx / nbh;
    bh = idx % nbh;
    return true;
  };
  auto tiles_of = [&](int qt) {                                           // key tiles with at least one key visible to the query tile
    int jmax = S - 1;
    if (mask_off >= 0) jmax = min(jmax, qt * 128 + 127 + mask_off);
    return jmax / 128 + 1;
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
    for (int g = 0; g < 2; g++) {
      mbar_init(q_full(g), 1); mbar_init(q_free(g), 1);
      mbar_init(s_full(g), 1); mbar_init(p_full(g), 4); mbar_init(o_done(g), 1); mbar_init(o_free(g), 4);
    }
    for (int s = 0; s < A8_NSTG; s++) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_addr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem = *tmem_ptr_gen;
  auto adr = [](uint32_t a) { return (uint64_t)((a & 0x3FFFFu) >> 4); };

  if (warp == 0) {
    // ===================== TMA producer (converged warp, one elected lane issues) =====================
    int it[2] = {0, 0};
    uint32_t c = 0;                                                      // ring position: K_j at even, V_j at odd counts
    for (int r = 0;; r++) {
      int bh, qp;
      if (!item_of(r, bh, qp)) break;
      const int b = bh / H, h = bh % H, qtA = 2 * qp;
      const bool hasB = qtA + 1 < nq;
      const int nmax = tiles_of(hasB ? qtA + 1 : qtA);
      for (int g = 0; g < (hasB ? 2 : 1); g++) {
        mbar_wait(q_free(g), ((uint32_t)it[g] & 1u) ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(q_full(g), L::TILE);
#pragma unroll
          for (int sub = 0; sub < L::NSUB; sub++)
            tma_load_3d(base + L::Q + g * L::TILE + sub * L::SUBT, &tmQ, q_full(g), h * DH + 64 * sub, (qtA + g) * 128, b);
        }
        __syncwarp();
        it[g]++;
      }
      for (int j = 0; j < nmax; j++) {
#pragma unroll
        for (int kv = 0; kv < 2; kv++, c++) {
          const int s = c % A8_NSTG;
          mbar_wait(kv_empty(s), ((c / A8_NSTG) & 1u) ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(kv_full(s), L::TILE);
#pragma unroll
            for (int sub = 0; sub < L::NSUB; sub++)
              tma_load_3d(base + L::KV + s * L::TILE + sub * L::SUBT, kv ? &tmV : &tmK, kv_full(s), h * DH + 64 * sub, j * 128, b);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp, one elected lane issues) =====================
    const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);            // S = Q K^T : both operands K-major
    const uint32_t idesc_o = umma_idesc_bf16(128, DH, 0, 1);             // O = P V   : A from TMEM, V MN-major (head dim contiguous)
    const uint64_t d_k = umma_desc(0, 16, 1024, BPM_SWZ_128B);           // K-major 128-byte rows: +32 B per k16, next sub-tile after 4
    const uint64_t d_v = umma_desc(0, L::SUBT, 1024, BPM_SWZ_128B);      // MN-major: LBO = pitch of the 64-column sub-tiles, +2048 B per k16
    int it[2] = {0, 0}, tcg[2] = {0, 0};
    uint32_t c = 0;
    for (int r = 0;; r++) {
      int bh, qp;
      if (!item_of(r, bh, qp)) break;
      const int qtA = 2 * qp;
      const bool hasB = qtA + 1 < nq;
      const int ng[2] = {tiles_of(qtA), hasB ? tiles_of(qtA + 1) : 0};
      const int nmax = max(ng[0], ng[1]);
      const int G2 = hasB ? 2 : 1;
      for (int g = 0; g < G2; g++) mbar_wait(q_full(g), (uint32_t)it[g] & 1u);
      const uint32_t c0 = c;
      // S_j(g) = Q_g K_j^T into the (S | P) region of g; K_j sits in ring slot `ks`
      auto issue_s = [&](int g, int ks, bool last_s, bool last_user) {
        tc_fence_after();
        if (elect_one()) {
          const uint64_t dq = d_k | adr(base + L::Q + g * L::TILE), dk = d_k | adr(base + L::KV + ks * L::TILE);
#pragma unroll
          for (int k = 0; k < DH / 16; k++) {
            const uint64_t off = (uint64_t)(((k >> 2) * L::SUBT + (k & 3) * 32) >> 4);
            umma_bf16(tmem + 128 * g, dq + off, dk + off, idesc_s, (uint32_t)k);
          }
          umma_commit(s_full(g));
          if (last_s) umma_commit(q_free(g));                              // this item's Q_g has been read for the last time
          if (last_user) umma_commit(kv_empty(ks));
        }
        __syncwarp();
      };
      {
        const int ks = c0 % A8_NSTG;
        mbar_wait(kv_full(ks), (c0 / A8_NSTG) & 1u);
        for (int g = 0; g < G2; g++) issue_s(g, ks, ng[g] == 1, g == G2 - 1);
      }
      for (int j = 0; j < nmax; j++) {
        const uint32_t cv = c0 + 2 * j + 1, ck = c0 + 2 * j + 2;
        const int vs = cv % A8_NSTG, ks = ck % A8_NSTG;
        bool v_ready = false, k_ready = false;
        for (int g = 0; g < G2; g++) {
          if (j >= ng[g]) continue;
          mbar_wait(p_full(g), (uint32_t)tcg[g] & 1u);                     // P_j(g) is in TMEM (and O_g has been rescaled if needed)
          if (!v_ready) { mbar_wait(kv_full(vs), (cv / A8_NSTG) & 1u); v_ready = true; }
          if (j == 0) mbar_wait(o_free(g), ((uint32_t)it[g] & 1u) ^ 1u);   // the previous item's epilogue has read O_g
          tc_fence_after();
          if (elect_one()) {
            const uint64_t dv = d_v | adr(base + L::KV + vs * L::TILE);
            const uint32_t tO = tmem + 256 + DH * g, tP = tmem + 128 * g;
#pragma unroll
            for (int k = 0; k < 8; k++) umma_bf16_ts(tO, tP + 8 * k, dv + (uint64_t)((k * 2048) >> 4), idesc_o, (uint32_t)(j > 0) | (uint32_t)k);
            if (j == ng[g] - 1) umma_commit(o_done(g));
            if (g == G2 - 1) umma_commit(kv_empty(vs));                    // (tile B sees every key tile that tile A sees: it is the last user)
          }
          __syncwarp();
          tcg[g]++;
          if (j + 1 < ng[g]) {
            if (!k_ready) { mbar_wait(kv_full(ks), (ck / A8_NSTG) & 1u); k_ready = true; }
            issue_s(g, ks, j + 2 == ng[g], g == G2 - 1);
          }
        }
      }
      c = c0 + 2 * nmax;
      for (int g = 0; g < G2; g++) it[g]++;
    }
  } else {
    // ===================== softmax warps: one thread per query row =====================
    const int g = (warp - 2) >> 2;                                         // query tile of the pair
    const int quarter = warp & 3;                                          // TMEM lane quarter this warp may access
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int rr = quarter * 32 + lane;
    const uint32_t tS = tmem + 128 * g + lane_off, tO = tmem + 256 + DH * g + lane_off;
    uint8_t* const stg = base_gen + L::STG + (warp - 2) * 4096;
    const uint32_t stg_s = base + L::STG + (warp - 2) * 4096;
    const DropCtx dc = make_drop(drop);
    const int W = (S + 31) >> 5;
    int it = 0, tc = 0;
    for (int r = 0;; r++) {
      int bh, qp;
      if (!item_of(r, bh, qp)) break;
      const int qt = 2 * qp + g;
      if (qt >= nq) continue;
      const int b = bh / H, h = bh % H, q0 = qt * 128, qi = q0 + rr, nt = tiles_of(qt);
      const int row_lim = (mask_off >= 0) ? min(qi + mask_off, S - 1) : S - 1;     // last visible key of this row
      // warp-uniform visibility of a 32-key chunk: keys <= vis_all are visible to all 32 rows of this warp, keys > vis_any to none
      const int vis_all = (mask_off >= 0) ? min(q0 + quarter * 32 + mask_off, S - 1) : S - 1;
      const int vis_any = (mask_off >= 0) ? min(q0 + quarter * 32 + 31 + mask_off, S - 1) : S - 1;
      const uint64_t ebase = ((uint64_t)bh * T + (uint64_t)min(qi, T - 1)) * (uint64_t)S;
      float m_used = -INFINITY, l = 0.f;                                   // reference maximum (log2 domain) of the exponentials so far; row sum
      for (int j = 0; j < nt; j++, tc++) {
        const int k0 = j * 128;
        const int nvis = max(0, min(4, (vis_any - k0 + 32) >> 5));         // chunks of this tile with at least one visible key
        uint32_t padm[4] = {0u, 0u, 0u, 0u};
        if (key_pad != nullptr) {
#pragma unroll
          for (int ci = 0; ci < 4; ci++) {
            const int key = k0 + ci * 32 + lane;
            padm[ci] = __ballot_sync(0xffffffffu, key < S && key_pad[(int64_t)b * S + key] != 0);
          }
        }
        mbar_wait(s_full(g), (uint32_t)tc & 1u);
        tc_fence_after();
        // ---- pass 1: row maximum (two 32-column chunks in flight per wait)
        float mx = -INFINITY;
        {
          float sa[32], sb[32];
          auto chunk_max = [&](float* sv, int ci) {
            const int c = ci * 32;
            if (k0 + c + 31 > vis_all || padm[ci]) a8_mask32(sv, row_lim - (k0 + c), padm[ci]);
            float m4[4] = {sv[0], sv[1], sv[2], sv[3]};
#pragma unroll
            for (int e = 4; e < 32; e += 4) {
              m4[0] = fmaxf(m4[0], sv[e]); m4[1] = fmaxf(m4[1], sv[e + 1]); m4[2] = fmaxf(m4[2], sv[e + 2]); m4[3] = fmaxf(m4[3], sv[e + 3]);
            }
            mx = fmaxf(mx, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
          };
#pragma unroll
          for (int cp = 0; cp < 2; cp++) {
            if (2 * cp < nvis) {
              tmem_ld32(tS + 64 * cp, sa);
              if (2 * cp + 1 < nvis) tmem_ld32(tS + 64 * cp + 32, sb);
              tmem_ld_wait();
              chunk_max(sa, 2 * cp);
              if (2 * cp + 1 < nvis) chunk_max(sb, 2 * cp + 1);
            }
          }
        }
        // ---- lazy rescale: only when the row maximum has outgrown the reference by more than 2^8
        const float cand = mx * LOG2E_F;
        if (j == 0) {
          m_used = cand;
        } else {
          const bool need = cand > m_used + 8.f;
          if (__any_sync(0xffffffffu, need)) {                             // (S_j complete => the previous P V into O_g is complete as well)
            const float alpha = need ? ex2f(m_used - cand) : 1.f;          // m_used = -inf: alpha = 0 (nothing accumulated yet)
#pragma unroll 1
            for (int oc = 0; oc < DH; oc += 32) {
              float ov[32];
              tmem_ld32(tO + oc, ov);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 32; e += 2) fmul2(ov[e], ov[e + 1], alpha);
              tmem_st32(tO + oc, (const uint32_t*)ov);
            }
            l *= alpha;
            if (need) m_used = cand;
          }
        }
        const float neg_m = (m_used == -INFINITY) ? 0.f : -m_used;         // a row that has not seen a visible key yet: exp2(-inf - 0) = 0
        // ---- pass 2: P = exp2(S log2e - m_used), row sum, dropout, packed bf16 over the consumed score columns
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll 1
        for (int ci = 0; ci < 4; ci++) {                                   // (not unrolled / prefetched: measured 10 % slower, the loop outgrows the instruction cache)
          const int c = ci * 32;
          uint32_t pk[16];
          if (ci >= nvis) {
#pragma unroll
            for (int u = 0; u < 16; u++) pk[u] = 0u;
            tmem_st16(tS + (uint32_t)(c >> 1), pk);
            continue;
          }
          float sv[32];
          tmem_ld32(tS + c, sv);
          tmem_ld_wait();
          if (k0 + c + 31 > vis_all || padm[ci]) a8_mask32(sv, row_lim - (k0 + c), padm[ci]);
          float r4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            ffma2(sv[e], sv[e + 1], LOG2E_F, neg_m);
            ffma2(sv[e + 2], sv[e + 3], LOG2E_F, neg_m);
#pragma unroll
            for (int u = 0; u < 4; u++) sv[e + u] = ex2f(sv[e + u]);
            fadd2(r4[0], r4[1], sv[e], sv[e + 1]);
            fadd2(r4[2], r4[3], sv[e + 2], sv[e + 3]);
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
          tmem_st16(tS + (uint32_t)(c >> 1), pk);
        }
        l += rs0 + rs1;
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(g));
      }
      // ---- epilogue: O_g / l -> bf16 -> staging (32 rows x 64 columns, SWIZZLE_128B) -> TMA store; rows beyond T are clipped by the map
      mbar_wait(o_done(g), (uint32_t)it & 1u);
      tc_fence_after();
      const float inv_l = (DROP ? dc.inv_keep : 1.f) / l;
#pragma unroll 1
      for (int sub = 0; sub < L::NSUB; sub++) {
        if (elect_one()) bulk_wait_read<0>();                              // the previous store has read the staging slice
        __syncwarp();
#pragma unroll
        for (int c2 = 0; c2 < 2; c2++) {
          float ov[32];
          tmem_ld32(tO + 64 * sub + 32 * c2, ov);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 4; u++) {
            uint4 w;
            w.x = pack_bf16x2(ov[u * 8 + 0] * inv_l, ov[u * 8 + 1] * inv_l); w.y = pack_bf16x2(ov[u * 8 + 2] * inv_l, ov[u * 8 + 3] * inv_l);
            w.z = pack_bf16x2(ov[u * 8 + 4] * inv_l, ov[u * 8 + 5] * inv_l); w.w = pack_bf16x2(ov[u * 8 + 6] * inv_l, ov[u * 8 + 7] * inv_l);
            *(uint4*)(stg + lane * 128 + (((c2 * 4 + u) ^ (lane & 7)) << 4)) = w;
          }
        }
        if (sub == L::NSUB - 1) {                                          // O_g is in registers / staged: the next item may overwrite it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(o_free(g));
        }
        fence_async_smem();
        __syncwarp();
        if (elect_one()) {
          tma_store_3d(&tmO, stg_s, h * DH + 64 * sub, q0 + quarter * 32, b);
          bulk_commit();
        }
        __syncwarp();
      }
      if (qi < T) lse[(int64_t)bh * T + qi] = (m_used + log2f(l)) * LN2_F;
      it++;
    }
    if (elect_one()) bulk_wait_read<0>();                                 // the staging slice must outlive the last TMA store's read
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// =====================================================================================================================
// Backward (head dim 128).  See the file header.  Warp roles:
//   warp 0     TMA producer: K_j / V_j once per item; per half tile Q_h / dO_h (64 queries x 128) through a 3-stage ring together with
//              that half's lse * log2e and delta rows (1-D bulk copies)
//   warp 1     issues S^T and dP^T (two TMEM generations: X_0 | X_1 and Y_0 | Y_1, 64 columns each)
//   warp 2     issues dV += P~^T dO_h (A = P~^T from TMEM), dK += dS^T Q_h, dQ_h^T = K_j^T dS_h (into the Y generation dP^T came from)
//   warps 3-10 compute: two warps per TMEM lane quarter (32 of the 64 query columns each); thread = key row; both tiles' 32 columns are
//              loaded up front (64 values in flight), so the exponentials of a row have 32-way instruction-level parallelism
//   warps 11-14 drain: one per lane quarter.  dQ_h^T of every half tile (thread = head dim) goes through two alternating 4 KB staging
//              slices as [32 query][32 dim] fp32 blocks into TMA reduce-adds; at the end of an item dK_j and dV_j leave the same way
//              (bf16, TMA stores).  The compute warps never wait for a drain or a store (measured before the split: 27 % of their time
//              waiting for the next S^T / dP^T, which could not be issued before they had drained dQ^T, 10 % for staging reads).
// TMEM: X_0 64 | X_1 64 | Y_0 64 | Y_1 64 | dK 128 | dV 128 = 512 columns.
// P~^T is packed over the S^T columns its own warp has consumed: queries [0,32) -> X columns [0,16), queries [32,64) -> [32,48).
// =====================================================================================================================
#define B8_CW 8                         // compute warps (3 .. 10)
#define B8_DW 4                         // drain warps (11 .. 14), one per TMEM lane quarter
#define B8_THREADS (96 + 32 * (B8_CW + B8_DW))
#define B8_QS 3

struct B8Smem {
  static constexpr int KV = 0;                                   // K_j 32 KB (2 sub-tiles of 128 rows x 128 B), V_j 32 KB
  static constexpr int QD = KV + 2 * 32768;                      // stages x (Q_h 16 KB: 2 sub-tiles of 64 rows x 128 B | dO_h 16 KB)
  static constexpr int DST = QD + B8_QS * 32768;                 // dS^T: 2 buffers x (128 keys x 64 queries bf16, 128-byte rows)
  static constexpr int STG = DST + 2 * 16384;                    // per drain warp 2 slices of 4 KB: dQ^T blocks / dK, dV slices on their way out
  static constexpr int LD = STG + B8_DW * 8192;                  // stages x (64 floats lse*log2e | 64 floats delta)
  static constexpr int BAR = LD + B8_QS * 512;
  static constexpr int NBAR = 2 + 2 * B8_QS + 14;
  static constexpr int TOTAL = BAR + 8 * NBAR + 16;
};
static_assert(B8Smem::TOTAL + 1024 <= 227 * 1024, "backward shared memory budget");

// delta = rowsum(dO * O), lse * log2e, and the zeroing of the fp32 dQ accumulator: 16 threads per (row, head), 8 columns each
__global__ void attn128_prep_kernel(const bf16* __restrict__ out, const bf16* __restrict__ dout, const float* __restrict__ lse, float* __restrict__ ws,
                                    float* __restrict__ dq_acc, int B, int T, int H) {
  pdl_trigger();
  pdl_wait();
  const int64_t n = (int64_t)B * T * H * 16;
  for (int64_t base = blockIdx.x * (int64_t)blockDim.x; base < n; base += (int64_t)gridDim.x * blockDim.x) {      // (warp-uniform trip count)
    const int64_t idx = base + threadIdx.x;
    const bool ok = idx < n;
    const int part = (int)(idx & 15);
    const int64_t rh = ok ? (idx >> 4) : 0;                      // (row, head)
    const int h = (int)(rh % H);
    const int64_t row = rh / H;
    const int64_t off = (row * H + h) * 128 + part * 8;
    float acc = 0.f;
    if (ok) {
      Vec8<bf16> a, c;
      a.load(out + off); c.load(dout + off);
#pragma unroll
      for (int j = 0; j < 8; j++) acc = fmaf(a.v[j], c.v[j], acc);
      *(float4*)(dq_acc + off) = make_float4(0.f, 0.f, 0.f, 0.f);
      *(float4*)(dq_acc + off + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (ok && part == 0) {
      const int t = (int)(row % T), b = (int)(row / T);
      const int64_t o_idx = ((int64_t)b * H + h) * T + t;
      ws[o_idx] = acc;
      ws[(int64_t)B * T * H + o_idx] = lse[o_idx] * LOG2E_F;
    }
  }
}

// dq (bf16) = dq_acc (fp32) * scale
__global__ void attn128_dq_cast_kernel(const float* __restrict__ acc, bf16* __restrict__ dq, int64_t n8, float scale) {
  pdl_trigger();
  pdl_wait();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    Vec8<float> a;
    a.load(acc + i * 8);
    Vec8<bf16> o;
#pragma unroll
    for (int j = 0; j < 8; j++) o.v[j] = a.v[j] * scale;
    o.store(dq + i * 8);
  }
}

template <bool DROP>
__global__ void __launch_bounds__(B8_THREADS, 1)
attn128_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const __grid_constant__ CUtensorMap tmdO, const __grid_constant__ CUtensorMap tmDQ, const __grid_constant__ CUtensorMap tmDK,
                   const __grid_constant__ CUtensorMap tmDV, const float* __restrict__ ws, bf16* __restrict__ dk, bf16* __restrict__ dv, int B, int T,
                   int S, int H, int mask_off, const uint8_t* __restrict__ key_pad, bpm_dropout_t drop, const uint32_t* __restrict__ drop_bits,
                   const int ld_dkv) {
  constexpr int DH = 128;
  using L = B8Smem;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + L::BAR;
  const uint32_t kv_full = bar0, kv_empty = bar0 + 8u;
  auto q_full = [&](int s) { return bar0 + 8u * (2 + s); };
  auto q_empty = [&](int s) { return bar0 + 8u * (2 + B8_QS + s); };
  constexpr int B1 = 2 + 2 * B8_QS;
  auto st_full = [&](int b) { return bar0 + 8u * (B1 + b); };          // S^T / dP^T of generation b are in TMEM
  auto x_free = [&](int b) { return bar0 + 8u * (B1 + 2 + b); };       // the dV MMAs have read P~^T of generation b
  auto y_free = [&](int b) { return bar0 + 8u * (B1 + 4 + b); };       // dQ^T of generation b has been drained
  auto pt_full = [&](int b) { return bar0 + 8u * (B1 + 6 + b); };      // P~^T (TMEM) and dS^T (smem buffer b) are written
  auto ds_free = [&](int b) { return bar0 + 8u * (B1 + 8 + b); };      // the dK / dQ MMAs have read dS^T buffer b
  auto dq_full = [&](int b) { return bar0 + 8u * (B1 + 10 + b); };
  const uint32_t dkv_full = bar0 + 8u * (B1 + 12), dkv_free = bar0 + 8u * (B1 + 13);
  const uint32_t tmem_ptr_addr = bar0 + 8u * L::NBAR;
  volatile uint32_t* tmem_ptr_gen = (volatile uint32_t*)(base_gen + L::BAR + 8 * L::NBAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nbh = B * H;
  const int64_t nrow = (int64_t)nbh * T;
  const int nqh = (T + 63) / 64, nkv = (S + 127) / 128;
  const int n_items = nbh * nkv;
  // first half tile of queries that can see key tile j (visible iff key <= q + off)
  auto ih_min_of = [&](int j) { const int qlo = j * 128 - mask_off; return (mask_off < 0 || qlo <= 0) ? 0 : qlo / 64; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO); tma_prefetch_desc(&tmDQ);
    mbar_init(kv_full, 1); mbar_init(kv_empty, 1);
    for (int s = 0; s < B8_QS; s++) { mbar_init(q_full(s), 1); mbar_init(q_empty(s), 1); }
    for (int b = 0; b < 2; b++) {
      mbar_init(st_full(b), 1); mbar_init(x_free(b), 1); mbar_init(y_free(b), B8_DW); mbar_init(pt_full(b), B8_CW);
      mbar_init(ds_free(b), 1); mbar_init(dq_full(b), 1);
    }
    mbar_init(dkv_full, 1); mbar_init(dkv_free, B8_DW);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_addr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem = *tmem_ptr_gen;
  const uint32_t tDK = tmem + 256, tDV = tmem + 384;
  auto tX = [&](int b) { return tmem + 64u * b; };
  auto tY = [&](int b) { return tmem + 128u + 64u * b; };
  auto adr = [](uint32_t a) { return (uint64_t)((a & 0x3FFFFu) >> 4); };

  if (warp == 0) {
    // ===================== TMA producer =====================
    uint32_t n = 0, ic = 0;
    for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x) {
      const int j = idx / nbh, bh = idx % nbh, b = bh / H, h = bh % H;
      const int ih0 = ih_min_of(j);
      if (ih0 >= nqh) continue;
      mbar_wait(kv_empty, (ic & 1u) ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(kv_full, 2 * 32768);
#pragma unroll
        for (int sub = 0; sub < 2; sub++) {
          tma_load_3d(base + L::KV + sub * 16384, &tmK, kv_full, h * DH + 64 * sub, j * 128, b);
          tma_load_3d(base + L::KV + 32768 + sub * 16384, &tmV, kv_full, h * DH + 64 * sub, j * 128, b);
        }
      }
      __syncwarp();
      ic++;
      for (int ih = ih0; ih < nqh; ih++, n++) {
        const int s = n % B8_QS;
        mbar_wait(q_empty(s), ((n / B8_QS) & 1u) ^ 1u);
        if (elect_one()) {
          const uint32_t nvalid = (uint32_t)min(64, T - ih * 64);
          mbar_expect_tx(q_full(s), 32768 + 2 * nvalid * 4);
          const uint32_t qa = base + L::QD + s * 32768;
#pragma unroll
          for (int sub = 0; sub < 2; sub++) {
            tma_load_3d(qa + sub * 8192, &tmQ, q_full(s), h * DH + 64 * sub, ih * 64, b);
            tma_load_3d(qa + 16384 + sub * 8192, &tmdO, q_full(s), h * DH + 64 * sub, ih * 64, b);
          }
          bulk_load_1d(base + L::LD + s * 512, ws + nrow + (int64_t)bh * T + ih * 64, nvalid * 4, q_full(s));
          bulk_load_1d(base + L::LD + s * 512 + 256, ws + (int64_t)bh * T + ih * 64, nvalid * 4, q_full(s));
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== S^T = K_j Q_h^T and dP^T = V_j dO_h^T (M128 N64 K128) =====================
    const uint32_t id_st = umma_idesc_bf16(128, 64, 0, 0);
    const uint64_t d_k = umma_desc(0, 16, 1024, BPM_SWZ_128B);            // K-major 128-byte rows; +32 B per k16 inside a 64-column sub-tile
    uint32_t n = 0, ic = 0;
    for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x) {
      const int j = idx / nbh;
      const int ih0 = ih_min_of(j);
      if (ih0 >= nqh) continue;
      mbar_wait(kv_full, ic & 1u);
      ic++;
      for (int ih = ih0; ih < nqh; ih++, n++) {
        const int s = n % B8_QS, b = n & 1;
        mbar_wait(q_full(s), (n / B8_QS) & 1u);
        mbar_wait(x_free(b), ((n >> 1) & 1u) ^ 1u);
        mbar_wait(y_free(b), ((n >> 1) & 1u) ^ 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t dka = d_k | adr(base + L::KV), dva = d_k | adr(base + L::KV + 32768);
          const uint64_t dqa = d_k | adr(base + L::QD + s * 32768), dga = d_k | adr(base + L::QD + s * 32768 + 16384);
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const uint64_t oa = (uint64_t)(((k >> 2) * 16384 + (k & 3) * 32) >> 4), ob = (uint64_t)(((k >> 2) * 8192 + (k & 3) * 32) >> 4);
            umma_bf16(tX(b), dka + oa, dqa + ob, id_st, (uint32_t)k);
          }
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const uint64_t oa = (uint64_t)(((k >> 2) * 16384 + (k & 3) * 32) >> 4), ob = (uint64_t)(((k >> 2) * 8192 + (k & 3) * 32) >> 4);
            umma_bf16(tY(b), dva + oa, dga + ob, id_st, (uint32_t)k);
          }
          umma_commit(st_full(b));
        }
        __syncwarp();
      }
    }
  } else if (warp == 2) {
    // ===================== dV_j += P~^T dO_h,  dK_j += dS^T Q_h,  dQ_h^T = K_j^T dS_h =====================
    const uint32_t id_kv = umma_idesc_bf16(128, 128, 0, 1);              // A K-major (TMEM / dS^T tile), B = dO_h / Q_h MN-major
    const uint32_t id_dq = umma_idesc_bf16(128, 64, 1, 1);               // A = K_j read MN-major (M = head dim), B = dS^T read MN-major (N = queries)
    const uint64_t d_qmn = umma_desc(0, 8192, 1024, BPM_SWZ_128B);       // Q_h / dO_h as MN-major B: 2 blocks of 64 head dims, +2048 B per k16 (16 queries)
    const uint64_t d_dsk = umma_desc(0, 16, 1024, BPM_SWZ_128B);         // dS^T K-major (queries contiguous): +32 B per k16
    const uint64_t d_kmn = umma_desc(0, 16384, 1024, BPM_SWZ_128B);      // K_j as MN-major A: 2 blocks of 64 head dims, +2048 B per k16 (16 keys)
    const uint64_t d_dsmn = umma_desc(0, 16384, 1024, BPM_SWZ_128B);     // dS^T as MN-major B (one block of 64 queries), +2048 B per k16 (16 keys)
    uint32_t n = 0, ic = 0;
    for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x) {
      const int j = idx / nbh;
      const int ih0 = ih_min_of(j);
      if (ih0 >= nqh) continue;
      for (int ih = ih0; ih < nqh; ih++, n++) {
        const int s = n % B8_QS, b = n & 1;
        mbar_wait(pt_full(b), (n >> 1) & 1u);
        if (ih == ih0) mbar_wait(dkv_free, (ic & 1u) ^ 1u);                // dK / dV of the previous item have been drained
        tc_fence_after();
        if (elect_one()) {
          const uint32_t acc = ih > ih0 ? 1u : 0u;
          const uint32_t qa = base + L::QD + s * 32768, da = base + L::DST + b * 16384;
          const uint64_t dgb = d_qmn | adr(qa + 16384), dqb = d_qmn | adr(qa);
          const uint64_t dda = d_dsk | adr(da), ddb = d_dsmn | adr(da), dka = d_kmn | adr(base + L::KV);
#pragma unroll
          for (int k = 0; k < 4; k++)                                      // queries 16k .. 16k+15: packed columns 8k (k < 2) / 32 + 8(k-2)
            umma_bf16_ts(tDV, tX(b) + (uint32_t)((k >> 1) * 32 + (k & 1) * 8), dgb + (uint64_t)((k * 2048) >> 4), id_kv, acc | (uint32_t)k);
          umma_commit(x_free(b));
#pragma unroll
          for (int k = 0; k < 4; k++) umma_bf16(tDK, dda + (uint64_t)((k * 32) >> 4), dqb + (uint64_t)((k * 2048) >> 4), id_kv, acc | (uint32_t)k);
#pragma unroll
          for (int k = 0; k < 8; k++) umma_bf16(tY(b), dka + (uint64_t)((k * 2048) >> 4), ddb + (uint64_t)((k * 2048) >> 4), id_dq, (uint32_t)k);
          umma_commit(dq_full(b));
          umma_commit(q_empty(s));
          umma_commit(ds_free(b));
          if (ih == nqh - 1) { umma_commit(dkv_full); umma_commit(kv_empty); }
        }
        __syncwarp();
      }
      ic++;
    }
  } else if (warp < 3 + B8_CW) {
    // ===================== compute warps =====================
    const int cw = warp - 3;
    const int quarter = warp & 3, colh = cw >> 2;                          // TMEM lane quarter; query-column half [32 colh, +32) of the half tile
    const int r = quarter * 32 + lane;                                     // key row of the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const DropCtx dc = make_drop(drop);
    const int W = (S + 31) >> 5;
    const int drop_mode = !DROP ? 0 : (!dc.on ? 0 : (drop_bits != nullptr ? 2 : 3));
    // dS^T row r: 128-byte rows, 8-row groups of 1024 B, 16-byte unit u at u ^ (r & 7)
    const uint32_t row_off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
    uint32_t n = 0;
    for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x) {
      const int j = idx / nbh, bh = idx % nbh, b = bh / H, h = bh % H;
      const int ih0 = ih_min_of(j);
      const int key = j * 128 + r;
      if (ih0 >= nqh) {                                                    // no query sees this key tile: dK = dV = 0
        if (key < S) {
          bf16* dst = (colh ? dv : dk) + ((int64_t)b * S + key) * ld_dkv + h * DH;
#pragma unroll
          for (int u = 0; u < DH / 8; u++) *(uint4*)(dst + u * 8) = make_uint4(0, 0, 0, 0);
        }
        continue;
      }
      const bool row_dead = key >= S || (key_pad != nullptr && key_pad[(int64_t)b * S + min(key, S - 1)] != 0);
      const bool any_dead = __any_sync(0xffffffffu, row_dead);
      const uint64_t e_row = (uint64_t)bh * (uint64_t)T * (uint64_t)S + (uint64_t)min(key, S - 1);
      for (int ih = ih0; ih < nqh; ih++, n++) {
        const int s = n % B8_QS, pb = n & 1;
        const int q0 = ih * 64, qc = q0 + 32 * colh;                       // first query of this warp's columns
        // columns (relative to qc) below cmin are masked for this key row; columns from cmax on are beyond T
        const int cmin = (mask_off >= 0) ? key - mask_off - qc : -0x40000000;
        const int cmax = T - qc;
        const bool need_mask = any_dead || cmax < 32 || (mask_off >= 0 && j * 128 + quarter * 32 + 31 - mask_off > qc);
        const float* lse_s = (const float*)(base_gen + L::LD + s * 512) + 32 * colh;
        const float* del_s = lse_s + 64;
        uint32_t mw = 0xFFFFFFFFu;                                         // keep bits of (query qc + lane, this warp's 32 keys)
        if (DROP && drop_mode == 2) {
          const int qq = qc + lane;
          mw = (qq < T && j * 4 + quarter < W) ? drop_bits[((int64_t)bh * T + qq) * W + j * 4 + quarter] : 0u;
        }
        mbar_wait(st_full(pb), (n >> 1) & 1u);
        tc_fence_after();
        float sv[32], dpv[32];                                             // both tiles' columns of this warp in flight at once
        tmem_ld32(tX(pb) + lane_off + 32 * colh, sv);
        tmem_ld32(tY(pb) + lane_off + 32 * colh, dpv);
        mbar_wait(ds_free(pb), ((n >> 1) & 1u) ^ 1u);                      // the MMAs of two half tiles ago have read this dS^T buffer
        uint8_t* const dtile = base_gen + L::DST + pb * 16384 + row_off;
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          const float4 l4 = *(const float4*)(lse_s + c);
          const float4 d4 = *(const float4*)(del_s + c);
          const float ls[4] = {l4.x, l4.y, l4.z, l4.w}, dl[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const int cc = c + e;
            float p = ex2f(fmaf(sv[cc], LOG2E_F, -ls[e]));
            float ds;
            if (!DROP || drop_mode == 0) {
              ds = p * (dpv[cc] - dl[e]);
            } else {
              float mult;
              if (drop_mode == 2) mult = ((__shfl_sync(0xffffffffu, mw, cc) >> lane) & 1u) ? dc.inv_keep : 0.f;
              else mult = drop_mult1(dc, ((uint64_t)min(qc + cc, T - 1)) * (uint64_t)S + e_row);
              ds = p * fmaf(dpv[cc], mult, -dl[e]);                        // dS^T
              p *= mult;                                                   // P~^T
            }
            // (the lse / delta rows of query columns beyond T are not loaded: their products may be anything, the selects discard them)
            const bool dead = need_mask && (row_dead || cc < cmin || cc >= cmax);
            sv[cc] = dead ? 0.f : p;
            dpv[cc] = dead ? 0.f : ds;
          }
        }
        uint32_t pk[16];
#pragma unroll
        for (int u = 0; u < 16; u++) pk[u] = pack_bf16x2(sv[2 * u], sv[2 * u + 1]);
        tmem_st16(tX(pb) + lane_off + (uint32_t)(32 * colh), pk);          // queries [32 colh, +32) of the half tile -> 16 packed columns
#pragma unroll
        for (int uu = 0; uu < 4; uu++) {
          const int u = colh * 4 + uu;                                     // 16-byte unit of the 128-byte dS^T row
          *(uint4*)(dtile + ((u ^ (r & 7)) << 4)) =
              make_uint4(pack_bf16x2(dpv[uu * 8], dpv[uu * 8 + 1]), pack_bf16x2(dpv[uu * 8 + 2], dpv[uu * 8 + 3]),
                         pack_bf16x2(dpv[uu * 8 + 4], dpv[uu * 8 + 5]), pack_bf16x2(dpv[uu * 8 + 6], dpv[uu * 8 + 7]));
        }
        tmem_st_wait();
        tc_fence_before();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(pt_full(pb));
      }
    }
  } else {
    // ===================== drain warps: dQ_h^T of every half tile (TMA reduce-add), dK_j / dV_j at the end of an item =====================
    const int dw = warp - (3 + B8_CW);
    const int quarter = warp & 3;                                          // lanes = head dims [32 quarter, +32) of dQ^T / key rows of dK, dV
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    uint8_t* const stg = base_gen + L::STG + dw * 8192;                    // two 4 KB slices, used alternately
    const uint32_t stg_s = base + L::STG + dw * 8192;
    uint32_t n = 0, ic = 0, sl = 0;
    for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x) {
      const int j = idx / nbh, bh = idx % nbh, b = bh / H, h = bh % H;
      const int ih0 = ih_min_of(j);
      if (ih0 >= nqh) continue;
      for (int ih = ih0; ih < nqh; ih++, n++) {
        const int pb = n & 1;
        mbar_wait(dq_full(pb), (n >> 1) & 1u);
        tc_fence_after();
        float acc[64];                                                     // lane = head dim, columns = the 64 queries of the half tile
        tmem_ld32(tY(pb) + lane_off, acc);
        tmem_ld32(tY(pb) + lane_off + 32, acc + 32);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(y_free(pb));
#pragma unroll
        for (int hq = 0; hq < 2; hq++, sl ^= 1u) {                         // 32 queries per slice: [query][dim] fp32, a warp writes 128 contiguous bytes
          if (elect_one()) bulk_wait_read<1>();                            // the store that last used this slice (two stores ago) has read it
          __syncwarp();
          uint8_t* const sp = stg + sl * 4096;
#pragma unroll
          for (int c = 0; c < 32; c++) *(float*)(sp + c * 128 + lane * 4) = acc[hq * 32 + c];
          fence_async_smem();
          __syncwarp();
          if (elect_one()) {
            tma_reduce_add_3d(&tmDQ, stg_s + sl * 4096, h * DH + 32 * quarter, ih * 64 + 32 * hq, b);
            bulk_commit();
          }
          __syncwarp();
        }
      }
      // ---- end of the item: dK_j then dV_j of this lane quarter's 32 key rows, 2 x (32 rows x 64 columns bf16) slices each
      mbar_wait(dkv_full, ic & 1u);
      tc_fence_after();
      ic++;
#pragma unroll 1
      for (int part = 0; part < 4; part++, sl ^= 1u) {                     // dK columns [0,64), [64,128), dV columns [0,64), [64,128)
        if (elect_one()) bulk_wait_read<1>();
        __syncwarp();
        uint8_t* const sp = stg + sl * 4096;
#pragma unroll
        for (int c2 = 0; c2 < 2; c2++) {
          float acc[32];
          tmem_ld32(((part >> 1) ? tDV : tDK) + lane_off + 64 * (part & 1) + 32 * c2, acc);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 4; u++)
            *(uint4*)(sp + lane * 128 + (((c2 * 4 + u) ^ (lane & 7)) << 4)) =
                make_uint4(pack_bf16x2(acc[u * 8], acc[u * 8 + 1]), pack_bf16x2(acc[u * 8 + 2], acc[u * 8 + 3]),
                           pack_bf16x2(acc[u * 8 + 4], acc[u * 8 + 5]), pack_bf16x2(acc[u * 8 + 6], acc[u * 8 + 7]));
        }
        if (part == 3) {                                                   // both accumulators are out of TMEM: the next item may start
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(dkv_free);
        }
        fence_async_smem();
        __syncwarp();
        if (elect_one()) {
          tma_store_3d((part >> 1) ? &tmDV : &tmDK, stg_s + sl * 4096, h * DH + 64 * (part & 1), j * 128 + quarter * 32, b);
          bulk_commit();
        }
        __syncwarp();
      }
    }
    if (elect_one()) bulk_wait_read<0>();
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------- host
// [B, rows, H*DH] bf16 tensor (row pitch `pitch` elements) as a 3-D map with {64 columns, box_rows, 1} boxes, SWIZZLE_128B
static int a8_map(CUtensorMap* m, const void* p, int B, int rows, int HP, int box_rows, int pitch) {
  if (pitch == 0) pitch = HP;
  uint64_t dims[3] = {(uint64_t)HP, (uint64_t)rows, (uint64_t)B};
  uint64_t str[2] = {(uint64_t)pitch * 2, (uint64_t)rows * pitch * 2};
  uint32_t box[3] = {64, (uint32_t)box_rows, 1};
  return bpm_make_tmap_bf16(m, p, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

// head dims served by the kernels of this file: forward 64 / 128 (stored pitch), backward 128
int bpm_xattn128_supported(const bpm_attn_t* a, int backward) {
  if (a->dtype != BPM_BF16) return 0;
  if (backward) return a->dhp == 128 && a->T % 4 == 0;                   // (per-half lse / delta rows travel as 16-byte bulk copies)
  return a->dhp == 128 || a->dhp == 64;
}

// floats of workspace the backward needs behind `delta`: [2, B, H, T] (delta | lse*log2e) + the fp32 dQ accumulator [B, T, H*dhp]
int64_t bpm_xattn128_ws_floats(const bpm_attn_t* a) {
  const int64_t n = 2 * (int64_t)a->B * a->H * a->T;
  return (n + 31) / 32 * 32 + (int64_t)a->B * a->T * a->H * a->dhp;
}

template <int DH, bool DROP>
static int a8_fwd_launch(const bpm_attn_t* a, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to, float* lse,
                         cudaStream_t stream) {
  const size_t smem = A8Fwd<DH>::TOTAL + 1024;
  auto kern = attn128_fwd_kernel<DH, DROP>;
  if (int rc = bpm_func_smem((const void*)kern, (int)smem, "xattn_fwd_tc128")) return rc;
  const int nq = bpm_cdiv(a->T, 128), n_items = a->B * a->H * ((nq + 1) / 2);
  const int ctas = min(bpm_num_sms(), n_items);
  cudaError_t le = bpm_launch(kern, dim3(ctas), dim3(A8_THREADS), smem, stream, tq, tk, tv, to, lse, a->B, a->T, a->S, a->H, a->mask_off, a->key_pad,
                              a->drop, a->drop_bits);
  if (le != cudaSuccess) { bpm_set_error("xattn_fwd_tc128: launch failed: %s", cudaGetErrorString(le)); (void)cudaGetLastError(); return BPM_ELAUNCH; }
  return BPM_OK;
}

int bpm_xattn_fwd_tc128(const bpm_attn_t* a, const void* q, const void* k, const void* v, void* out, float* lse, cudaStream_t stream) {
  BPM_REQUIRE(((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) % 16 == 0, "xattn_fwd: pointers must be 16-byte aligned");
  BPM_REQUIRE(a->ld_kv % 8 == 0, "xattn_fwd: ld_kv must be a multiple of 8 elements");
  const int HP = a->H * a->dhp;
  CUtensorMap tq, tk, tv, to;
  int rc;
  if ((rc = a8_map(&tq, q, a->B, a->T, HP, 128, 0))) return rc;
  if ((rc = a8_map(&tk, k, a->B, a->S, HP, 128, a->ld_kv))) return rc;
  if ((rc = a8_map(&tv, v, a->B, a->S, HP, 128, a->ld_kv))) return rc;
  if ((rc = a8_map(&to, out, a->B, a->T, HP, 32, 0))) return rc;             // output: {64 columns, 32 rows} store boxes
  const bool dr = a->drop.p > 0.f;
  if (a->dhp == 128) return dr ? a8_fwd_launch<128, true>(a, tq, tk, tv, to, lse, stream) : a8_fwd_launch<128, false>(a, tq, tk, tv, to, lse, stream);
  return dr ? a8_fwd_launch<64, true>(a, tq, tk, tv, to, lse, stream) : a8_fwd_launch<64, false>(a, tq, tk, tv, to, lse, stream);
}

int bpm_xattn_bwd_tc128(const bpm_attn_t* a, const void* q, const void* k, const void* v, const void* out, const void* dout, const float* lse,
                        float* ws, void* dq, float dq_scale, void* dk, void* dv, cudaStream_t stream) {
  BPM_REQUIRE(((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out | (uintptr_t)dout | (uintptr_t)dq | (uintptr_t)dk | (uintptr_t)dv) % 16 == 0 &&
                  (uintptr_t)ws % 128 == 0, "xattn_bwd: pointers must be 16-byte aligned (workspace: 128)");
  BPM_REQUIRE(a->ld_kv % 8 == 0 && a->ld_dkv % 8 == 0, "xattn_bwd: ld_kv / ld_dkv must be multiples of 8 elements");
  const int HP = a->H * a->dhp;
  const int64_t n_ws = 2 * (int64_t)a->B * a->H * a->T;
  float* dq_acc = ws + (n_ws + 31) / 32 * 32;
  const int64_t n_dq = (int64_t)a->B * a->T * HP;
  {
    const int64_t n = (int64_t)a->B * a->T * a->H * 16;
    const int grid = (int)min((n + 255) / 256, (int64_t)bpm_num_sms() * 8);
    cudaError_t le = bpm_launch(attn128_prep_kernel, dim3(grid), dim3(256), 0, stream, (const bf16*)out, (const bf16*)dout, lse, ws, dq_acc, a->B, a->T, a->H);
    if (le != cudaSuccess) { bpm_set_error("xattn_prep128: launch failed: %s", cudaGetErrorString(le)); (void)cudaGetLastError(); return BPM_ELAUNCH; }
  }
  CUtensorMap tq, tk, tv, tg, tdq, tdk, tdv;
  int rc;
  if ((rc = a8_map(&tq, q, a->B, a->T, HP, 64, 0))) return rc;                // half tiles of 64 queries
  if ((rc = a8_map(&tg, dout, a->B, a->T, HP, 64, 0))) return rc;
  if ((rc = a8_map(&tk, k, a->B, a->S, HP, 128, a->ld_kv))) return rc;
  if ((rc = a8_map(&tv, v, a->B, a->S, HP, 128, a->ld_kv))) return rc;
  if ((rc = a8_map(&tdk, dk, a->B, a->S, HP, 32, a->ld_dkv))) return rc;      // {64 columns, 32 rows} store boxes
  if ((rc = a8_map(&tdv, dv, a->B, a->S, HP, 32, a->ld_dkv))) return rc;
  {
    uint64_t dims[3] = {(uint64_t)HP, (uint64_t)a->T, (uint64_t)a->B};
    uint64_t str[2] = {(uint64_t)HP * 4, (uint64_t)a->T * HP * 4};
    uint32_t box[3] = {32, 32, 1};                                            // [32 queries][32 head dims] fp32 reduce-add blocks
    if ((rc = bpm_make_tmap_f32(&tdq, dq_acc, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
  }
  const size_t smem = B8Smem::TOTAL + 1024;
  const bool dr = a->drop.p > 0.f;
  auto kern = dr ? attn128_bwd_kernel<true> : attn128_bwd_kernel<false>;
  if ((rc = bpm_func_smem((const void*)kern, (int)smem, "xattn_bwd_tc128"))) return rc;
  const int n_items = a->B * a->H * bpm_cdiv(a->S, 128);
  const int ctas = min(n_items, bpm_num_sms());
  cudaError_t le = bpm_launch(kern, dim3(ctas), dim3(B8_THREADS), smem, stream, tq, tk, tv, tg, tdq, tdk, tdv, (const float*)ws, (bf16*)dk, (bf16*)dv, a->B,
                              a->T, a->S, a->H, a->mask_off, a->key_pad, a->drop, (const uint32_t*)a->drop_bits, a->ld_dkv ? a->ld_dkv : HP);
  if (le != cudaSuccess) { bpm_set_error("xattn_bwd_tc128: launch failed: %s", cudaGetErrorString(le)); (void)cudaGetLastError(); return BPM_ELAUNCH; }
  {
    const int64_t n8 = n_dq / 8;
    const int grid = (int)min((n8 + 255) / 256, (int64_t)bpm_num_sms() * 8);
    le = bpm_launch(attn128_dq_cast_kernel, dim3(grid), dim3(256), 0, stream, (const float*)dq_acc, (bf16*)dq, n8, dq_scale);
    if (le != cudaSuccess) { bpm_set_error("xattn_dq_cast128: launch failed: %s", cudaGetErrorString(le)); (void)cudaGetLastError(); return BPM_ELAUNCH; }
  }
  return BPM_OK;
}
