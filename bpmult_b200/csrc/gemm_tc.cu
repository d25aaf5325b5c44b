// bf16 GEMM on 5th-gen tensor cores (sm_100a): TMA (SWIZZLE_128B) -> shared memory ring -> tcgen05.mma (cta_group::1,
// M = 128, N = BN <= 256, K = 16 per instruction) -> fp32 accumulator in TMEM -> tcgen05.ld -> fused epilogue -> TMA store.
//
//  * Persistent: one CTA per SM walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... (n-tile fastest, so the CTAs that run
//    together share A tiles through L2).  Two TMEM accumulators alternate between tiles: the MMA warp starts tile t+1 while
//    the epilogue warps drain tile t; the TMA producer keeps the operand ring full across tile boundaries.
//  * Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer (descriptors are
//    formed once and advanced by adds: the issuing thread is latency-bound, every instruction in its loop counts),
//    warps 2..9 = epilogue.  K is only 320..1216 here (5..19 k-blocks per tile), so the kernel lives or dies by its epilogue:
//    each epilogue warp owns the 32 accumulator rows of its TMEM lane quarter (warp % 4) and every second 128-byte column
//    chunk, and runs its OWN pipeline with no CTA-wide barrier: tcgen05.ld -> math -> swizzled staging buffer (conflict-free
//    16-byte stores) -> one TMA store (or TMA reduce-add for split-K wgrad) of a {chunk, 32 rows} box.  Two warps per
//    scheduler hide each other's TMEM / shared-memory latency.
//  * A residual (or relu-backward gate) tile is fetched by the same warp with a TMA load INTO its staging buffer two chunks
//    ahead, combined in place, and stored from there: every global access of the epilogue is a full 128-byte line and
//    M / N tails are clipped by the TMA unit.
//  * All four operand layouts without any transposed copies: K-major operands use the canonical SW128 K-major layout
//    (SBO = 1024 B); "transposed" operands (dgrad's W, wgrad's dY^T and X) are loaded as 64-wide MN chunks and described to
//    the MMA as MN-major (LBO = BK*128 B between chunks, SBO = 1024 B between 8-row k groups).
//  * wgrad (K = B*T rows) is split along K across tiles.  The bias gradient (column sums of dY) is fused into wgrad as one
//    extra N=16 MMA per k-step against a tile of ones (dY^T * 1), so dY is never re-read for it.
#include "gemm_epilogue.cuh"
#include "tc_common.cuh"

#define TC_BM 128
#define TC_BK 64
#define TC_EPI_WARPS 8
#define TC_THREADS (64 + 32 * TC_EPI_WARPS)
#define TC_MAX_STAGES 8
#define TC_MAX_NB 4

struct TcGemmParams {
  int M, N, K;
  int BN;                 // multiple of 32, <= 256
  int a_mn, b_mn;         // 1: operand is MN-major in memory ("transposed")
  int stages;
  int a_bytes, b_bytes;   // per-stage bytes (multiples of 1024)
  int ring_bytes;
  int ebuf_off, misc_off;  // byte offsets of the staging buffers and of the ones tile / barriers behind the 1024-aligned base
  int kb_per_split;
  int gx, gy, split;
  int tmem_cols;
  uint32_t idesc, idesc_ones;
  int n_chunks;           // column chunks per tile
  int nb;                 // staging buffers per epilogue warp (2..4)
  int in_mode;            // 0 none, 1 residual (added), 2 gate (relu-backward mask from a saved activation)
  int reduce_add;
  float* colsum;          // fused bias gradient (wgrad only), or null
  unsigned long long* trace;   // BPM_GEMM_TRACE builds only: per-role event log of CTA 0 (scripts/trace_gemm.py)
  int dbg;                // diagnostic knobs (bpm_debug_set slot 0): 1 no TMA loads, 2 no MMAs, 4 no epilogue math, 8 no staging/store, 16 no tcgen05.ld, 32 MMA issuer skips the full-barrier wait, 64 no empty-barrier traffic, 128 never pair CTAs
  const float* bias; float alpha; int act; float gate_scale; int ldc;
  bpm_dropout_t drop;
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)m), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)m), "r"(smem_src),
               "r"(c0), "r"(c1)
               : "memory");
}

// byte offset of 16-byte unit u of row r inside a swizzled [32 rows x CB] box (TMA SWIZZLE_128B / SWIZZLE_64B)
template <int CB> __device__ __forceinline__ uint32_t swz_off(int r, int u) {
  return CB == 128 ? (uint32_t)(r * 128 + ((u ^ (r & 7)) << 4)) : (uint32_t)(r * 64 + ((u ^ ((r >> 1) & 3)) << 4));
}

// ELEM: bytes per output element (2 = bf16, 4 = fp32); CB: bytes per row of one epilogue chunk (128, or 64 when BN * ELEM is not
// a multiple of 128).  CW = CB / ELEM accumulator columns per chunk.
// CG: 1 = one CTA per tile (M = 128);  2 = CTA pair (cluster of 2, tcgen05 cta_group::2): the pair owns a 256-row tile, each CTA loads
// its 128 rows of A and HALF of the B tile, so the per-SM operand traffic from L2 (the measured bound of the 1-CTA main loop,
// ~64 B/clk/SM) drops by a third; the leader CTA issues the MMAs, both CTAs run producer and epilogue warps on their own halves.
// EPI: epilogue specialisation.  0 = everything decided at run time; the others compile only what one hot shape of the encoder layer
// needs (a shorter instruction stream per chunk; the epilogue warps are instruction-fetch sensitive):
//   1 plain / bias / bias + relu     2 + dropout (fc1)     3 relu-backward gate tile (dgrad into fc1)
//   4 dropout + residual tile, fp32 out (out-proj, fc2)     5 fp32 reduce-add with fused column sums (wgrad)
#ifdef BPM_GEMM_TRACE
#define GT_N 1024
#define GT(role, id)                                                                                         \
  do {                                                                                                       \
    if (p.trace != nullptr && blockIdx.x == 0 && lane == 0 && gt_n < GT_N) {                                 \
      p.trace[(role) * GT_N + gt_n] = ((unsigned long long)clock64() << 8) | (unsigned long long)(id);      \
      gt_n++;                                                                                                \
    }                                                                                                        \
  } while (0)
#else
#define GT(role, id) do {} while (0)
#endif

template <int ELEM, int CB, int CG, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
               const __grid_constant__ CUtensorMap tmI, const TcGemmParams p) {
  const int in_mode_ = EPI == 0 ? p.in_mode : (EPI == 3 ? 2 : (EPI == 4 ? 1 : 0));
  float* const colsum_ = (EPI == 0 || EPI == 5) ? p.colsum : nullptr;

  constexpr int CW = CB / ELEM;
  constexpr int EB = 32 * CB;                                             // bytes of one staging buffer ({CW cols, 32 rows} box)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;      // 1024-byte alignment for SWIZZLE_128B tiles
  uint8_t* const base_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int stage_bytes = p.a_bytes + p.b_bytes;
  // layout: [operand ring | staging buffers: 8 warps x nb x EB | ones tile 2 KB | barriers]; split-K wgrads whose CTAs own ONE tile each
  // put the staging buffers ON the ring (it is idle by the time the only epilogue starts): the 64 KB go to ring stages instead
  const int ebuf_base = p.ebuf_off;
  const int misc = p.misc_off;
  const uint32_t ones_addr = smem_base + misc;
  const uint32_t bar_base = smem_base + misc + 2048;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (TC_MAX_STAGES + s); };
  auto tmem_full = [&](int a) { return bar_base + 8u * (2 * TC_MAX_STAGES + a); };
  auto tmem_empty = [&](int a) { return bar_base + 8u * (2 * TC_MAX_STAGES + 2 + a); };
  auto in_full = [&](int w, int b) { return bar_base + 8u * (2 * TC_MAX_STAGES + 4 + w * TC_MAX_NB + b); };
  constexpr int NBARS = 2 * TC_MAX_STAGES + 4 + TC_EPI_WARPS * TC_MAX_NB;
  const uint32_t tmem_ptr_addr = bar_base + 8u * NBARS;
  volatile uint32_t* tmem_ptr_gen = (volatile uint32_t*)(base_gen + misc + 2048 + 8 * NBARS);

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef BPM_GEMM_TRACE
  int gt_n = 0;
#endif
  const int num_kb_total = (p.K + TC_BK - 1) / TC_BK;
  const int tiles_mn = p.gx * p.gy;                  // p.gy counts (CG * 128)-row tiles
  const int total_tiles = tiles_mn * p.split;
  const int acc_stride = p.tmem_cols >> 1;
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;       // 0 = leader of the pair
  const int first_tile = (int)blockIdx.x / CG, tile_step = (int)gridDim.x / CG;
  const int m_sub = (int)cta_rank * TC_BM;                          // this CTA's rows inside the tile
  const int bn_cta = p.BN / CG;                                     // B columns held by this CTA
  const int n_sub = (int)cta_rank * bn_cta;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmC);
    if (in_mode_) tma_prefetch_desc(&tmI);
    for (int s = 0; s < p.stages; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; a++) { mbar_init(tmem_full(a), 1); mbar_init(tmem_empty(a), CG * TC_EPI_WARPS); }
    for (int w = 0; w < TC_EPI_WARPS; w++)
      for (int b = 0; b < p.nb; b++) mbar_init(in_full(w, b), 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    if (CG == 2) tmem_alloc_2cta(tmem_ptr_addr, (uint32_t)p.tmem_cols);
    else tmem_alloc(tmem_ptr_addr, (uint32_t)p.tmem_cols);
  }
  if (warp >= 2 && colsum_ != nullptr) {                                   // 16 x 64 tile of bf16 ones (B operand of the colsum MMA)
    uint32_t* o = (uint32_t*)(base_gen + misc);
    for (int c = threadIdx.x - 64; c < 512; c += 32 * TC_EPI_WARPS) o[c] = 0x3F803F80u;
    fence_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();               // the peer's barriers are initialised before anything can arrive on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  pdl_wait();                                   // everything above overlapped the previous kernel's tail; global memory from here on

  if (warp == 0) {
    // ===================== TMA producer (converged warp, one elected lane issues) =====================
    {
      const int b_chunks = (bn_cta + 63) / 64;
      const uint32_t tx = (uint32_t)stage_bytes;
      int s = 0;
      uint32_t ph = 0;
      for (int t = first_tile; t < total_tiles; t += tile_step) {
        const int n0 = (t % p.gx) * p.BN + n_sub, m0 = ((t / p.gx) % p.gy) * (CG * TC_BM) + m_sub, z = t / tiles_mn;
        const int kb0 = z * p.kb_per_split, kb1 = min(num_kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; kb++) {
          if (!(p.dbg & 64)) mbar_wait(empty_bar(s), ph ^ 1u);
          GT(0, 1);
          const uint32_t sa = smem_base + s * stage_bytes, sb = sa + p.a_bytes;
          const uint32_t fb = full_bar(s);
          if (elect_one()) {
            if (p.dbg & 1) {
              mbar_arrive(fb);
            } else if (CG == 1) {
              mbar_expect_tx(fb, tx);
              const int k = kb * TC_BK;
              if (!p.a_mn) tma_load_2d(sa, &tmA, fb, k, m0);                                  // box {64 k, 128 m}
              else { tma_load_2d(sa, &tmA, fb, m0, k); tma_load_2d(sa + 8192, &tmA, fb, m0 + 64, k); }   // box {64 m, 64 k} x2
              if (!p.b_mn) tma_load_2d(sb, &tmB, fb, k, n0);                                  // box {64 k, BN n}
              else
                for (int c = 0; c < b_chunks; c++) tma_load_2d(sb + c * 8192, &tmB, fb, n0 + 64 * c, k);   // box {64 n, 64 k}
            } else {
              // pair: the leader's barrier collects the bytes of BOTH CTAs (the leader alone arrives, expecting 2x; each CTA's loads
              // land in its own shared memory and complete on the leader's barrier)
              if (cta_rank == 0) mbar_expect_tx(fb, 2 * tx);
              const uint32_t fl = mapa_rank(fb, 0);                                           // the leader's full barrier
              const int k = kb * TC_BK;
              if (!p.a_mn) tma_load_2d_2cta(sa, &tmA, fl, k, m0);
              else { tma_load_2d_2cta(sa, &tmA, fl, m0, k); tma_load_2d_2cta(sa + 8192, &tmA, fl, m0 + 64, k); }
              if (!p.b_mn) tma_load_2d_2cta(sb, &tmB, fl, k, n0);                             // box {64 k, BN/2 n}
              else
                for (int c = 0; c < b_chunks; c++) tma_load_2d_2cta(sb + c * 8192, &tmB, fl, n0 + 64 * c, k);
            }
          }
          __syncwarp();
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (converged warp; one elected lane issues each k-block's batch); pair: leader CTA only =====================
    if (CG == 1 || cta_rank == 0) {
      // descriptor templates: K-major SW128 (LBO 16, SBO 1024; +32 B per k16 step) or MN-major (LBO BK*128, SBO 1024; +2048 B per step)
      const uint64_t da_t = p.a_mn ? umma_desc(0, TC_BK * 128, 1024, BPM_SWZ_128B) : umma_desc(0, 16, 1024, BPM_SWZ_128B);
      const uint64_t db_t = p.b_mn ? umma_desc(0, TC_BK * 128, 1024, BPM_SWZ_128B) : umma_desc(0, 16, 1024, BPM_SWZ_128B);
      const uint64_t da_k = p.a_mn ? (2048 >> 4) : (32 >> 4), db_k = p.b_mn ? (2048 >> 4) : (32 >> 4);
      const uint64_t d_ones = umma_desc(ones_addr, 16, 1024, BPM_SWZ_128B);
      const bool no_mma = (p.dbg & 2) != 0;
      int s = 0, tc = 0;
      uint32_t ph = 0;
      // tile state is prepared one tile AHEAD (while the tensor core still has the last k-block of the current tile queued): the
      // divisions and the accumulator hand-shake of a tile boundary would otherwise idle the MMA pipe for several hundred cycles
      int t = first_tile;
      int kb0 = 0, kb1 = 0;
      bool do_colsum = false;
      auto tile_setup = [&](int tt, int tcc) {
        const int z = tt / tiles_mn;
        kb0 = z * p.kb_per_split;
        kb1 = min(num_kb_total, kb0 + p.kb_per_split);
        do_colsum = colsum_ != nullptr && (tt % p.gx) == 0;
        mbar_wait(tmem_empty(tcc & 1), ((uint32_t)(tcc >> 1) & 1u) ^ 1u);      // the epilogue has drained this accumulator
        tc_fence_after();
      };
      if (t < total_tiles) tile_setup(t, 0);
      while (t < total_tiles) {
        const int a = tc & 1;
        const uint32_t acc = tmem_base + (uint32_t)(a * acc_stride);
        const int kb_end = kb1;
        const bool colsum_now = do_colsum;
        uint32_t accum = 0;
        for (int kb = kb0; kb < kb_end; kb++) {
          if (kb == kb_end - 1 && t + tile_step < total_tiles) tile_setup(t + tile_step, tc + 1);
          if (!(p.dbg & 32)) {
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
          }
          GT(1, 2);
          const uint32_t sa = smem_base + s * stage_bytes;
          uint64_t da = da_t | (uint64_t)((sa & 0x3FFFFu) >> 4), db = db_t | (uint64_t)(((sa + p.a_bytes) & 0x3FFFFu) >> 4);
          if (elect_one()) {
            if (!no_mma) {
#pragma unroll
              for (int k = 0; k < TC_BK / 16; k++) {
                if (CG == 2) umma_bf16_2cta(acc, da, db, p.idesc, accum | (uint32_t)k);
                else umma_bf16(acc, da, db, p.idesc, accum | (uint32_t)k);
                // column sums of dY: dY^T (this A tile) times a tile of ones -> 16 identical columns behind the accumulator
                if (colsum_now) {
                  if (CG == 2) umma_bf16_2cta(acc + p.BN, da, d_ones + 2 * k, p.idesc_ones, accum | (uint32_t)k);
                  else umma_bf16(acc + p.BN, da, d_ones + 2 * k, p.idesc_ones, accum | (uint32_t)k);
                }
                da += da_k; db += db_k;
              }
            }
            if (!(p.dbg & 64)) {                                      // frees the smem slot (in both CTAs) once these MMAs have read it
              if (CG == 2) umma_commit_2cta(empty_bar(s));
              else umma_commit(empty_bar(s));
            }
          }
          __syncwarp();
          GT(1, 3);
          accum = 1;
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
        if (elect_one()) {                                        // accumulator complete (both CTAs' epilogues)
          if (CG == 2) umma_commit_2cta(tmem_full(a));
          else umma_commit(tmem_full(a));
        }
        __syncwarp();
        t += tile_step;
        tc++;
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 2;                                               // 0..7
    const int quarter = warp & 3;                                          // TMEM lane quarter this warp may access
    const int half = ew >> 2;                                              // takes chunks half, half + 2, ...
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const DropCtx dc = make_drop(p.drop);
    const bool drop_on = EPI == 0 ? dc.on : (EPI == 2 || EPI == 4);
    const float alpha = drop_on ? p.alpha * dc.inv_keep : p.alpha;
    const int nb = p.nb;
    const int dist = nb - 1;                                               // residual / gate prefetch distance (chunks)
    uint8_t* const ebuf_gen = base_gen + ebuf_base + ew * nb * EB;
    const uint32_t ebuf_s = smem_base + ebuf_base + ew * nb * EB;
    const bool skip_ld = (p.dbg & 16) != 0, skip_math = (p.dbg & 4) != 0, skip_store = (p.dbg & 8) != 0;

    // this warp's chunk sequence over its tiles: (t, c) with c = half, half+2, ... while the chunk starts inside N
    auto chunk_ok = [&](int t, int c) { return c < p.n_chunks && (t % p.gx) * p.BN + c * CW < p.N; };
    auto advance = [&](int& t, int& c) {
      c += 2;
      while (t < total_tiles && !chunk_ok(t, c)) { t += tile_step; c = half; }
    };
    auto issue_in = [&](int t, int c, int idx) {                           // elected lane only
      const int b = idx % nb;
      const int n = (t % p.gx) * p.BN + c * CW, m = ((t / p.gx) % p.gy) * (CG * TC_BM) + m_sub + quarter * 32;
      mbar_expect_tx(in_full(ew, b), (uint32_t)EB);
      tma_load_2d(ebuf_s + b * EB, &tmI, in_full(ew, b), n, m);
    };
    int pt = first_tile, pc = half - 2, issued = 0;                        // prefetch cursor
    if (in_mode_) {
      advance(pt, pc);
      for (int i = 0; i < dist && pt < total_tiles; i++) {
        if (elect_one()) issue_in(pt, pc, issued);
        __syncwarp();
        issued++;
        advance(pt, pc);
      }
    }

    int done = 0, tc = 0;                                                  // chunks processed by this warp; tiles seen
    for (int t = first_tile; t < total_tiles; t += tile_step, tc++) {
      const int n0 = (t % p.gx) * p.BN, m0 = ((t / p.gx) % p.gy) * (CG * TC_BM) + m_sub + quarter * 32;
      const int row = m0 + lane;
      const int a = tc & 1;
      const uint32_t acc = tmem_base + (uint32_t)(a * acc_stride) + lane_off;
      const bool do_colsum = colsum_ != nullptr && n0 == 0 && half == 0;
      int last_c = -1;
      for (int c = half; chunk_ok(t, c); c += 2) last_c = c;
      mbar_wait(tmem_full(a), (uint32_t)(tc >> 1) & 1u);
      tc_fence_after();
      if (ew == 0) GT(2, 4);
      if (last_c < 0) {                                                    // nothing for this warp in this tile
        if (do_colsum) {
          float cs[16];
          tmem_ld16(acc + (uint32_t)p.BN, cs);
          tmem_ld_wait();
          if (row < p.M) atomicAdd(colsum_ + row, cs[0]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (CG == 2 && cta_rank != 0) mbar_arrive_remote(tmem_empty(a), 0); else mbar_arrive(tmem_empty(a)); }
        continue;
      }
      for (int c = half; c <= last_c; c += 2, done++) {
        const int b = done % nb;
        const int col0 = c * CW, n = n0 + col0;
        float v[CW];
        if (!skip_ld) {
          if constexpr (CW == 64) { tmem_ld32(acc + (uint32_t)col0, v); tmem_ld32(acc + (uint32_t)col0 + 32, v + 32); }
          else if constexpr (CW == 32) tmem_ld32(acc + (uint32_t)col0, v);
          else tmem_ld16(acc + (uint32_t)col0, v);
          tmem_ld_wait();
        }
        if (ew == 0) GT(2, 6);
        if (c == last_c) {
          float cs0 = 0.f;
          if (do_colsum) {
            float cs[16];
            tmem_ld16(acc + (uint32_t)p.BN, cs);
            tmem_ld_wait();
            cs0 = cs[0];
          }
          tc_fence_before();                                                // this warp's part of the accumulator is in registers
          __syncwarp();
          if (lane == 0) { if (CG == 2 && cta_rank != 0) mbar_arrive_remote(tmem_empty(a), 0); else mbar_arrive(tmem_empty(a)); }
          if (do_colsum && row < p.M) atomicAdd(colsum_ + row, cs0);
        }
        // ---- bias, alpha, relu, dropout
        if (!skip_math) {
          if (p.bias != nullptr) {
#pragma unroll
            for (int g4 = 0; g4 < CW / 4; g4++) {
              float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
              if (n + g4 * 4 + 4 <= p.N) bb = __ldg((const float4*)(p.bias + n + g4 * 4));
              v[g4 * 4] += bb.x; v[g4 * 4 + 1] += bb.y; v[g4 * 4 + 2] += bb.z; v[g4 * 4 + 3] += bb.w;
            }
          }
          if (alpha != 1.f) {          // (the dropout scale 1/(1-p) is folded into alpha: relu is positively homogeneous)
#pragma unroll
            for (int e = 0; e < CW; e++) v[e] *= alpha;
          }
          if (p.act == 1) {
#pragma unroll
            for (int e = 0; e < CW; e++) v[e] = fmaxf(v[e], 0.f);
          }
          if (drop_on) {
            const uint64_t pair0 = ((uint64_t)row * (uint64_t)p.ldc + (uint64_t)n) >> 1;
            const uint32_t lo0 = (uint32_t)pair0;
            if (lo0 + (uint32_t)(CW / 2) >= lo0) {                           // no carry into the high counter word inside this chunk
              const uint32_t hk = dc.k0 ^ ((uint32_t)(pair0 >> 32) * 0x85ebca77u);
#pragma unroll
              for (int j = 0; j < CW / 2; j++) {
                uint32_t x = (lo0 + (uint32_t)j) * 0x7feb352du + hk;
                x ^= x >> 16;
                x = x * 0x846ca68bu + dc.k1;
                v[2 * j] = drop_keep_lo(dc, x) ? v[2 * j] : 0.f;
                v[2 * j + 1] = drop_keep_hi(dc, x) ? v[2 * j + 1] : 0.f;
              }
            } else {
#pragma unroll
              for (int j = 0; j < CW / 2; j++) {
                const uint32_t x = drop_rand_pair(dc, pair0 + j);
                v[2 * j] = drop_keep_lo(dc, x) ? v[2 * j] : 0.f;
                v[2 * j + 1] = drop_keep_hi(dc, x) ? v[2 * j + 1] : 0.f;
              }
            }
          }
        }
        if (ew == 0) GT(2, 7);
        if (skip_store) continue;
        uint8_t* const st = ebuf_gen + b * EB;
        if (in_mode_) {
          // ---- relu-backward gate or residual from the tile the TMA unit placed in this warp's staging buffer
          mbar_wait(in_full(ew, b), (uint32_t)(done / nb) & 1u);
#pragma unroll
          for (int u = 0; u < CB / 16; u++) {
            const uint4 w = *(const uint4*)(st + swz_off<CB>(lane, u));
            if constexpr (ELEM == 2) {
              const __nv_bfloat162* h2 = (const __nv_bfloat162*)&w;
#pragma unroll
              for (int e = 0; e < 4; e++) {
                const float2 f = __bfloat1622float2(h2[e]);
                if (in_mode_ == 2) {
                  v[u * 8 + 2 * e] = f.x > 0.f ? v[u * 8 + 2 * e] * p.gate_scale : 0.f;
                  v[u * 8 + 2 * e + 1] = f.y > 0.f ? v[u * 8 + 2 * e + 1] * p.gate_scale : 0.f;
                } else {
                  v[u * 8 + 2 * e] += f.x;
                  v[u * 8 + 2 * e + 1] += f.y;
                }
              }
            } else {
              const float* f = (const float*)&w;
#pragma unroll
              for (int e = 0; e < 4; e++) {
                if (in_mode_ == 2) v[u * 4 + e] = f[e] > 0.f ? v[u * 4 + e] * p.gate_scale : 0.f;
                else v[u * 4 + e] += f[e];
              }
            }
          }
          // the buffer of chunk done + dist is the one chunk done - 1 was stored from: once the TMA unit has read it, refill it
          if (pt < total_tiles) {
            if (elect_one()) {
              bulk_wait_read<0>();
              issue_in(pt, pc, issued);
            }
            __syncwarp();
            issued++;
            advance(pt, pc);
          }
        } else {
          if (elect_one()) bulk_wait_read_n(nb - 1);                        // the store that last used this buffer has read it
          __syncwarp();
        }
        if (ew == 0) GT(2, 8);
        // ---- stage the chunk (swizzled, in place over the residual / gate tile) and hand it to the TMA unit
#pragma unroll
        for (int u = 0; u < CB / 16; u++) {
          uint4 w;
          if constexpr (ELEM == 2) {
            __nv_bfloat162* h2 = (__nv_bfloat162*)&w;
#pragma unroll
            for (int e = 0; e < 4; e++) h2[e] = __floats2bfloat162_rn(v[u * 8 + 2 * e], v[u * 8 + 2 * e + 1]);
          } else {
            w = make_uint4(__float_as_uint(v[u * 4]), __float_as_uint(v[u * 4 + 1]), __float_as_uint(v[u * 4 + 2]), __float_as_uint(v[u * 4 + 3]));
          }
          *(uint4*)(st + swz_off<CB>(lane, u)) = w;
        }
        fence_async_smem();
        __syncwarp();
        if (elect_one()) {
          if (p.reduce_add) tma_reduce_add_2d(&tmC, ebuf_s + b * EB, n, m0);
          else tma_store_2d(&tmC, ebuf_s + b * EB, n, m0);
          bulk_commit();
        }
        __syncwarp();
        if (ew == 0) GT(2, 9);
      }
    }
    if (elect_one()) bulk_wait_read<0>();                                   // smem must outlive the last TMA store's read
    __syncwarp();
    if (ew == 0) GT(2, 5);
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();               // no CTA of the pair may exit (or free TMEM) while the other can still signal it
  if (warp == 1) {
    if (CG == 2) tmem_dealloc_2cta(tmem_base, (uint32_t)p.tmem_cols);
    else tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ---------------------------------------------------------------- host
bpm_encode_tiled_fn bpm_get_encode_tiled() {
  static bpm_encode_tiled_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      (void)cudaGetLastError();
      return nullptr;
    }
    fn = (bpm_encode_tiled_fn)p;
  }
  return fn;
}

static int make_tmap(CUtensorMap* out, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, CUtensorMapSwizzle swz) {
  bpm_encode_tiled_fn enc = bpm_get_encode_tiled();
  if (!enc) { bpm_set_error("cuTensorMapEncodeTiled entry point unavailable"); return BPM_ELAUNCH; }
  cuuint64_t gd[5]; cuuint64_t gs[5]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; i++) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; i++) gs[i] = strides_bytes[i];
  CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    bpm_set_error("cuTensorMapEncodeTiled failed (%d): base %p rank %d dims %llu,%llu stride %llu box %u,%u", (int)r, base, rank,
                  (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)strides_bytes[0], box[0], box[1]);
    return BPM_EINVAL;
  }
  return BPM_OK;
}

int bpm_make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                       CUtensorMapSwizzle swz) {
  return make_tmap(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, swz);
}

int bpm_make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                      CUtensorMapSwizzle swz) {
  return make_tmap(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box, swz);
}

// [rows, cols] row-major map for the epilogue (C / residual / gate): box {chunk columns, 32 rows}
static int make_epi_map(CUtensorMap* out, const void* base, int rows, int cols, int ld, int elem, int chunk_bytes) {
  uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows}, str[1] = {(uint64_t)ld * elem};
  uint32_t box[2] = {(uint32_t)(chunk_bytes / elem), 32};
  return make_tmap(out, elem == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, 2, dims, str, box,
                   chunk_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
}

static int pick_bn(int N, int max_bn) {
  // fewest tiles first, then least padding; BN multiple of 32 in [32, max_bn]
  int tiles = bpm_cdiv(N, max_bn);
  int bn = bpm_cdiv(bpm_cdiv(N, tiles), 32) * 32;
  return bn < 32 ? 32 : bn;
}

template <int ELEM, int CB, int CG, int EPI>
static int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmI, const TcGemmParams& p, int ctas,
                     size_t smem, cudaStream_t stream) {
  if (int rc = bpm_func_smem((const void*)gemm_tc_kernel<ELEM, CB, CG, EPI>, 227 * 1024, "gemm_tc")) return rc;
  cudaError_t le = bpm_launch_cluster(CG, gemm_tc_kernel<ELEM, CB, CG, EPI>, dim3(ctas), dim3(TC_THREADS), smem, stream, tmA, tmB, tmC, tmI, p);
  if (le != cudaSuccess) { bpm_set_error("gemm_tc: launch failed: %s", cudaGetErrorString(le)); (void)cudaGetLastError(); return BPM_ELAUNCH; }
  return BPM_OK;
}

int bpm_gemm_tc(const bpm_gemm_t* g, cudaStream_t stream) {
  BPM_REQUIRE(((uintptr_t)g->A % 16 == 0) && ((uintptr_t)g->B % 16 == 0) && g->lda % 8 == 0 && g->ldb % 8 == 0,
              "gemm(bf16): A/B must be 16-byte aligned with pitches multiple of 8 elements (lda %d ldb %d)", g->lda, g->ldb);
  BPM_REQUIRE(!g->accumulate || g->c_dtype == BPM_F32, "gemm: accumulate needs fp32 C");
  const int elem = g->c_dtype == BPM_BF16 ? 2 : 4;
  BPM_REQUIRE(((uintptr_t)g->C % 16 == 0) && ((int64_t)g->ldc * elem) % 16 == 0, "gemm(bf16): C must be 16-byte aligned with a 16-byte multiple pitch");
  BPM_REQUIRE(!g->residual || (g->res_dtype == g->c_dtype && ((uintptr_t)g->residual % 16 == 0) && ((int64_t)g->ldr * elem) % 16 == 0),
              "gemm(bf16): residual must have C's dtype and 16-byte alignment");
  BPM_REQUIRE(!g->gate || (g->gate_dtype == g->c_dtype && ((uintptr_t)g->gate % 16 == 0) && ((int64_t)g->ldg * elem) % 16 == 0),
              "gemm(bf16): gate must have C's dtype and 16-byte alignment");
  BPM_REQUIRE(!(g->gate && g->residual), "gemm(bf16): gate and residual cannot be combined in one epilogue");
  BPM_REQUIRE(!g->colsum_out || g->ta == 1, "gemm: the fused column sum is defined for ta = 1 (wgrad) only");
  BPM_REQUIRE(!g->bias || (g->N % 4 == 0 && (uintptr_t)g->bias % 16 == 0), "gemm(bf16): bias needs N %% 4 == 0 and 16-byte alignment");
  BPM_REQUIRE(g->drop.p <= 0.f || g->ldc % 8 == 0, "gemm(bf16): dropout needs ldc %% 8 == 0");
  const bool plain = !g->bias && g->act == 0 && g->drop.p == 0.f && !g->gate && !g->residual;
  BPM_REQUIRE(!g->accumulate || plain, "gemm(bf16): accumulate supports the plain (alpha-only) epilogue");
  TcGemmParams p;
  p.M = g->M; p.N = g->N; p.K = g->K;
  p.colsum = g->colsum_out;
  p.dbg = bpm_debug_get(0);
  p.trace = (unsigned long long*)bpm_debug_get_ptr();
  int max_bn = p.colsum ? 224 : 256;               // 2 x (224 + 16) TMEM columns still fit in 512
  if (bpm_debug_get(2) >= 32) max_bn = min(max_bn, bpm_debug_get(2));
  p.BN = pick_bn(g->N, max_bn);
  // CTA pairs (256-row tiles) when the row count does not waste much more than 128-row tiles would (bpm_debug slot 0 bit 128: off)
  const int waste1 = bpm_cdiv(g->M, TC_BM) * TC_BM, waste2 = bpm_cdiv(g->M, 2 * TC_BM) * 2 * TC_BM;
  // ... and the main loop is long enough to matter: with K <= 384 the kernel is bound by its epilogue and pairing only couples the two
  // CTAs' epilogues (measured: fc1 45 -> 50 us paired, fc2 49 -> 44 us, dgrad K = 1216 41 -> 34 us)
  const int cg = (!(p.dbg & 128) && g->M > TC_BM && waste2 * 20 <= waste1 * 21 && bpm_num_sms() % 2 == 0 && (g->K >= 768 || (p.dbg & 256))) ? 2 : 1;
  const int bn_cta = p.BN / cg;                     // B columns per CTA (BN is a multiple of 32)
  p.a_mn = g->ta ? 1 : 0;
  p.b_mn = g->tb ? 1 : 0;
  p.a_bytes = TC_BM * TC_BK * 2;
  p.b_bytes = p.b_mn ? bpm_cdiv(bn_cta, 64) * 8192 : bpm_cdiv(bn_cta * 128, 1024) * 1024;
  const int stage_bytes = p.a_bytes + p.b_bytes;
  p.in_mode = g->residual ? 1 : (g->gate ? 2 : 0);
  p.reduce_add = g->accumulate ? 1 : 0;
  const int cb = (p.BN * elem) % 128 == 0 ? 128 : 64;
  p.n_chunks = p.BN * elem / cb;
  const int eb = 32 * cb;
  const int fixed = 2048 + 8 * (2 * TC_MAX_STAGES + 4 + TC_EPI_WARPS * TC_MAX_NB) + 16 + 1024;      // ones tile, barriers, alignment slack
  const int budget = 227 * 1024 - fixed;
  // staging buffers per epilogue warp: with a residual / gate tile prefetched through them 3 (two chunks ahead) when the operand
  // ring still gets 3 stages AND the main loop is short (K <= 512: epilogue-bound, out-proj 28.1 -> 27.3 us); with a long K the
  // shared memory is worth more as a fourth ring stage (fc2, K = 1216: 42.3 -> 38.9 us with 2), else 2
  p.nb = 2;
  if (p.in_mode && g->K <= 512 && (budget - TC_EPI_WARPS * 3 * eb) / stage_bytes >= 3) p.nb = 3;
  if (bpm_debug_get(3) >= 2 && bpm_debug_get(3) <= TC_MAX_NB) p.nb = bpm_debug_get(3);
  // two accumulators (one per in-flight tile); each BN (+16 for the fused column sum) columns wide
  p.tmem_cols = 64;
  while (p.tmem_cols < 2 * (p.BN + (p.colsum ? 16 : 0))) p.tmem_cols *= 2;
  p.idesc = umma_idesc_bf16(cg * TC_BM, p.BN, p.a_mn, p.b_mn);
  p.idesc_ones = umma_idesc_bf16(cg * TC_BM, 16, p.a_mn, 0);
  p.bias = g->bias; p.alpha = g->alpha; p.act = g->act; p.gate_scale = g->gate_scale; p.ldc = g->ldc; p.drop = g->drop;
  int num_kb = bpm_cdiv(g->K, TC_BK);
  int gx = bpm_cdiv(g->N, p.BN), gy = bpm_cdiv(g->M, cg * TC_BM);
  int split = 1;
  if (g->accumulate) {
    if (g->split_k > 0) {
      split = g->split_k;
    } else {
      // split K so that the persistent CTAs finish together: minimise (rounds of tiles per SM) x (k-blocks per tile + a few
      // k-block times of per-tile epilogue), instead of leaving a half-empty last round
      const int sms = bpm_num_sms() / cg, tiles_mn = gx * gy;
      int best_cost = 1 << 30;
      for (int s_ = 1; s_ <= max(1, min(num_kb / 4, (4 * sms) / tiles_mn)); s_++) {
        const int kbps = bpm_cdiv(num_kb, s_), real = bpm_cdiv(num_kb, kbps);
        const int rounds = bpm_cdiv(tiles_mn * real, sms);
        const int cost = rounds * (kbps + 3);
        if (cost < best_cost) { best_cost = cost; split = real; }
      }
    }
  }
  p.kb_per_split = bpm_cdiv(num_kb, split);
  split = bpm_cdiv(num_kb, p.kb_per_split);
  p.gx = gx; p.gy = gy; p.split = split;
  // A DRAM-streaming main loop runs at (bytes in flight) / (loaded DRAM latency, ~3000 clk) per SM (measured: 5 stages x 32 KB give 38 B/clk):
  // when no CTA sees a second tile the epilogue's staging buffers can live on the drained ring, and its 64 KB become ring stages.
  const bool alias = g->accumulate && gx * gy * split <= bpm_num_sms() / cg && !(p.dbg & 32768);
  const int stage_area = TC_EPI_WARPS * p.nb * eb;
  p.stages = max(2, min(TC_MAX_STAGES, (budget - (alias ? 0 : stage_area)) / stage_bytes));
  p.ring_bytes = p.stages * stage_bytes;
  p.ebuf_off = alias ? 0 : p.ring_bytes;
  p.misc_off = alias ? max(p.ring_bytes, stage_area) : p.ring_bytes + stage_area;

  CUtensorMap tmA, tmB, tmC, tmI;
  {
    // A: ta == 0 -> stored [M, K]: dims {K, M}, box {64, 128}.  ta == 1 -> stored [K, M]: dims {M, K}, box {64, 64}
    uint64_t dims[2], str[1]; uint32_t box[2];
    if (!p.a_mn) { dims[0] = g->K; dims[1] = g->M; box[0] = TC_BK; box[1] = TC_BM; }
    else { dims[0] = g->M; dims[1] = g->K; box[0] = 64; box[1] = TC_BK; }
    str[0] = (uint64_t)g->lda * 2;
    int rc = bpm_make_tmap_bf16(&tmA, g->A, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    // B: tb == 0 -> stored [N, K]: dims {K, N}, box {64, BN}.  tb == 1 -> stored [K, N]: dims {N, K}, box {64, 64}
    if (!p.b_mn) { dims[0] = g->K; dims[1] = g->N; box[0] = TC_BK; box[1] = bn_cta; }
    else { dims[0] = g->N; dims[1] = g->K; box[0] = 64; box[1] = TC_BK; }
    str[0] = (uint64_t)g->ldb * 2;
    rc = bpm_make_tmap_bf16(&tmB, g->B, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    if ((rc = make_epi_map(&tmC, g->C, g->M, g->N, g->ldc, elem, cb))) return rc;
    tmI = tmC;
    if (g->residual && (rc = make_epi_map(&tmI, g->residual, g->M, g->N, g->ldr, elem, cb))) return rc;
    if (g->gate && (rc = make_epi_map(&tmI, g->gate, g->M, g->N, g->ldg, elem, cb))) return rc;
  }
  size_t smem = (size_t)p.misc_off + fixed;
  BPM_REQUIRE(smem <= 227 * 1024 && p.tmem_cols <= 512, "gemm_tc: smem %zu / tmem %d too large", smem, p.tmem_cols);
  int total_tiles = gx * gy * split;
  int ctas = cg * min(total_tiles, bpm_num_sms() / cg);
  const bool has_drop = g->drop.p > 0.f;
  int epi = 0;
  if (!(bpm_debug_get(0) & 512)) {
    if (elem == 2) {
      if (!has_drop && p.in_mode == 0 && p.colsum == nullptr) epi = 1;
      else if (has_drop && p.in_mode == 0 && p.colsum == nullptr) epi = 2;
      else if (!has_drop && p.in_mode == 2 && p.colsum == nullptr) epi = 3;
    } else {
      if (has_drop && p.in_mode == 1 && p.colsum == nullptr) epi = 4;
      else if (!has_drop && p.in_mode == 0) epi = 5;
    }
  }
#define TC_GO(E, C, G, X) return launch_tc<E, C, G, X>(tmA, tmB, tmC, tmI, p, ctas, smem, stream)
#define TC_BF16(C, G) do { if (epi == 1) TC_GO(2, C, G, 1); if (epi == 2) TC_GO(2, C, G, 2); if (epi == 3) TC_GO(2, C, G, 3); TC_GO(2, C, G, 0); } while (0)
#define TC_F32(G) do { if (epi == 4) TC_GO(4, 128, G, 4); if (epi == 5) TC_GO(4, 128, G, 5); TC_GO(4, 128, G, 0); } while (0)
  if (cg == 2) {
    if (elem == 4) TC_F32(2);
    if (cb == 128) TC_BF16(128, 2);
    TC_BF16(64, 2);
  }
  if (elem == 4) TC_F32(1);
  if (cb == 128) TC_BF16(128, 1);
  TC_BF16(64, 1);
#undef TC_GO
#undef TC_BF16
#undef TC_F32

}
