// bf16 GEMM on 5th-gen tensor cores (sm_100a): TMA (SWIZZLE_128B) -> shared memory ring -> tcgen05.mma (cta_group::1,
// M = 128, N = BN <= 256, K = 16 per instruction) -> fp32 accumulator in TMEM -> tcgen05.ld -> fused epilogue -> TMA store.
//
//  * Warp roles (192 threads): warp 0 = TMA producer (operands, then the epilogue's residual / gate tiles), warp 1 = TMEM
//    allocator + single-thread MMA issuer, warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4, one thread per row).
//  * All four operand layouts without any transposed copies: K-major operands use the canonical SW128 K-major layout
//    (SBO = 1024 B); "transposed" operands (dgrad's W, wgrad's dY^T and X) are loaded as 64-wide MN chunks and described to
//    the MMA as MN-major (LBO = BK*128 B between chunks, SBO = 1024 B between 8-row k groups).
//  * Epilogue: K is only 300..1216 here (5..19 k-blocks per tile), so the kernel lives or dies by its epilogue.  The
//    accumulator is drained in 128-byte-wide column chunks; residual / relu-gate tiles arrive by TMA into shared memory
//    (prefetched one chunk ahead), the bias sits in shared memory, results are staged in shared memory (swizzled, conflict
//    free) and leave with ONE TMA store per chunk (or one TMA reduce-add for split-K wgrad) -- every global access is a
//    full 128-byte line and M / N tails are clipped by the TMA unit.  The staging buffers alias the drained operand ring.
//  * Two CTAs are co-resident per SM (<= 111 KB smem, <= 256 TMEM columns each): one CTA's epilogue overlaps the other's
//    main loop.
//  * wgrad (K = B*T rows) is split along K across blockIdx.z.  The bias gradient (column sums of dY) is fused into wgrad as
//    one extra N=16 MMA per k-step against a tile of ones (dY^T * 1), so dY is never re-read for it.
#include "gemm_epilogue.cuh"
#include "tc_common.cuh"

#define TC_BM 128
#define TC_BK 64
#define TC_THREADS 192
#define TC_SLOT 16384          // one 128-row x 128-byte epilogue tile

struct TcGemmParams {
  int M, N, K;
  int BN;                 // multiple of 32, <= 256
  int a_mn, b_mn;         // 1: operand is MN-major in memory ("transposed")
  int stages;
  int a_bytes, b_bytes;   // per-stage bytes (multiples of 1024)
  int ring_bytes;         // operand ring (>= the epilogue's slot needs)
  int kb_per_split;
  int gx, gy, split;
  int tmem_cols;
  uint32_t idesc, idesc_ones;
  int elem;               // output / residual / gate element size (2 or 4)
  int chunk_bytes;        // 128 or 64: bytes per row of one epilogue chunk
  int n_chunks;
  int has_res, has_gate, reduce_add;
  float* colsum;          // fused bias gradient (wgrad only), or null
  EpiParams ep;
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
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// byte offset of 16-byte unit u of row r inside a swizzled [128 rows x chunk_bytes] tile (TMA SWIZZLE_128B / SWIZZLE_64B)
__device__ __forceinline__ uint32_t swz_off(int r, int u, int chunk_bytes) {
  return chunk_bytes == 128 ? (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((u ^ (r & 7)) << 4))
                            : (uint32_t)((r >> 3) * 512 + (r & 7) * 64 + ((u ^ ((r >> 1) & 3)) << 4));
}

// Persistent kernel: one CTA per SM walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...  (n-tile fastest, so the CTAs that
// run together share A tiles through L2).  Two TMEM accumulators alternate between tiles: the MMA warp starts tile t+1 while the
// epilogue warps drain tile t, and the TMA producer simply keeps the operand ring full across tile boundaries.
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
               const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmG, const TcGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;      // 1024-byte alignment for SWIZZLE_128B tiles
  uint8_t* const base_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int stage_bytes = p.a_bytes + p.b_bytes;
  // layout: [operand ring | epilogue slots (out x2, res x2, gate x2) | ones tile 2 KB | bias 1 KB | barriers]
  const int slot_base = p.ring_bytes;
  const int n_slots = 2 + 2 * p.has_res + 2 * p.has_gate;
  const int misc = slot_base + n_slots * TC_SLOT;
  const uint32_t ones_addr = smem_base + misc;
  float* const bias_s = (float*)(base_gen + misc + 2048);
  const uint32_t bar_base = smem_base + misc + 2048 + 1024;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (4 + s); };
  auto tmem_full = [&](int a) { return bar_base + 8u * (8 + a); };
  auto tmem_empty = [&](int a) { return bar_base + 8u * (10 + a); };
  auto in_full = [&](int s) { return bar_base + 8u * (12 + s); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * 14;
  volatile uint32_t* tmem_ptr_gen = (volatile uint32_t*)(base_gen + misc + 2048 + 1024 + 8 * 14);
  const int out_slot = slot_base, res_slot = slot_base + 2 * TC_SLOT, gate_slot = slot_base + (2 + 2 * p.has_res) * TC_SLOT;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb_total = (p.K + TC_BK - 1) / TC_BK;
  const int tiles_mn = p.gx * p.gy;
  const int total_tiles = tiles_mn * p.split;
  const int acc_stride = p.tmem_cols >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmC);
    for (int s = 0; s < p.stages; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; a++) { mbar_init(tmem_full(a), 1); mbar_init(tmem_empty(a), 4); mbar_init(in_full(a), 1); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_addr, (uint32_t)p.tmem_cols);
  if (warp >= 2 && p.colsum != nullptr) {                                   // 16 x 64 tile of bf16 ones (B operand of the colsum MMA)
    uint32_t* o = (uint32_t*)(base_gen + misc);
    for (int c = threadIdx.x - 64; c < 512; c += 128) o[c] = 0x3F803F80u;
    fence_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int b_chunks = (p.BN + 63) / 64;
      int it = 0;                                                          // running k-block counter (ring position)
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int n0 = (t % p.gx) * p.BN, m0 = ((t / p.gx) % p.gy) * TC_BM, z = t / tiles_mn;
        const int kb0 = z * p.kb_per_split, kb1 = min(num_kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; kb++, it++) {
          int s = it % p.stages;
          uint32_t ph = (uint32_t)(it / p.stages) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
          uint32_t sa = smem_base + s * stage_bytes, sb = sa + p.a_bytes;
          mbar_expect_tx(full_bar(s), (uint32_t)(p.a_bytes + p.b_bytes));
          int k = kb * TC_BK;
          if (!p.a_mn) tma_load_2d(sa, &tmA, full_bar(s), k, m0);                        // box {64 k, 128 m}
          else { tma_load_2d(sa, &tmA, full_bar(s), m0, k); tma_load_2d(sa + 8192, &tmA, full_bar(s), m0 + 64, k); }   // box {64 m, 64 k} x2
          if (!p.b_mn) tma_load_2d(sb, &tmB, full_bar(s), k, n0);                        // box {64 k, BN n}
          else
            for (int c = 0; c < b_chunks; c++) tma_load_2d(sb + c * 8192, &tmB, full_bar(s), n0 + 64 * c, k);        // box {64 n, 64 k}
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      int it = 0, tc = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, tc++) {
        const int z = t / tiles_mn;
        const int kb0 = z * p.kb_per_split, kb1 = min(num_kb_total, kb0 + p.kb_per_split);
        const bool do_colsum = p.colsum != nullptr && (t % p.gx) == 0;
        const int a = tc & 1;
        const uint32_t acc = tmem_base + (uint32_t)(a * acc_stride);
        mbar_wait(tmem_empty(a), ((uint32_t)(tc >> 1) & 1u) ^ 1u);           // the epilogue has drained this accumulator
        tc_fence_after();
        for (int kb = kb0; kb < kb1; kb++, it++) {
          int s = it % p.stages;
          uint32_t ph = (uint32_t)(it / p.stages) & 1u;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          uint32_t sa = smem_base + s * stage_bytes, sb = sa + p.a_bytes;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; k++) {
            // K-major: +32 B per 16-element k step inside the 128 B swizzle row; MN-major: +16 rows * 128 B
            uint64_t da = p.a_mn ? umma_desc(sa + k * 2048, TC_BK * 128, 1024, BPM_SWZ_128B) : umma_desc(sa + k * 32, 16, 1024, BPM_SWZ_128B);
            uint64_t db = p.b_mn ? umma_desc(sb + k * 2048, TC_BK * 128, 1024, BPM_SWZ_128B) : umma_desc(sb + k * 32, 16, 1024, BPM_SWZ_128B);
            const uint32_t accum = (kb > kb0 || k > 0) ? 1u : 0u;
            umma_bf16(acc, da, db, p.idesc, accum);
            if (do_colsum)    // column sums of dY: dY^T (this A tile) times a tile of ones -> 16 identical columns behind the accumulator
              umma_bf16(acc + p.BN, da, umma_desc(ones_addr + k * 32, 16, 1024, BPM_SWZ_128B), p.idesc_ones, accum);
          }
          umma_commit(empty_bar(s));            // frees the smem slot once these MMAs have read it
        }
        umma_commit(tmem_full(a));              // accumulator complete
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int et = threadIdx.x - 64;                                       // 0..127
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    DropCtx dc = make_drop(p.ep.drop);
    const int cw = p.chunk_bytes / p.elem;                                  // columns per chunk: 16 / 32 / 64
    const int upr = p.chunk_bytes / 16;                                     // 16-byte units per row
    const bool has_in = p.has_res || p.has_gate;
    const uint32_t in_bytes = 128u * (uint32_t)p.chunk_bytes * (uint32_t)(p.has_res + p.has_gate);
    // residual / gate tiles are fetched by TMA one chunk AHEAD (running chunk counter cc, 2 slots), also across tile boundaries
    auto issue_in = [&](int t, int c, int cc) {
      const int s = cc & 1;
      const int n0 = (t % p.gx) * p.BN, m0 = ((t / p.gx) % p.gy) * TC_BM;
      mbar_expect_tx(in_full(s), in_bytes);
      if (p.has_res) tma_load_2d(smem_base + res_slot + s * TC_SLOT, &tmR, in_full(s), n0 + c * cw, m0);
      if (p.has_gate) tma_load_2d(smem_base + gate_slot + s * TC_SLOT, &tmG, in_full(s), n0 + c * cw, m0);
    };
    int cc = 0, tc = 0;
    if (has_in && et == 0 && (int)blockIdx.x < total_tiles) issue_in(blockIdx.x, 0, 0);
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, tc++) {
      const int n0 = (t % p.gx) * p.BN, m0 = ((t / p.gx) % p.gy) * TC_BM;
      const int row = m0 + r;
      const int a = tc & 1;
      const uint32_t acc = tmem_base + (uint32_t)(a * acc_stride) + lane_off;
      const bool do_colsum = p.colsum != nullptr && (t % p.gx) == 0;
      epi_bar();                                                            // previous tile's readers of bias_s are done
      for (int c = et; c < p.BN; c += 128) bias_s[c] = (p.ep.bias && n0 + c < p.N) ? p.ep.bias[n0 + c] : 0.f;
      epi_bar();
      mbar_wait(tmem_full(a), (uint32_t)(tc >> 1) & 1u);
      tc_fence_after();
      for (int c = 0; c < p.n_chunks; c++, cc++) {
        const int s = cc & 1;
        const int col0 = c * cw;
        float v[64];
        {
          if (cw >= 32) tmem_ld32(acc + (uint32_t)col0, v); else tmem_ld16(acc + (uint32_t)col0, v);
          if (cw == 64) tmem_ld32(acc + (uint32_t)col0 + 32, v + 32);
          tmem_ld_wait();
        }
        if (c == p.n_chunks - 1) {
          float cs[16];
          if (do_colsum) {
            tmem_ld16(acc + (uint32_t)p.BN, cs);
            tmem_ld_wait();
          }
          tc_fence_before();                                                // this tile's accumulator is now entirely in registers
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty(a));
          if (do_colsum && row < p.M) atomicAdd(p.colsum + row, cs[0]);
        }
        // ---- bias, alpha, relu, dropout
#pragma unroll
        for (int g8 = 0; g8 < 8; g8++) {
          if (g8 * 8 < cw) {
            float mlt[8];
            const int n = n0 + col0 + g8 * 8;
            drop_mult8(dc, (uint64_t)row * (uint64_t)p.ep.ldc + (uint64_t)n, mlt);
            const float4 b0 = *(const float4*)(bias_s + col0 + g8 * 8), b1 = *(const float4*)(bias_s + col0 + g8 * 8 + 4);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int e = 0; e < 8; e++) {
              float x = (v[g8 * 8 + e] + bb[e]) * p.ep.alpha;
              if (p.ep.act == 1) x = fmaxf(x, 0.f);
              v[g8 * 8 + e] = x * mlt[e];
            }
          }
        }
        // ---- relu-backward gate and residual from the TMA-staged tiles
        if (has_in) {
          mbar_wait(in_full(s), (uint32_t)(cc >> 1) & 1u);
          const uint8_t* rs = base_gen + res_slot + s * TC_SLOT;
          const uint8_t* gs = base_gen + gate_slot + s * TC_SLOT;
#pragma unroll
          for (int u = 0; u < 8; u++) {
            if (u < upr) {
              const uint32_t off = swz_off(r, u, p.chunk_bytes);
              if (p.has_gate) {
                const uint4 w = *(const uint4*)(gs + off);
                if (p.elem == 2) {
                  const __nv_bfloat162* h2 = (const __nv_bfloat162*)&w;
#pragma unroll
                  for (int e = 0; e < 4; e++) {
                    const float2 f = __bfloat1622float2(h2[e]);
                    v[u * 8 + 2 * e] = f.x > 0.f ? v[u * 8 + 2 * e] * p.ep.gate_scale : 0.f;
                    v[u * 8 + 2 * e + 1] = f.y > 0.f ? v[u * 8 + 2 * e + 1] * p.ep.gate_scale : 0.f;
                  }
                } else {
                  const float* f = (const float*)&w;
#pragma unroll
                  for (int e = 0; e < 4; e++) v[u * 4 + e] = f[e] > 0.f ? v[u * 4 + e] * p.ep.gate_scale : 0.f;
                }
              }
              if (p.has_res) {
                const uint4 w = *(const uint4*)(rs + off);
                if (p.elem == 2) {
                  const __nv_bfloat162* h2 = (const __nv_bfloat162*)&w;
#pragma unroll
                  for (int e = 0; e < 4; e++) { const float2 f = __bfloat1622float2(h2[e]); v[u * 8 + 2 * e] += f.x; v[u * 8 + 2 * e + 1] += f.y; }
                } else {
                  const float* f = (const float*)&w;
#pragma unroll
                  for (int e = 0; e < 4; e++) v[u * 4 + e] += f[e];
                }
              }
            }
          }
        }
        // ---- stage the chunk (swizzled) and hand it to the TMA unit
        if (et == 0) bulk_wait_read<1>();                                   // the store that last used this staging slot has read it
        epi_bar();                                                          // (also: every thread has finished reading in-slot s^1's predecessor)
        if (has_in && et == 0) {                                            // prefetch the next chunk's residual / gate tile
          int nt = t, nc = c + 1;
          if (nc == p.n_chunks) { nt = t + gridDim.x; nc = 0; }
          if (nt < total_tiles) issue_in(nt, nc, cc + 1);
        }
        uint8_t* st = base_gen + out_slot + s * TC_SLOT;
#pragma unroll
        for (int u = 0; u < 8; u++) {
          if (u < upr) {
            uint4 w;
            if (p.elem == 2) {
              __nv_bfloat162* h2 = (__nv_bfloat162*)&w;
#pragma unroll
              for (int e = 0; e < 4; e++) h2[e] = __floats2bfloat162_rn(v[u * 8 + 2 * e], v[u * 8 + 2 * e + 1]);
            } else {
              w = make_uint4(__float_as_uint(v[u * 4]), __float_as_uint(v[u * 4 + 1]), __float_as_uint(v[u * 4 + 2]), __float_as_uint(v[u * 4 + 3]));
            }
            *(uint4*)(st + swz_off(r, u, p.chunk_bytes)) = w;
          }
        }
        fence_async_smem();
        epi_bar();
        if (et == 0) {
          if (p.reduce_add) tma_reduce_add_2d(&tmC, smem_base + out_slot + s * TC_SLOT, n0 + col0, m0);
          else tma_store_2d(&tmC, smem_base + out_slot + s * TC_SLOT, n0 + col0, m0);
          bulk_commit();
        }
      }
    }
    if (et == 0) bulk_wait_read<0>();                                       // smem must outlive the last TMA store's read
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
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

static int make_tmap(CUtensorMap* out, CUtensorMapDataType dt, int elem, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz) {
  bpm_encode_tiled_fn enc = bpm_get_encode_tiled();
  if (!enc) { bpm_set_error("cuTensorMapEncodeTiled entry point unavailable"); return BPM_ELAUNCH; }
  cuuint64_t gd[5]; cuuint64_t gs[5]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; i++) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; i++) gs[i] = strides_bytes[i];
  (void)elem;
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
  return make_tmap(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, rank, dims, strides_bytes, box, swz);
}

// [rows, cols] row-major tile map for the epilogue (C / residual / gate): box {chunk columns, 128 rows}
static int make_epi_map(CUtensorMap* out, const void* base, int rows, int cols, int ld, int elem, int chunk_bytes) {
  uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows}, str[1] = {(uint64_t)ld * elem};
  uint32_t box[2] = {(uint32_t)(chunk_bytes / elem), TC_BM};
  return make_tmap(out, elem == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, elem, base, 2, dims, str, box,
                   chunk_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
}

static int pick_bn(int N, int max_bn) {
  // fewest tiles first, then least padding; BN multiple of 32 in [32, max_bn]
  int tiles = bpm_cdiv(N, max_bn);
  int bn = bpm_cdiv(bpm_cdiv(N, tiles), 32) * 32;
  return bn < 32 ? 32 : bn;
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
  BPM_REQUIRE(!g->colsum_out || g->ta == 1, "gemm: the fused column sum is defined for ta = 1 (wgrad) only");
  const bool plain = !g->bias && g->act == 0 && g->drop.p == 0.f && !g->gate && !g->residual;
  BPM_REQUIRE(!g->accumulate || plain, "gemm(bf16): accumulate supports the plain (alpha-only) epilogue");
  TcGemmParams p;
  p.M = g->M; p.N = g->N; p.K = g->K;
  p.colsum = g->colsum_out;
  p.BN = pick_bn(g->N, p.colsum ? 224 : 256);      // 2 x (224 + 16) TMEM columns still fit in 512
  p.a_mn = g->ta ? 1 : 0;
  p.b_mn = g->tb ? 1 : 0;
  p.a_bytes = TC_BM * TC_BK * 2;
  p.b_bytes = p.b_mn ? bpm_cdiv(p.BN, 64) * 8192 : bpm_cdiv(p.BN * 128, 1024) * 1024;
  int stage_bytes = p.a_bytes + p.b_bytes;
  p.has_res = g->residual ? 1 : 0;
  p.has_gate = g->gate ? 1 : 0;
  p.reduce_add = g->accumulate ? 1 : 0;
  const int slots = 2 + 2 * p.has_res + 2 * p.has_gate;
  const int fixed = slots * TC_SLOT + 2048 + 1024 + 8 * 16 + 1024;         // epilogue slots, ones tile, bias, barriers, alignment slack
  p.stages = max(2, min(4, (227 * 1024 - fixed) / stage_bytes));
  p.ring_bytes = p.stages * stage_bytes;
  // two accumulators (one per in-flight tile); each BN (+16 for the fused column sum) columns wide
  p.tmem_cols = 64;
  while (p.tmem_cols < 2 * (p.BN + (p.colsum ? 16 : 0))) p.tmem_cols *= 2;
  p.idesc = umma_idesc_bf16(TC_BM, p.BN, p.a_mn, p.b_mn);
  p.idesc_ones = umma_idesc_bf16(TC_BM, 16, p.a_mn, 0);
  p.elem = elem;
  p.chunk_bytes = (p.BN * elem) % 128 == 0 ? 128 : 64;
  p.n_chunks = p.BN * elem / p.chunk_bytes;
  p.ep = make_epi(g);
  int num_kb = bpm_cdiv(g->K, TC_BK);
  int gx = bpm_cdiv(g->N, p.BN), gy = bpm_cdiv(g->M, TC_BM);
  int split = 1;
  if (g->accumulate) split = g->split_k > 0 ? g->split_k : max(1, min(num_kb / 4, (2 * bpm_num_sms()) / max(1, gx * gy)));
  p.kb_per_split = bpm_cdiv(num_kb, split);
  split = bpm_cdiv(num_kb, p.kb_per_split);
  p.gx = gx; p.gy = gy; p.split = split;

  CUtensorMap tmA, tmB, tmC, tmR, tmG;
  {
    // A: ta == 0 -> stored [M, K]: dims {K, M}, box {64, 128}.  ta == 1 -> stored [K, M]: dims {M, K}, box {64, 64}
    uint64_t dims[2], str[1]; uint32_t box[2];
    if (!p.a_mn) { dims[0] = g->K; dims[1] = g->M; box[0] = TC_BK; box[1] = TC_BM; }
    else { dims[0] = g->M; dims[1] = g->K; box[0] = 64; box[1] = TC_BK; }
    str[0] = (uint64_t)g->lda * 2;
    int rc = bpm_make_tmap_bf16(&tmA, g->A, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    // B: tb == 0 -> stored [N, K]: dims {K, N}, box {64, BN}.  tb == 1 -> stored [K, N]: dims {N, K}, box {64, 64}
    if (!p.b_mn) { dims[0] = g->K; dims[1] = g->N; box[0] = TC_BK; box[1] = p.BN; }
    else { dims[0] = g->N; dims[1] = g->K; box[0] = 64; box[1] = TC_BK; }
    str[0] = (uint64_t)g->ldb * 2;
    rc = bpm_make_tmap_bf16(&tmB, g->B, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    if ((rc = make_epi_map(&tmC, g->C, g->M, g->N, g->ldc, elem, p.chunk_bytes))) return rc;
    tmR = tmC; tmG = tmC;
    if (g->residual && (rc = make_epi_map(&tmR, g->residual, g->M, g->N, g->ldr, elem, p.chunk_bytes))) return rc;
    if (g->gate && (rc = make_epi_map(&tmG, g->gate, g->M, g->N, g->ldg, elem, p.chunk_bytes))) return rc;
  }
  size_t smem = (size_t)p.ring_bytes + fixed;
  BPM_REQUIRE(smem <= 227 * 1024 && p.tmem_cols <= 512, "gemm_tc: smem %zu / tmem %d too large", smem, p.tmem_cols);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
    if (e != cudaSuccess) { bpm_set_error("gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return BPM_ELAUNCH; }
    attr_set = true;
  }
  int total_tiles = gx * gy * split;
  int ctas = min(total_tiles, bpm_num_sms());
  gemm_tc_kernel<<<ctas, TC_THREADS, smem, stream>>>(tmA, tmB, tmC, tmR, tmG, p);
  BPM_CHECK_LAUNCH("gemm_tc");
  return BPM_OK;
}
