// bf16 GEMM on 5th-gen tensor cores (sm_100a): TMA (SWIZZLE_128B) -> shared memory ring -> tcgen05.mma (cta_group::1,
// M = 128, N = BN <= 256, K = 16 per instruction) -> fp32 accumulator in TMEM -> tcgen05.ld -> fused epilogue.
//
//  * Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
//    warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4).
//  * All four operand layouts without any transposed copies: K-major operands use the canonical SW128 K-major
//    layout (SBO = 1024 B); "transposed" operands (dgrad's W, wgrad's dY^T and X) are loaded as 64-wide MN chunks
//    and described to the MMA as MN-major (LBO = BK*128 B between chunks, SBO = 1024 B between 8-row k groups).
//  * K = 300..1216 here, i.e. only 5..19 k-blocks per tile: the kernel is prologue/epilogue dominated, so it is sized
//    for TWO co-resident CTAs per SM (<= 110 KB smem, <= 256 TMEM columns each): one CTA's epilogue overlaps the
//    other's main loop.  wgrad (K = B*T rows) is split along K across blockIdx.z and accumulated with vector
//    fp32 reductions (red.global.add.v4.f32).
#include "gemm_epilogue.cuh"
#include "tc_common.cuh"

#define TC_BM 128
#define TC_BK 64
#define TC_THREADS 192

struct TcGemmParams {
  int M, N, K;
  int BN;            // multiple of 16, <= 256
  int a_mn, b_mn;    // 1: operand is MN-major in memory ("transposed")
  int stages;
  int a_bytes, b_bytes;   // per-stage bytes (multiples of 1024)
  int kb_per_split;
  int tmem_cols;
  int plain_acc;     // split-K accumulate with a pure (alpha-only) epilogue
  uint32_t idesc;
  EpiParams ep;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(TC_THREADS, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B tiles
  uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int stage_bytes = p.a_bytes + p.b_bytes;
  uint32_t bar_base = smem_base + p.stages * stage_bytes;      // full[stages], empty[stages], tmem_full, tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  uint32_t tmem_full_bar = bar_base + 8u * (2 * p.stages);
  uint32_t tmem_ptr_addr = bar_base + 8u * (2 * p.stages + 1);
  volatile uint32_t* tmem_ptr_gen = (volatile uint32_t*)(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * p.BN, m0 = blockIdx.y * TC_BM;
  const int num_kb_total = (p.K + TC_BK - 1) / TC_BK;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int kb1 = min(num_kb_total, kb0 + p.kb_per_split);
  const int num_kb = kb1 - kb0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_addr, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (num_kb > 0) {
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        const int b_chunks = (p.BN + 63) / 64;
        for (int i = 0; i < num_kb; i++) {
          int s = i % p.stages;
          uint32_t ph = (uint32_t)(i / p.stages) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
          uint32_t sa = smem_base + s * stage_bytes, sb = sa + p.a_bytes;
          mbar_expect_tx(full_bar(s), (uint32_t)(p.a_bytes + p.b_bytes));
          int k = (kb0 + i) * TC_BK;
          if (!p.a_mn) tma_load_2d(sa, &tmA, full_bar(s), k, m0);                        // box {64 k, 128 m}
          else { tma_load_2d(sa, &tmA, full_bar(s), m0, k); tma_load_2d(sa + 8192, &tmA, full_bar(s), m0 + 64, k); }   // box {64 m, 64 k} x2
          if (!p.b_mn) tma_load_2d(sb, &tmB, full_bar(s), k, n0);                        // box {64 k, BN n}
          else
            for (int c = 0; c < b_chunks; c++) tma_load_2d(sb + c * 8192, &tmB, full_bar(s), n0 + 64 * c, k);        // box {64 n, 64 k}
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer (one thread) =====================
      if (lane == 0) {
        for (int i = 0; i < num_kb; i++) {
          int s = i % p.stages;
          uint32_t ph = (uint32_t)(i / p.stages) & 1u;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          uint32_t sa = smem_base + s * stage_bytes, sb = sa + p.a_bytes;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; k++) {
            // K-major: +32 B per 16-element k step inside the 128 B swizzle row; MN-major: +16 rows * 128 B
            uint64_t da = p.a_mn ? umma_desc(sa + k * 2048, TC_BK * 128, 1024, BPM_SWZ_128B) : umma_desc(sa + k * 32, 16, 1024, BPM_SWZ_128B);
            uint64_t db = p.b_mn ? umma_desc(sb + k * 2048, TC_BK * 128, 1024, BPM_SWZ_128B) : umma_desc(sb + k * 32, 16, 1024, BPM_SWZ_128B);
            umma_bf16(tmem_base, da, db, p.idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(s));            // frees the smem slot once these MMAs have read it
        }
        umma_commit(tmem_full_bar);             // accumulator complete
      }
    } else {
      // ===================== epilogue warps =====================
      const int quarter = warp & 3;
      const int row = m0 + quarter * 32 + lane;
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
      DropCtx dc = make_drop(p.ep.drop);
      const bool vec_ok = (p.ep.ldc % 8 == 0) && (!p.ep.gate || p.ep.ldg % 8 == 0) && (!p.ep.residual || p.ep.ldr % 8 == 0);
      for (int c0 = 0; c0 < p.BN; c0 += 32) {
        float v[32];
        uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0;
        if (c0 + 32 <= p.BN) tmem_ld32(taddr, v);
        else tmem_ld16(taddr, v);               // BN % 32 == 16 tail
        tmem_ld_wait();
        if (row < p.M) {
          int ncols = min(32, p.BN - c0);
          for (int j = 0; j < ncols; j += 8) {
            int n = n0 + c0 + j;
            if (n >= p.N) break;
            if (vec_ok && n + 8 <= p.N) {
              if (p.plain_acc) {
                float* c = (float*)p.ep.C + (int64_t)row * p.ep.ldc + n;
                red_add_v4(c, v[j] * p.ep.alpha, v[j + 1] * p.ep.alpha, v[j + 2] * p.ep.alpha, v[j + 3] * p.ep.alpha);
                red_add_v4(c + 4, v[j + 4] * p.ep.alpha, v[j + 5] * p.ep.alpha, v[j + 6] * p.ep.alpha, v[j + 7] * p.ep.alpha);
              } else {
                epi_store8(p.ep, dc, row, n, v + j);
              }
            } else {
              for (int jj = 0; jj < 8 && n + jj < p.N; jj++) epi_store1(p.ep, dc, row, n + jj, v[j + jj]);
            }
          }
        }
      }
    }
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

int bpm_make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                       CUtensorMapSwizzle swz) {
  bpm_encode_tiled_fn enc = bpm_get_encode_tiled();
  if (!enc) { bpm_set_error("cuTensorMapEncodeTiled entry point unavailable"); return BPM_ELAUNCH; }
  cuuint64_t gd[5]; cuuint64_t gs[5]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; i++) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; i++) gs[i] = strides_bytes[i];
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    bpm_set_error("cuTensorMapEncodeTiled failed (%d): base %p rank %d dims %llu,%llu stride %llu box %u,%u", (int)r, base, rank,
                  (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)strides_bytes[0], box[0], box[1]);
    return BPM_EINVAL;
  }
  return BPM_OK;
}

static int pick_bn(int N) {
  // fewest tiles first, then least padding; BN multiple of 16 in [32, 256]
  int tiles = bpm_cdiv(N, 256);
  int bn = bpm_cdiv(bpm_cdiv(N, tiles), 16) * 16;
  return bn < 32 ? 32 : bn;
}

int bpm_gemm_tc(const bpm_gemm_t* g, cudaStream_t stream) {
  BPM_REQUIRE(((uintptr_t)g->A % 16 == 0) && ((uintptr_t)g->B % 16 == 0) && g->lda % 8 == 0 && g->ldb % 8 == 0,
              "gemm(bf16): A/B must be 16-byte aligned with pitches multiple of 8 elements (lda %d ldb %d)", g->lda, g->ldb);
  BPM_REQUIRE(!g->accumulate || g->c_dtype == BPM_F32, "gemm: accumulate needs fp32 C");
  TcGemmParams p;
  p.M = g->M; p.N = g->N; p.K = g->K;
  p.BN = pick_bn(g->N);
  p.a_mn = g->ta ? 1 : 0;
  p.b_mn = g->tb ? 1 : 0;
  p.a_bytes = TC_BM * TC_BK * 2;
  p.b_bytes = p.b_mn ? bpm_cdiv(p.BN, 64) * 8192 : bpm_cdiv(p.BN * 128, 1024) * 1024;
  int stage_bytes = p.a_bytes + p.b_bytes;
  p.stages = max(2, min(4, (108 * 1024) / stage_bytes));
  p.tmem_cols = p.BN <= 32 ? 32 : p.BN <= 64 ? 64 : p.BN <= 128 ? 128 : 256;
  p.idesc = umma_idesc_bf16(TC_BM, p.BN, p.a_mn, p.b_mn);
  p.ep = make_epi(g);
  int num_kb = bpm_cdiv(g->K, TC_BK);
  int gx = bpm_cdiv(g->N, p.BN), gy = bpm_cdiv(g->M, TC_BM);
  int split = 1;
  p.plain_acc = g->accumulate && !g->bias && g->act == 0 && g->drop.p == 0.f && !g->gate && !g->residual;
  if (p.plain_acc) {
    split = g->split_k > 0 ? g->split_k : max(1, min(num_kb / 4, (2 * bpm_num_sms()) / max(1, gx * gy)));
  }
  p.kb_per_split = bpm_cdiv(num_kb, split);
  split = bpm_cdiv(num_kb, p.kb_per_split);

  CUtensorMap tmA, tmB;
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
  }
  size_t smem = (size_t)p.stages * stage_bytes + 1024 + 8 * (2 * p.stages + 2);
  static size_t smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(111 * 1024));
    if (e != cudaSuccess) { bpm_set_error("gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return BPM_ELAUNCH; }
    smem_set = 111 * 1024;
  }
  BPM_REQUIRE(smem <= 111 * 1024, "gemm_tc: smem %zu too large", smem);
  dim3 grid(gx, gy, split);
  gemm_tc_kernel<<<grid, TC_THREADS, smem, stream>>>(tmA, tmB, p);
  BPM_CHECK_LAUNCH("gemm_tc");
  return BPM_OK;
}
