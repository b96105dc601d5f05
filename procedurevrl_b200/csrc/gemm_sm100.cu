// Persistent warp-specialised bf16 GEMM for sm_100a: TMA (128B-swizzled tiles) -> 4-stage smem ring ->
// tcgen05.mma (128x256x16, fp32 accumulators double-buffered in TMEM) -> fused epilogues straight from
// TMEM to global memory.  One CTA per SM; warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM
// allocator, warps 4..11 = epilogue (each owns a 32-lane TMEM quarter and one 128-column half).
//
// Covers every dense contraction of the TimeSformer block (SURVEY.md 2a K1,K4,K6,K8,K10 and their
// backward K15): "NT" for y = x W^T (+ dX through pre-transposed weights) and "TN" (MN-major operands,
// split-K + fp32 red.add) for dW = dY^T X.
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "gemm_common.cuh"

namespace pvrl {
namespace {

constexpr int STAGES = 4;
// EW = epilogue warps: 8 (two 128- / 96-column groups per TMEM lane quarter) or, for the GELU / gelu' epilogues whose
// per-element arithmetic makes the epilogue the longer phase of a K = 768 tile, 12 with BN = 192 (three 64-column groups):
// half the elements per thread, 3 instead of 2 warps per scheduler to hide the MUFU / FMA latencies.
template <int BN, int EW>
struct Cfg {
  static constexpr int B_BYTES = BN * BK * 2;                // 32 KB (BN 256) / 24 KB (BN 192)
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
  static constexpr int THREADS = 128 + 32 * EW;
  static constexpr int GROUP_COLS = BN / (EW / 4);
  static constexpr int SMEM_BYTES = PIPE_BYTES + EW * EPI_STAGE_BYTES + 1024 /*align*/ + 128 /*barriers*/;
  static_assert(GROUP_COLS % 32 == 0 && SMEM_BYTES <= 232448, "tile configuration");
};


template <int EPI, typename OutT, bool TN, int BN, int EW>
__global__ void __launch_bounds__(128 + 32 * EW, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs p) {
  using C = Cfg<BN, EW>;
  constexpr int STAGE_BYTES = C::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t tiles_addr = (raw_addr + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024 B alignment
  uint8_t* smem = smem_raw + (tiles_addr - raw_addr);
  const uint32_t epi_addr = tiles_addr + C::PIPE_BYTES;
  constexpr int BAR_OFF = C::PIPE_BYTES + EW * EPI_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  const uint32_t bars_addr = tiles_addr + BAR_OFF;
  // barrier slots: full[0..3], empty[4..7], tmem_full[8..9], tmem_empty[10..11], tmem ptr at slot 12
  auto full_bar = [&](int s) { return bars_addr + 8u * s; };
  auto empty_bar = [&](int s) { return bars_addr + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bars_addr + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bars_addr + 8u * (2 * STAGES + 2 + s); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (p.M + BM - 1) / BM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int num_kb = (p.K + BK - 1) / BK;
  const int total_tiles = m_tiles * n_tiles * p.k_splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), EW);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (PDL_EARLY_TRIGGER) pdl_launch_dependents();
  pdl_wait();   // everything above overlapped the tail of the previous kernel; from here on its outputs are read

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int m_blk, n_blk, ks;
        decode_tile<TN>(tile, m_tiles, n_tiles, p.k_splits, m_blk, n_blk, ks);
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(num_kb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), STAGE_BYTES);
          const uint32_t sA = tiles_addr + stage * STAGE_BYTES;
          const uint32_t sB = sA + A_BYTES;
          if (!TN) {
            tma_load_2d(sA, &tmA, full_bar(stage), kb * BK, m_blk * BM);
            tma_load_2d(sB, &tmB, full_bar(stage), kb * BK, n_blk * BN);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c)
              tma_load_2d(sA + c * CHUNK_BYTES, &tmA, full_bar(stage), m_blk * BM + c * 64, kb * BK);
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_2d(sB + c * CHUNK_BYTES, &tmB, full_bar(stage), n_blk * BN + c * 64, kb * BK);
          }
          if (++stage == STAGES) stage = 0, phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: the whole warp runs the loop
    // converged and one elected lane issues (pvrl_ptx.cuh: a single-lane branch routes every descriptor through R2UR
    // moves and an ELECT / BRA.U.ANY loop per instruction -- ~100 cycles per tcgen05.mma against the 96 / 128 it executes)
    {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, TN ? 1 : 0, TN ? 1 : 0);
      // descriptor advance per UMMA_K = 16: K-major +32 B inside the swizzle atom, MN-major +16 k-rows
      constexpr uint32_t kstep = TN ? 16u * 128u : 32u;
      constexpr uint32_t lbo = TN ? CHUNK_BYTES : 16u;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int m_blk, n_blk, ks;
        decode_tile<TN>(tile, m_tiles, n_tiles, p.k_splits, m_blk, n_blk, ks);
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(num_kb, kb0 + p.kb_per_split);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sA = tiles_addr + stage * STAGE_BYTES;
          const uint32_t sB = sA + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = make_smem_desc(sA + k * kstep, lbo, 1024);
            const uint64_t bdesc = make_smem_desc(sB + k * kstep, lbo, 1024);
            umma_bf16_e(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit_e(empty_bar(stage));  // smem slot reusable once these MMAs have read it
          if (++stage == STAGES) stage = 0, phase ^= 1u;
        }
        umma_commit_e(tfull_bar(acc));      // accumulator complete -> epilogue
        if (++acc == 2) acc = 0, acc_phase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: TMEM -> regs -> smem transpose -> global
    const int quarter = warp & 3;          // TMEM lanes [32*quarter, +32) are the only ones this warp may read
    const int half = (warp - 4) >> 2;      // column group of the BN-wide accumulator
    constexpr int HALF_COLS = C::GROUP_COLS;
    uint8_t* stg = smem + C::PIPE_BYTES + (warp - 4) * EPI_STAGE_BYTES;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int m_blk, n_blk, ks;
      decode_tile<TN>(tile, m_tiles, n_tiles, p.k_splits, m_blk, n_blk, ks);
      const int m_base = m_blk * BM + quarter * 32;
      epilogue_tile<EPI, OutT, HALF_COLS>(p, stg, tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN +
                                                       half * HALF_COLS,
                                          m_base, n_blk * BN + half * HALF_COLS, tfull_bar(acc), acc_phase, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) acc = 0, acc_phase ^= 1u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

struct MapKey {
  const void* ptr;
  uint64_t inner, outer, ld;
  uint32_t box_inner, box_outer;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && ld == o.ld && box_inner == o.box_inner &&
           box_outer == o.box_outer;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h = h * 1000003u ^ k.inner;
    h = h * 1000003u ^ k.outer;
    h = h * 1000003u ^ k.ld;
    h = h * 1000003u ^ (k.box_inner << 16 | k.box_outer);
    return h;
  }
};

}  // namespace

// ---- host: TMA descriptor cache ---------------------------------------------------------------------
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

// 2-D bf16 tensor map over a row-major [outer, inner] matrix with leading dimension ld (elements),
// box = [box_outer, box_inner], 128-byte swizzle, zero fill out of bounds.
int make_tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                      uint32_t box_inner, uint32_t box_outer) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{ptr, inner, outer, ld, box_inner, box_outer};
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) return fail(-2, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15u) != 0 || (ld * 2) % 16 != 0)
    return fail(-1, "TMA operand must be 16-byte aligned with a leading dimension multiple of 8 elements");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(-3, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  std::lock_guard<std::mutex> lock(mu);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return 0;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

namespace {

template <int EPI, typename OutT, bool TN, int BN, int EW>
int launch_gemm_bn(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& a, cudaStream_t stream) {
  auto kern = gemm_bf16_kernel<EPI, OutT, TN, BN, EW>;
  constexpr int SMEM_BYTES = Cfg<BN, EW>::SMEM_BYTES;
  static bool configured = false;
  if (!configured) {
    PVRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  const int m_tiles = (a.M + BM - 1) / BM, n_tiles = (a.N + BN - 1) / BN;
  const int total = m_tiles * n_tiles * a.k_splits;
  const int grid = total < num_sms() ? total : num_sms();
  PVRL_CUDA(launch_pdl(kern, dim3(grid), dim3(Cfg<BN, EW>::THREADS), SMEM_BYTES, stream, ta, tb, a));
  return launched("gemm_bf16_kernel");
}

template <int EPI, typename OutT, bool TN>
int launch_gemm(int bn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& a, cudaStream_t stream) {
  // bf16 GELU / gelu' epilogues on 192-wide tiles: 12 epilogue warps (see Cfg)
  constexpr bool HEAVY = (EPI == PVRL_EPI_GELU || EPI == PVRL_EPI_DGELU) && sizeof(OutT) == 2;
  if (bn == 192) {
    if (HEAVY) return launch_gemm_bn<EPI, OutT, TN, 192, HEAVY ? 12 : 8>(ta, tb, a, stream);
    return launch_gemm_bn<EPI, OutT, TN, 192, 8>(ta, tb, a, stream);
  }
  return launch_gemm_bn<EPI, OutT, TN, 256, 8>(ta, tb, a, stream);
}

// Tile width and split-K count for one problem: the (BN, splits) pair whose tiles fill whole waves of SMs best
// (useful columns / padded columns x busy SM slots / scheduled SM slots).  Ties go to the wider tile (fewer A re-reads).
void pick_tiling(int M, int N, int num_kb, bool splitk, int forced_splits, int forced_bn, int* bn_out, int* splits_out) {
  const int sms = num_sms();
  if (forced_splits > num_kb) forced_splits = num_kb;
  double best = -1.0;
  *bn_out = 256, *splits_out = 1;
  for (int bn : {256, 192}) {
    if (forced_bn != 0 && bn != forced_bn) continue;
    const int n_tiles = (N + bn - 1) / bn;
    const int tiles = ((M + BM - 1) / BM) * n_tiles;
    // measured: the 192-wide tile moves ~5 % more operand bytes per FLOP through L2 -> it must win that back in waves
    const double col_eff = static_cast<double>(N) / (n_tiles * bn) * (bn == 192 ? 0.95 : 1.0);
    const int s_max = splitk ? (forced_splits > 0 ? forced_splits : 64) : 1;
    for (int s = (splitk && forced_splits > 0) ? forced_splits : 1; s <= s_max && s <= num_kb; ++s) {
      if (forced_splits <= 0 && s > 1 && (num_kb + s - 1) / s < 4) break;
      const int work = tiles * s;
      const double util = col_eff * work / (static_cast<double>((work + sms - 1) / sms) * sms);
      if (util > best + 0.02) best = util, *bn_out = bn, *splits_out = s;
    }
  }
}

}  // namespace
}  // namespace pvrl

namespace pvrl {
int gemm2_dispatch(const pvrl_gemm_t* d, cudaStream_t stream);   // gemm2_sm100.cu: 256 x 256 tiles on CTA pairs
}

extern "C" int pvrl_gemm_bf16(const pvrl_gemm_t* d, void* stream_) {
  using namespace pvrl;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PVRL_CHECK_ARG(d != nullptr, "pvrl_gemm_bf16: null descriptor");
  PVRL_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0, "pvrl_gemm_bf16: empty problem M=%d N=%d K=%d", d->M, d->N, d->K);
  PVRL_CHECK_ARG(d->N % 32 == 0, "pvrl_gemm_bf16: N=%d must be a multiple of 32", d->N);
  PVRL_CHECK_ARG(d->A && d->B && d->out, "pvrl_gemm_bf16: null operand");
  PVRL_CHECK_ARG(d->epilogue >= PVRL_EPI_STORE && d->epilogue <= PVRL_EPI_ATOMIC, "pvrl_gemm_bf16: bad epilogue %d",
                 d->epilogue);
  if (d->trans == 1) PVRL_CHECK_ARG(d->epilogue == PVRL_EPI_ATOMIC, "pvrl_gemm_bf16: TN form supports ATOMIC only");
  if (d->epilogue == PVRL_EPI_GELU) PVRL_CHECK_ARG(d->out2 != nullptr, "pvrl_gemm_bf16: GELU needs out2");
  if (d->epilogue == PVRL_EPI_DGELU) PVRL_CHECK_ARG(d->aux != nullptr, "pvrl_gemm_bf16: DGELU needs aux");
  if (d->rowscale) PVRL_CHECK_ARG(d->rs_div > 0, "pvrl_gemm_bf16: rowscale needs rs_div > 0");
  if (d->colsum)
    PVRL_CHECK_ARG(d->epilogue == PVRL_EPI_STORE || d->epilogue == PVRL_EPI_DGELU,
                   "pvrl_gemm_bf16: colsum is fused into the STORE / DGELU epilogues only");
  if (d->map == PVRL_MAP_SPATIAL)
    PVRL_CHECK_ARG(d->epilogue == PVRL_EPI_RESID && d->out2 != nullptr,
                   "pvrl_gemm_bf16: MAP_SPATIAL needs the RESID epilogue and a cls side buffer");

  // CTA-pair tiles (gemm2_sm100.cu) for the long contractions (dW, K >= 1536), single-CTA 128 x 256 / 128 x 192 tiles
  // for the K = 768 layers whose time is mostly epilogue and which gain from the finer wave granularity (measured,
  // profiles/): PVRL_GEMM_2CTA = 0 / 1 pins one of the two.
  static const int mode_2cta = [] {
    const char* e = getenv("PVRL_GEMM_2CTA");
    return e ? atoi(e) : 2;
  }();
  // bf16 GELU / gelu' epilogues (fc1 forward, fc2 input gradient; K = 768, N = 3072): measured on the CTA-pair kernel
  // 154 -> 141 us and 143 -> 131 us even with its two-group epilogue; PVRL_GEMM_HEAVY_2CTA=0 keeps them on 128 x 192 tiles.
  static const bool heavy_2cta = [] {
    const char* e = getenv("PVRL_GEMM_HEAVY_2CTA");
    return e == nullptr || atoi(e) != 0;
  }();
  const bool heavy = heavy_2cta && (d->epilogue == PVRL_EPI_GELU || d->epilogue == PVRL_EPI_DGELU) &&
                     d->out_dtype == PVRL_BF16 && d->N % 256 == 0;
  if (mode_2cta == 1 || (mode_2cta == 2 && (d->trans == 1 || d->K >= 1536 || heavy))) return gemm2_dispatch(d, stream);
  GemmArgs a = make_gemm_args(d);

  const int num_kb = (d->K + BK - 1) / BK;
  int splits = 1, bn = 256;
  static const int forced_bn = [] {
    const char* e = getenv("PVRL_GEMM_BN");   // development knob: pin the tile width (192 / 256)
    return e ? atoi(e) : 0;
  }();
  int want_bn = forced_bn;
  if (want_bn == 0 && (d->epilogue == PVRL_EPI_GELU || d->epilogue == PVRL_EPI_DGELU) && d->out_dtype == PVRL_BF16 &&
      d->N % 192 == 0)
    want_bn = 192;   // the 12-epilogue-warp variant
  pick_tiling(d->M, d->N, num_kb, d->epilogue == PVRL_EPI_ATOMIC, d->k_splits, want_bn, &bn, &splits);
  if (splits > num_kb) splits = num_kb;
  a.kb_per_split = (num_kb + splits - 1) / splits;
  a.k_splits = (num_kb + a.kb_per_split - 1) / a.kb_per_split;  // no empty split

  CUtensorMap ta, tb;
  int rc;
  if (d->trans == 0) {
    if ((rc = make_tmap_2d_bf16(&ta, d->A, d->K, d->M, d->lda, BK, BM))) return rc;
    if ((rc = make_tmap_2d_bf16(&tb, d->B, d->K, d->N, d->ldb, BK, bn))) return rc;
  } else {
    if ((rc = make_tmap_2d_bf16(&ta, d->A, d->M, d->K, d->lda, 64, BK))) return rc;
    if ((rc = make_tmap_2d_bf16(&tb, d->B, d->N, d->K, d->ldb, 64, BK))) return rc;
  }

  const bool f32 = d->out_dtype == PVRL_F32;
  switch (d->epilogue) {
    case PVRL_EPI_STORE:
      return f32 ? launch_gemm<PVRL_EPI_STORE, float, false>(bn, ta, tb, a, stream)
                 : launch_gemm<PVRL_EPI_STORE, __nv_bfloat16, false>(bn, ta, tb, a, stream);
    case PVRL_EPI_GELU:
      return f32 ? launch_gemm<PVRL_EPI_GELU, float, false>(bn, ta, tb, a, stream)
                 : launch_gemm<PVRL_EPI_GELU, __nv_bfloat16, false>(bn, ta, tb, a, stream);
    case PVRL_EPI_DGELU:
      return f32 ? launch_gemm<PVRL_EPI_DGELU, float, false>(bn, ta, tb, a, stream)
                 : launch_gemm<PVRL_EPI_DGELU, __nv_bfloat16, false>(bn, ta, tb, a, stream);
    case PVRL_EPI_RESID:
      return d->add_pos != nullptr ? launch_gemm<EPI_RESID_POS, float, false>(bn, ta, tb, a, stream)
                                   : launch_gemm<PVRL_EPI_RESID, float, false>(bn, ta, tb, a, stream);
    default:
      return d->trans ? launch_gemm<PVRL_EPI_ATOMIC, float, true>(bn, ta, tb, a, stream)
                      : launch_gemm<PVRL_EPI_ATOMIC, float, false>(bn, ta, tb, a, stream);
  }
}
