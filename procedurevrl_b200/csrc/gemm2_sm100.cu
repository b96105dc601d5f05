// 2-CTA ("cta_group::2") variant of the persistent tcgen05 GEMM: a cluster of two CTAs on the two SMs of a TPC computes
// one 256 x 256 output tile.  CTA r loads rows [128r, 128r+128) of the A tile and rows [128r, 128r+128) of the B tile
// (16 KB + 16 KB per 64-deep stage instead of 16 + 32), the leader CTA's single MMA thread issues
// tcgen05.mma.cta_group::2 (256 x 256 x 16) which reads A from each CTA's own shared memory and shares the two B halves
// between the pair, and each CTA ends up with its own 128 x 256 fp32 accumulator in its own TMEM.
//
// Why: the 1-CTA kernel (gemm_sm100.cu) moves 96 KB through each SM's shared memory per 64-deep k-block (48 KB written
// by TMA + 48 KB read by the MMAs) against 512 tensor-core cycles -- measured, it runs at the shared-memory rate
// (~0.45 us per k-block), not the tensor rate.  The CTA pair needs 64 KB per SM for the same FLOPs, and the smaller
// stages leave room for a 6-deep ring.
//
// Barrier protocol (per stage s / accumulator buffer a; "leader" = cluster rank 0):
//   full[s]   leader's only, 1 arrival: the leader's producer arms it with the bytes of BOTH CTAs (arrive.expect_tx) and
//             both CTAs' TMA loads complete_tx on it (cp.async.bulk.tensor.cta_group::2).  The peer's producer never
//             arrives: a per-stage remote release-arrive from its single thread measured ~0.9 us per k-block; its loads
//             for phase n+1 of a stage cannot start before the commit that follows the MMAs of phase n, so its
//             complete_tx can only run ahead of the leader's expect_tx inside the same phase (tx-count may go negative).
//   empty[s]  one per CTA, 1 arrival: tcgen05.commit.cta_group::2 multicast to both CTAs once the MMAs have read stage s.
//   tfull[a]  one per CTA, 1 arrival: commit multicast when the accumulator is complete -> both CTAs' epilogue warps.
//   tempty[a] leader's only, 16 arrivals: the 8 epilogue warps of each CTA (the peer's arrive remotely).
// Epilogue, row maps and fused epilogues are the shared code of gemm_common.cuh.
#include <cstdlib>

#include "gemm_common.cuh"

namespace pvrl {
namespace {

constexpr int TILE2_M = 256, TILE2_N = 256;
constexpr int B2_BYTES = (TILE2_N / 2) * BK * 2;             // this CTA's half of the B tile: 16 KB
constexpr int STAGE2_BYTES = A_BYTES + B2_BYTES;             // 32 KB
// EW = epilogue warps per CTA: 8 (two 128-column groups per TMEM lane quarter, 6-stage ring; the default everywhere) or
// 16 (four 64-column groups, 5-stage ring: an experiment for the bf16 GELU / gelu' epilogues, PVRL_GEMM2_EW=16, slower).
template <int EW>
struct Cfg2 {
  static constexpr int STAGES = EW == 8 ? 6 : 5;
  static constexpr int PIPE_BYTES = STAGES * STAGE2_BYTES;   // 192 / 160 KB
  static constexpr int THREADS = 128 + 32 * EW;
  static constexpr int GROUP_COLS = TILE2_N / (EW / 4);
  static constexpr int SMEM_BYTES = PIPE_BYTES + EW * EPI_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static_assert(SMEM_BYTES <= 232448 && GROUP_COLS % 32 == 0, "tile configuration");
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA's layout) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion is signalled on an mbarrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem) {  // one whole warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at this offset in every CTA of `mask` once all MMAs issued so far by this thread have completed
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask)
               : "memory");
}

// warp-converged forms (the whole warp executes the call with warp-uniform operands, one elected lane issues)
__device__ __forceinline__ void umma_bf16_2sm_e(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm_e(uint32_t bar, uint16_t mask) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
      ::"r"(bar), "h"(mask)
      : "memory");
}

template <int EPI, typename OutT, bool TN, int EW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128 + 32 * EW, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs p) {
  using C = Cfg2<EW>;
  constexpr int STAGES2 = C::STAGES, PIPE2_BYTES = C::PIPE_BYTES, NUM_EPI_WARPS = EW;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t tiles_addr = (raw_addr + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024 B alignment
  uint8_t* smem = smem_raw + (tiles_addr - raw_addr);
  constexpr int BAR_OFF = PIPE2_BYTES + NUM_EPI_WARPS * EPI_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  const uint32_t bars_addr = tiles_addr + BAR_OFF;
  // barrier slots: full[0..5], empty[6..11], tmem_full[12..13], tmem_empty[14..15], tmem ptr at slot 16
  auto full_bar = [&](int s) { return bars_addr + 8u * s; };
  auto empty_bar = [&](int s) { return bars_addr + 8u * (STAGES2 + s); };
  auto tfull_bar = [&](int s) { return bars_addr + 8u * (2 * STAGES2 + s); };
  auto tempty_bar = [&](int s) { return bars_addr + 8u * (2 * STAGES2 + 2 + s); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(bars + 2 * STAGES2 + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  const int m_tiles = (p.M + TILE2_M - 1) / TILE2_M;
  const int n_tiles = (p.N + TILE2_N - 1) / TILE2_N;
  const int num_kb = (p.K + BK - 1) / BK;
  const int total_tiles = m_tiles * n_tiles * p.k_splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES2; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 2 * NUM_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_2sm<512>(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)));
  tc_fence_before();
  cluster_sync_all();   // both CTAs' barriers are initialised before any remote arrive / TMA completion can reach them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (PDL_EARLY_TRIGGER) pdl_launch_dependents();
  pdl_wait();   // the set-up above overlapped the tail of the previous kernel

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs, own halves)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
        int m_blk, n_blk, ks;
        decode_tile<TN>(tile, m_tiles, n_tiles, p.k_splits, m_blk, n_blk, ks);
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(num_kb, kb0 + p.kb_per_split);
        const int m0 = m_blk * TILE2_M + static_cast<int>(rank) * 128;
        const int n0 = n_blk * TILE2_N + static_cast<int>(rank) * 128;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t leader_full = mapa_rank(full_bar(stage), 0);
          if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * STAGE2_BYTES);
          const uint32_t sA = tiles_addr + stage * STAGE2_BYTES;
          const uint32_t sB = sA + A_BYTES;
          if (!TN) {
            tma_load_2d_2sm(sA, &tmA, leader_full, kb * BK, m0);
            tma_load_2d_2sm(sB, &tmB, leader_full, kb * BK, n0);
          } else {
#pragma unroll
            for (int c = 0; c < 2; ++c) tma_load_2d_2sm(sA + c * CHUNK_BYTES, &tmA, leader_full, m0 + c * 64, kb * BK);
#pragma unroll
            for (int c = 0; c < 2; ++c) tma_load_2d_2sm(sB + c * CHUNK_BYTES, &tmB, leader_full, n0 + c * 64, kb * BK);
          }
          if (++stage == STAGES2) stage = 0, phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp 1 of the leader CTA, converged;
    // one elected lane issues)
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(TILE2_M, TILE2_N, TN ? 1 : 0, TN ? 1 : 0);
      constexpr uint32_t kstep = TN ? 16u * 128u : 32u;
      constexpr uint32_t lbo = TN ? CHUNK_BYTES : 16u;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
        int m_blk, n_blk, ks;
        decode_tile<TN>(tile, m_tiles, n_tiles, p.k_splits, m_blk, n_blk, ks);
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(num_kb, kb0 + p.kb_per_split);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * TILE2_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sA = tiles_addr + stage * STAGE2_BYTES;
          const uint32_t sB = sA + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = make_smem_desc(sA + k * kstep, lbo, 1024);
            const uint64_t bdesc = make_smem_desc(sB + k * kstep, lbo, 1024);
            umma_bf16_2sm_e(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit_2sm_e(empty_bar(stage), 0b11);   // both CTAs' stage s is free once these MMAs have read it
          if (++stage == STAGES2) stage = 0, phase ^= 1u;
        }
        umma_commit_2sm_e(tfull_bar(acc), 0b11);       // accumulator complete -> both CTAs' epilogue warps
        if (++acc == 2) acc = 0, acc_phase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (both CTAs, own 128 rows)
    const int quarter = warp & 3;
    const int half = (warp - 4) >> 2;
    constexpr int HALF_COLS = C::GROUP_COLS;
    uint8_t* stg = smem + PIPE2_BYTES + (warp - 4) * EPI_STAGE_BYTES;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
      int m_blk, n_blk, ks;
      decode_tile<TN>(tile, m_tiles, n_tiles, p.k_splits, m_blk, n_blk, ks);
      const int m_base = m_blk * TILE2_M + static_cast<int>(rank) * 128 + quarter * 32;
      epilogue_tile<EPI, OutT, HALF_COLS>(
          p, stg, tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * TILE2_N + half * HALF_COLS, m_base,
          n_blk * TILE2_N + half * HALF_COLS, tfull_bar(acc), acc_phase, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(tempty_bar(acc));
        else mbar_arrive_cluster(mapa_rank(tempty_bar(acc), 0));
      }
      if (++acc == 2) acc = 0, acc_phase ^= 1u;
    }
  }
  tc_fence_before();
  cluster_sync_all();   // no CTA leaves (or frees TMEM) while its peer may still signal its barriers / read its smem
  if (warp == 2) tmem_dealloc_2sm<512>(tmem_base);
}

// CTA pairs that can be co-resident (the persistent tile loop strides by the number of launched pairs, so launching
// more pairs than fit at once would serialise them into waves).
template <typename Kern>
int max_active_pairs(Kern kern, int threads, int smem_bytes) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(num_sms(), 1, 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 2, attr.val.clusterDim.y = 1, attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = num_sms() / 2;
  }
  return n;
}

int g_pairs_override = [] {
  const char* e = getenv("PVRL_GEMM2_PAIRS");   // development knob
  return e ? atoi(e) : 0;
}();
int g_last_max_pairs = 0;

template <int EPI, typename OutT, bool TN, int EW>
int launch_gemm2_ew(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& a, cudaStream_t stream) {
  auto kern = gemm2_bf16_kernel<EPI, OutT, TN, EW>;
  constexpr int SMEM2_BYTES = Cfg2<EW>::SMEM_BYTES, NUM_THREADS = Cfg2<EW>::THREADS;
  static int pairs_max = 0;
  if (pairs_max == 0) {
    PVRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES));
    pairs_max = max_active_pairs(kern, NUM_THREADS, SMEM2_BYTES);
    g_last_max_pairs = pairs_max;
  }
  const int m_tiles = (a.M + TILE2_M - 1) / TILE2_M, n_tiles = (a.N + TILE2_N - 1) / TILE2_N;
  const int total = m_tiles * n_tiles * a.k_splits;
  const int cap = g_pairs_override > 0 ? g_pairs_override : pairs_max;
  const int clusters = total < cap ? total : cap;
  PVRL_CUDA(launch_pdl(kern, dim3(2 * clusters), dim3(NUM_THREADS), SMEM2_BYTES, stream, ta, tb, a));
  return launched("gemm2_bf16_kernel");
}

template <int EPI, typename OutT, bool TN>
int launch_gemm2(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& a, cudaStream_t stream) {
  constexpr bool HEAVY = (EPI == PVRL_EPI_GELU || EPI == PVRL_EPI_DGELU) && sizeof(OutT) == 2;
  static const int heavy_ew = [] {
    // development knob: 16 selects the four-group epilogue for the GELU GEMMs.  Measured (K = 768, N = 3072): 150 / 150 us
    // against 141 / 130 us with two groups -- the epilogue is not what bounds these tiles on a CTA pair, and the fifth
    // and sixth ring stages are worth more than the extra warps.
    const char* e = getenv("PVRL_GEMM2_EW");
    return e ? atoi(e) : 8;
  }();
  if (HEAVY && heavy_ew == 16) return launch_gemm2_ew<EPI, OutT, TN, HEAVY ? 16 : 8>(ta, tb, a, stream);
  return launch_gemm2_ew<EPI, OutT, TN, 8>(ta, tb, a, stream);
}

}  // namespace

// Entry used by pvrl_gemm_bf16 (gemm_sm100.cu) after argument validation.
int gemm2_dispatch(const pvrl_gemm_t* d, cudaStream_t stream) {
  GemmArgs a = make_gemm_args(d);
  const int num_kb = (d->K + BK - 1) / BK;
  const int tiles = ((d->M + TILE2_M - 1) / TILE2_M) * ((d->N + TILE2_N - 1) / TILE2_N);
  int splits = 1;
  if (d->epilogue == PVRL_EPI_ATOMIC) {
    splits = d->k_splits;
    if (splits <= 0) {
      // Fill one whole wave of CTA pairs (PVRL_GEMM2_WAVES = 2 / 3: more, shorter tiles per pair so that the atomic
      // epilogue of one hides behind the MMAs of the next -- measured within noise of a single wave, at twice the
      // red.add traffic, so one wave stays the default).
      static const int waves_env = [] {
        const char* e = getenv("PVRL_GEMM2_WAVES");
        return e ? atoi(e) : 1;
      }();
      const int pairs = num_sms() / 2;
      double best = -1.0;
      splits = 1;
      for (int s = 1; s <= 64 && s <= num_kb; ++s) {
        if (s > 1 && (num_kb + s - 1) / s < 8) break;
        const int work = tiles * s;
        const int waves = (work + pairs - 1) / pairs;
        if (waves > waves_env) break;
        const double util = static_cast<double>(work) / (waves * pairs) + 0.05 * (waves == waves_env);
        if (util > best + 0.02) best = util, splits = s;
      }
    }
    if (splits > num_kb) splits = num_kb;
  }
  a.kb_per_split = (num_kb + splits - 1) / splits;
  a.k_splits = (num_kb + a.kb_per_split - 1) / a.kb_per_split;

  CUtensorMap ta, tb;
  int rc;
  if (d->trans == 0) {
    if ((rc = make_tmap_2d_bf16(&ta, d->A, d->K, d->M, d->lda, BK, 128))) return rc;
    if ((rc = make_tmap_2d_bf16(&tb, d->B, d->K, d->N, d->ldb, BK, 128))) return rc;
  } else {
    if ((rc = make_tmap_2d_bf16(&ta, d->A, d->M, d->K, d->lda, 64, BK))) return rc;
    if ((rc = make_tmap_2d_bf16(&tb, d->B, d->N, d->K, d->ldb, 64, BK))) return rc;
  }
  const bool f32 = d->out_dtype == PVRL_F32;
  switch (d->epilogue) {
    case PVRL_EPI_STORE:
      return f32 ? launch_gemm2<PVRL_EPI_STORE, float, false>(ta, tb, a, stream)
                 : launch_gemm2<PVRL_EPI_STORE, __nv_bfloat16, false>(ta, tb, a, stream);
    case PVRL_EPI_GELU:
      return f32 ? launch_gemm2<PVRL_EPI_GELU, float, false>(ta, tb, a, stream)
                 : launch_gemm2<PVRL_EPI_GELU, __nv_bfloat16, false>(ta, tb, a, stream);
    case PVRL_EPI_DGELU:
      return f32 ? launch_gemm2<PVRL_EPI_DGELU, float, false>(ta, tb, a, stream)
                 : launch_gemm2<PVRL_EPI_DGELU, __nv_bfloat16, false>(ta, tb, a, stream);
    case PVRL_EPI_RESID:
      return d->add_pos != nullptr ? launch_gemm2<EPI_RESID_POS, float, false>(ta, tb, a, stream)
                                   : launch_gemm2<PVRL_EPI_RESID, float, false>(ta, tb, a, stream);
    default:
      return d->trans ? launch_gemm2<PVRL_EPI_ATOMIC, float, true>(ta, tb, a, stream)
                      : launch_gemm2<PVRL_EPI_ATOMIC, float, false>(ta, tb, a, stream);
  }
}

}  // namespace pvrl

// development aid: co-resident CTA pairs reported by cudaOccupancyMaxActiveClusters for the last configured variant
extern "C" int pvrl_debug_gemm2_max_pairs(void) { return pvrl::g_last_max_pairs; }
