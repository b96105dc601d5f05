// Spatial attention (seq = 1 + H*W = 197 tokens per frame, head_dim 64) on the 5th-gen tensor cores:
// TMA loads the per-head Q/K/V (and dO) slices straight out of the fused QKV GEMM output through 3-D tensor
// maps (no head-split / transpose copies), tcgen05.mma computes Q K^T, P V and the four backward contractions
// with fp32 accumulators in TMEM, softmax runs one thread per query row on tcgen05.ld'ed scores and writes P
// (and dS) back to shared memory in the 128B-swizzled layout the next MMA reads.
// Attention.forward vit.py:84-88 and its autograd.  bf16 only, seq <= 256.
#include "pvrl_host.h"
#include "pvrl_ptx.cuh"

namespace pvrl {
namespace {

constexpr int TILE_BYTES = 128 * 128;  // [128 rows][64 bf16] SWIZZLE_128B tile
constexpr float LOG2E = 1.4426950408889634f;

// byte offset of 16-byte piece `piece` (8 bf16) of row r inside a [rows][128 B] SWIZZLE_128B tile
__device__ __forceinline__ uint32_t swz(int r, int piece) { return r * 128 + ((piece ^ (r & 7)) << 4); }

__device__ __forceinline__ void st_piece(uint8_t* tile, int r, int piece, const float* v) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]);
  u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(tile + swz(r, piece)) = u;
}

__device__ __forceinline__ void store_row64_bf16(__nv_bfloat16* dst, const float* v) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint4 u;
    u.x = pack_bf16x2(v[8 * i], v[8 * i + 1]);
    u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
    u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
    u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
    reinterpret_cast<uint4*>(dst)[i] = u;
  }
}

// =====================================================================================================
// forward: one CTA per (sequence, head, 128-query tile).  160 threads: warps 0-3 softmax / epilogue (one
// query row per thread), warp 4 = TMA + MMA issuer.  smem: [Q | K] (overlaid by P once S is done) + V.
// =====================================================================================================
__global__ void __launch_bounds__(160)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                   __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int seq, int H, float scale, int npad) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  // layout: P region 4 tiles (64 KB) holding Q at +0 and K at +16 KB until S is done; V after it
  const uint32_t sQ = base, sK = base + TILE_BYTES, sP = base, sV = base + 4 * TILE_BYTES;
  const uint32_t bars = sV + 256 * 128;
  const uint32_t bar_load = bars, bar_s = bars + 8, bar_p = bars + 16, bar_o = bars + 24;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + (bars - base) + 32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // grid (query tiles, pairs): the two query tiles of a (sequence, head) run back to back, so the second finds K / V in L2;
  // last sequences first: the freshest rows of the QKV GEMM output are still in L2
  const int pair = gridDim.y - 1 - blockIdx.y, mt = blockIdx.x;
  const int s_idx = pair / H, h = pair % H;
  const int C = H * 64;

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmKV);
      mbar_init(bar_load, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, 128);
      mbar_init(bar_o, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<256>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (PDL_EARLY_TRIGGER) pdl_launch_dependents();
  pdl_wait();   // barrier init / TMEM allocation above overlapped the previous kernel's tail

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(bar_load, TILE_BYTES + 2 * npad * 128);
      tma_load_3d(sQ, &tmQ, bar_load, h * 64, mt * 128, s_idx);
      tma_load_3d(sK, &tmKV, bar_load, C + h * 64, 0, s_idx);
      tma_load_3d(sV, &tmKV, bar_load, 2 * C + h * 64, 0, s_idx);
      mbar_wait(bar_load, 0);
      tc_fence_after();
      const uint32_t idesc_s = make_idesc_bf16(128, npad, 0, 0);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem, make_smem_desc(sQ + k * 32, 16, 1024), make_smem_desc(sK + k * 32, 16, 1024), idesc_s, k > 0);
      umma_commit(bar_s);
      mbar_wait(bar_p, 0);
      tc_fence_after();
      const uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);
      for (int k = 0; k < npad / 16; ++k)
        umma_bf16(tmem, make_smem_desc(sP + (k >> 2) * TILE_BYTES + (k & 3) * 32, 16, 1024),
                  make_smem_desc(sV + k * 2048, 8192, 1024), idesc_o, k > 0);
      umma_commit(bar_o);
    }
  } else {
    const int r = warp * 32 + lane;
    const int qi = mt * 128 + r;
    const uint32_t trow = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    mbar_wait(bar_s, 0);
    tc_fence_after();
    // Only the last 32-column chunk can hold padded keys (zero-filled K rows -> score 0): the full chunks run without
    // any per-element predicate, and four independent max / sum accumulators keep the FMNMX / FADD chains short.
    // TMEM reads are software-pipelined: the load of chunk c+1 is in flight while chunk c is processed.
    const int nch = npad >> 5, nfull = seq >> 5;
    uint32_t raw[2][32];
    float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    tmem_ld32(trow, raw[0]);
#pragma unroll 1
    for (int c = 0; c < nch; c += 2) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (c + u < nch) {
          tmem_ld_wait_on(raw[u]);
          if (c + u + 1 < nch) tmem_ld32(trow + (c + u + 1) * 32, raw[u ^ 1]);
          if (c + u < nfull) {
#pragma unroll
            for (int j = 0; j < 32; ++j) mx4[j & 3] = fmaxf(mx4[j & 3], __uint_as_float(raw[u][j]));
          } else {
            const int rem = seq - (c + u) * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < rem) mx4[j & 3] = fmaxf(mx4[j & 3], __uint_as_float(raw[u][j]));
          }
        }
      }
    }
    const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
    const float sl2 = scale * LOG2E;
    const float mxs = mx * sl2;
    float sum4[4] = {0.f, 0.f, 0.f, 0.f};
    tmem_ld32(trow, raw[0]);
#pragma unroll 1
    for (int c = 0; c < nch; c += 2) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (c + u < nch) {
          tmem_ld_wait_on(raw[u]);
          if (c + u + 1 < nch) tmem_ld32(trow + (c + u + 1) * 32, raw[u ^ 1]);
          float p[32];
          if (c + u < nfull) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              p[j] = ex2_approx(fmaf(__uint_as_float(raw[u][j]), sl2, -mxs));
              sum4[j & 3] += p[j];
            }
          } else {
            const int rem = seq - (c + u) * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              p[j] = j < rem ? ex2_approx(fmaf(__uint_as_float(raw[u][j]), sl2, -mxs)) : 0.f;
              sum4[j & 3] += p[j];
            }
          }
          const int c0 = (c + u) * 32;
          uint8_t* tile = smem + (c0 >> 6) * TILE_BYTES;
#pragma unroll
          for (int q = 0; q < 4; ++q) st_piece(tile, r, ((c0 & 63) >> 3) + q, p + 8 * q);
        }
      }
    }
    const float sum = (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(bar_p);
    mbar_wait(bar_o, 0);
    tc_fence_after();
    float o[64];
    tmem_ld32(trow, raw[0]);
    tmem_ld32(trow + 32, raw[1]);
    tmem_ld_wait_on(raw[0]);
    tmem_ld_wait_on(raw[1]);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(raw[0][j]) * inv;
#pragma unroll
    for (int j = 0; j < 32; ++j) o[32 + j] = __uint_as_float(raw[1][j]) * inv;
    if (qi < seq) {
      store_row64_bf16(out + ((long long)s_idx * seq + qi) * C + h * 64, o);
      if (lse != nullptr) lse[(long long)pair * seq + qi] = mx * scale + __logf(sum);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<256>(tmem);
}

// Coalesced store of a [128 rows][64 bf16] result tile: each thread drops its 32 fp32 values (row r, dims [32*half, +32))
// into a swizzled staging tile, then the 8 warps write whole 128-byte rows (8 lanes per row).  `rows_valid` rows exist.
__device__ __forceinline__ void store_tile_coalesced(uint8_t* stage, const float (&v)[32], int r, int half, int warp,
                                                     int lane, __nv_bfloat16* gbase, long long gpitch, int rows_valid) {
#pragma unroll
  for (int q = 0; q < 4; ++q) st_piece(stage, r, half * 4 + q, v + 8 * q);
  asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = warp * 16 + i * 4 + (lane >> 3), piece = lane & 7;
    if (row < rows_valid)
      *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(gbase + row * gpitch) + piece * 16) =
          *reinterpret_cast<const uint4*>(stage + swz(row, piece));
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");   // staging tile reusable
}

// development trace (see attention_sp.cu): the LAST CTA (the first problem) stamps clock64() into tr[(warp * 16 + it) * 8 + slot]
__device__ __forceinline__ void btrace(long long* tr, int warp, int it, int slot) {
  if (tr != nullptr && blockIdx.x == gridDim.x - 1 && (threadIdx.x & 31) == 0) tr[(warp * 16 + it) * 8 + slot] = clock64();
}

// =====================================================================================================
// backward: ONE persistent CTA per SM streams (sequence, head) problems; Q, K, V, dO, O resident (two 128-row tiles each),
// loop over 128-key tiles kt and 128-query tiles mt:   S = Q K^T, dP = dO V^T  ->  P = exp(S*scale - lse),
// dS = P*(dP - delta)*scale  ->  dQ_mt += dS K,  dV_kt += P^T dO,  dK_kt += dS^T Q.
// TMEM: S 0..127 | dP 128..255 | dK 256..319 | dV 320..383 | dQ_0 384..447 | dQ_1 448..511.
// 288 threads: warps 0-7 = softmax-gradient threads (TMEM quarter = warp % 4, column half = warp / 4),
// warp 8 = TMA + MMA issuer (the whole warp runs the issue loop converged, one elected lane issues).  The S/dP MMAs of
// iteration it+1 are issued before the dQ/dV/dK MMAs of iteration it, so the threads' TMEM loads and exp/FMA work overlap
// the tensor-core work of the previous tile pair.  As soon as the last MMAs of a problem have completed, the issuer
// refills the operand buffers with the NEXT problem (O has its own buffer; P / dS double as the output staging tiles), so
// the ~7500 cycles of TMA latency + delta prologue that a one-problem CTA exposed (measured with PVRL_SP_TRACE) run under
// the dK / dV / dQ stores of the previous problem.
// =====================================================================================================
constexpr int BWD_THREADS = 288;
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                   const __grid_constant__ CUtensorMap tmO, const float* __restrict__ lse,
                   __nv_bfloat16* __restrict__ dqkv, int seq, int H, float scale, int total, long long* __restrict__ tr) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t sQ = base, sK = base + 2 * TILE_BYTES, sV = base + 4 * TILE_BYTES, sdO = base + 6 * TILE_BYTES;
  const uint32_t sP = base + 8 * TILE_BYTES, sdS = base + 10 * TILE_BYTES, sO = base + 12 * TILE_BYTES;
  uint8_t* pP = smem + 8 * TILE_BYTES;
  uint8_t* pdS = smem + 10 * TILE_BYTES;
  const uint8_t* pO = smem + 12 * TILE_BYTES;
  const uint32_t bars = base + 14 * TILE_BYTES;
  const uint32_t bar_load0 = bars, bar_load1 = bars + 8, bar_sdp = bars + 16, bar_pds = bars + 24, bar_mma2 = bars + 32;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + 14 * TILE_BYTES + 48);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = H * 64;
  const int n_tiles = (seq + 127) / 128;   // 1 or 2 (seq <= 256)
  const int n_iter = n_tiles * n_tiles;
  const int G = gridDim.x;
  const int n_my = (total - static_cast<int>(blockIdx.x) + G - 1) / G;

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tmQKV);
      tma_prefetch_desc(&tmDO);
      tma_prefetch_desc(&tmO);
      mbar_init(bar_load0, 1);
      mbar_init(bar_load1, 1);
      mbar_init(bar_sdp, 1);
      mbar_init(bar_pds, 256);
      mbar_init(bar_mma2, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<512>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (PDL_EARLY_TRIGGER) pdl_launch_dependents();
  pdl_wait();   // barrier init / TMEM allocation above overlapped the previous kernel's tail
  constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DK = 256, COL_DV = 320, COL_DQ = 384;

  if (warp == 8) {
    if (tmem != 0) __trap();          // a 512-column allocation is all of TMEM: base column 0, lane 0
    // tile 0 of every operand first (all the first iteration needs), tile 1 behind it
    auto load_problem = [&](int j) {
      if (lane == 0) {
        const int pair = total - 1 - (static_cast<int>(blockIdx.x) + j * G);   // last sequences first (dO was just written)
        const int s_idx = pair / H, h = pair - s_idx * H;
        mbar_expect_tx(bar_load0, 5 * TILE_BYTES);
        tma_load_3d(sO, &tmO, bar_load0, h * 64, 0, s_idx);
        tma_load_3d(sdO, &tmDO, bar_load0, h * 64, 0, s_idx);
        tma_load_3d(sQ, &tmQKV, bar_load0, h * 64, 0, s_idx);
        tma_load_3d(sK, &tmQKV, bar_load0, C + h * 64, 0, s_idx);
        tma_load_3d(sV, &tmQKV, bar_load0, 2 * C + h * 64, 0, s_idx);
        if (n_tiles > 1) {
          mbar_expect_tx(bar_load1, 5 * TILE_BYTES);
          tma_load_3d(sO + TILE_BYTES, &tmO, bar_load1, h * 64, 128, s_idx);
          tma_load_3d(sdO + TILE_BYTES, &tmDO, bar_load1, h * 64, 128, s_idx);
          tma_load_3d(sQ + TILE_BYTES, &tmQKV, bar_load1, h * 64, 128, s_idx);
          tma_load_3d(sK + TILE_BYTES, &tmQKV, bar_load1, C + h * 64, 128, s_idx);
          tma_load_3d(sV + TILE_BYTES, &tmQKV, bar_load1, 2 * C + h * 64, 128, s_idx);
        }
      }
      __syncwarp();
    };
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_dq = make_idesc_bf16(128, 64, 0, 1);
    constexpr uint32_t idesc_dkv = make_idesc_bf16(128, 64, 1, 1);
    auto issue_sdp = [&](int kt, int mt) {
      const uint32_t q_t = sQ + mt * TILE_BYTES, do_t = sdO + mt * TILE_BYTES;
      const uint32_t k_t = sK + kt * TILE_BYTES, v_t = sV + kt * TILE_BYTES;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16_e(COL_S, make_smem_desc(q_t + k * 32, 16, 1024), make_smem_desc(k_t + k * 32, 16, 1024), idesc_s, k > 0);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16_e(COL_DP, make_smem_desc(do_t + k * 32, 16, 1024), make_smem_desc(v_t + k * 32, 16, 1024), idesc_s, k > 0);
      umma_commit_e(bar_sdp);
    };
    load_problem(0);
    for (int j = 0; j < n_my; ++j) {
      const int g0 = j * n_iter;       // running tile-pair count: phase parities of bar_sdp / bar_pds / bar_mma2
      mbar_wait(bar_load0, j & 1);
      __syncwarp();
      tc_fence_after();
      btrace(tr, 8, 15, 1);
      issue_sdp(0, 0);
      for (int it = 0; it < n_iter; ++it) {
        const int kt = it / n_tiles, mt = it % n_tiles;
        const uint32_t q_t = sQ + mt * TILE_BYTES, do_t = sdO + mt * TILE_BYTES, k_t = sK + kt * TILE_BYTES;
        mbar_wait(bar_pds, (g0 + it) & 1);   // P / dS of this iteration are in smem, S / dP columns are free again
        __syncwarp();
        tc_fence_after();
        btrace(tr, 8, it, 0);
        if (it + 1 < n_iter) {
          if (it == 0) {                 // first touch of the second tiles
            mbar_wait(bar_load1, j & 1);
            __syncwarp();
            tc_fence_after();
          }
          issue_sdp((it + 1) / n_tiles, (it + 1) % n_tiles);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)   // dQ_mt += dS[q, keys] K[keys, d]
          umma_bf16_e(COL_DQ + mt * 64, make_smem_desc(sdS + (k >> 2) * TILE_BYTES + (k & 3) * 32, 16, 1024),
                      make_smem_desc(k_t + k * 2048, TILE_BYTES, 1024), idesc_dq, (kt > 0 || k > 0));
#pragma unroll
        for (int k = 0; k < 8; ++k)   // dV_kt += P^T[keys, q] dO[q, d]
          umma_bf16_e(COL_DV, make_smem_desc(sP + k * 2048, TILE_BYTES, 1024),
                      make_smem_desc(do_t + k * 2048, TILE_BYTES, 1024), idesc_dkv, (mt > 0 || k > 0));
#pragma unroll
        for (int k = 0; k < 8; ++k)   // dK_kt += dS^T[keys, q] Q[q, d]
          umma_bf16_e(COL_DK, make_smem_desc(sdS + k * 2048, TILE_BYTES, 1024),
                      make_smem_desc(q_t + k * 2048, TILE_BYTES, 1024), idesc_dkv, (mt > 0 || k > 0));
        umma_commit_e(bar_mma2);
        btrace(tr, 8, it, 1);
      }
      if (j + 1 < n_my) {
        // every MMA of this problem has read its operands: refill the buffers while the threads store dK / dV / dQ
        mbar_wait(bar_mma2, (g0 + n_iter - 1) & 1);
        __syncwarp();
        load_problem(j + 1);
      }
    }
  } else {
    const int quarter = warp & 3, half = warp >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t trow = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    const long long pitch = 3LL * C;
    const uint8_t* pdO = smem + 6 * TILE_BYTES;
    const float sl2 = scale * LOG2E;
    for (int j = 0; j < n_my; ++j) {
      const int pair = total - 1 - (static_cast<int>(blockIdx.x) + j * G);
      const int s_idx = pair / H, h = pair - s_idx * H;
      const int g0 = j * n_iter;
      // per-row constants for both query tiles: lse and delta = dO . O (both operands from shared memory)
      float lse_l2[2], delta[2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int qi = mt * 128 + r;
        lse_l2[mt] = 0.f, delta[mt] = 0.f;
        if (mt < n_tiles) {
          mbar_wait(mt == 0 ? bar_load0 : bar_load1, j & 1);
          float acc = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint4 a = *reinterpret_cast<const uint4*>(pO + mt * TILE_BYTES + swz(r, i));
            const uint4 b = *reinterpret_cast<const uint4*>(pdO + mt * TILE_BYTES + swz(r, i));
            const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
            const float2 b0 = unpack_bf16x2(b.x), b1 = unpack_bf16x2(b.y), b2 = unpack_bf16x2(b.z), b3 = unpack_bf16x2(b.w);
            acc += a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2.x * b2.x + a2.y * b2.y + a3.x * b3.x +
                   a3.y * b3.y;
          }
          delta[mt] = acc;                       // rows past seq are zero-filled -> delta = 0
          if (qi < seq) lse_l2[mt] = lse[(long long)pair * seq + qi] * LOG2E;
        }
      }
      btrace(tr, warp, 15, 1);
      for (int it = 0; it < n_iter; ++it) {
        const int kt = it / n_tiles, mt = it % n_tiles;
        btrace(tr, warp, it, 0);
        mbar_wait(bar_sdp, (g0 + it) & 1);
        tc_fence_after();
        btrace(tr, warp, it, 1);
        // No per-element masks: padded query rows have Q = dO = 0 (TMA zero fill) and lse = delta = 0, so P = 1 but
        // dS = 0 and P^T dO = 0; padded key columns have K = V = 0, so their dS meets K = 0 in dQ and their dK / dV rows
        // are never stored.
        const float l2 = lse_l2[mt], dls = delta[mt] * scale;
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          const int c0 = half * 64 + cc * 32;
          uint32_t rs[32], rp[32];
          tmem_ld32(trow + COL_S + c0, rs);
          tmem_ld32(trow + COL_DP + c0, rp);
          tmem_ld_wait_on(rs);
          tmem_ld_wait_on(rp);
          float p[32], ds[32];
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            p[jj] = ex2_approx(fmaf(__uint_as_float(rs[jj]), sl2, -l2));
            ds[jj] = p[jj] * fmaf(__uint_as_float(rp[jj]), scale, -dls);
          }
          // previous P / dS tiles consumed by the MMAs (the first tile pair of the kernel has no predecessor; the first
          // pair of a later problem waits for the last pair of the previous problem, whose stores used P as staging)
          if (cc == 0 && g0 + it > 0) mbar_wait(bar_mma2, (g0 + it - 1) & 1);
          const int tile = c0 >> 6, piece0 = (c0 & 63) >> 3;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            st_piece(pP + tile * TILE_BYTES, r, piece0 + q, p + 8 * q);
            st_piece(pdS + tile * TILE_BYTES, r, piece0 + q, ds + 8 * q);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(bar_pds);
        btrace(tr, warp, it, 2);
        if (mt == n_tiles - 1) {   // key tile finished: dK_kt / dV_kt, this thread = key kt*128 + r, 32 of the 64 dims
          mbar_wait(bar_mma2, (g0 + it) & 1);   // all MMAs of this iteration are done: the P / dS tiles are free as staging
          tc_fence_after();
          const int rows_valid = min(128, seq - kt * 128);
          __nv_bfloat16* gk = dqkv + ((long long)s_idx * seq + kt * 128) * pitch + C + h * 64;
          uint32_t raw[32];
          float v[32];
          tmem_ld32(trow + COL_DK + half * 32, raw);
          tmem_ld_wait();
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) v[jj] = __uint_as_float(raw[jj]);
          store_tile_coalesced(pP, v, r, half, warp, lane, gk, pitch, rows_valid);
          tmem_ld32(trow + COL_DV + half * 32, raw);
          tmem_ld_wait();
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) v[jj] = __uint_as_float(raw[jj]);
          store_tile_coalesced(pP, v, r, half, warp, lane, gk + C, pitch, rows_valid);
          tc_fence_before();
          btrace(tr, warp, it, 3);
        }
      }
      // dQ tiles (the last bar_mma2 wait above covers every MMA)
      for (int mt = 0; mt < n_tiles; ++mt) {
        uint32_t raw[32];
        float v[32];
        tmem_ld32(trow + COL_DQ + mt * 64 + half * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) v[jj] = __uint_as_float(raw[jj]);
        store_tile_coalesced(pP, v, r, half, warp, lane, dqkv + ((long long)s_idx * seq + mt * 128) * pitch + h * 64,
                             pitch, min(128, seq - mt * 128));
      }
      tc_fence_before();     // the next problem's first MMAs overwrite the accumulators these loads just read
      btrace(tr, warp, 15, 2);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<512>(tmem);
}

}  // namespace

// 3-D bf16 tensor map over a row-major [n_seq, seq, cols] view, box [1, box_rows, box_cols], zero OOB fill.
int make_tmap_3d_bf16_box(CUtensorMap* out, const void* ptr, uint64_t cols, uint64_t seq, uint64_t n_seq, uint32_t box_cols,
                          uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) return fail(-2, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15u) != 0 || (cols * 2) % 16 != 0)
    return fail(-1, "attention operand must be 16-byte aligned with a row pitch multiple of 8 elements");
  if (box_cols != 64 && box_cols != 32) return fail(-1, "attention tensor map: box of 32 or 64 columns expected");
  cuuint64_t dims[3] = {cols, seq, n_seq};
  cuuint64_t strides[2] = {cols * 2, seq * cols * 2};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(-3, "cuTensorMapEncodeTiled (3-D) failed with CUresult %d", static_cast<int>(r));
  return 0;
}

// 3-D bf16 tensor map over a row-major [n_seq, seq, cols] view, box [1, box_rows, 64], 128B swizzle, zero OOB fill.
int make_tmap_3d_bf16(CUtensorMap* out, const void* ptr, uint64_t cols, uint64_t seq, uint64_t n_seq, uint32_t box_rows) {
  return make_tmap_3d_bf16_box(out, ptr, cols, seq, n_seq, 64, box_rows);
}

long long* sp_trace_buffer();
int attn_sp_fwd_launch(const void* qkv, void* out, float* lse, int n_seq, int seq, int H, float scale, cudaStream_t stream);

namespace {
inline int make_tmap_3d(CUtensorMap* out, const void* ptr, uint64_t cols, uint64_t seq, uint64_t n_seq, uint32_t box_rows) {
  return make_tmap_3d_bf16(out, ptr, cols, seq, n_seq, box_rows);
}
// PVRL_ATTN_SP=0 selects the first-generation one-shot kernels (kept for A/B runs)
inline bool use_sp() {
  static const bool on = [] {
    const char* e = getenv("PVRL_ATTN_SP");
    return e == nullptr || atoi(e) != 0;
  }();
  return on;
}
}  // namespace
}  // namespace pvrl

using namespace pvrl;

extern "C" int pvrl_attn_tc_fwd(const void* qkv, void* out, float* lse, int32_t n_seq, int32_t seq, int32_t H,
                                float scale, void* stream) {
  PVRL_CHECK_ARG(qkv && out && n_seq > 0 && H > 0, "pvrl_attn_tc_fwd: bad arguments");
  PVRL_CHECK_ARG(seq > 0 && seq <= 256, "pvrl_attn_tc_fwd: seq=%d must be in [1, 256]", seq);

  if (use_sp()) return attn_sp_fwd_launch(qkv, out, lse, n_seq, seq, H, scale, static_cast<cudaStream_t>(stream));
  const int npad = ((seq + 31) / 32) * 32;
  CUtensorMap tq, tkv;
  int rc;
  if ((rc = make_tmap_3d(&tq, qkv, 3ull * H * 64, seq, n_seq, 128))) return rc;
  if ((rc = make_tmap_3d(&tkv, qkv, 3ull * H * 64, seq, n_seq, npad))) return rc;
  const size_t smem = 4 * TILE_BYTES + 256 * 128 + 64 + 1024;
  static bool configured = false;
  if (!configured) {
    PVRL_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  dim3 grid((seq + 127) / 128, n_seq * H);
  PVRL_CUDA(launch_pdl(attn_tc_fwd_kernel, grid, dim3(160), smem, static_cast<cudaStream_t>(stream), tq, tkv,
                       static_cast<__nv_bfloat16*>(out), lse, seq, H, scale, npad));
  return launched("attn_tc_fwd_kernel");
}

extern "C" int pvrl_attn_tc_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                                int32_t n_seq, int32_t seq, int32_t H, float scale, void* stream) {
  PVRL_CHECK_ARG(qkv && out && dout && lse && dqkv && n_seq > 0 && H > 0, "pvrl_attn_tc_bwd: bad arguments");
  PVRL_CHECK_ARG(seq > 0 && seq <= 256, "pvrl_attn_tc_bwd: seq=%d must be in [1, 256]", seq);
  CUtensorMap tqkv, tdo, to;
  int rc;
  if ((rc = make_tmap_3d(&tqkv, qkv, 3ull * H * 64, seq, n_seq, 128))) return rc;
  if ((rc = make_tmap_3d(&tdo, dout, 1ull * H * 64, seq, n_seq, 128))) return rc;
  if ((rc = make_tmap_3d(&to, out, 1ull * H * 64, seq, n_seq, 128))) return rc;
  const size_t smem = 14 * TILE_BYTES + 64 + 1024;
  static bool configured = false;
  if (!configured) {
    PVRL_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int total = n_seq * H;
  const int grid = total < num_sms() ? total : num_sms();
  PVRL_CUDA(launch_pdl(attn_tc_bwd_kernel, dim3(grid), dim3(BWD_THREADS), smem, static_cast<cudaStream_t>(stream), tqkv,
                       tdo, to, lse, static_cast<__nv_bfloat16*>(dqkv), seq, H, scale, total, sp_trace_buffer()));
  return launched("attn_tc_bwd_kernel");
}
