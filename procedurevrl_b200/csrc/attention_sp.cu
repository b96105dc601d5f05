// Spatial attention forward, second generation: ONE persistent CTA per SM that streams (frame, head) problems
// (seq = 1 + H*W = 197 tokens, head_dim 64) through a warp-specialised pipeline.  Attention.forward vit.py:84-88.
//
//   warp 8      TMA producer: Q (two 128-row tiles), K, V of problem i+1 land in the second smem stage while problem i
//               computes.  K / V are fetched ONCE per (frame, head) and serve both query tiles.
//   warps 9,10  tcgen05 issuers, one per query tile (independent blocking waits, no polling loop): S = Q K^T
//               (128 x npad x 64) as soon as the tile's TMEM region is free, O = P V as soon as its probabilities are
//               in TMEM.
//   warps 0-3   softmax + epilogue of query tile 0;  warps 4-7 the same for query tile 1.  One query row per thread:
//               row max and exp2 on tcgen05.ld'ed scores (two passes over TMEM), the bf16 probabilities go back INTO
//               TMEM over the dead scores (tcgen05.st) and feed the P V MMA as its A operand -- P never touches shared
//               memory.  The two groups run on different tiles, so one group's MUFU work overlaps the other group's
//               MMA / epilogue phases, and S of problem i+1 is issued while the other tile of problem i is in softmax.
//               O leaves through a per-warp swizzled staging tile and a TMA store (whole 128-byte rows, rows >= seq
//               clipped by the tensor map).
//
// TMEM (512 columns): tile t owns columns [256 t, 256 t + 256): S fp32 in [0, npad), P bf16x2 in [0, npad / 2)
// (thread-private row, written behind the read pointer), O fp32 in [128, 192) once S is dead.
// Padded keys: TMA zero-fills K / V rows >= seq, so their scores are 0 (masked out of max / sum) and their V rows are 0.
#include "pvrl_host.h"
#include "pvrl_ptx.cuh"

namespace pvrl {
constexpr int TR_WARPS = 20, TR_ITERS = 16, TR_SLOTS = 8;
namespace {

constexpr int SP_TILE_BYTES = 128 * 128;   // [128 rows][64 bf16], SWIZZLE_128B
constexpr int SP_THREADS = 352;
constexpr int SP_STAGE_OUT = 4096;         // per-warp output staging: [32 rows][128 B]
constexpr float SP_LOG2E = 1.4426950408889634f;

// Development aid (PVRL_SP_TRACE=1): CTA 0 records clock64() at the phase boundaries of its first 16 problems into a
// device buffer that pvrl_debug_sp_trace() copies out; a null pointer (the default) costs one predicate per phase.
__device__ __forceinline__ void trace(long long* tr, int warp, int it, int slot) {
  if (tr != nullptr && blockIdx.x == 0 && it < TR_ITERS && (threadIdx.x & 31) == 0)
    tr[(warp * TR_ITERS + it) * TR_SLOTS + slot] = clock64();
}

__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }   // one FMNMX3

// running max over a chunk of W score columns; `live` = number of real keys in the chunk (W when the chunk is full)
template <int W>
__device__ __forceinline__ void chunk_max(const uint32_t (&r)[W], int live, float (&mx)[4]) {
  if (live >= W) {
#pragma unroll
    for (int j = 0; j < W; j += 8) {
      mx[0] = max3(mx[0], __uint_as_float(r[j]), __uint_as_float(r[j + 1]));
      mx[1] = max3(mx[1], __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
      mx[2] = max3(mx[2], __uint_as_float(r[j + 4]), __uint_as_float(r[j + 5]));
      mx[3] = max3(mx[3], __uint_as_float(r[j + 6]), __uint_as_float(r[j + 7]));
    }
  } else {
#pragma unroll
    for (int j = 0; j < W; ++j)
      if (j < live) mx[j & 3] = fmaxf(mx[j & 3], __uint_as_float(r[j]));
  }
}

// p = 2^(s * sl2 - mxs) for a chunk of W columns, packed to bf16 pairs; padded keys get p = 0.  The scale-and-shift and
// the row-sum accumulation are packed fp32x2 instructions (one FFMA2 + one FADD2 per two columns).
template <int W>
__device__ __forceinline__ void chunk_exp(const uint32_t (&r)[W], int live, uint64_t sl2_2, uint64_t nmxs_2, uint64_t (&sum)[4],
                                          uint32_t (&pk)[W / 2]) {
#pragma unroll
  for (int j = 0; j < W; j += 2) {
    float x0, x1;
    f2_unpack(f2_fma(f2_pack(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), sl2_2, nmxs_2), x0, x1);
    float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
    if (live < W) {                      // only the last chunk of a row: padded keys contribute nothing
      p0 = j < live ? p0 : 0.f;
      p1 = j + 1 < live ? p1 : 0.f;
    }
    sum[(j >> 1) & 3] = f2_add(sum[(j >> 1) & 3], f2_pack(p0, p1));
    pk[j >> 1] = pack_bf16x2(p0, p1);
  }
}

// MMA issue loop of query tile T (warp 9 + T).  The whole warp runs it converged with warp-uniform operands (tile index a
// template constant, the 512-column TMEM allocation starts at column 0 / lane 0) so that the descriptors live in
// uniform registers and one elected lane issues (see umma_bf16_e).
template <int T>
__device__ __forceinline__ void sp_issue_loop(uint32_t base, uint32_t bars, int stage_bytes, int kv_bytes, int npad, int n_my,
                                              int n_tiles, long long* tr) {
  const uint32_t bar_full = bars, bar_empty = bars + 16, bar_s = bars + 32 + 8 * T, bar_p = bars + 48 + 8 * T,
                 bar_o = bars + 64 + 8 * T, bar_free = bars + 80 + 8 * T;
  const uint32_t idesc_s = make_idesc_bf16(128, npad, 0, 0);
  constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);
  constexpr uint32_t tcol = T * 256;
  const int nk = npad >> 4;
  for (int it = 0; it < n_my; ++it) {
    const int s = it & 1;
    const uint32_t sQ = base + s * stage_bytes + T * SP_TILE_BYTES, sK = base + s * stage_bytes + 2 * SP_TILE_BYTES;
    const uint32_t sV = sK + kv_bytes;
    mbar_wait(bar_free, (it & 1) ^ 1);                      // the tile's TMEM region has been drained (passes at it = 0)
    mbar_wait(bar_full + 8 * s, (it >> 1) & 1);
    __syncwarp();
    tc_fence_after();
    trace(tr, 9 + T, it, 0);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_bf16_e(tcol, make_smem_desc(sQ + k * 32, 16, 1024), make_smem_desc(sK + k * 32, 16, 1024), idesc_s, k > 0);
    umma_commit_e(bar_s);
    mbar_wait(bar_p, it & 1);
    __syncwarp();
    tc_fence_after();
    trace(tr, 9 + T, it, 1);
#pragma unroll 4
    for (int k = 0; k < nk; ++k)
      umma_bf16_ts_e(tcol + 128, tcol + k * 8, make_smem_desc(sV + k * 2048, 8192, 1024), idesc_o, k > 0);
    umma_commit_e(bar_o);
    umma_commit_e(bar_empty + 8 * s);                       // this tile is done with the stage (count = n_tiles)
  }
}

__global__ void __launch_bounds__(SP_THREADS, 1)
attn_sp_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                   const __grid_constant__ CUtensorMap tmO, float* __restrict__ lse, int seq, int H, float scale,
                   int npad, int total, int flags, long long* __restrict__ tr) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const int kv_bytes = npad * 128;
  const int stage_bytes = 2 * SP_TILE_BYTES + 2 * kv_bytes;
  const uint32_t out_off = 2 * stage_bytes;                 // 8 x 4 KB output staging
  const uint32_t bars = base + out_off + 8 * SP_STAGE_OUT;
  // barriers: full[2] empty[2] s_ready[2] p_ready[2] o_ready[2] t_free[2] turn[2]
  const uint32_t bar_full = bars, bar_empty = bars + 16, bar_s = bars + 32, bar_p = bars + 48, bar_o = bars + 64,
                 bar_free = bars + 80, bar_turn = bars + 96;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + out_off + 8 * SP_STAGE_OUT + 112);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (seq + 127) >> 7;
  const int C = H * 64;
  const int G = gridDim.x;
  const int n_my = (total - static_cast<int>(blockIdx.x) + G - 1) / G;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmO);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int rows = min(128, seq - i * 128);              // live query rows of tile i
      const int live_threads = rows > 0 ? ((rows + 31) >> 5) * 32 : 32;
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, n_tiles);
      mbar_init(bar_s + 8 * i, 1);
      mbar_init(bar_p + 8 * i, live_threads);
      mbar_init(bar_o + 8 * i, 1);
      mbar_init(bar_free + 8 * i, live_threads);
      // turn[i]: group i may start its exp pass; the OTHER group's live threads arrive when theirs is done
      const int other = min(128, seq - (i ^ 1) * 128);
      mbar_init(bar_turn + 8 * i, other > 0 ? ((other + 31) >> 5) * 32 : 32);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<512>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (PDL_EARLY_TRIGGER) pdl_launch_dependents();
  pdl_wait();

  if (warp == 8) {
    // ------------------------------------------------------------------------------- TMA producer
    if (lane == 0) {
      const uint32_t tx = n_tiles * SP_TILE_BYTES + 2 * kv_bytes;
      for (int it = 0; it < n_my; ++it) {
        const int s = it & 1;
        const int pair = total - 1 - (static_cast<int>(blockIdx.x) + it * G);   // freshest rows of the QKV GEMM first
        const int s_idx = pair / H, h = pair - s_idx * H;
        const uint32_t sQ = base + s * stage_bytes, sK = sQ + 2 * SP_TILE_BYTES, sV = sK + kv_bytes;
        mbar_wait(bar_empty + 8 * s, ((it >> 1) & 1) ^ 1);
        trace(tr, warp, it, 0);
        mbar_expect_tx(bar_full + 8 * s, tx);
        tma_load_3d(sQ, &tmQ, bar_full + 8 * s, h * 64, 0, s_idx);
        tma_load_3d(sK, &tmKV, bar_full + 8 * s, C + h * 64, 0, s_idx);
        if (n_tiles > 1) tma_load_3d(sQ + SP_TILE_BYTES, &tmQ, bar_full + 8 * s, h * 64, 128, s_idx);
        tma_load_3d(sV, &tmKV, bar_full + 8 * s, 2 * C + h * 64, 0, s_idx);
      }
    }
  } else if (warp >= 9) {
    // ------------------------------------------------------------------------------- MMA issuers: warp 9 -> tile 0, warp 10 -> tile 1
    if (tmem != 0) __trap();
    if (warp == 9) {
      sp_issue_loop<0>(base, bars, stage_bytes, kv_bytes, npad, n_my, n_tiles, tr);
    } else if (n_tiles > 1) {
      sp_issue_loop<1>(base, bars, stage_bytes, kv_bytes, npad, n_my, n_tiles, tr);
    }
  } else {
    // ------------------------------------------------------------------------------- softmax + epilogue
    const int t = warp >> 2, qd = warp & 3;
    const int row0 = t * 128 + qd * 32;
    if (t < n_tiles && row0 < seq) {
      const uint32_t tb = tmem + (static_cast<uint32_t>(qd * 32) << 16) + t * 256;
      const int qi = row0 + lane;
      uint8_t* stg = smem + out_off + warp * SP_STAGE_OUT;
      const uint32_t stg_u32 = base + out_off + warp * SP_STAGE_OUT;
      const int n32 = npad >> 5;
      const bool tail16 = (npad & 16) != 0;
      const float sl2 = scale * SP_LOG2E;
      const bool turns = n_tiles > 1 && (flags & 1);
      for (int it = 0; it < n_my; ++it) {
        const int pair = total - 1 - (static_cast<int>(blockIdx.x) + it * G);
        const int s_idx = pair / H, h = pair - s_idx * H;
        trace(tr, warp, it, 0);
        mbar_wait(bar_s + 8 * t, it & 1);
        tc_fence_after();
        trace(tr, warp, it, 1);
        // ---- pass 1: row max (TMEM reads software-pipelined: chunk c + 1 is in flight while chunk c is reduced)
        uint32_t ra[32], rb[32], rt[16];
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        tmem_ld32(tb, ra);
#pragma unroll 1
        for (int c = 0; c < n32; c += 2) {
          tmem_ld_wait_on(ra);
          if (c + 1 < n32) tmem_ld32(tb + (c + 1) * 32, rb);
          chunk_max<32>(ra, seq - c * 32, mx4);
          if (c + 1 < n32) {
            tmem_ld_wait_on(rb);
            if (c + 2 < n32) tmem_ld32(tb + (c + 2) * 32, ra);
            chunk_max<32>(rb, seq - (c + 1) * 32, mx4);
          }
        }
        if (tail16) {
          tmem_ld16(tb + n32 * 32, rt);
          tmem_ld_wait_on(rt);
          chunk_max<16>(rt, seq - n32 * 32, mx4);
        }
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        const float mxs = mx * sl2;
        trace(tr, warp, it, 2);
        // ---- pass 2: p = 2^(s * scale * log2e - max), bf16 pairs written back over the scores.
        // The exp pass is MUFU-bound and the two groups share the SM's MUFU pipes: taking turns (group 0, group 1,
        // group 0, ...) keeps them in anti-phase, so one group's exponentials run at the full MUFU rate while the
        // other group sits in its MMA / epilogue / row-max phases -- left alone they fall into lockstep.
        if (turns) mbar_wait(bar_turn + 8 * t, t == 0 ? (it & 1) ^ 1 : (it & 1));
        uint64_t sum4[4] = {0ull, 0ull, 0ull, 0ull};
        const uint64_t sl2_2 = f2_pack(sl2, sl2), nmxs_2 = f2_pack(-mxs, -mxs);
        uint32_t pk[16];
        tmem_ld32(tb, ra);
#pragma unroll 1
        for (int c = 0; c < n32; c += 2) {
          tmem_ld_wait_on(ra);
          if (c + 1 < n32) tmem_ld32(tb + (c + 1) * 32, rb);
          chunk_exp<32>(ra, seq - c * 32, sl2_2, nmxs_2, sum4, pk);
          tmem_st16(tb + c * 16, pk);
          if (c + 1 < n32) {
            tmem_ld_wait_on(rb);
            if (c + 2 < n32) tmem_ld32(tb + (c + 2) * 32, ra);
            chunk_exp<32>(rb, seq - (c + 1) * 32, sl2_2, nmxs_2, sum4, pk);
            tmem_st16(tb + (c + 1) * 16, pk);
          }
        }
        if (tail16) {
          uint32_t pk8[8];
          tmem_ld16(tb + n32 * 32, rt);
          tmem_ld_wait_on(rt);
          chunk_exp<16>(rt, seq - n32 * 32, sl2_2, nmxs_2, sum4, pk8);
          tmem_st8(tb + n32 * 16, pk8);
        }
        float sum_lo, sum_hi;
        f2_unpack(f2_add(f2_add(sum4[0], sum4[1]), f2_add(sum4[2], sum4[3])), sum_lo, sum_hi);
        const float sum = sum_lo + sum_hi;
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar_p + 8 * t);
        if (turns) mbar_arrive(bar_turn + 8 * (t ^ 1));
        trace(tr, warp, it, 3);
        // ---- epilogue: O / sum -> bf16 -> staging tile -> TMA store
        mbar_wait(bar_o + 8 * t, it & 1);
        tc_fence_after();
        trace(tr, warp, it, 4);
        tmem_ld32(tb + 128, ra);
        tmem_ld32(tb + 160, rb);
        tmem_ld_wait_on(ra);
        tmem_ld_wait_on(rb);
        tc_fence_before();
        mbar_arrive(bar_free + 8 * t);                       // the tile's TMEM region may take the next S
        trace(tr, warp, it, 5);
        if (lane == 0) bulk_wait_read0();                    // previous TMA store has finished reading the staging tile
        __syncwarp();
        const float inv = 1.0f / sum;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(ra[8 * i]) * inv, __uint_as_float(ra[8 * i + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(ra[8 * i + 2]) * inv, __uint_as_float(ra[8 * i + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(ra[8 * i + 4]) * inv, __uint_as_float(ra[8 * i + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(ra[8 * i + 6]) * inv, __uint_as_float(ra[8 * i + 7]) * inv);
          *reinterpret_cast<uint4*>(stg + lane * 128 + ((i ^ (lane & 7)) << 4)) = u;
          u.x = pack_bf16x2(__uint_as_float(rb[8 * i]) * inv, __uint_as_float(rb[8 * i + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(rb[8 * i + 2]) * inv, __uint_as_float(rb[8 * i + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(rb[8 * i + 4]) * inv, __uint_as_float(rb[8 * i + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(rb[8 * i + 6]) * inv, __uint_as_float(rb[8 * i + 7]) * inv);
          *reinterpret_cast<uint4*>(stg + lane * 128 + (((4 + i) ^ (lane & 7)) << 4)) = u;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&tmO, stg_u32, h * 64, row0, s_idx);
          bulk_commit();
        }
        if (lse != nullptr && qi < seq) lse[static_cast<long long>(pair) * seq + qi] = mx * scale + __logf(sum);
        trace(tr, warp, it, 6);
      }
      if (lane == 0) bulk_wait0();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem);
}

}  // namespace

int make_tmap_3d_bf16(CUtensorMap* out, const void* ptr, uint64_t cols, uint64_t seq, uint64_t n_seq, uint32_t box_rows);

// PVRL_SP_TRACE=1: device buffer the kernels' trace() calls write into (nullptr otherwise)
long long* sp_trace_buffer() {
  static long long* buf = [] {
    const char* e = getenv("PVRL_SP_TRACE");
    long long* p = nullptr;
    if (e != nullptr && atoi(e) != 0 && cudaMalloc(&p, sizeof(long long) * TR_WARPS * TR_ITERS * TR_SLOTS) == cudaSuccess)
      cudaMemset(p, 0, sizeof(long long) * TR_WARPS * TR_ITERS * TR_SLOTS);
    return p;
  }();
  return buf;
}

int attn_sp_fwd_launch(const void* qkv, void* out, float* lse, int n_seq, int seq, int H, float scale, cudaStream_t stream) {
  const int npad = ((seq + 15) / 16) * 16;
  CUtensorMap tq, tkv, to;
  int rc;
  if ((rc = make_tmap_3d_bf16(&tq, qkv, 3ull * H * 64, seq, n_seq, 128))) return rc;
  if ((rc = make_tmap_3d_bf16(&tkv, qkv, 3ull * H * 64, seq, n_seq, npad))) return rc;
  if ((rc = make_tmap_3d_bf16(&to, out, 1ull * H * 64, seq, n_seq, 32))) return rc;
  const size_t smem = 2 * (2 * SP_TILE_BYTES + 2 * npad * 128) + 8 * SP_STAGE_OUT + 128 + 1024;
  static const int flags = [] {
    const char* e = getenv("PVRL_SP_TURNS");
    return (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }();
  static size_t configured = 0;
  if (configured < smem) {
    PVRL_CUDA(cudaFuncSetAttribute(attn_sp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int total = n_seq * H;
  const int grid = total < num_sms() ? total : num_sms();
  PVRL_CUDA(launch_pdl(attn_sp_fwd_kernel, dim3(grid), dim3(SP_THREADS), smem, stream, tq, tkv, to, lse, seq, H, scale,
                       npad, total, flags, sp_trace_buffer()));
  return launched("attn_sp_fwd_kernel");
}

}  // namespace pvrl

// Development aid: copies the PVRL_SP_TRACE buffer ([20 warps][16 problems][8 slots] clock64 stamps of CTA 0) to `host_out`
// and clears it; returns the number of int64 values written, 0 when tracing is off.
extern "C" int pvrl_debug_sp_trace(long long* host_out) {
  long long* buf = pvrl::sp_trace_buffer();
  if (buf == nullptr || host_out == nullptr) return 0;
  const size_t n = pvrl::TR_WARPS * pvrl::TR_ITERS * pvrl::TR_SLOTS;
  if (cudaMemcpy(host_out, buf, n * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
  cudaMemset(buf, 0, n * sizeof(long long));
  return static_cast<int>(n);
}
