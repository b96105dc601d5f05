// Pieces shared by the 1-CTA (gemm_sm100.cu) and the 2-CTA (gemm2_sm100.cu) tcgen05 GEMM kernels: the argument block and
// the fused epilogue (TMEM -> registers -> shared-memory transpose -> coalesced global access).
#pragma once
#include "pvrl_host.h"
#include "pvrl_ptx.cuh"

namespace pvrl {
namespace {

constexpr int BM = 128, BK = 64;
constexpr int A_BYTES = BM * BK * 2;          // 16 KB
constexpr int CHUNK_BYTES = 64 * BK * 2;      // one 64-wide MN chunk of a TN tile (8 KB)
constexpr int NUM_THREADS = 384;
constexpr int NUM_EPI_WARPS = 8;
constexpr int EPI_STAGE_BYTES = 32 * 128;     // per epilogue warp: 32 rows x 32 fp32 columns, 16 B pieces xor-swizzled
// Internal epilogue id (not part of the ABI): PVRL_EPI_RESID with the pos / time embedding adds of the patch-embedding
// GEMM.  A separate instantiation keeps their index arithmetic (two integer divisions per row) out of the code of the 48
// residual GEMMs per step that never take it: that epilogue is latency- and instruction-cache-bound, not a place for
// dead branches (ncu: 4000 SASS instructions per tile and warp before the split, 11 % no-instruction stalls).
constexpr int EPI_RESID_POS = 100;
struct FalseT { static constexpr bool value = false; };
struct TrueT { static constexpr bool value = true; };
template <int EPI>
__host__ __device__ constexpr bool is_resid() { return EPI == PVRL_EPI_RESID || EPI == EPI_RESID_POS; }
struct GemmArgs {
  int M, N, K;
  int k_splits, kb_per_split;
  void* out;
  long long ldo;
  void* out2;
  const float* bias;
  const float* rowscale;
  int rs_div;
  int map;
  const void* aux;
  long long ld_aux;
  const float* resid;
  const float* add_pos;
  const float* add_time;
  float* colsum;
  Geom g;
};

// ---- epilogue helpers --------------------------------------------------------------------------------
// The accumulator leaves TMEM one row per thread (tcgen05.ld 32x32b); a row-per-thread global access would touch 32
// different 128-byte lines per instruction, so every 32x32 fp32 block is transposed through a 4 KB per-warp staging
// buffer first: afterwards a lane owns 4 consecutive columns of one row, 8 lanes cover a full 128-byte line and every
// global load / store / red of the epilogue (output, residual, DGELU pre-activations) is coalesced.
template <typename T>
__device__ __forceinline__ void st_vec4(T* dst, const float4& v);
template <>
__device__ __forceinline__ void st_vec4<float>(float* dst, const float4& v) {
  *reinterpret_cast<float4*>(dst) = v;
}
template <>
__device__ __forceinline__ void st_vec4<__nv_bfloat16>(__nv_bfloat16* dst, const float4& v) {
  uint2 u;
  u.x = pack_bf16x2(v.x, v.y);
  u.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(dst) = u;
}
template <typename T>
__device__ __forceinline__ float4 ld_vec4(const T* src);
template <>
__device__ __forceinline__ float4 ld_vec4<float>(const float* src) {
  return __ldg(reinterpret_cast<const float4*>(src));
}
template <>
__device__ __forceinline__ float4 ld_vec4<__nv_bfloat16>(const __nv_bfloat16* src) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(src));
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void add4(float4& a, const float4& b) { a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w; }
__device__ __forceinline__ void mul4(float4& a, float s) { a.x *= s, a.y *= s, a.z *= s, a.w *= s; }

// what the transposed epilogue needs to know about one output row (computed once per tile by the lane that owns the
// row in TMEM order, then handed to the lanes that own it in the transposed order by warp shuffles)
struct RowCtx {
  int orow;         // residual-stream / output row (map_row); < 0 = cls row of a spatial sequence
  float rs;         // DropPath factor of the row (1 when none)
  // Addresses of this lane's 4 columns of the row in chunk 0 of the tile; chunk c adds the constant 32 * c elements, so
  // the unrolled epilogue addresses every access as pointer + immediate (no per-store 64-bit arithmetic).
  uint8_t* o1;        // out (or, for the cls rows of a spatial sequence, the side buffer out2)
  uint8_t* o2;        // GELU: out2
  const uint8_t* sd;  // RESID: residual, DGELU: saved gelu'
};

// "side" operand of one 4-column piece: the fp32 residual (RESID) or the saved GELU derivative (DGELU).  It does not
// depend on the accumulator, so the epilogue fetches it one 32-column chunk ahead (see the kernel) and the HBM latency
// of these loads overlaps the MMAs / the previous chunk instead of sitting between the TMEM read and the store.
template <int EPI, typename OutT>
__device__ __forceinline__ float4 load_side(const GemmArgs& p, int m, const RowCtx& rc, int coff) {
  if (is_resid<EPI>()) {   // rc.sd == nullptr: no residual operand, or a cls row of a spatial sequence (parked, not added)
    if (rc.sd != nullptr && m < p.M) return ld_vec4<float>(reinterpret_cast<const float*>(rc.sd) + coff);
  } else if (EPI == PVRL_EPI_DGELU) {
    if (m < p.M) return ld_vec4<OutT>(reinterpret_cast<const OutT*>(rc.sd) + coff);
  }
  return make_float4(0.f, 0.f, 0.f, 0.f);
}

template <int EPI, typename OutT>
__device__ __forceinline__ void epilogue_vec4(const GemmArgs& p, int m, const RowCtx& rc, int coff, int n, float4& v,
                                              const float4& bias4, const float4& side) {
  if (EPI == PVRL_EPI_ATOMIC) {
    float* dst = reinterpret_cast<float*>(rc.o1) + coff;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
    return;
  }
  add4(v, bias4);
  if (EPI == PVRL_EPI_GELU) {   // out2 = gelu(z), out = gelu'(z): the backward GEMM then only multiplies (EPI_DGELU)
    float4 a, d;
    gelu_both<OutT>(v.x, a.x, d.x), gelu_both<OutT>(v.y, a.y, d.y), gelu_both<OutT>(v.z, a.z, d.z),
        gelu_both<OutT>(v.w, a.w, d.w);
    st_vec4<OutT>(reinterpret_cast<OutT*>(rc.o1) + coff, d);
    st_vec4<OutT>(reinterpret_cast<OutT*>(rc.o2) + coff, a);
    return;
  }
  if (EPI == PVRL_EPI_DGELU) v.x *= side.x, v.y *= side.y, v.z *= side.z, v.w *= side.w;
  mul4(v, rc.rs);
  if (is_resid<EPI>()) {
    // cls rows of a spatial sequence are parked in the side buffer for the mean over frames (vit.py:147-149): their o1
    // points into out2 and their `side` is zero (load_side), so they take the same straight-line code as every other row
    add4(v, side);
    if (EPI == EPI_RESID_POS) {  // MAP_PATCH (patch embedding only): + pos_embed[1+n] + time_embed[t]  (vit.py:373-404)
      const int bt = m / p.g.HW;
      add4(v, ld_vec4<float>(p.add_pos + (long long)(1 + m - bt * p.g.HW) * p.N + n));
      add4(v, ld_vec4<float>(p.add_time + (long long)(bt % p.g.T) * p.N + n));
    }
    st_vec4<float>(reinterpret_cast<float*>(rc.o1) + coff, v);
    return;
  }
  // STORE / DGELU
  st_vec4<OutT>(reinterpret_cast<OutT*>(rc.o1) + coff, v);
}

// pvrl_gemm_t -> kernel argument block (the split-K fields are filled in by the launcher)
inline GemmArgs make_gemm_args(const pvrl_gemm_t* d) {
  GemmArgs a;
  a.M = d->M, a.N = d->N, a.K = d->K;
  a.k_splits = 1, a.kb_per_split = 0;
  a.out = d->out, a.ldo = d->ldo, a.out2 = d->out2;
  a.bias = d->bias, a.rowscale = d->rowscale, a.rs_div = d->rs_div > 0 ? d->rs_div : 1;
  a.map = d->map, a.aux = d->aux, a.ld_aux = d->ld_aux, a.resid = d->resid;
  a.add_pos = d->add_pos, a.add_time = d->add_time;
  a.colsum = d->colsum;
  a.g = Geom(d->g.T > 0 ? d->g.T : 1, d->g.HW > 0 ? d->g.HW : 1);
  return a;
}

// tile index -> (m block, n block, k split).  NT: n fastest, so the CTAs of a wave share A tiles through L2 and the
// (small) weight matrix stays resident.  TN (dW, split-K): the k split is the SLOWEST index, so at any time the
// resident CTAs cover all output tiles of one or two contraction slices and every slice of dY / X is fetched from
// HBM once instead of once per output tile.
template <bool TN>
__device__ __forceinline__ void decode_tile(int tile, int m_tiles, int n_tiles, int k_splits, int& m_blk, int& n_blk,
                                            int& ks) {
  if (TN) {
    const int per = m_tiles * n_tiles;
    ks = tile / per;
    const int rest = tile - ks * per;
    n_blk = rest % n_tiles, m_blk = rest / n_tiles;
  } else {
    ks = tile % k_splits;
    const int rest = tile / k_splits;
    n_blk = rest % n_tiles, m_blk = rest / n_tiles;
  }
}


// The whole epilogue of one accumulator tile for one epilogue warp: rows [m_base, m_base + 32) (the warp's TMEM lane
// quarter), columns [n_base, n_base + HALF_COLS) of the output (TMEM columns tmem_cols ...).  Waits for the
// accumulator on `tfull` AFTER the first side-operand loads are in flight.
template <int EPI, typename OutT, int HALF_COLS>
__device__ __forceinline__ void epilogue_tile(const GemmArgs& p, uint8_t* stg, uint32_t tmem_cols, int m_base, int n_base,
                                              uint32_t tfull, uint32_t tfull_phase, int lane) {
  const int rsub = lane >> 3, piece = lane & 7;   // transposed ownership: row 4*i + rsub, columns [4*piece, +4)
  // per-row context, computed by the lane that owns the row in TMEM order ...
  RowCtx own;
  {
    const int m = min(m_base + lane, p.M - 1);
    own.orow = (EPI == PVRL_EPI_ATOMIC || EPI == PVRL_EPI_GELU) ? m : static_cast<int>(map_row(p.map, m, p.g));
    own.rs = (EPI != PVRL_EPI_ATOMIC && EPI != PVRL_EPI_GELU && p.rowscale != nullptr)
                 ? __ldg(p.rowscale + m / p.rs_div) : 1.0f;
  }
  // ... and handed to the transposed owners
  RowCtx rc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int src = 4 * i + rsub;
    rc[i].orow = __shfl_sync(0xffffffffu, own.orow, src);
    rc[i].rs = (EPI != PVRL_EPI_ATOMIC && EPI != PVRL_EPI_GELU) ? __shfl_sync(0xffffffffu, own.rs, src) : 1.0f;
    constexpr int OSZ = (is_resid<EPI>() || EPI == PVRL_EPI_ATOMIC) ? 4 : static_cast<int>(sizeof(OutT));
    const long long oel = (long long)(rc[i].orow >= 0 ? rc[i].orow : -rc[i].orow - 1) * p.ldo + n_base + piece * 4;
    rc[i].o1 = reinterpret_cast<uint8_t*>((is_resid<EPI>() && rc[i].orow < 0) ? p.out2 : p.out) + oel * OSZ;
    rc[i].o2 = EPI == PVRL_EPI_GELU ? reinterpret_cast<uint8_t*>(p.out2) + oel * OSZ : nullptr;
    rc[i].sd = nullptr;
    if (is_resid<EPI>() && p.resid != nullptr && rc[i].orow >= 0)
      rc[i].sd = reinterpret_cast<const uint8_t*>(p.resid) + oel * 4;
    if (EPI == PVRL_EPI_DGELU)
      rc[i].sd = reinterpret_cast<const uint8_t*>(p.aux) +
                 ((long long)(m_base + src) * p.ld_aux + n_base + piece * 4) * static_cast<int>(sizeof(OutT));
  }
  constexpr int NCH = HALF_COLS / 32;
  constexpr bool HAS_SIDE = is_resid<EPI>() || EPI == PVRL_EPI_DGELU;
  constexpr bool HAS_COLSUM = EPI == PVRL_EPI_STORE || EPI == PVRL_EPI_DGELU;
  // bf16 gelu' (DGELU): the side operand of the WHOLE tile is fetched packed (2 registers per piece) before the
  // accumulator is waited for -- one chunk of lookahead (~0.4 us of work) does not cover an HBM round trip.
  // fp32 residuals (RESID) / fp32 gelu' keep the two-deep chunk pipeline (a whole tile would need 128 registers).
  constexpr bool SIDE_PACKED = EPI == PVRL_EPI_DGELU && sizeof(OutT) == 2;
  const int n_first = n_base + piece * 4;   // this lane's columns in chunk 0
  float4 side[SIDE_PACKED ? 1 : 2][8];
  uint2 sidep[SIDE_PACKED ? NCH : 1][8];
  if (SIDE_PACKED) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int n = n_first + c * 32;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int m = m_base + 4 * i + rsub;
        sidep[c][i] = (m < p.M && n < p.N)
                          ? __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(rc[i].sd) + c * 32))
                          : make_uint2(0u, 0u);
      }
    }
  } else if (HAS_SIDE && n_first < p.N) {
#pragma unroll
    for (int i = 0; i < 8; ++i) side[0][i] = load_side<EPI, OutT>(p, m_base + 4 * i + rsub, rc[i], 0);
  }
  mbar_wait(tfull, tfull_phase);
  tc_fence_after();
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int n0 = n_base + c * 32;
    if (n0 < p.N) {  // warp-uniform
      if (HAS_SIDE && !SIDE_PACKED && c + 1 < NCH && n0 + 32 < p.N) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          side[(c + 1) & 1][i] = load_side<EPI, OutT>(p, m_base + 4 * i + rsub, rc[i], (c + 1) * 32);
      }
      uint32_t raw[32];
      tmem_ld32(tmem_cols + c * 32, raw);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) =
            make_uint4(raw[4 * j], raw[4 * j + 1], raw[4 * j + 2], raw[4 * j + 3]);
      __syncwarp();
      const int n = n0 + piece * 4;
      float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (EPI != PVRL_EPI_ATOMIC && p.bias != nullptr) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
      float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
      // All but the last row block of the matrix are full: their eight rows run as straight-line code (no per-row
      // bounds branch), so the shared-memory reads of all rows are in flight before the first row's arithmetic.
      auto rows = [&](auto checked) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = 4 * i + rsub;
          float4 v = *reinterpret_cast<const float4*>(stg + row * 128 + ((piece ^ (row & 7)) << 4));
          const int m = m_base + row;
          if (!decltype(checked)::value || m < p.M) {
            float4 sd = bias4;
            if (SIDE_PACKED) {
              const float2 lo = unpack_bf16x2(sidep[c][i].x), hi = unpack_bf16x2(sidep[c][i].y);
              sd = make_float4(lo.x, lo.y, hi.x, hi.y);
            } else if (HAS_SIDE) {
              sd = side[c & 1][i];
            }
            epilogue_vec4<EPI, OutT>(p, m, rc[i], c * 32, n, v, bias4, sd);
            if (HAS_COLSUM) add4(csum, v);
          }
        }
      };
      if (m_base + 32 <= p.M) rows(FalseT{});
      else rows(TrueT{});
      if (HAS_COLSUM && p.colsum != nullptr) {   // fused bias gradient: column sums of what was just stored
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
          csum.x += __shfl_xor_sync(0xffffffffu, csum.x, o), csum.y += __shfl_xor_sync(0xffffffffu, csum.y, o);
          csum.z += __shfl_xor_sync(0xffffffffu, csum.z, o), csum.w += __shfl_xor_sync(0xffffffffu, csum.w, o);
        }
        if (rsub == 0)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.colsum + n), "f"(csum.x), "f"(csum.y),
                       "f"(csum.z), "f"(csum.w)
                       : "memory");
      }
      __syncwarp();
    }
  }
}

}  // namespace
}  // namespace pvrl
