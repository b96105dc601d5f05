// Pooled attention forward of the MViTv2 path on the warp-level tensor-core MMA (mma.sync.m16n8k16, bf16 operands, fp32
// accumulate): same contract as pooled_attn_fwd_kernel of mvit.cu (include/pvrl.h, pvrl_pooled_attn_fwd) for bf16 I/O.
//
// A block of 4 warps owns 64 consecutive queries of one (clip, head); each warp 16 of them.  Keys / values stream through
// shared memory in chunks of 64 rows (row pitch 104 bf16 = 208 B: the 8 rows an MMA fragment or an ldmatrix phase touches
// fall into 8 disjoint bank groups).  Per chunk and warp:
//   S = Q K^T      48 MMAs: A = the warp's Q rows, loaded once from HBM straight into fragment registers (lane (g, t) holds
//                  dims 2t, 2t+1 (+8) of rows g, g+8 per 16-dim tile); B = K rows as they lie in shared memory
//                  (B[k = dim][n = key]: the pair (dims 2t, 2t+1) of key g is one 32-bit word);
//   scores         S * scale + bq[i, kt] + bq[i, Kt + kh] + bq[i, Kt + Kh + kw] in fp32 (component columns of the 64 keys
//                  decoded once per chunk into shared memory), keys beyond Nk masked;
//   softmax        online (running max / sum per row, quad shuffles), FlashAttention-2 style: the S accumulator layout
//                  (row g / g+8, keys 2t, 2t+1 per 8-key tile) IS the A-fragment layout of the next MMA, so P never
//                  leaves registers;
//   O += P V       48 MMAs: B = V through ldmatrix.x4.trans (B[k = key][n = dim] needs two keys per register).
// Epilogue: O / sum (+ q, residual pooling), written as [B, Nq, heads * 96]; lse = max + log(sum).
// The fragment algebra is checked lane by lane on the CPU by tests/test_mvit_mma_emulation.py (a numpy model of
// mma.sync / ldmatrix executing this file's index arithmetic) -- see DESIGN.md section 9.
#include <cuda_bf16.h>

#include <cfloat>

#include "pvrl_host.h"

namespace pvrl {
namespace {

typedef __nv_bfloat16 bf16;

constexpr int MM_C = 96;        // head width
constexpr int MM_KC = 64;       // keys per chunk
constexpr int MM_PITCH = 104;   // shared-memory row pitch in bf16
constexpr int MM_WARPS = 4;
constexpr int MM_KT = MM_C / 16;    // 16-dim contraction tiles of Q K^T
constexpr int MM_NT = MM_KC / 8;    // 8-key score tiles per chunk
constexpr int MM_OT = MM_C / 8;     // 8-dim output tiles

struct MmaAttnArgs {
  int heads, Nq, Nk, Kt, Kh, Kw, resid;
  float scale;
};

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void ldsm_x4_trans(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  const uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);     // .x = lo (low half), .y = hi
  return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

__global__ void __launch_bounds__(MM_WARPS * 32) pooled_attn_fwd_mma_kernel(const bf16* __restrict__ q,
                                                                            const bf16* __restrict__ k,
                                                                            const bf16* __restrict__ v,
                                                                            const float* __restrict__ bq,
                                                                            bf16* __restrict__ out, float* __restrict__ lse,
                                                                            MmaAttnArgs g) {
  __shared__ __align__(16) bf16 ks[MM_KC][MM_PITCH];
  __shared__ __align__(16) bf16 vs[MM_KC][MM_PITCH];
  __shared__ int kcomp[MM_KC];   // bq columns of each key of the chunk: kt | (Kt + kh) << 8 | (Kt + Kh + kw) << 16; -1: no bias
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int gq = lane >> 2, t = lane & 3;
  const int bh = blockIdx.y, KB = g.Kt + g.Kh + g.Kw;
  const int r0 = blockIdx.x * (MM_WARPS * 16) + wid * 16 + gq, r1 = r0 + 8;      // this lane's two query rows
  const bf16* qb = q + (size_t)bh * g.Nq * MM_C;
  const bf16* kb = k + (size_t)bh * g.Nk * MM_C;
  const bf16* vb = v + (size_t)bh * g.Nk * MM_C;

  uint32_t qa[MM_KT][4];
#pragma unroll
  for (int kk = 0; kk < MM_KT; ++kk) {
    const int d = kk * 16 + 2 * t;
    qa[kk][0] = r0 < g.Nq ? *reinterpret_cast<const uint32_t*>(qb + (size_t)r0 * MM_C + d) : 0u;
    qa[kk][1] = r1 < g.Nq ? *reinterpret_cast<const uint32_t*>(qb + (size_t)r1 * MM_C + d) : 0u;
    qa[kk][2] = r0 < g.Nq ? *reinterpret_cast<const uint32_t*>(qb + (size_t)r0 * MM_C + d + 8) : 0u;
    qa[kk][3] = r1 < g.Nq ? *reinterpret_cast<const uint32_t*>(qb + (size_t)r1 * MM_C + d + 8) : 0u;
  }
  const float* bq0 = (r0 > 0 && r0 < g.Nq) ? bq + ((size_t)bh * (g.Nq - 1) + (r0 - 1)) * KB : nullptr;
  const float* bq1 = (r1 > 0 && r1 < g.Nq) ? bq + ((size_t)bh * (g.Nq - 1) + (r1 - 1)) * KB : nullptr;

  float m0 = -FLT_MAX, m1 = -FLT_MAX, l0 = 0.f, l1 = 0.f;     // l: this lane's share of the row sums (reduced at the end)
  float o[MM_OT][4];
#pragma unroll
  for (int n = 0; n < MM_OT; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;

  for (int j0 = 0; j0 < g.Nk; j0 += MM_KC) {
    __syncthreads();                                          // every warp is done with the previous chunk
    for (int idx = threadIdx.x; idx < MM_KC * (MM_C / 8); idx += MM_WARPS * 32) {
      const int row = idx / (MM_C / 8), piece = idx - row * (MM_C / 8);
      const int j = j0 + row;
      uint4 kv = make_uint4(0u, 0u, 0u, 0u), vv = kv;         // rows beyond Nk are zero: 0 * p contributes nothing
      if (j < g.Nk) {
        kv = *reinterpret_cast<const uint4*>(kb + (size_t)j * MM_C + piece * 8);
        vv = *reinterpret_cast<const uint4*>(vb + (size_t)j * MM_C + piece * 8);
      }
      *reinterpret_cast<uint4*>(&ks[row][piece * 8]) = kv;
      *reinterpret_cast<uint4*>(&vs[row][piece * 8]) = vv;
    }
    if (threadIdx.x < MM_KC) {
      const int j = j0 + threadIdx.x;
      int c = -1;
      if (j > 0 && j < g.Nk) {
        const int jj = j - 1;
        const int kw = jj % g.Kw, kh = (jj / g.Kw) % g.Kh, kt = jj / (g.Kw * g.Kh);
        c = kt | ((g.Kt + kh) << 8) | ((g.Kt + g.Kh + kw) << 16);
      }
      kcomp[threadIdx.x] = c;
    }
    __syncthreads();

    float s[MM_NT][4];
#pragma unroll
    for (int n = 0; n < MM_NT; ++n) {
      s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < MM_KT; ++kk) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&ks[n * 8 + gq][kk * 16 + 2 * t]);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&ks[n * 8 + gq][kk * 16 + 8 + 2 * t]);
        mma16816(s[n], qa[kk], b0, b1);
      }
    }
    float mx0 = -FLT_MAX, mx1 = -FLT_MAX;
#pragma unroll
    for (int n = 0; n < MM_NT; ++n) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int jl = n * 8 + 2 * t + e;
        const int c = kcomp[jl];
        float a0 = s[n][e] * g.scale, a1 = s[n][2 + e] * g.scale;
        if (c >= 0) {
          const int ct = c & 0xff, ch = (c >> 8) & 0xff, cw = (c >> 16) & 0xff;
          if (bq0 != nullptr) a0 += __ldg(bq0 + ct) + __ldg(bq0 + ch) + __ldg(bq0 + cw);
          if (bq1 != nullptr) a1 += __ldg(bq1 + ct) + __ldg(bq1 + ch) + __ldg(bq1 + cw);
        }
        if (j0 + jl >= g.Nk) a0 = a1 = -FLT_MAX;
        s[n][e] = a0, s[n][2 + e] = a1;
        mx0 = fmaxf(mx0, a0), mx1 = fmaxf(mx1, a1);
      }
    }
    const float mn0 = fmaxf(m0, quad_max(mx0)), mn1 = fmaxf(m1, quad_max(mx1));
    const float corr0 = expf(m0 - mn0), corr1 = expf(m1 - mn1);
    m0 = mn0, m1 = mn1;
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int n = 0; n < MM_NT; ++n) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float p0 = expf(s[n][e] - mn0), p1 = expf(s[n][2 + e] - mn1);     // masked keys: exp(-FLT_MAX - mn) = 0
        s[n][e] = p0, s[n][2 + e] = p1;
        sum0 += p0, sum1 += p1;
      }
    }
    l0 = l0 * corr0 + sum0, l1 = l1 * corr1 + sum1;
#pragma unroll
    for (int n = 0; n < MM_OT; ++n) o[n][0] *= corr0, o[n][1] *= corr0, o[n][2] *= corr1, o[n][3] *= corr1;

#pragma unroll
    for (int kk = 0; kk < MM_KC / 16; ++kk) {
      uint32_t pa[4];
      pa[0] = pack2(s[2 * kk][0], s[2 * kk][1]);              // row g,     keys 16 kk + 2t, +1
      pa[1] = pack2(s[2 * kk][2], s[2 * kk][3]);              // row g + 8
      pa[2] = pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]);      // row g,     keys 16 kk + 8 + 2t, +1
      pa[3] = pack2(s[2 * kk + 1][2], s[2 * kk + 1][3]);      // row g + 8
#pragma unroll
      for (int np = 0; np < MM_OT / 2; ++np) {
        // four 8 x 8 blocks of V: (keys 16kk .. +7 | +8 .. +15) x (dims 16np .. +7 | +8 .. +15); lane l addresses row l & 7 of block l >> 3
        const int mtx = lane >> 3, row = lane & 7;
        uint32_t b0, b1, b2, b3;
        ldsm_x4_trans(b0, b1, b2, b3, &vs[kk * 16 + (mtx & 1) * 8 + row][np * 16 + (mtx >> 1) * 8]);
        mma16816(o[2 * np], pa, b0, b1);
        mma16816(o[2 * np + 1], pa, b2, b3);
      }
    }
  }

  l0 = quad_sum(l0), l1 = quad_sum(l1);
  const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
  const int b = bh / g.heads, h = bh % g.heads;
  const size_t pitch = (size_t)g.heads * MM_C;
#pragma unroll
  for (int n = 0; n < MM_OT; ++n) {
    const int d = n * 8 + 2 * t;
    if (r0 < g.Nq) {
      float x = o[n][0] * inv0, y = o[n][1] * inv0;
      if (g.resid && r0 > 0) {
        const float2 qq = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(qb + (size_t)r0 * MM_C + d));
        x += qq.x, y += qq.y;
      }
      *reinterpret_cast<uint32_t*>(out + ((size_t)b * g.Nq + r0) * pitch + (size_t)h * MM_C + d) = pack2(x, y);
    }
    if (r1 < g.Nq) {
      float x = o[n][2] * inv1, y = o[n][3] * inv1;
      if (g.resid && r1 > 0) {
        const float2 qq = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(qb + (size_t)r1 * MM_C + d));
        x += qq.x, y += qq.y;
      }
      *reinterpret_cast<uint32_t*>(out + ((size_t)b * g.Nq + r1) * pitch + (size_t)h * MM_C + d) = pack2(x, y);
    }
  }
  if (t == 0) {
    if (r0 < g.Nq) lse[(size_t)bh * g.Nq + r0] = m0 + logf(l0);
    if (r1 < g.Nq) lse[(size_t)bh * g.Nq + r1] = m1 + logf(l1);
  }
}


// ------------------------------------------------------------------------------------------------ backward (opt-in)
// Same building blocks as the forward.  Chunks of 32 keys (dQ pass) / 32 queries (dK/dV pass) keep the two score-shaped
// accumulators (S and dP) at 16 registers each next to the 48-register output accumulators.
//
// dQ pass: a warp owns 16 queries and walks the keys twice.  Walk 1: S = Q K^T, dP = dO V^T, p = exp(S scale + bias - lse),
// delta = sum_j p dP (fp32, exact: nothing is derived from the rounded forward output).  Walk 2: the same S / dP / p, then
// dS = p (dP - delta) rounded to bf16 feeds dQ += dS K (K through ldmatrix.trans) and dbq += dS Sel, where Sel [key][component]
// is the one-hot matrix of the three bias columns a key selects -- the bias gradient is one more MMA instead of atomics.
constexpr int MB_C = 32;         // keys (dQ pass) / queries (dK/dV pass) per chunk
constexpr int MB_NT = MB_C / 8;
constexpr int MB_SELP = 72;      // row pitch of Sel in bf16 (144 B: ldmatrix rows fall into disjoint bank groups)
constexpr int MB_KBT = 8;        // 8-column tiles of the bias gradient (Kt + Kh + Kw <= 64)

__device__ __forceinline__ void load_rows_to_smem(bf16 (*dst)[MM_PITCH], const bf16* src, size_t row_pitch, int first, int limit,
                                                  int rows) {
  for (int idx = threadIdx.x; idx < rows * (MM_C / 8); idx += MM_WARPS * 32) {
    const int row = idx / (MM_C / 8), piece = idx - row * (MM_C / 8);
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (first + row < limit) val = *reinterpret_cast<const uint4*>(src + (size_t)(first + row) * row_pitch + piece * 8);
    *reinterpret_cast<uint4*>(&dst[row][piece * 8]) = val;
  }
}

__device__ __forceinline__ void load_afrag(uint32_t (&a)[MM_KT][4], const bf16* row0, const bf16* row1, int t) {
#pragma unroll
  for (int kk = 0; kk < MM_KT; ++kk) {
    const int d = kk * 16 + 2 * t;
    a[kk][0] = row0 != nullptr ? *reinterpret_cast<const uint32_t*>(row0 + d) : 0u;
    a[kk][1] = row1 != nullptr ? *reinterpret_cast<const uint32_t*>(row1 + d) : 0u;
    a[kk][2] = row0 != nullptr ? *reinterpret_cast<const uint32_t*>(row0 + d + 8) : 0u;
    a[kk][3] = row1 != nullptr ? *reinterpret_cast<const uint32_t*>(row1 + d + 8) : 0u;
  }
}

__global__ void __launch_bounds__(MM_WARPS * 32) pooled_attn_bwd_q_mma_kernel(
    const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v, const float* __restrict__ bq,
    const bf16* __restrict__ dout, const float* __restrict__ lse, bf16* __restrict__ dq, float* __restrict__ dbq,
    float* __restrict__ delta, MmaAttnArgs g) {
  __shared__ __align__(16) bf16 ks[MB_C][MM_PITCH];
  __shared__ __align__(16) bf16 vs[MB_C][MM_PITCH];
  __shared__ __align__(16) bf16 sel[MB_C][MB_SELP];
  __shared__ int kcomp[MB_C];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int gq = lane >> 2, t = lane & 3;
  const int bh = blockIdx.y, KB = g.Kt + g.Kh + g.Kw;
  const int b = bh / g.heads, h = bh % g.heads;
  const int r0 = blockIdx.x * (MM_WARPS * 16) + wid * 16 + gq, r1 = r0 + 8;
  const bf16* qb = q + (size_t)bh * g.Nq * MM_C;
  const bf16* kb = k + (size_t)bh * g.Nk * MM_C;
  const bf16* vb = v + (size_t)bh * g.Nk * MM_C;
  const size_t pitch = (size_t)g.heads * MM_C;
  const bf16* do0 = r0 < g.Nq ? dout + ((size_t)b * g.Nq + r0) * pitch + (size_t)h * MM_C : nullptr;
  const bf16* do1 = r1 < g.Nq ? dout + ((size_t)b * g.Nq + r1) * pitch + (size_t)h * MM_C : nullptr;
  uint32_t qa[MM_KT][4], da[MM_KT][4];
  load_afrag(qa, r0 < g.Nq ? qb + (size_t)r0 * MM_C : nullptr, r1 < g.Nq ? qb + (size_t)r1 * MM_C : nullptr, t);
  load_afrag(da, do0, do1, t);
  const float* bq0 = (r0 > 0 && r0 < g.Nq) ? bq + ((size_t)bh * (g.Nq - 1) + (r0 - 1)) * KB : nullptr;
  const float* bq1 = (r1 > 0 && r1 < g.Nq) ? bq + ((size_t)bh * (g.Nq - 1) + (r1 - 1)) * KB : nullptr;
  const float ls0 = r0 < g.Nq ? lse[(size_t)bh * g.Nq + r0] : 0.f, ls1 = r1 < g.Nq ? lse[(size_t)bh * g.Nq + r1] : 0.f;
  const int nbt = (KB + 7) / 8;
  float dl0 = 0.f, dl1 = 0.f;
  float dqa[MM_OT][4], dba[MB_KBT][4];
#pragma unroll
  for (int n = 0; n < MM_OT; ++n) dqa[n][0] = dqa[n][1] = dqa[n][2] = dqa[n][3] = 0.f;
#pragma unroll
  for (int n = 0; n < MB_KBT; ++n) dba[n][0] = dba[n][1] = dba[n][2] = dba[n][3] = 0.f;

  for (int pass = 0; pass < 2; ++pass) {
    for (int j0 = 0; j0 < g.Nk; j0 += MB_C) {
      __syncthreads();
      load_rows_to_smem(ks, kb, MM_C, j0, g.Nk, MB_C);
      load_rows_to_smem(vs, vb, MM_C, j0, g.Nk, MB_C);
      if (pass == 1)
        for (int idx = threadIdx.x; idx < MB_C * MB_SELP / 8; idx += MM_WARPS * 32)
          reinterpret_cast<uint4*>(&sel[0][0])[idx] = make_uint4(0u, 0u, 0u, 0u);
      int c = -1;
      if (threadIdx.x < MB_C) {
        const int j = j0 + threadIdx.x;
        if (j > 0 && j < g.Nk) {
          const int jj = j - 1;
          const int kw = jj % g.Kw, kh = (jj / g.Kw) % g.Kh, kt = jj / (g.Kw * g.Kh);
          c = kt | ((g.Kt + kh) << 8) | ((g.Kt + g.Kh + kw) << 16);
        }
        kcomp[threadIdx.x] = c;
      }
      __syncthreads();
      if (pass == 1 && c >= 0) {
        const bf16 one = __float2bfloat16(1.0f);
        sel[threadIdx.x][c & 0xff] = one, sel[threadIdx.x][(c >> 8) & 0xff] = one, sel[threadIdx.x][(c >> 16) & 0xff] = one;
      }
      if (pass == 1) __syncthreads();

      float s[MB_NT][4], dp[MB_NT][4];
#pragma unroll
      for (int n = 0; n < MB_NT; ++n) {
        s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
        dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < MM_KT; ++kk) {
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&ks[n * 8 + gq][kk * 16 + 2 * t]);
          const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&ks[n * 8 + gq][kk * 16 + 8 + 2 * t]);
          mma16816(s[n], qa[kk], b0, b1);
          const uint32_t c0 = *reinterpret_cast<const uint32_t*>(&vs[n * 8 + gq][kk * 16 + 2 * t]);
          const uint32_t c1 = *reinterpret_cast<const uint32_t*>(&vs[n * 8 + gq][kk * 16 + 8 + 2 * t]);
          mma16816(dp[n], da[kk], c0, c1);
        }
      }
#pragma unroll
      for (int n = 0; n < MB_NT; ++n) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int jl = n * 8 + 2 * t + e;
          const int cc = kcomp[jl];
          float a0 = s[n][e] * g.scale, a1 = s[n][2 + e] * g.scale;
          if (cc >= 0) {
            const int ct = cc & 0xff, ch = (cc >> 8) & 0xff, cw = (cc >> 16) & 0xff;
            if (bq0 != nullptr) a0 += __ldg(bq0 + ct) + __ldg(bq0 + ch) + __ldg(bq0 + cw);
            if (bq1 != nullptr) a1 += __ldg(bq1 + ct) + __ldg(bq1 + ch) + __ldg(bq1 + cw);
          }
          const bool live = j0 + jl < g.Nk;
          const float p0 = (live && r0 < g.Nq) ? expf(a0 - ls0) : 0.f, p1 = (live && r1 < g.Nq) ? expf(a1 - ls1) : 0.f;
          if (pass == 0) {
            dl0 += p0 * dp[n][e], dl1 += p1 * dp[n][2 + e];
          } else {
            s[n][e] = p0 * (dp[n][e] - dl0), s[n][2 + e] = p1 * (dp[n][2 + e] - dl1);     // dS
          }
        }
      }
      if (pass == 1) {
#pragma unroll
        for (int kk = 0; kk < MB_C / 16; ++kk) {
          uint32_t pa[4];
          pa[0] = pack2(s[2 * kk][0], s[2 * kk][1]);
          pa[1] = pack2(s[2 * kk][2], s[2 * kk][3]);
          pa[2] = pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
          pa[3] = pack2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
          const int mtx = lane >> 3, row = lane & 7;
#pragma unroll
          for (int np = 0; np < MM_OT / 2; ++np) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4_trans(b0, b1, b2, b3, &ks[kk * 16 + (mtx & 1) * 8 + row][np * 16 + (mtx >> 1) * 8]);
            mma16816(dqa[2 * np], pa, b0, b1);
            mma16816(dqa[2 * np + 1], pa, b2, b3);
          }
#pragma unroll
          for (int nb = 0; nb < MB_KBT / 2; ++nb) {
            if (2 * nb < nbt) {
              uint32_t b0, b1, b2, b3;
              ldsm_x4_trans(b0, b1, b2, b3, &sel[kk * 16 + (mtx & 1) * 8 + row][nb * 16 + (mtx >> 1) * 8]);
              mma16816(dba[2 * nb], pa, b0, b1);
              mma16816(dba[2 * nb + 1], pa, b2, b3);
            }
          }
        }
      }
    }
    if (pass == 0) {
      dl0 = quad_sum(dl0), dl1 = quad_sum(dl1);
      if (t == 0) {
        if (r0 < g.Nq) delta[(size_t)bh * g.Nq + r0] = dl0;
        if (r1 < g.Nq) delta[(size_t)bh * g.Nq + r1] = dl1;
      }
    }
  }
#pragma unroll
  for (int n = 0; n < MM_OT; ++n) {
    const int d = n * 8 + 2 * t;
    if (r0 < g.Nq) {
      float x = dqa[n][0] * g.scale, y = dqa[n][1] * g.scale;
      if (g.resid && r0 > 0) {
        const float2 dd = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(do0 + d));
        x += dd.x, y += dd.y;
      }
      *reinterpret_cast<uint32_t*>(dq + ((size_t)bh * g.Nq + r0) * MM_C + d) = pack2(x, y);
    }
    if (r1 < g.Nq) {
      float x = dqa[n][2] * g.scale, y = dqa[n][3] * g.scale;
      if (g.resid && r1 > 0) {
        const float2 dd = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(do1 + d));
        x += dd.x, y += dd.y;
      }
      *reinterpret_cast<uint32_t*>(dq + ((size_t)bh * g.Nq + r1) * MM_C + d) = pack2(x, y);
    }
  }
#pragma unroll
  for (int n = 0; n < MB_KBT; ++n) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = n * 8 + 2 * t + e;
      if (col < KB) {
        if (r0 > 0 && r0 < g.Nq) dbq[((size_t)bh * (g.Nq - 1) + (r0 - 1)) * KB + col] = dba[n][e];
        if (r1 > 0 && r1 < g.Nq) dbq[((size_t)bh * (g.Nq - 1) + (r1 - 1)) * KB + col] = dba[n][2 + e];
      }
    }
  }
}

// dK / dV pass: a warp owns 16 keys (rows of the transposed problem) and walks a slice of the queries in chunks of 32:
// S^T = K Q^T and dP^T = V dO^T (B operands = Q / dO rows as they lie in shared memory), p and dS as above with lse / delta
// per column, then dV += P^T dO and dK += dS^T Q (Q / dO through ldmatrix.trans).  Partial sums of the query slices meet
// in fp32 atomics on zero-initialised dk / dv.
__global__ void __launch_bounds__(MM_WARPS * 32) pooled_attn_bwd_kv_mma_kernel(
    const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v, const float* __restrict__ bq,
    const bf16* __restrict__ dout, const float* __restrict__ lse, const float* __restrict__ delta, float* __restrict__ dk,
    float* __restrict__ dv, MmaAttnArgs g, int qsplit) {
  __shared__ __align__(16) bf16 qs[MB_C][MM_PITCH];
  __shared__ __align__(16) bf16 dos[MB_C][MM_PITCH];
  __shared__ float lss[MB_C], dls[MB_C];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int gq = lane >> 2, t = lane & 3;
  const int bh = blockIdx.y, KB = g.Kt + g.Kh + g.Kw;
  const int b = bh / g.heads, h = bh % g.heads;
  const int j_0 = blockIdx.x * (MM_WARPS * 16) + wid * 16 + gq, j_1 = j_0 + 8;
  const bf16* qb = q + (size_t)bh * g.Nq * MM_C;
  const bf16* kb = k + (size_t)bh * g.Nk * MM_C;
  const bf16* vb = v + (size_t)bh * g.Nk * MM_C;
  const size_t pitch = (size_t)g.heads * MM_C;
  const bf16* dob = dout + (size_t)b * g.Nq * pitch + (size_t)h * MM_C;
  uint32_t ka[MM_KT][4], va[MM_KT][4];
  load_afrag(ka, j_0 < g.Nk ? kb + (size_t)j_0 * MM_C : nullptr, j_1 < g.Nk ? kb + (size_t)j_1 * MM_C : nullptr, t);
  load_afrag(va, j_0 < g.Nk ? vb + (size_t)j_0 * MM_C : nullptr, j_1 < g.Nk ? vb + (size_t)j_1 * MM_C : nullptr, t);
  int c0t = -1, c0h = 0, c0w = 0, c1t = -1, c1h = 0, c1w = 0;       // bias columns of the two keys (-1: cls / out of range)
  if (j_0 > 0 && j_0 < g.Nk) {
    const int jj = j_0 - 1;
    c0w = g.Kt + g.Kh + jj % g.Kw, c0h = g.Kt + (jj / g.Kw) % g.Kh, c0t = jj / (g.Kw * g.Kh);
  }
  if (j_1 > 0 && j_1 < g.Nk) {
    const int jj = j_1 - 1;
    c1w = g.Kt + g.Kh + jj % g.Kw, c1h = g.Kt + (jj / g.Kw) % g.Kh, c1t = jj / (g.Kw * g.Kh);
  }
  float dka[MM_OT][4], dva[MM_OT][4];
#pragma unroll
  for (int n = 0; n < MM_OT; ++n) {
    dka[n][0] = dka[n][1] = dka[n][2] = dka[n][3] = 0.f;
    dva[n][0] = dva[n][1] = dva[n][2] = dva[n][3] = 0.f;
  }
  const int per = (((g.Nq + qsplit - 1) / qsplit) + MB_C - 1) / MB_C * MB_C;
  const int ibeg = blockIdx.z * per, iend = min(g.Nq, ibeg + per);
  for (int i0 = ibeg; i0 < iend; i0 += MB_C) {
    __syncthreads();
    load_rows_to_smem(qs, qb, MM_C, i0, iend, MB_C);
    load_rows_to_smem(dos, dob, pitch, i0, iend, MB_C);
    if (threadIdx.x < MB_C) {
      const int i = i0 + threadIdx.x;
      lss[threadIdx.x] = i < iend ? lse[(size_t)bh * g.Nq + i] : 0.f;
      dls[threadIdx.x] = i < iend ? delta[(size_t)bh * g.Nq + i] : 0.f;
    }
    __syncthreads();
    float st[MB_NT][4], dpt[MB_NT][4];
#pragma unroll
    for (int n = 0; n < MB_NT; ++n) {
      st[n][0] = st[n][1] = st[n][2] = st[n][3] = 0.f;
      dpt[n][0] = dpt[n][1] = dpt[n][2] = dpt[n][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < MM_KT; ++kk) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&qs[n * 8 + gq][kk * 16 + 2 * t]);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&qs[n * 8 + gq][kk * 16 + 8 + 2 * t]);
        mma16816(st[n], ka[kk], b0, b1);
        const uint32_t e0 = *reinterpret_cast<const uint32_t*>(&dos[n * 8 + gq][kk * 16 + 2 * t]);
        const uint32_t e1 = *reinterpret_cast<const uint32_t*>(&dos[n * 8 + gq][kk * 16 + 8 + 2 * t]);
        mma16816(dpt[n], va[kk], e0, e1);
      }
    }
#pragma unroll
    for (int n = 0; n < MB_NT; ++n) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int il = n * 8 + 2 * t + e;
        const int i = i0 + il;
        float a0 = st[n][e] * g.scale, a1 = st[n][2 + e] * g.scale;
        if (i > 0 && i < iend) {
          const float* brow = bq + ((size_t)bh * (g.Nq - 1) + (i - 1)) * KB;
          if (c0t >= 0) a0 += __ldg(brow + c0t) + __ldg(brow + c0h) + __ldg(brow + c0w);
          if (c1t >= 0) a1 += __ldg(brow + c1t) + __ldg(brow + c1h) + __ldg(brow + c1w);
        }
        const float p0 = (i < iend && j_0 < g.Nk) ? expf(a0 - lss[il]) : 0.f;
        const float p1 = (i < iend && j_1 < g.Nk) ? expf(a1 - lss[il]) : 0.f;
        const float d0 = p0 * (dpt[n][e] - dls[il]), d1 = p1 * (dpt[n][2 + e] - dls[il]);
        st[n][e] = p0, st[n][2 + e] = p1;
        dpt[n][e] = d0, dpt[n][2 + e] = d1;
      }
    }
#pragma unroll
    for (int kk = 0; kk < MB_C / 16; ++kk) {
      uint32_t pa[4], sa[4];
      pa[0] = pack2(st[2 * kk][0], st[2 * kk][1]), pa[1] = pack2(st[2 * kk][2], st[2 * kk][3]);
      pa[2] = pack2(st[2 * kk + 1][0], st[2 * kk + 1][1]), pa[3] = pack2(st[2 * kk + 1][2], st[2 * kk + 1][3]);
      sa[0] = pack2(dpt[2 * kk][0], dpt[2 * kk][1]), sa[1] = pack2(dpt[2 * kk][2], dpt[2 * kk][3]);
      sa[2] = pack2(dpt[2 * kk + 1][0], dpt[2 * kk + 1][1]), sa[3] = pack2(dpt[2 * kk + 1][2], dpt[2 * kk + 1][3]);
      const int mtx = lane >> 3, row = lane & 7;
#pragma unroll
      for (int np = 0; np < MM_OT / 2; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_trans(b0, b1, b2, b3, &dos[kk * 16 + (mtx & 1) * 8 + row][np * 16 + (mtx >> 1) * 8]);
        mma16816(dva[2 * np], pa, b0, b1);
        mma16816(dva[2 * np + 1], pa, b2, b3);
        ldsm_x4_trans(b0, b1, b2, b3, &qs[kk * 16 + (mtx & 1) * 8 + row][np * 16 + (mtx >> 1) * 8]);
        mma16816(dka[2 * np], sa, b0, b1);
        mma16816(dka[2 * np + 1], sa, b2, b3);
      }
    }
  }
#pragma unroll
  for (int n = 0; n < MM_OT; ++n) {
    const int d = n * 8 + 2 * t;
    if (j_0 < g.Nk) {
      const size_t off = ((size_t)bh * g.Nk + j_0) * MM_C + d;
      atomicAdd(dk + off, dka[n][0] * g.scale), atomicAdd(dk + off + 1, dka[n][1] * g.scale);
      atomicAdd(dv + off, dva[n][0]), atomicAdd(dv + off + 1, dva[n][1]);
    }
    if (j_1 < g.Nk) {
      const size_t off = ((size_t)bh * g.Nk + j_1) * MM_C + d;
      atomicAdd(dk + off, dka[n][2] * g.scale), atomicAdd(dk + off + 1, dka[n][3] * g.scale);
      atomicAdd(dv + off, dva[n][2]), atomicAdd(dv + off + 1, dva[n][3]);
    }
  }
}

}  // namespace

// Called by pvrl_pooled_attn_fwd (mvit.cu) for bf16 problems after argument validation.
int pooled_attn_fwd_mma_launch(const void* q, const void* k, const void* v, const float* bq, void* out, float* lse, int B,
                               int heads, int Nq, int Nk, int Kt, int Kh, int Kw, float scale, int resid,
                               cudaStream_t stream) {
  PVRL_CHECK_ARG(Kt + Kh + Kw <= 255, "pvrl_pooled_attn_fwd: Kt + Kh + Kw = %d exceeds 255", Kt + Kh + Kw);
  MmaAttnArgs g;
  g.heads = heads, g.Nq = Nq, g.Nk = Nk, g.Kt = Kt, g.Kh = Kh, g.Kw = Kw, g.resid = resid, g.scale = scale;
  const dim3 grid((Nq + MM_WARPS * 16 - 1) / (MM_WARPS * 16), B * heads);
  pooled_attn_fwd_mma_kernel<<<grid, MM_WARPS * 32, 0, stream>>>(static_cast<const bf16*>(q), static_cast<const bf16*>(k),
                                                                 static_cast<const bf16*>(v), bq, static_cast<bf16*>(out),
                                                                 lse, g);
  return launched("pooled_attn_fwd_mma_kernel");
}


// Called by pvrl_pooled_attn_bwd (mvit.cu) for bf16 problems when PVRL_MVIT_ATTN_MMA_BWD = 1 (opt-in: checked by the CPU
// emulation only so far).  dk / dv: fp32, zero-initialised by the caller.
int pooled_attn_bwd_mma_launch(const void* q, const void* k, const void* v, const float* bq, const void* dout, const float* lse,
                               void* dq, float* dk, float* dv, float* dbq, float* delta, int B, int heads, int Nq, int Nk,
                               int Kt, int Kh, int Kw, float scale, int resid, cudaStream_t stream) {
  PVRL_CHECK_ARG(Kt + Kh + Kw <= 8 * MB_KBT, "pvrl_pooled_attn_bwd: Kt + Kh + Kw = %d exceeds %d", Kt + Kh + Kw, 8 * MB_KBT);
  MmaAttnArgs g;
  g.heads = heads, g.Nq = Nq, g.Nk = Nk, g.Kt = Kt, g.Kh = Kh, g.Kw = Kw, g.resid = resid, g.scale = scale;
  const int BH = B * heads;
  const dim3 gq((Nq + MM_WARPS * 16 - 1) / (MM_WARPS * 16), BH);
  pooled_attn_bwd_q_mma_kernel<<<gq, MM_WARPS * 32, 0, stream>>>(
      static_cast<const bf16*>(q), static_cast<const bf16*>(k), static_cast<const bf16*>(v), bq, static_cast<const bf16*>(dout),
      lse, static_cast<bf16*>(dq), dbq, delta, g);
  int rc = launched("pooled_attn_bwd_q_mma_kernel");
  if (rc) return rc;
  const int kblocks = (Nk + MM_WARPS * 16 - 1) / (MM_WARPS * 16);
  int split = (num_sms() * 4 + kblocks * BH - 1) / (kblocks * BH);
  const int max_split = (Nq + 255) / 256;
  if (split > max_split) split = max_split;
  if (split < 1) split = 1;
  if (split > 65535) split = 65535;
  const dim3 gk(kblocks, BH, split);
  pooled_attn_bwd_kv_mma_kernel<<<gk, MM_WARPS * 32, 0, stream>>>(
      static_cast<const bf16*>(q), static_cast<const bf16*>(k), static_cast<const bf16*>(v), bq, static_cast<const bf16*>(dout),
      lse, delta, dk, dv, g, split);
  return launched("pooled_attn_bwd_kv_mma_kernel");
}

}  // namespace pvrl
