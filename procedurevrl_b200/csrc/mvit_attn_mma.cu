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

}  // namespace pvrl
