// Temporal attention for T <= 8 frames (the benchmark's 8x224 clips; vit.py:130-133 -> :84-88), bf16, head_dim 64.
// 42 336 (sequence, head) problems of 8x8 scores per 18 clips: pure HBM traffic (read qkv once, write o once), so the
// kernel is built to execute as few instructions per byte as possible:
//   * a warp owns TWO (sequence, head) pairs; lane (g = lane/4, t = lane%4) loads dims [16t, 16t+16) of row g of
//     q / k / v with two 16-byte loads each -- a quad covers one contiguous 128-byte head row;
//   * the loaded bf16 pairs ARE the mma.sync fragments: S = Q K^T runs as m16n8k16 (pair A in rows 0-7, pair B in rows
//     8-15; the contraction order over the 64 dims is permuted identically for both operands, which a dot product does
//     not see), softmax happens on the 2 scores per lane with quad shuffles, P feeds the next MMA straight from
//     registers, and V / K / Q / dO become "col" operands through movmatrix (8x8 b16 register transpose);
//   * the m16n8 results land as 16 contiguous dims per lane again -> two 16-byte stores per row.
// No shared memory, no unpacking to fp32, ~150 (forward) / ~350 (backward) instructions per warp and pair of problems.
// Tensor work is mma.sync (legacy warp-level MMA): 8x8 problems cannot fill a 128-row tcgen05 tile, and the op is
// bandwidth-bound anyway.  Sequences of 9..32 frames and the fp32 parity mode use attention_small.cu.
#include "pvrl_host.h"
#include "pvrl_ptx.cuh"

namespace pvrl {
namespace {

constexpr float LOG2E_F = 1.4426950408889634f;
constexpr int T8_THREADS = 128;

__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_1688(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {   // d = a b (no accumulate)
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%7, %7, %7, %7};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a0), "r"(a1), "r"(b0), "f"(0.f));
}
// 8x8 b16 matrix held one 32-bit register per lane (lane (g, t) = row g, columns 2t, 2t+1) -> its transpose, same layout
__device__ __forceinline__ uint32_t movm_trans(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

struct Row16 {   // 16 bf16 of one head row: r[i] = dims (16t + 2i, 16t + 2i + 1)
  uint32_t r[8];
};
__device__ __forceinline__ Row16 load_row16(const __nv_bfloat16* p, bool valid) {
  Row16 v;
  if (valid) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(p)), b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
    v.r[0] = a.x, v.r[1] = a.y, v.r[2] = a.z, v.r[3] = a.w, v.r[4] = b.x, v.r[5] = b.y, v.r[6] = b.z, v.r[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v.r[i] = 0u;
  }
  return v;
}
__device__ __forceinline__ void store_row16(__nv_bfloat16* p, const uint32_t (&r)[8]) {
  reinterpret_cast<uint4*>(p)[0] = make_uint4(r[0], r[1], r[2], r[3]);
  reinterpret_cast<uint4*>(p)[1] = make_uint4(r[4], r[5], r[6], r[7]);
}

// rows 0-7 <- X_a Y_a^T (d[0], d[1]) and rows 8-15 <- X_b Y_b^T (d[2], d[3]): two m16n8k16 chains over the 64 dims; the
// off-diagonal halves (X_b Y_a^T, X_a Y_b^T) are computed and dropped.
__device__ __forceinline__ void pair_scores(const Row16& xa, const Row16& xb, const Row16& ya, const Row16& yb, float& a0,
                                            float& a1, float& b0, float& b1) {
  float da[4] = {0.f, 0.f, 0.f, 0.f}, db[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    mma_16816(da, xa.r[2 * s], xb.r[2 * s], xa.r[2 * s + 1], xb.r[2 * s + 1], ya.r[2 * s], ya.r[2 * s + 1]);
    mma_16816(db, xa.r[2 * s], xb.r[2 * s], xa.r[2 * s + 1], xb.r[2 * s + 1], yb.r[2 * s], yb.r[2 * s + 1]);
  }
  a0 = da[0], a1 = da[1], b0 = db[2], b1 = db[3];
}

// out_a[g][16t..] = W_a[g][:] Y_a[:][16t..] (rows 0-7, keys in the contraction), same for b: per 8-dim block one
// movmatrix per pair and one m16n8k8 per pair.  wa / wb: this lane's packed (row g, keys 2t, 2t+1) weights.
__device__ __forceinline__ void pair_apply(uint32_t wa, uint32_t wb, const Row16& ya, const Row16& yb, uint32_t (&oa)[8],
                                           uint32_t (&ob)[8]) {
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    float da[4], db[4];
    mma_1688(da, wa, wb, movm_trans(ya.r[m]));
    mma_1688(db, wa, wb, movm_trans(yb.r[m]));
    oa[m] = pack_bf16x2(da[0], da[1]);
    ob[m] = pack_bf16x2(db[2], db[3]);
  }
}

__global__ void __launch_bounds__(T8_THREADS)
attn_t8_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, float* __restrict__ lse,
                   int n_pairs, int seq, int H, float scale) {
  if (PDL_EARLY_TRIGGER) pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  // pairs are visited from the last to the first: the GEMM that produced qkv / dO wrote its highest rows last, and they
  // are still in L2
  const int pa = ((n_pairs + 1) / 2 - 1 - (blockIdx.x * (T8_THREADS / 32) + (threadIdx.x >> 5))) * 2;   // n_pairs < 2^31
  if (pa < 0) return;
  const int pb = pa + 1;
  const bool has_b = pb < n_pairs, row_ok = g < seq;
  const int C = H * 64;
  const long long pitch = 3LL * C;
  const int sa_i = pa / H, sb_i = has_b ? pb / H : sa_i;
  const int ha = pa - sa_i * H, hb = has_b ? pb - sb_i * H : ha;
  const long long sa = sa_i, sb = sb_i;
  const __nv_bfloat16* ba = qkv + (sa * seq + g) * pitch + ha * 64 + 16 * t;
  const __nv_bfloat16* bb = qkv + (sb * seq + g) * pitch + hb * 64 + 16 * t;
  const bool va_ok = row_ok, vb_ok = row_ok && has_b;
  const Row16 qa = load_row16(ba, va_ok), ka = load_row16(ba + C, va_ok), va = load_row16(ba + 2 * C, va_ok);
  const Row16 qb = load_row16(bb, vb_ok), kb = load_row16(bb + C, vb_ok), vb = load_row16(bb + 2 * C, vb_ok);

  float s[4];   // S_a[g][2t], S_a[g][2t+1], S_b[g][2t], S_b[g][2t+1]
  pair_scores(qa, qb, ka, kb, s[0], s[1], s[2], s[3]);
  const float sl2 = scale * LOG2E_F;
  const bool c0 = 2 * t < seq, c1 = 2 * t + 1 < seq;
  uint32_t w[2];
  float lse_v[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const float x0 = c0 ? s[2 * p] : -INFINITY, x1 = c1 ? s[2 * p + 1] : -INFINITY;
    const float mx = quad_max(fmaxf(x0, x1));
    const float e0 = ex2_approx((x0 - mx) * sl2), e1 = ex2_approx((x1 - mx) * sl2);
    const float sum = quad_sum(e0 + e1);
    const float inv = 1.0f / sum;
    w[p] = pack_bf16x2(e0 * inv, e1 * inv);
    lse_v[p] = mx * scale + __logf(sum);
  }
  uint32_t oa[8], ob[8];
  pair_apply(w[0], w[1], va, vb, oa, ob);
  if (va_ok) {
    store_row16(out + (sa * seq + g) * C + ha * 64 + 16 * t, oa);
    if (lse != nullptr && t == 0) lse[(long long)pa * seq + g] = lse_v[0];
  }
  if (vb_ok) {
    store_row16(out + (sb * seq + g) * C + hb * 64 + 16 * t, ob);
    if (lse != nullptr && t == 0) lse[(long long)pb * seq + g] = lse_v[1];
  }
}

__global__ void __launch_bounds__(T8_THREADS)
attn_t8_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ out,
                   const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse,
                   __nv_bfloat16* __restrict__ dqkv, int n_pairs, int seq, int H, float scale) {
  if (PDL_EARLY_TRIGGER) pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  // pairs are visited from the last to the first: the GEMM that produced qkv / dO wrote its highest rows last, and they
  // are still in L2
  const int pa = ((n_pairs + 1) / 2 - 1 - (blockIdx.x * (T8_THREADS / 32) + (threadIdx.x >> 5))) * 2;   // n_pairs < 2^31
  if (pa < 0) return;
  const int pb = pa + 1;
  const bool has_b = pb < n_pairs, row_ok = g < seq;
  const int C = H * 64;
  const long long pitch = 3LL * C;
  const int sa_i = pa / H, sb_i = has_b ? pb / H : sa_i;
  const int ha = pa - sa_i * H, hb = has_b ? pb - sb_i * H : ha;
  const long long sa = sa_i, sb = sb_i;
  const bool va_ok = row_ok, vb_ok = row_ok && has_b;
  const long long qoa = (sa * seq + g) * pitch + ha * 64 + 16 * t, qob = (sb * seq + g) * pitch + hb * 64 + 16 * t;
  const long long ooa = (sa * seq + g) * C + ha * 64 + 16 * t, oob = (sb * seq + g) * C + hb * 64 + 16 * t;
  const Row16 qa = load_row16(qkv + qoa, va_ok), ka = load_row16(qkv + qoa + C, va_ok), va = load_row16(qkv + qoa + 2 * C, va_ok);
  const Row16 qb = load_row16(qkv + qob, vb_ok), kb = load_row16(qkv + qob + C, vb_ok), vb = load_row16(qkv + qob + 2 * C, vb_ok);
  const Row16 oa = load_row16(out + ooa, va_ok), ob = load_row16(out + oob, vb_ok);
  const Row16 ga = load_row16(dout + ooa, va_ok), gb = load_row16(dout + oob, vb_ok);
  const float la = va_ok ? __ldg(lse + (long long)pa * seq + g) : 0.f, lb = vb_ok ? __ldg(lse + (long long)pb * seq + g) : 0.f;

  // delta_g = dO_g . O_g = diagonal of dO O^T: element (g, g) sits in lane (g, g / 2), slot g % 2
  float d[4];
  pair_scores(ga, gb, oa, ob, d[0], d[1], d[2], d[3]);
  const int diag_lane = (lane & ~3) | (g >> 1);
  const float delta_a = __shfl_sync(0xffffffffu, (g & 1) ? d[1] : d[0], diag_lane);
  const float delta_b = __shfl_sync(0xffffffffu, (g & 1) ? d[3] : d[2], diag_lane);

  float s[4], dp[4];
  pair_scores(qa, qb, ka, kb, s[0], s[1], s[2], s[3]);      // S = Q K^T
  pair_scores(ga, gb, va, vb, dp[0], dp[1], dp[2], dp[3]);  // dP = dO V^T
  const float sl2 = scale * LOG2E_F;
  const bool c0 = 2 * t < seq, c1 = 2 * t + 1 < seq;
  const float l2a = la * LOG2E_F, l2b = lb * LOG2E_F;
  const float p0 = c0 ? ex2_approx(fmaf(s[0], sl2, -l2a)) : 0.f, p1 = c1 ? ex2_approx(fmaf(s[1], sl2, -l2a)) : 0.f;
  const float p2 = c0 ? ex2_approx(fmaf(s[2], sl2, -l2b)) : 0.f, p3 = c1 ? ex2_approx(fmaf(s[3], sl2, -l2b)) : 0.f;
  const uint32_t P_a = pack_bf16x2(p0, p1), P_b = pack_bf16x2(p2, p3);
  const uint32_t dS_a = pack_bf16x2(p0 * (dp[0] - delta_a) * scale, p1 * (dp[1] - delta_a) * scale);
  const uint32_t dS_b = pack_bf16x2(p2 * (dp[2] - delta_b) * scale, p3 * (dp[3] - delta_b) * scale);

  uint32_t ra[8], rb[8];
  pair_apply(dS_a, dS_b, ka, kb, ra, rb);                          // dQ = dS K
  if (va_ok) store_row16(dqkv + qoa, ra);
  if (vb_ok) store_row16(dqkv + qob, rb);
  pair_apply(movm_trans(dS_a), movm_trans(dS_b), qa, qb, ra, rb);  // dK = dS^T Q
  if (va_ok) store_row16(dqkv + qoa + C, ra);
  if (vb_ok) store_row16(dqkv + qob + C, rb);
  pair_apply(movm_trans(P_a), movm_trans(P_b), ga, gb, ra, rb);    // dV = P^T dO
  if (va_ok) store_row16(dqkv + qoa + 2 * C, ra);
  if (vb_ok) store_row16(dqkv + qob + 2 * C, rb);
}

}  // namespace

int attn_t8_fwd_launch(const void* qkv, void* out, float* lse, int n_seq, int seq, int H, float scale,
                       cudaStream_t stream) {
  if ((long long)n_seq * H >= (1LL << 31) - 2) return fail(-1, "attn_t8: too many (sequence, head) pairs");
  const int n_pairs = n_seq * H;
  const int warps = (n_pairs + 1) / 2;
  const unsigned grid = static_cast<unsigned>((warps + T8_THREADS / 32 - 1) / (T8_THREADS / 32));
  PVRL_CUDA(launch_pdl(attn_t8_fwd_kernel, dim3(grid), dim3(T8_THREADS), 0, stream, static_cast<const __nv_bfloat16*>(qkv),
                       static_cast<__nv_bfloat16*>(out), lse, n_pairs, seq, H, scale));
  return launched("attn_t8_fwd_kernel");
}

int attn_t8_bwd_launch(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int n_seq,
                       int seq, int H, float scale, cudaStream_t stream) {
  if ((long long)n_seq * H >= (1LL << 31) - 2) return fail(-1, "attn_t8: too many (sequence, head) pairs");
  const int n_pairs = n_seq * H;
  const int warps = (n_pairs + 1) / 2;
  const unsigned grid = static_cast<unsigned>((warps + T8_THREADS / 32 - 1) / (T8_THREADS / 32));
  PVRL_CUDA(launch_pdl(attn_t8_bwd_kernel, dim3(grid), dim3(T8_THREADS), 0, stream,
                       static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(out),
                       static_cast<const __nv_bfloat16*>(dout), lse, static_cast<__nv_bfloat16*>(dqkv), n_pairs, seq, H,
                       scale));
  return launched("attn_t8_bwd_kernel");
}

}  // namespace pvrl
