// Temporal attention for 9..32 frames (BASELINE config 4: TimeSformer-B 32x224; vit.py:130-133 -> :84-88), bf16,
// head_dim 64.  42 336 (sequence, head) problems of up to 32x32 scores per 18 clips: 11 GFLOP against 0.7 GB of HBM
// traffic (read qkv once, write o once) -- bandwidth-bound, like the T <= 8 case, and built the same way
// (attention_t8.cu), one problem per warp:
//   * lane (g = lane/4, t = lane%4) loads dims [16t, 16t+16) of rows g, 8+g, 16+g, 24+g of q / k / v with two 16-byte
//     loads each -- a quad covers one contiguous 128-byte head row, nothing is staged in shared memory;
//   * the loaded bf16 pairs ARE mma.sync.m16n8k16 fragments: for S = Q K^T the contraction order over the 64 dims is
//     permuted identically for both operands (a dot product does not see it); the 32x32 scores sit in 2 x 4 accumulator
//     tiles whose (row g / g+8, keys 2t, 2t+1) layout is exactly the A-operand layout of the next MMA, so the softmax
//     (quad shuffles) feeds P V straight from registers;
//   * V / K / Q / dO enter the second MMAs as "col" operands through movmatrix (8x8 b16 register transpose), P and dS
//     are transposed the same way for dV = P^T dO and dK = dS^T Q;
//   * every result lands as 16 contiguous dims per lane again -> two 16-byte stores per row.
// The backward forms delta_i = sum_j P_ij dP_ij from the tiles it already holds (= rowsum(dO * O) exactly), so `out` is
// not read at all.  NT = number of 8-row tiles (2: seq <= 16, 4: seq <= 32); rows / keys >= seq are zero-filled / masked.
// Tensor work is mma.sync (legacy warp-level MMA): a 32x32 problem cannot fill a 128-row tcgen05 tile either.
#include "pvrl_host.h"
#include "pvrl_ptx.cuh"

namespace pvrl {
namespace {

constexpr float LOG2E_F = 1.4426950408889634f;
constexpr int T32_THREADS = 128;

__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
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

// acc[jt] += X[rows 16M .. 16M+15] Y[rows 8jt .. 8jt+7]^T over the 64 dims: tile (M, jt) of X Y^T.
// acc[jt][0..1] = (row 16M+g, cols 8jt+2t, +1), acc[jt][2..3] = (row 16M+8+g, same cols).
template <int NT>
__device__ __forceinline__ void score_tiles(const Row16 (&x)[NT], const Row16 (&y)[NT], int M, float (&acc)[NT][4]) {
#pragma unroll
  for (int jt = 0; jt < NT; ++jt) {
#pragma unroll
    for (int s = 0; s < 4; ++s)
      mma_16816(acc[jt], x[2 * M].r[2 * s], x[2 * M + 1].r[2 * s], x[2 * M].r[2 * s + 1], x[2 * M + 1].r[2 * s + 1],
                y[jt].r[2 * s], y[jt].r[2 * s + 1]);
  }
}

// Rows 16M+g (lo) and 16M+8+g (hi) of W Y, dims [16t, 16t+16) per lane.  a[ks][0..3] = the m16n8k16 A fragments of W for
// the contraction slice 16ks .. 16ks+15: (row g, k 2t), (row g+8, k 2t), (row g, k 8+2t), (row g+8, k 8+2t).
template <int NT>
__device__ __forceinline__ void apply_tiles(const uint32_t (&a)[NT / 2][4], const Row16 (&y)[NT], uint32_t (&lo)[8],
                                            uint32_t (&hi)[8]) {
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ks = 0; ks < NT / 2; ++ks)
      mma_16816(acc, a[ks][0], a[ks][1], a[ks][2], a[ks][3], movm_trans(y[2 * ks].r[m]), movm_trans(y[2 * ks + 1].r[m]));
    lo[m] = pack_bf16x2(acc[0], acc[1]);
    hi[m] = pack_bf16x2(acc[2], acc[3]);
  }
}

template <int NT>
__global__ void __launch_bounds__(T32_THREADS)
attn_t32_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, float* __restrict__ lse,
                    int n_pairs, int seq, int H, float scale) {
  constexpr int MT = NT / 2;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  // problems are visited from the last to the first: the GEMM that produced qkv wrote its highest rows last (still in L2)
  const int p = n_pairs - 1 - (blockIdx.x * (T32_THREADS / 32) + (threadIdx.x >> 5));
  if (p < 0) return;
  const int C = H * 64;
  const long long pitch = 3LL * C;
  const int s_i = p / H, h = p - s_i * H;
  const long long row0 = (long long)s_i * seq;
  const __nv_bfloat16* base = qkv + row0 * pitch + h * 64 + 16 * t;
  Row16 q[NT], k[NT], v[NT];
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    const int r = 8 * i + g;
    const bool ok = r < seq;
    const __nv_bfloat16* ptr = base + r * pitch;
    q[i] = load_row16(ptr, ok), k[i] = load_row16(ptr + C, ok), v[i] = load_row16(ptr + 2 * C, ok);
  }
  const float sl2 = scale * LOG2E_F;
#pragma unroll
  for (int M = 0; M < MT; ++M) {
    float s[NT][4];
#pragma unroll
    for (int jt = 0; jt < NT; ++jt) s[jt][0] = s[jt][1] = s[jt][2] = s[jt][3] = 0.f;
    score_tiles<NT>(q, k, M, s);
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int jt = 0; jt < NT; ++jt) {
      const bool c0 = 8 * jt + 2 * t < seq, c1 = 8 * jt + 2 * t + 1 < seq;
      s[jt][0] = c0 ? s[jt][0] : -INFINITY, s[jt][1] = c1 ? s[jt][1] : -INFINITY;
      s[jt][2] = c0 ? s[jt][2] : -INFINITY, s[jt][3] = c1 ? s[jt][3] : -INFINITY;
      mx0 = fmaxf(mx0, fmaxf(s[jt][0], s[jt][1])), mx1 = fmaxf(mx1, fmaxf(s[jt][2], s[jt][3]));
    }
    mx0 = quad_max(mx0), mx1 = quad_max(mx1);      // key 0 is always valid: finite
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int jt = 0; jt < NT; ++jt) {
      s[jt][0] = ex2_approx((s[jt][0] - mx0) * sl2), s[jt][1] = ex2_approx((s[jt][1] - mx0) * sl2);
      s[jt][2] = ex2_approx((s[jt][2] - mx1) * sl2), s[jt][3] = ex2_approx((s[jt][3] - mx1) * sl2);
      sum0 += s[jt][0] + s[jt][1], sum1 += s[jt][2] + s[jt][3];
    }
    sum0 = quad_sum(sum0), sum1 = quad_sum(sum1);
    const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
    uint32_t a[MT][4];
#pragma unroll
    for (int ks = 0; ks < MT; ++ks) {
      a[ks][0] = pack_bf16x2(s[2 * ks][0] * inv0, s[2 * ks][1] * inv0);
      a[ks][1] = pack_bf16x2(s[2 * ks][2] * inv1, s[2 * ks][3] * inv1);
      a[ks][2] = pack_bf16x2(s[2 * ks + 1][0] * inv0, s[2 * ks + 1][1] * inv0);
      a[ks][3] = pack_bf16x2(s[2 * ks + 1][2] * inv1, s[2 * ks + 1][3] * inv1);
    }
    uint32_t lo[8], hi[8];
    apply_tiles<NT>(a, v, lo, hi);
    const int r0 = 16 * M + g, r1 = r0 + 8;
    if (r0 < seq) {
      store_row16(out + (row0 + r0) * C + h * 64 + 16 * t, lo);
      if (lse != nullptr && t == 0) lse[(long long)p * seq + r0] = mx0 * scale + __logf(sum0);
    }
    if (r1 < seq) {
      store_row16(out + (row0 + r1) * C + h * 64 + 16 * t, hi);
      if (lse != nullptr && t == 0) lse[(long long)p * seq + r1] = mx1 * scale + __logf(sum1);
    }
  }
}

template <int NT>
__global__ void __launch_bounds__(T32_THREADS)
attn_t32_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout,
                    const float* __restrict__ lse, __nv_bfloat16* __restrict__ dqkv, int n_pairs, int seq, int H,
                    float scale) {
  constexpr int MT = NT / 2;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int p = n_pairs - 1 - (blockIdx.x * (T32_THREADS / 32) + (threadIdx.x >> 5));
  if (p < 0) return;
  const int C = H * 64;
  const long long pitch = 3LL * C;
  const int s_i = p / H, h = p - s_i * H;
  const long long row0 = (long long)s_i * seq;
  const long long qoff = row0 * pitch + h * 64 + 16 * t, ooff = row0 * C + h * 64 + 16 * t;
  Row16 q[NT], k[NT], v[NT], go[NT];
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    const int r = 8 * i + g;
    const bool ok = r < seq;
    const __nv_bfloat16* ptr = qkv + qoff + r * pitch;
    q[i] = load_row16(ptr, ok), k[i] = load_row16(ptr + C, ok), v[i] = load_row16(ptr + 2 * C, ok);
    go[i] = load_row16(dout + ooff + (long long)r * C, ok);
  }
  const float sl2 = scale * LOG2E_F;
  // P and dS as 8x8 blocks: [M][half][jt] = rows 16M + 8*half + g, keys 8jt + 2t, +1 (bf16 pairs)
  uint32_t P[MT][2][NT], dS[MT][2][NT];
#pragma unroll
  for (int M = 0; M < MT; ++M) {
    float s[NT][4], dp[NT][4];
#pragma unroll
    for (int jt = 0; jt < NT; ++jt) {
      s[jt][0] = s[jt][1] = s[jt][2] = s[jt][3] = 0.f;
      dp[jt][0] = dp[jt][1] = dp[jt][2] = dp[jt][3] = 0.f;
    }
    score_tiles<NT>(q, k, M, s);       // S = Q K^T
    score_tiles<NT>(go, v, M, dp);     // dP = dO V^T
    const int r0 = 16 * M + g, r1 = r0 + 8;
    const bool ok0 = r0 < seq, ok1 = r1 < seq;
    const float l0 = ok0 ? __ldg(lse + (long long)p * seq + r0) * LOG2E_F : 0.f;
    const float l1 = ok1 ? __ldg(lse + (long long)p * seq + r1) * LOG2E_F : 0.f;
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int jt = 0; jt < NT; ++jt) {
      const bool c0 = 8 * jt + 2 * t < seq, c1 = 8 * jt + 2 * t + 1 < seq;
      s[jt][0] = (c0 && ok0) ? ex2_approx(fmaf(s[jt][0], sl2, -l0)) : 0.f;
      s[jt][1] = (c1 && ok0) ? ex2_approx(fmaf(s[jt][1], sl2, -l0)) : 0.f;
      s[jt][2] = (c0 && ok1) ? ex2_approx(fmaf(s[jt][2], sl2, -l1)) : 0.f;
      s[jt][3] = (c1 && ok1) ? ex2_approx(fmaf(s[jt][3], sl2, -l1)) : 0.f;
      d0 = fmaf(s[jt][0], dp[jt][0], fmaf(s[jt][1], dp[jt][1], d0));
      d1 = fmaf(s[jt][2], dp[jt][2], fmaf(s[jt][3], dp[jt][3], d1));
    }
    d0 = quad_sum(d0), d1 = quad_sum(d1);      // delta = rowsum(P * dP) = rowsum(dO * O)
#pragma unroll
    for (int jt = 0; jt < NT; ++jt) {
      P[M][0][jt] = pack_bf16x2(s[jt][0], s[jt][1]);
      P[M][1][jt] = pack_bf16x2(s[jt][2], s[jt][3]);
      dS[M][0][jt] = pack_bf16x2(s[jt][0] * (dp[jt][0] - d0) * scale, s[jt][1] * (dp[jt][1] - d0) * scale);
      dS[M][1][jt] = pack_bf16x2(s[jt][2] * (dp[jt][2] - d1) * scale, s[jt][3] * (dp[jt][3] - d1) * scale);
    }
  }
  uint32_t lo[8], hi[8];
  // dQ = dS K: query rows 16M.., contraction over the keys
#pragma unroll
  for (int M = 0; M < MT; ++M) {
    uint32_t a[MT][4];
#pragma unroll
    for (int ks = 0; ks < MT; ++ks)
      a[ks][0] = dS[M][0][2 * ks], a[ks][1] = dS[M][1][2 * ks], a[ks][2] = dS[M][0][2 * ks + 1], a[ks][3] = dS[M][1][2 * ks + 1];
    apply_tiles<NT>(a, k, lo, hi);
    const int r0 = 16 * M + g, r1 = r0 + 8;
    if (r0 < seq) store_row16(dqkv + qoff + r0 * pitch, lo);
    if (r1 < seq) store_row16(dqkv + qoff + r1 * pitch, hi);
  }
  // dK = dS^T Q and dV = P^T dO: key rows 16KM.., contraction over the queries; the A fragments are the transposed blocks:
  // (key 16KM+g, q 16qs+2t) = block (q tile 2qs, key tile 2KM)^T, (key +8) = key tile 2KM+1, (q +8) = q tile 2qs+1
#pragma unroll
  for (int KM = 0; KM < MT; ++KM) {
    uint32_t a[MT][4];
    const int r0 = 16 * KM + g, r1 = r0 + 8;
#pragma unroll
    for (int qs = 0; qs < MT; ++qs) {
      a[qs][0] = movm_trans(dS[qs][0][2 * KM]), a[qs][1] = movm_trans(dS[qs][0][2 * KM + 1]);
      a[qs][2] = movm_trans(dS[qs][1][2 * KM]), a[qs][3] = movm_trans(dS[qs][1][2 * KM + 1]);
    }
    apply_tiles<NT>(a, q, lo, hi);
    if (r0 < seq) store_row16(dqkv + qoff + C + r0 * pitch, lo);
    if (r1 < seq) store_row16(dqkv + qoff + C + r1 * pitch, hi);
#pragma unroll
    for (int qs = 0; qs < MT; ++qs) {
      a[qs][0] = movm_trans(P[qs][0][2 * KM]), a[qs][1] = movm_trans(P[qs][0][2 * KM + 1]);
      a[qs][2] = movm_trans(P[qs][1][2 * KM]), a[qs][3] = movm_trans(P[qs][1][2 * KM + 1]);
    }
    apply_tiles<NT>(a, go, lo, hi);
    if (r0 < seq) store_row16(dqkv + qoff + 2 * C + r0 * pitch, lo);
    if (r1 < seq) store_row16(dqkv + qoff + 2 * C + r1 * pitch, hi);
  }
}

}  // namespace

int attn_t32_fwd_launch(const void* qkv, void* out, float* lse, int n_seq, int seq, int H, float scale,
                        cudaStream_t stream) {
  if ((long long)n_seq * H >= (1LL << 31) - 8) return fail(-1, "attn_t32: too many (sequence, head) pairs");
  const int n_pairs = n_seq * H;
  const unsigned grid = static_cast<unsigned>((n_pairs + T32_THREADS / 32 - 1) / (T32_THREADS / 32));
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(qkv);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
  if (seq <= 16)
    attn_t32_fwd_kernel<2><<<grid, T32_THREADS, 0, stream>>>(x, o, lse, n_pairs, seq, H, scale);
  else
    attn_t32_fwd_kernel<4><<<grid, T32_THREADS, 0, stream>>>(x, o, lse, n_pairs, seq, H, scale);
  return launched("attn_t32_fwd_kernel");
}

int attn_t32_bwd_launch(const void* qkv, const void* dout, const float* lse, void* dqkv, int n_seq, int seq, int H,
                        float scale, cudaStream_t stream) {
  if ((long long)n_seq * H >= (1LL << 31) - 8) return fail(-1, "attn_t32: too many (sequence, head) pairs");
  const int n_pairs = n_seq * H;
  const unsigned grid = static_cast<unsigned>((n_pairs + T32_THREADS / 32 - 1) / (T32_THREADS / 32));
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(qkv);
  const __nv_bfloat16* g = static_cast<const __nv_bfloat16*>(dout);
  __nv_bfloat16* dx = static_cast<__nv_bfloat16*>(dqkv);
  if (seq <= 16)
    attn_t32_bwd_kernel<2><<<grid, T32_THREADS, 0, stream>>>(x, g, lse, dx, n_pairs, seq, H, scale);
  else
    attn_t32_bwd_kernel<4><<<grid, T32_THREADS, 0, stream>>>(x, g, lse, dx, n_pairs, seq, H, scale);
  return launched("attn_t32_bwd_kernel");
}

}  // namespace pvrl
