// Clip-order ("diffusion") transformer of the pre-training branch, reference lib/models/tfm_model.py:32-53 (residual
// attention block), :165-204 (denoising levels), vit.py:330 (call site).  4 levels x 4 blocks over B*S = 18 tokens of
// width 512: < 0.1 % of the step's FLOPs, but ~1500 eager kernels per step when expressed op by op.  Here every block
// is 5 forward / 11 backward launches of fp32 kernels specialised for "a handful of rows":
//   * ot_linear_fwd: y = [resid +] prologue(x) W^T + b, prologue = identity | LayerNorm | QuickGELU.  One warp per output
//     column streams the weight row (coalesced float4), the <= 32 activation rows sit in shared memory, the 32 row sums
//     are folded across lanes with a 31-shuffle transpose-reduce.
//   * ot_linear_dx / ot_linear_dw: dA = dY W (optionally x QuickGELU'), dW += dY^T a, db += colsum(dY).
//   * ot_ln_bwd, ot_attn_fwd / ot_attn_bwd (S <= 16 tokens per sequence, key-padding mask), ot_embed_fwd / ot_embed_bwd.
// All weight traffic (52 MB of fp32 parameters) stays L2-resident across the 16 (level, block) steps.
#include "pvrl_host.h"
#include "pvrl_ptx.cuh"

namespace pvrl {
namespace {

constexpr int OT_MT = 32;      // activation rows per block tile
constexpr int OT_KC = 512;     // contraction chunk held in shared memory
constexpr int OT_KMAX = 2048;  // largest contraction of the forward kernel (weight row cached in registers)
constexpr int OT_THREADS = 256;

__device__ __forceinline__ float quick_gelu(float u) { return u / (1.0f + __expf(-1.702f * u)); }
__device__ __forceinline__ float quick_gelu_grad(float u) {
  const float s = 1.0f / (1.0f + __expf(-1.702f * u));
  return s * (1.0f + 1.702f * u * (1.0f - s));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ----------------------------------------------------------------------------------------------------- linear forward
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// grid (ceil(N / 8), ceil(M / mt)), 256 threads: warp w owns output column n = 8 * blockIdx.x + w.
// x_mode: 0 = x as is, 1 = LayerNorm(x) (block column 0 also writes xhat / rstd for the backward), 2 = QuickGELU(x).
// These kernels are latency-bound (a few dozen rows): the row tile [mt][K] is fetched with cp.async (every 16-byte
// piece in flight at once, no register staging) while the warp's whole weight row streams into registers.
template <int KJ>   // K <= 128 * KJ: KJ = 4 keeps two blocks per SM resident (one wave for N <= 2368 columns)
__global__ void __launch_bounds__(OT_THREADS, KJ <= 4 ? 2 : 1)
ot_linear_fwd_kernel(const float* __restrict__ x, int x_mode, const float* __restrict__ ln_w,
                     const float* __restrict__ ln_b, float eps, float* __restrict__ xhat_out,
                     float* __restrict__ rstd_out, const float* __restrict__ W, const float* __restrict__ bias,
                     const float* __restrict__ resid, float* __restrict__ y, float* __restrict__ act_out, int M, int N,
                     int K, int mt) {
  extern __shared__ __align__(16) float xs[];   // [mt][K]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * mt;
  const int rows = min(mt, M - m0);
  const int n = blockIdx.x * 8 + warp;
  const int k4n = K >> 2;
  for (int i = threadIdx.x; i < rows * k4n; i += OT_THREADS) cp_async16(xs + 4 * i, x + (long long)m0 * K + 4 * i);
  float4 wreg[KJ];
#pragma unroll
  for (int j = 0; j < KJ; ++j)
    if (n < N && j * 128 < K) wreg[j] = __ldg(reinterpret_cast<const float4*>(W + (long long)n * K + j * 128 + lane * 4));
  float acc[OT_MT];
#pragma unroll
  for (int m = 0; m < OT_MT; ++m) acc[m] = 0.f;
  cp_async_wait_all();
  if (x_mode == 2)   // QuickGELU on the pieces this thread fetched itself (visible to it after the wait)
    for (int i = threadIdx.x; i < rows * k4n; i += OT_THREADS) {
      float4 v = *reinterpret_cast<float4*>(xs + 4 * i);
      v.x = quick_gelu(v.x), v.y = quick_gelu(v.y), v.z = quick_gelu(v.z), v.w = quick_gelu(v.w);
      *reinterpret_cast<float4*>(xs + 4 * i) = v;
    }
  __syncthreads();
  if (x_mode == 1) {                              // LayerNorm over the K features of each row, one warp per row
    for (int m = warp; m < rows; m += OT_THREADS / 32) {
      float* xr = xs + m * K;
      float s = 0.f;
      for (int k = lane; k < K; k += 32) s += xr[k];
      const float mean = warp_sum(s) / K;
      float q = 0.f;
      for (int k = lane; k < K; k += 32) {
        const float d = xr[k] - mean;
        q += d * d;
      }
      const float rstd = rsqrtf(warp_sum(q) / K + eps);
      for (int k = lane; k < K; k += 32) {
        const float xh = (xr[k] - mean) * rstd;
        xr[k] = xh * __ldg(ln_w + k) + __ldg(ln_b + k);
        if (blockIdx.x == 0 && xhat_out != nullptr) xhat_out[(long long)(m0 + m) * K + k] = xh;
      }
      if (blockIdx.x == 0 && lane == 0 && rstd_out != nullptr) rstd_out[m0 + m] = rstd;
    }
    __syncthreads();
  }
  if (n < N) {
#pragma unroll
    for (int j = 0; j < KJ; ++j) {
      if (j * 128 < K) {
        const float4 w4 = wreg[j];
#pragma unroll
        for (int m = 0; m < OT_MT; ++m)
          if (m < rows) {
            const float4 a4 = *reinterpret_cast<const float4*>(xs + m * K + j * 128 + lane * 4);
            acc[m] = fmaf(w4.x, a4.x, fmaf(w4.y, a4.y, fmaf(w4.z, a4.z, fmaf(w4.w, a4.w, acc[m]))));
          }
      }
    }
  }
  // transpose-reduce: afterwards lane m holds the sum over all lanes of acc[m]
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? acc[i] : acc[i + off];
      const float keep = up ? acc[i + off] : acc[i];
      acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  if (n < N && lane < rows) {
    const long long o = (long long)(m0 + lane) * N + n;
    float v = acc[0] + (bias != nullptr ? __ldg(bias + n) : 0.f);
    if (resid != nullptr) v += resid[o];
    y[o] = v;
    // QuickGELU(y) once, here, for the Linear that consumes it and for that Linear's dW -- not once per output column of
    // the consumer (x_mode 2) and once per weight row of its dW (a_mode 2), where the exponentials dominated the kernels
    if (act_out != nullptr) act_out[o] = quick_gelu(v);
  }
}

// ----------------------------------------------------------------------------------------------------- linear dX
// dA[m, k] = (sum_n dY[m, n] W[n, k]) * (pre ? QuickGELU'(pre[m, k]) : 1).
// grid (K / 32, ceil(M / 32), n_splits), 256 threads = 32 k-columns x 8 n-slices (warp = slice): a warp reads 128
// contiguous bytes of one weight row per step (16 rows in flight per thread), dY rows are broadcast from shared memory.
// With n_splits > 1 every block covers N / n_splits weight rows and the partial sums meet in fp32 atomics on a
// zeroed dA (only without `pre`): enough blocks to cover the L2 latency instead of 16 blocks walking 2048 rows each.
__global__ void __launch_bounds__(OT_THREADS)
ot_linear_dx_kernel(const float* __restrict__ dY, const float* __restrict__ W, const float* __restrict__ pre,
                    float* __restrict__ dA, int M, int N, int K, int n_per_split) {
  extern __shared__ __align__(16) float ys[];   // [OT_MT][OT_KC] chunk of dY; reused as [8][OT_MT][32] for the slice reduction
  const int slice = threadIdx.x >> 5, kcol = threadIdx.x & 31;
  const int m0 = blockIdx.y * OT_MT;
  const int rows = min(OT_MT, M - m0);
  const int k = blockIdx.x * 32 + kcol;
  const int n_begin = blockIdx.z * n_per_split;
  const int n_end = min(N, n_begin + n_per_split);
  float acc[OT_MT];
#pragma unroll
  for (int m = 0; m < OT_MT; ++m) acc[m] = 0.f;
  for (int nc = n_begin; nc < n_end; nc += OT_KC) {
    const int nw = min(OT_KC, n_end - nc);
    __syncthreads();
    for (int i = threadIdx.x; i < rows * (nw >> 2); i += OT_THREADS) {
      const int m = i / (nw >> 2), n4 = (i - m * (nw >> 2)) << 2;
      cp_async16(ys + m * OT_KC + n4, dY + (long long)(m0 + m) * N + nc + n4);
    }
    cp_async_wait_all();
    __syncthreads();
    if (k < K)
      for (int nb = slice; nb < nw; nb += 128) {
        float w[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) w[u] = nb + 8 * u < nw ? __ldg(W + (long long)(nc + nb + 8 * u) * K + k) : 0.f;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int n = min(nb + 8 * u, nw - 1);
#pragma unroll
          for (int m = 0; m < OT_MT; ++m)
            if (m < rows) acc[m] = fmaf(ys[m * OT_KC + n], w[u], acc[m]);
        }
      }
  }
  __syncthreads();
#pragma unroll
  for (int m = 0; m < OT_MT; ++m) ys[(slice * OT_MT + m) * 32 + kcol] = acc[m];
  __syncthreads();
  for (int m = slice; m < rows; m += 8) {
    float s = 0.f;
#pragma unroll
    for (int sl = 0; sl < 8; ++sl) s += ys[(sl * OT_MT + m) * 32 + kcol];
    if (k < K) {
      const long long o = (long long)(m0 + m) * K + k;
      if (gridDim.z > 1) {
        atomicAdd(dA + o, s);
      } else {
        if (pre != nullptr) s *= quick_gelu_grad(__ldg(pre + o));
        dA[o] = s;
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------------- linear dW / db
// dW[n, k] += sum_m dY[m, n] a[m, k],  db[n] += sum_m dY[m, n];  a = A | A * ln_w + ln_b (A = xhat) | QuickGELU(A).
// grid (K / 128, N / 8), 256 threads: warp = one weight row n, lane = 4 consecutive k.
__global__ void __launch_bounds__(OT_THREADS)
ot_linear_dw_kernel(const float* __restrict__ dY, const float* __restrict__ A, int a_mode,
                    const float* __restrict__ ln_w, const float* __restrict__ ln_b, float* __restrict__ dW,
                    float* __restrict__ db, int M, int N, int K) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.y * 8 + warp;
  const int k = blockIdx.x * 128 + lane * 4;
  if (n >= N || k >= K) return;
  float4 w4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (a_mode == 1) w4 = __ldg(reinterpret_cast<const float4*>(ln_w + k)), b4 = __ldg(reinterpret_cast<const float4*>(ln_b + k));
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float sb = 0.f;
  float4* dst = reinterpret_cast<float4*>(dW + (long long)n * K + k);
  const float4 cur0 = *dst;                        // in flight together with the first batch of rows
  for (int mb = 0; mb < M; mb += 8) {              // 8 rows per batch: 16 independent loads in flight per thread
    float dy[8];
    float4 a[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int m = min(mb + u, M - 1);
      dy[u] = mb + u < M ? __ldg(dY + (long long)m * N + n) : 0.f;
      a[u] = __ldg(reinterpret_cast<const float4*>(A + (long long)m * K + k));
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      float4 v = a[u];
      if (a_mode == 1) v.x = fmaf(v.x, w4.x, b4.x), v.y = fmaf(v.y, w4.y, b4.y), v.z = fmaf(v.z, w4.z, b4.z), v.w = fmaf(v.w, w4.w, b4.w);
      if (a_mode == 2) v.x = quick_gelu(v.x), v.y = quick_gelu(v.y), v.z = quick_gelu(v.z), v.w = quick_gelu(v.w);
      acc.x = fmaf(dy[u], v.x, acc.x), acc.y = fmaf(dy[u], v.y, acc.y), acc.z = fmaf(dy[u], v.z, acc.z), acc.w = fmaf(dy[u], v.w, acc.w);
      sb += dy[u];
    }
  }
  float4 cur = cur0;
  cur.x += acc.x, cur.y += acc.y, cur.z += acc.z, cur.w += acc.w;
  *dst = cur;
  if (db != nullptr && blockIdx.x == 0 && lane == 0) db[n] += sb;
}

// ----------------------------------------------------------------------------------------------------- LayerNorm backward
// dh[m] += rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dA * w;  dw += sum_m dA * xhat;  db += sum_m dA.
// grid M, 128 threads (one row per block; the parameter gradients meet in fp32 atomics, M is a few dozen).
__global__ void __launch_bounds__(128)
ot_ln_bwd_kernel(const float* __restrict__ dA, const float* __restrict__ xhat, const float* __restrict__ rstd,
                 const float* __restrict__ w, float* __restrict__ dh, float* __restrict__ dw, float* __restrict__ db,
                 int C) {
  __shared__ float red[2][4];
  const int m = blockIdx.x;
  const float* ga = dA + (long long)m * C;
  const float* xh = xhat + (long long)m * C;
  float s1 = 0.f, s2 = 0.f;
  for (int c = threadIdx.x; c < C; c += 128) {
    const float g = ga[c] * __ldg(w + c);
    s1 += g, s2 += g * xh[c];
  }
  s1 = warp_sum(s1), s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = s1, red[1][threadIdx.x >> 5] = s2;
  __syncthreads();
  const float c1 = (red[0][0] + red[0][1] + red[0][2] + red[0][3]) / C;
  const float c2 = (red[1][0] + red[1][1] + red[1][2] + red[1][3]) / C;
  const float r = rstd[m];
  for (int c = threadIdx.x; c < C; c += 128) {
    const float a = ga[c], x = xh[c];
    dh[(long long)m * C + c] += r * (a * __ldg(w + c) - c1 - x * c2);
    atomicAdd(dw + c, a * x);
    atomicAdd(db + c, a);
  }
}

// ----------------------------------------------------------------------------------------------------- attention
// nn.MultiheadAttention (tfm_model.py:36,46-48): heads of 64, q scaled by 1/8, key_padding_mask = (s >= pad_start[b]).
// One block per (b, head); S <= 16 tokens.  rows are b-major: token (b, s) = row b * S + s.
constexpr int OT_SMAX = 16, OT_HD = 64;

__global__ void __launch_bounds__(128)
ot_attn_fwd_kernel(const float* __restrict__ qkv, const long long* __restrict__ pad_start, float* __restrict__ probs,
                   float* __restrict__ o, int S, int H) {
  __shared__ float q[OT_SMAX][OT_HD + 1], k[OT_SMAX][OT_HD + 1], v[OT_SMAX][OT_HD + 1], p[OT_SMAX][OT_SMAX + 1];
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int C = H * OT_HD;
  const int ps = pad_start != nullptr ? static_cast<int>(pad_start[b]) : S;
  for (int i = threadIdx.x; i < S * OT_HD; i += 128) {
    const int s = i / OT_HD, d = i % OT_HD;
    const float* row = qkv + (long long)(b * S + s) * 3 * C + h * OT_HD + d;
    q[s][d] = row[0] * 0.125f, k[s][d] = row[C], v[s][d] = row[2 * C];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < S * S; t += 128) {
    const int i = t / S, j = t % S;
    float s = 0.f;
#pragma unroll 16
    for (int d = 0; d < OT_HD; ++d) s = fmaf(q[i][d], k[j][d], s);
    p[i][j] = j < ps ? s : -INFINITY;
  }
  __syncthreads();
  if (threadIdx.x < S) {
    const int i = threadIdx.x;
    float mx = -INFINITY;
    for (int j = 0; j < S; ++j) mx = fmaxf(mx, p[i][j]);
    float sum = 0.f;
    for (int j = 0; j < S; ++j) {
      const float e = __expf(p[i][j] - mx);
      p[i][j] = e, sum += e;
    }
    const float inv = 1.0f / sum;
    for (int j = 0; j < S; ++j) {
      p[i][j] *= inv;
      probs[((long long)blockIdx.x * S + i) * S + j] = p[i][j];
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < S * OT_HD; t += 128) {
    const int i = t / OT_HD, d = t % OT_HD;
    float s = 0.f;
    for (int j = 0; j < S; ++j) s = fmaf(p[i][j], v[j][d], s);
    o[(long long)(b * S + i) * C + h * OT_HD + d] = s;
  }
}

__global__ void __launch_bounds__(128)
ot_attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ probs, const float* __restrict__ dO,
                   float* __restrict__ dqkv, int S, int H) {
  __shared__ float q[OT_SMAX][OT_HD + 1], k[OT_SMAX][OT_HD + 1], v[OT_SMAX][OT_HD + 1], go[OT_SMAX][OT_HD + 1];
  __shared__ float p[OT_SMAX][OT_SMAX + 1], ds[OT_SMAX][OT_SMAX + 1];
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int C = H * OT_HD;
  for (int i = threadIdx.x; i < S * OT_HD; i += 128) {
    const int s = i / OT_HD, d = i % OT_HD;
    const float* row = qkv + (long long)(b * S + s) * 3 * C + h * OT_HD + d;
    q[s][d] = row[0] * 0.125f, k[s][d] = row[C], v[s][d] = row[2 * C];
    go[s][d] = dO[(long long)(b * S + s) * C + h * OT_HD + d];
  }
  for (int t = threadIdx.x; t < S * S; t += 128) p[t / S][t % S] = probs[(long long)blockIdx.x * S * S + t];
  __syncthreads();
  for (int t = threadIdx.x; t < S * S; t += 128) {   // dP
    const int i = t / S, j = t % S;
    float s = 0.f;
#pragma unroll 16
    for (int d = 0; d < OT_HD; ++d) s = fmaf(go[i][d], v[j][d], s);
    ds[i][j] = s;
  }
  __syncthreads();
  if (threadIdx.x < S) {                              // dS = P * (dP - sum_j P dP)
    const int i = threadIdx.x;
    float dot = 0.f;
    for (int j = 0; j < S; ++j) dot = fmaf(p[i][j], ds[i][j], dot);
    for (int j = 0; j < S; ++j) ds[i][j] = p[i][j] * (ds[i][j] - dot);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < S * OT_HD; t += 128) {
    const int i = t / OT_HD, d = t % OT_HD;
    float dq = 0.f, dk = 0.f, dv = 0.f;
    for (int j = 0; j < S; ++j) {
      dq = fmaf(ds[i][j], k[j][d], dq);             // d(q scaled) -> * 1/8 below
      dk = fmaf(ds[j][i], q[j][d], dk);             // q already carries the 1/8
      dv = fmaf(p[j][i], go[j][d], dv);
    }
    float* row = dqkv + (long long)(b * S + i) * 3 * C + h * OT_HD + d;
    row[0] = dq * 0.125f, row[C] = dk, row[2 * C] = dv;
  }
}

// ----------------------------------------------------------------------------------------------------- level input
// h[(b, s)] = in + type_emb[s == mask_b] + pos_emb[s] + tvec, in = noisy_b (s == mask_b) | pad_emb (s >= pad_start_b) |
// video_emb[(b, s)], noisy_b = ca * src_b + cb * noise_b   (tfm_model.py:171-191, ennoise :291-302).
__global__ void ot_embed_fwd_kernel(const float* __restrict__ video, const float* __restrict__ src,
                                    const float* __restrict__ noise, float ca, float cb,
                                    const long long* __restrict__ mask_inds, const long long* __restrict__ pad_start,
                                    const float* __restrict__ type_w, const float* __restrict__ pos_w,
                                    const float* __restrict__ pad_w, const float* __restrict__ tvec,
                                    float* __restrict__ h, int S, int C) {
  const int row = blockIdx.x, b = row / S, s = row % S;
  const bool is_mask = s == static_cast<int>(mask_inds[b]);
  const bool is_pad = s >= static_cast<int>(pad_start[b]);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float in;
    if (is_mask) in = ca * src[(long long)b * C + c] + cb * noise[(long long)b * C + c];
    else if (is_pad) in = pad_w[c];
    else in = video[(long long)row * C + c];
    h[(long long)row * C + c] = in + type_w[(is_mask ? C : 0) + c] + pos_w[(long long)s * C + c] + tvec[c];
  }
}

// one thread per feature column; loops over the B*S rows (a few dozen)
__global__ void ot_embed_bwd_kernel(const float* __restrict__ dh, const long long* __restrict__ mask_inds,
                                    const long long* __restrict__ pad_start, float* __restrict__ dvideo,
                                    float* __restrict__ dtype, float* __restrict__ dpos, float* __restrict__ dpad,
                                    float* __restrict__ dtvec, int B, int S, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float t0 = 0.f, t1 = 0.f, pd = 0.f, tv = 0.f;
  for (int b = 0; b < B; ++b) {
    const int mk = static_cast<int>(mask_inds[b]), ps = static_cast<int>(pad_start[b]);
    for (int s = 0; s < S; ++s) {
      const float g = dh[(long long)(b * S + s) * C + c];
      tv += g;
      dpos[(long long)s * C + c] += g;
      if (s == mk) t1 += g;
      else {
        t0 += g;
        if (s >= ps) pd += g;
        else dvideo[(long long)(b * S + s) * C + c] += g;
      }
    }
  }
  dtype[c] += t0, dtype[C + c] += t1, dpad[c] += pd, dtvec[c] = tv;
}

}  // namespace
}  // namespace pvrl

using namespace pvrl;
#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int pvrl_ot_linear_fwd(const float* x, int32_t x_mode, const float* ln_w, const float* ln_b, float eps,
                                  float* xhat_out, float* rstd_out, const float* W, const float* bias,
                                  const float* resid, float* y, float* act_out, int32_t M, int32_t N, int32_t K,
                                  void* stream) {
  PVRL_CHECK_ARG(x && W && y && M > 0 && N > 0 && K > 0, "pvrl_ot_linear_fwd: bad arguments");
  PVRL_CHECK_ARG(K % 128 == 0 && K <= OT_KMAX, "pvrl_ot_linear_fwd: K=%d must be a multiple of 128, <= 2048", K);
  PVRL_CHECK_ARG(x_mode >= 0 && x_mode <= 2, "pvrl_ot_linear_fwd: bad x_mode %d", x_mode);
  if (x_mode == 1) PVRL_CHECK_ARG(ln_w && ln_b, "pvrl_ot_linear_fwd: the LayerNorm prologue needs ln_w / ln_b");
  // rows per block tile: as many as fit next to each other in 200 KB of shared memory (32 for K <= 1536, 25 for K = 2048)
  int mt = (200 * 1024) / (K * 4);
  mt = mt > OT_MT ? OT_MT : mt;
  mt = mt > M ? M : mt;
  const int smem = mt * K * 4;
  static bool configured = false;
  if (!configured) {
    PVRL_CUDA(cudaFuncSetAttribute(ot_linear_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    PVRL_CUDA(cudaFuncSetAttribute(ot_linear_fwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  dim3 grid((N + 7) / 8, (M + mt - 1) / mt);
  if (K <= 512)
    ot_linear_fwd_kernel<4><<<grid, OT_THREADS, smem, STREAM>>>(x, x_mode, ln_w, ln_b, eps, xhat_out, rstd_out, W, bias,
                                                                resid, y, act_out, M, N, K, mt);
  else
    ot_linear_fwd_kernel<16><<<grid, OT_THREADS, smem, STREAM>>>(x, x_mode, ln_w, ln_b, eps, xhat_out, rstd_out, W, bias,
                                                                 resid, y, act_out, M, N, K, mt);
  return launched("ot_linear_fwd_kernel");
}

extern "C" int pvrl_ot_linear_dx(const float* dY, const float* W, const float* pre, float* dA, int32_t M, int32_t N,
                                 int32_t K, void* stream) {
  PVRL_CHECK_ARG(dY && W && dA && M > 0 && N > 0 && K > 0, "pvrl_ot_linear_dx: bad arguments");
  PVRL_CHECK_ARG(N % 4 == 0, "pvrl_ot_linear_dx: N=%d must be a multiple of 4", N);
  const int smem = OT_MT * OT_KC * 4;
  static bool configured = false;
  if (!configured) {
    PVRL_CUDA(cudaFuncSetAttribute(ot_linear_dx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  // split N across blocks until ~256 blocks are in flight (never with the QuickGELU' epilogue, which needs full sums)
  const int kb = (K + 31) / 32, mb = (M + OT_MT - 1) / OT_MT;
  int splits = 1;
  if (pre == nullptr)
    while (kb * mb * splits < 256 && N / (splits * 2) >= 128 && (N % (splits * 2 * 4)) == 0) splits *= 2;
  const int n_per_split = (N + splits - 1) / splits;
  if (splits > 1) PVRL_CUDA(cudaMemsetAsync(dA, 0, sizeof(float) * (size_t)M * K, STREAM));
  dim3 grid(kb, mb, splits);
  ot_linear_dx_kernel<<<grid, OT_THREADS, smem, STREAM>>>(dY, W, pre, dA, M, N, K, n_per_split);
  return launched("ot_linear_dx_kernel");
}

extern "C" int pvrl_ot_linear_dw(const float* dY, const float* A, int32_t a_mode, const float* ln_w, const float* ln_b,
                                 float* dW, float* db, int32_t M, int32_t N, int32_t K, void* stream) {
  PVRL_CHECK_ARG(dY && A && dW && M > 0 && N > 0 && K > 0, "pvrl_ot_linear_dw: bad arguments");
  PVRL_CHECK_ARG(K % 4 == 0, "pvrl_ot_linear_dw: K=%d must be a multiple of 4", K);
  PVRL_CHECK_ARG(a_mode >= 0 && a_mode <= 2 && (a_mode != 1 || (ln_w && ln_b)), "pvrl_ot_linear_dw: bad a_mode");
  dim3 grid((K + 127) / 128, (N + 7) / 8);
  ot_linear_dw_kernel<<<grid, OT_THREADS, 0, STREAM>>>(dY, A, a_mode, ln_w, ln_b, dW, db, M, N, K);
  return launched("ot_linear_dw_kernel");
}

extern "C" int pvrl_ot_ln_bwd(const float* dA, const float* xhat, const float* rstd, const float* w, float* dh,
                              float* dw, float* db, int32_t M, int32_t C, void* stream) {
  PVRL_CHECK_ARG(dA && xhat && rstd && w && dh && dw && db && M > 0 && C > 0, "pvrl_ot_ln_bwd: bad arguments");
  ot_ln_bwd_kernel<<<M, 128, 0, STREAM>>>(dA, xhat, rstd, w, dh, dw, db, C);
  return launched("ot_ln_bwd_kernel");
}

extern "C" int pvrl_ot_attn_fwd(const float* qkv, const int64_t* pad_start, float* probs, float* o, int32_t B, int32_t S,
                                int32_t H, void* stream) {
  PVRL_CHECK_ARG(qkv && probs && o && B > 0 && H > 0, "pvrl_ot_attn_fwd: bad arguments");
  PVRL_CHECK_ARG(S > 0 && S <= OT_SMAX, "pvrl_ot_attn_fwd: S=%d must be in [1, 16]", S);
  ot_attn_fwd_kernel<<<B * H, 128, 0, STREAM>>>(qkv, reinterpret_cast<const long long*>(pad_start), probs, o, S, H);
  return launched("ot_attn_fwd_kernel");
}

extern "C" int pvrl_ot_attn_bwd(const float* qkv, const float* probs, const float* dO, float* dqkv, int32_t B, int32_t S,
                                int32_t H, void* stream) {
  PVRL_CHECK_ARG(qkv && probs && dO && dqkv && B > 0 && H > 0, "pvrl_ot_attn_bwd: bad arguments");
  PVRL_CHECK_ARG(S > 0 && S <= OT_SMAX, "pvrl_ot_attn_bwd: S=%d must be in [1, 16]", S);
  ot_attn_bwd_kernel<<<B * H, 128, 0, STREAM>>>(qkv, probs, dO, dqkv, S, H);
  return launched("ot_attn_bwd_kernel");
}

extern "C" int pvrl_ot_embed_fwd(const float* video, const float* src, const float* noise, float ca, float cb,
                                 const int64_t* mask_inds, const int64_t* pad_start, const float* type_w,
                                 const float* pos_w, const float* pad_w, const float* tvec, float* h, int32_t B,
                                 int32_t S, int32_t C, void* stream) {
  PVRL_CHECK_ARG(video && src && noise && mask_inds && pad_start && type_w && pos_w && pad_w && tvec && h && B > 0 &&
                     S > 0 && C > 0, "pvrl_ot_embed_fwd: bad arguments");
  ot_embed_fwd_kernel<<<B * S, 128, 0, STREAM>>>(video, src, noise, ca, cb, reinterpret_cast<const long long*>(mask_inds),
                                                 reinterpret_cast<const long long*>(pad_start), type_w, pos_w, pad_w,
                                                 tvec, h, S, C);
  return launched("ot_embed_fwd_kernel");
}

extern "C" int pvrl_ot_embed_bwd(const float* dh, const int64_t* mask_inds, const int64_t* pad_start, float* dvideo,
                                 float* dtype, float* dpos, float* dpad, float* dtvec, int32_t B, int32_t S, int32_t C,
                                 void* stream) {
  PVRL_CHECK_ARG(dh && mask_inds && pad_start && dvideo && dtype && dpos && dpad && dtvec && B > 0 && S > 0 && C > 0,
                 "pvrl_ot_embed_bwd: bad arguments");
  ot_embed_bwd_kernel<<<(C + 127) / 128, 128, 0, STREAM>>>(dh, reinterpret_cast<const long long*>(mask_inds),
                                                           reinterpret_cast<const long long*>(pad_start), dvideo, dtype,
                                                           dpos, dpad, dtvec, B, S, C);
  return launched("ot_embed_bwd_kernel");
}
