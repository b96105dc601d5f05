// MViTv2 video encoder (BASELINE config 5): the ops around its Linear layers -- first sm_100a path.
//
//   pvrl_ln_any_fwd / _bwd          LayerNorm over any width <= 1024 (96 / 192 / 384 / 768 here)      attention.py:239-280,530,556
//   pvrl_pool3d_fwd / _bwd          depth-wise Conv3d pooling of Q / K / V per head, cls bypass         attention.py:14-48
//   pvrl_maxpool3d_fwd / _bwd       MaxPool3d of the skip path, cls bypass                              attention.py:521-543
//   pvrl_im2col3d                   rows of the (3,7,7)/(2,4,4) Conv3d stem for the tcgen05 GEMM        stem_helper.py:290-322
//   pvrl_pooled_attn_fwd / _bwd     softmax(q k^T * scale + decomposed rel-pos bias) v + q (residual pooling),
//                                   queries up to 25 089, 393 / 1 569 pooled keys, 96-wide heads          attention.py:282-411
//
// Every kernel here is a plain CUDA-core kernel: a warp owns a row (LayerNorm, pooling) or a few queries / keys
// (attention, flash-style: nothing of size queries x keys ever reaches HBM).  The 75 % of the encoder's FLOPs that are
// Linear layers run on the tcgen05 GEMM (tc_functional.py); the attention contractions (23 %) are the part a later round
// moves to tcgen05 -- these kernels are then the parity baseline for it.
#include <cuda_bf16.h>

#include <cfloat>

#include "pvrl_host.h"

namespace pvrl {
namespace {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(bf16* p, float v) { *p = __float2bfloat16(v); }

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 8 consecutive elements of a row -> fp32 (one 16-byte load for bf16, two for fp32); p must be 16-byte aligned.
__device__ __forceinline__ void ld8(const bf16* p, float* f) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x, f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void ld8(const float* p, float* f) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x, f[1] = a.y, f[2] = a.z, f[3] = a.w, f[4] = b.x, f[5] = b.y, f[6] = b.z, f[7] = b.w;
}

// ------------------------------------------------------------------------------------------------ LayerNorm, any width
constexpr int LNA_MAXV = 32;   // D <= 32 * LNA_MAXV

template <typename InT, typename OutT>
__global__ void __launch_bounds__(256) ln_any_fwd_kernel(const InT* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ b, OutT* __restrict__ y,
                                                         float* __restrict__ stats, int M, int D, float eps) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int m = warp; m < M; m += nwarps) {
    const InT* xr = x + (size_t)m * D;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s += ldf(xr + c);
    const float mean = wsum(s) / D;
    float q = 0.f;
    for (int c = lane; c < D; c += 32) {
      const float d = ldf(xr + c) - mean;
      q += d * d;
    }
    const float rstd = 1.0f / sqrtf(wsum(q) / D + eps);
    OutT* yr = y + (size_t)m * D;
    for (int c = lane; c < D; c += 32) stf(yr + c, (ldf(xr + c) - mean) * rstd * w[c] + b[c]);
    if (lane == 0 && stats != nullptr) stats[2 * m] = mean, stats[2 * m + 1] = rstd;
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * w;  dw += sum_rows dy * xhat, db += sum_rows dy (atomics).
template <typename DyT, typename InT, typename DxT>
__global__ void __launch_bounds__(256) ln_any_bwd_kernel(const DyT* __restrict__ dy, const InT* __restrict__ x,
                                                         const float* __restrict__ w, const float* __restrict__ stats,
                                                         DxT* __restrict__ dx, float* __restrict__ dw,
                                                         float* __restrict__ db, int M, int D) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  float aw[LNA_MAXV], ab[LNA_MAXV];
#pragma unroll
  for (int i = 0; i < LNA_MAXV; ++i) aw[i] = 0.f, ab[i] = 0.f;
  for (int m = warp; m < M; m += nwarps) {
    const float mean = stats[2 * m], rstd = stats[2 * m + 1];
    const DyT* dyr = dy + (size_t)m * D;
    const InT* xr = x + (size_t)m * D;
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < D; c += 32) {
      const float g = ldf(dyr + c) * w[c], xh = (ldf(xr + c) - mean) * rstd;
      s1 += g * xh, s2 += g;
    }
    const float c1 = wsum(s1) / D, c2 = wsum(s2) / D;
    DxT* dxr = dx + (size_t)m * D;
#pragma unroll
    for (int i = 0; i < LNA_MAXV; ++i) {
      const int c = lane + 32 * i;
      if (c < D) {
        const float d = ldf(dyr + c), xh = (ldf(xr + c) - mean) * rstd;
        stf(dxr + c, rstd * (d * w[c] - c2 - xh * c1));
        aw[i] += d * xh, ab[i] += d;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < LNA_MAXV; ++i) {
    const int c = lane + 32 * i;
    if (c < D) {
      if (dw != nullptr) atomicAdd(dw + c, aw[i]);
      if (db != nullptr) atomicAdd(db + c, ab[i]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ depth-wise Conv3d pooling
struct PoolGeom {
  int B, heads, C;
  int T, H, W;      // input grid
  int KT, KH, KW;   // kernel
  int ST, SH, SW;   // stride
  int PT, PH, PW;   // padding
  int OT, OH, OW;   // output grid
  long long ld;     // row pitch (elements) of the token-major buffer [B, 1 + T*H*W, ld]; head h owns columns [h*C, h*C + C)
};
constexpr int POOL_MAXV = 4;   // C <= 128

// out[b, h, 0] = in[b, 0, h*C..]  (cls bypass);  out[b, h, 1 + o] = sum_k w[c, k] * in[b, 1 + window_k(o), h*C + c]
// (cross-correlation with zero padding, as nn.Conv3d(groups = C, bias = False)).  w == nullptr: plain re-layout (no pooling).
template <typename T>
__global__ void __launch_bounds__(256) pool3d_fwd_kernel(const T* __restrict__ in, const float* __restrict__ w,
                                                         T* __restrict__ out, PoolGeom g) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int L = g.T * g.H * g.W, Lo = g.OT * g.OH * g.OW, KK = g.KT * g.KH * g.KW;
  const long long total = (long long)g.B * g.heads * (1 + Lo);
  for (long long idx = warp; idx < total; idx += nwarps) {
    const int o = (int)(idx % (1 + Lo));
    const int bh = (int)(idx / (1 + Lo)), h = bh % g.heads, b = bh / g.heads;
    T* dst = out + idx * g.C;
    const T* src = in + (size_t)b * (1 + L) * g.ld + (size_t)h * g.C;
    if (o == 0 || w == nullptr) {
      for (int c = lane; c < g.C; c += 32) dst[c] = src[(size_t)o * g.ld + c];
      continue;
    }
    const int ow = (o - 1) % g.OW, oh = ((o - 1) / g.OW) % g.OH, ot = (o - 1) / (g.OW * g.OH);
    float acc[POOL_MAXV];
#pragma unroll
    for (int i = 0; i < POOL_MAXV; ++i) acc[i] = 0.f;
    for (int dt = 0; dt < g.KT; ++dt) {
      const int it = ot * g.ST - g.PT + dt;
      if (it < 0 || it >= g.T) continue;
      for (int dh = 0; dh < g.KH; ++dh) {
        const int ih = oh * g.SH - g.PH + dh;
        if (ih < 0 || ih >= g.H) continue;
        for (int dw_ = 0; dw_ < g.KW; ++dw_) {
          const int iw = ow * g.SW - g.PW + dw_;
          if (iw < 0 || iw >= g.W) continue;
          const T* row = src + (size_t)(1 + (it * g.H + ih) * g.W + iw) * g.ld;
          const int k = (dt * g.KH + dh) * g.KW + dw_;
#pragma unroll
          for (int i = 0; i < POOL_MAXV; ++i) {
            const int c = lane + 32 * i;
            if (c < g.C) acc[i] += __ldg(w + c * KK + k) * ldf(row + c);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < POOL_MAXV; ++i) {
      const int c = lane + 32 * i;
      if (c < g.C) stf(dst + c, acc[i]);
    }
  }
}

// Input gradient, gather form: din[b, 1 + i, h*C + c] = sum over the outputs o whose window holds i of w[c, k(o, i)] * dout[b, h, 1 + o, c].
template <typename T>
__global__ void __launch_bounds__(256) pool3d_bwd_in_kernel(const T* __restrict__ dout, const float* __restrict__ w,
                                                            T* __restrict__ din, PoolGeom g) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int L = g.T * g.H * g.W, Lo = g.OT * g.OH * g.OW, KK = g.KT * g.KH * g.KW;
  const long long total = (long long)g.B * g.heads * (1 + L);
  for (long long idx = warp; idx < total; idx += nwarps) {
    const int i = (int)(idx % (1 + L));
    const int bh = (int)(idx / (1 + L)), h = bh % g.heads, b = bh / g.heads;
    T* dst = din + ((size_t)b * (1 + L) + i) * g.ld + (size_t)h * g.C;
    const T* src = dout + (size_t)bh * (1 + Lo) * g.C;
    if (i == 0 || w == nullptr) {
      for (int c = lane; c < g.C; c += 32) dst[c] = src[(size_t)i * g.C + c];
      continue;
    }
    const int iw = (i - 1) % g.W, ih = ((i - 1) / g.W) % g.H, it = (i - 1) / (g.W * g.H);
    float acc[POOL_MAXV];
#pragma unroll
    for (int v = 0; v < POOL_MAXV; ++v) acc[v] = 0.f;
    for (int dt = 0; dt < g.KT; ++dt) {
      const int nt = it + g.PT - dt;
      if (nt < 0 || nt % g.ST != 0 || nt / g.ST >= g.OT) continue;
      for (int dh = 0; dh < g.KH; ++dh) {
        const int nh = ih + g.PH - dh;
        if (nh < 0 || nh % g.SH != 0 || nh / g.SH >= g.OH) continue;
        for (int dw_ = 0; dw_ < g.KW; ++dw_) {
          const int nw = iw + g.PW - dw_;
          if (nw < 0 || nw % g.SW != 0 || nw / g.SW >= g.OW) continue;
          const T* row = src + (size_t)(1 + ((nt / g.ST) * g.OH + nh / g.SH) * g.OW + nw / g.SW) * g.C;
          const int k = (dt * g.KH + dh) * g.KW + dw_;
#pragma unroll
          for (int v = 0; v < POOL_MAXV; ++v) {
            const int c = lane + 32 * v;
            if (c < g.C) acc[v] += __ldg(w + c * KK + k) * ldf(row + c);
          }
        }
      }
    }
#pragma unroll
    for (int v = 0; v < POOL_MAXV; ++v) {
      const int c = lane + 32 * v;
      if (c < g.C) stf(dst + c, acc[v]);
    }
  }
}

// Weight gradient dw[c, k] += sum_{b, h, o} dout[b, h, 1 + o, c] * in[b, 1 + window_k(o), h*C + c] for 3 x 3 x 3 kernels: per-lane
// partial sums in registers, one shared-memory reduction per block, one global atomic per (block, element).
template <typename T>
__global__ void __launch_bounds__(256) pool3d_bwd_w_kernel(const T* __restrict__ dout, const T* __restrict__ in,
                                                           float* __restrict__ dw, PoolGeom g) {
  extern __shared__ float sacc[];   // [C * 27]
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int L = g.T * g.H * g.W, Lo = g.OT * g.OH * g.OW;
  for (int e = threadIdx.x; e < g.C * 27; e += blockDim.x) sacc[e] = 0.f;
  __syncthreads();
  float acc[POOL_MAXV][27];
#pragma unroll
  for (int v = 0; v < POOL_MAXV; ++v)
#pragma unroll
    for (int k = 0; k < 27; ++k) acc[v][k] = 0.f;
  const long long total = (long long)g.B * g.heads * Lo;
  for (long long idx = warp; idx < total; idx += nwarps) {
    const int o = (int)(idx % Lo);
    const int bh = (int)(idx / Lo), h = bh % g.heads, b = bh / g.heads;
    const T* drow = dout + ((size_t)bh * (1 + Lo) + 1 + o) * g.C;
    const T* src = in + (size_t)b * (1 + L) * g.ld + (size_t)h * g.C;
    const int ow = o % g.OW, oh = (o / g.OW) % g.OH, ot = o / (g.OW * g.OH);
    float d[POOL_MAXV];
#pragma unroll
    for (int v = 0; v < POOL_MAXV; ++v) d[v] = (lane + 32 * v < g.C) ? ldf(drow + lane + 32 * v) : 0.f;
#pragma unroll
    for (int dt = 0; dt < 3; ++dt) {
      const int it = ot * g.ST - g.PT + dt;
      if (it < 0 || it >= g.T) continue;
#pragma unroll
      for (int dh = 0; dh < 3; ++dh) {
        const int ih = oh * g.SH - g.PH + dh;
        if (ih < 0 || ih >= g.H) continue;
#pragma unroll
        for (int dw_ = 0; dw_ < 3; ++dw_) {
          const int iw = ow * g.SW - g.PW + dw_;
          if (iw < 0 || iw >= g.W) continue;
          const T* row = src + (size_t)(1 + (it * g.H + ih) * g.W + iw) * g.ld;
#pragma unroll
          for (int v = 0; v < POOL_MAXV; ++v)
            if (lane + 32 * v < g.C) acc[v][(dt * 3 + dh) * 3 + dw_] += d[v] * ldf(row + lane + 32 * v);
        }
      }
    }
  }
#pragma unroll
  for (int v = 0; v < POOL_MAXV; ++v)
#pragma unroll
    for (int k = 0; k < 27; ++k)
      if (lane + 32 * v < g.C) atomicAdd(&sacc[(lane + 32 * v) * 27 + k], acc[v][k]);
  __syncthreads();
  for (int e = threadIdx.x; e < g.C * 27; e += blockDim.x) atomicAdd(dw + e, sacc[e]);
}

// Variant (default since its hardware run; PVRL_POOL_DW = 1 selects the kernel above): the same sums with ONE channel per lane -- blockIdx.y picks the
// 32-channel group -- so a lane carries 27 accumulators instead of 108 (254 registers, one block per SM, 310 us per launch
// in the first launch list): ~5x the resident warps for the same loads.
template <typename T>
__global__ void __launch_bounds__(256) pool3d_bwd_w_cg_kernel(const T* __restrict__ dout, const T* __restrict__ in,
                                                              float* __restrict__ dw, PoolGeom g) {
  __shared__ float sacc[32 * 27];
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.y * 32 + lane;                      // this lane's channel
  const bool live = c < g.C;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int L = g.T * g.H * g.W, Lo = g.OT * g.OH * g.OW;
  for (int e = threadIdx.x; e < 32 * 27; e += blockDim.x) sacc[e] = 0.f;
  __syncthreads();
  float acc[27];
#pragma unroll
  for (int k = 0; k < 27; ++k) acc[k] = 0.f;
  const long long total = (long long)g.B * g.heads * Lo;
  for (long long idx = warp; idx < total; idx += nwarps) {
    const int o = (int)(idx % Lo);
    const int bh = (int)(idx / Lo), h = bh % g.heads, b = bh / g.heads;
    const T* drow = dout + ((size_t)bh * (1 + Lo) + 1 + o) * g.C;
    const T* src = in + (size_t)b * (1 + L) * g.ld + (size_t)h * g.C;
    const int ow = o % g.OW, oh = (o / g.OW) % g.OH, ot = o / (g.OW * g.OH);
    const float d = live ? ldf(drow + c) : 0.f;
#pragma unroll
    for (int dt = 0; dt < 3; ++dt) {
      const int it = ot * g.ST - g.PT + dt;
      if (it < 0 || it >= g.T) continue;
#pragma unroll
      for (int dh = 0; dh < 3; ++dh) {
        const int ih = oh * g.SH - g.PH + dh;
        if (ih < 0 || ih >= g.H) continue;
#pragma unroll
        for (int dw_ = 0; dw_ < 3; ++dw_) {
          const int iw = ow * g.SW - g.PW + dw_;
          if (iw < 0 || iw >= g.W) continue;
          if (live) acc[(dt * 3 + dh) * 3 + dw_] += d * ldf(src + (size_t)(1 + (it * g.H + ih) * g.W + iw) * g.ld + c);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 27; ++k) atomicAdd(&sacc[lane * 27 + k], acc[k]);
  __syncthreads();
  for (int e = threadIdx.x; e < 32 * 27; e += blockDim.x) {
    const int cc = blockIdx.y * 32 + e / 27;
    if (cc < g.C) atomicAdd(dw + (size_t)cc * 27 + e % 27, sacc[e]);
  }
}

// ------------------------------------------------------------------------------------------------ MaxPool3d skip
// x [B, 1 + L, D] -> y [B, 1 + Lo, D] (cls row copied), arg [B, Lo, D] = the input token (0 .. L-1) that won each window:
// the first maximum in (t, h, w) scan order, as ATen's max_pool3d.
template <typename T>
__global__ void __launch_bounds__(256) maxpool3d_fwd_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                            int32_t* __restrict__ arg, PoolGeom g) {
  const int D = g.C, L = g.T * g.H * g.W, Lo = g.OT * g.OH * g.OW;
  const long long total = (long long)g.B * (1 + Lo) * D;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % D);
    const long long r = e / D;
    const int o = (int)(r % (1 + Lo)), b = (int)(r / (1 + Lo));
    const T* xb = x + (size_t)b * (1 + L) * D + c;
    if (o == 0) {
      y[e] = xb[0];
      continue;
    }
    const int ow = (o - 1) % g.OW, oh = ((o - 1) / g.OW) % g.OH, ot = (o - 1) / (g.OW * g.OH);
    float best = -FLT_MAX;
    int bi = -1;
    for (int dt = 0; dt < g.KT; ++dt) {
      const int it = ot * g.ST - g.PT + dt;
      if (it < 0 || it >= g.T) continue;
      for (int dh = 0; dh < g.KH; ++dh) {
        const int ih = oh * g.SH - g.PH + dh;
        if (ih < 0 || ih >= g.H) continue;
        for (int dw_ = 0; dw_ < g.KW; ++dw_) {
          const int iw = ow * g.SW - g.PW + dw_;
          if (iw < 0 || iw >= g.W) continue;
          const int i = (it * g.H + ih) * g.W + iw;
          const float v = ldf(xb + (size_t)(1 + i) * D);
          if (bi < 0 || v > best || v != v) best = v, bi = i;
        }
      }
    }
    stf(y + e, best);
    arg[((size_t)b * Lo + (o - 1)) * D + c] = bi;
  }
}

// dx (fp32, zero-initialised by the caller) += scatter of dy through arg; windows overlap, hence the atomics.
template <typename T>
__global__ void __launch_bounds__(256) maxpool3d_bwd_kernel(const T* __restrict__ dy, const int32_t* __restrict__ arg,
                                                            float* __restrict__ dx, PoolGeom g) {
  const int D = g.C, L = g.T * g.H * g.W, Lo = g.OT * g.OH * g.OW;
  const long long total = (long long)g.B * (1 + Lo) * D;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % D);
    const long long r = e / D;
    const int o = (int)(r % (1 + Lo)), b = (int)(r / (1 + Lo));
    float* dxb = dx + (size_t)b * (1 + L) * D + c;
    const float d = ldf(dy + e);
    if (o == 0) {
      dxb[0] = d;
      continue;
    }
    const int i = arg[((size_t)b * Lo + (o - 1)) * D + c];
    if (i >= 0) atomicAdd(dxb + (size_t)(1 + i) * D, d);
  }
}

// ------------------------------------------------------------------------------------------------ Conv3d stem im2col
// frames fp32 [B, Cin, T, H, W] -> rows [B * OT * OH * OW, Kpad]: column ((c * KT + dt) * KH + dh) * KW + dw (the order of
// Conv3d's weight.reshape(D, -1)), zero beyond Cin * KT * KH * KW and where the window leaves the clip.
template <typename T>
__global__ void __launch_bounds__(256) im2col3d_kernel(const float* __restrict__ frames, T* __restrict__ out, PoolGeom g,
                                                       int Cin, int Kpad) {
  const int Lo = g.OT * g.OH * g.OW, KK = g.KT * g.KH * g.KW;
  const long long total = (long long)g.B * Lo * Kpad;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(e % Kpad);
    const long long r = e / Kpad;
    const int o = (int)(r % Lo), b = (int)(r / Lo);
    float v = 0.f;
    if (k < Cin * KK) {
      const int c = k / KK, kk = k - c * KK;
      const int dw_ = kk % g.KW, dh = (kk / g.KW) % g.KH, dt = kk / (g.KW * g.KH);
      const int ow = o % g.OW, oh = (o / g.OW) % g.OH, ot = o / (g.OW * g.OH);
      const int it = ot * g.ST - g.PT + dt, ih = oh * g.SH - g.PH + dh, iw = ow * g.SW - g.PW + dw_;
      if (it >= 0 && it < g.T && ih >= 0 && ih < g.H && iw >= 0 && iw < g.W)
        v = frames[((((size_t)b * Cin + c) * g.T + it) * g.H + ih) * g.W + iw];
    }
    stf(out + e, v);
  }
}

// ------------------------------------------------------------------------------------------------ pooled attention
// q [BH, Nq, C], k / v [BH, Nk, C] (BH = clips x heads, row 0 = cls), bq [BH, Nq - 1, KB] fp32 = the query's projections on
// the relative-position tables, KB = Kt + Kh + Kw: the bias of (query i > 0, key j > 0, j - 1 = (kt, kh, kw)) is
// bq[i-1, kt] + bq[i-1, Kt + kh] + bq[i-1, Kt + Kh + kw] (attention.py:136-159); cls row / cls column carry none.
struct AttnGeom {
  int BH, heads, Nq, Nk;
  int Kt, Kh, Kw;
  float scale;
  int resid;     // residual pooling: out[i > 0] += q[i]  (attention.py:397-401)
  int qsplit;    // dKV kernel: the query range is cut into this many slices (gridDim.z)
};
constexpr int AT_C = 96;      // head width of every MViTv2 stage
constexpr int AT_V = AT_C / 32;
constexpr int AT_QPW = 4;     // queries (dQ) / keys (dKV) per warp
constexpr int AT_WARPS = 4;
constexpr int AT_KBMAX = 64;
#ifndef PVRL_MVIT_ATTN_MMA_DEFAULT
#define PVRL_MVIT_ATTN_MMA_DEFAULT 1   // bf16 forward on mma.sync (122 vs 618 us on the 9 x 4 x 1569 x 393 blocks); 0 = CUDA cores
#endif

__device__ __forceinline__ float attn_bias(const float* __restrict__ bqrow, int j, const AttnGeom& g) {
  const int jj = j - 1;
  const int kw = jj % g.Kw, kh = (jj / g.Kw) % g.Kh, kt = jj / (g.Kw * g.Kh);
  return __ldg(bqrow + kt) + __ldg(bqrow + g.Kt + kh) + __ldg(bqrow + g.Kt + g.Kh + kw);
}

// out [B, Nq, heads * C] (the layout the projection GEMM reads), lse [BH, Nq].
template <typename T>
__global__ void __launch_bounds__(AT_WARPS * 32) pooled_attn_fwd_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                                        const T* __restrict__ v,
                                                                        const float* __restrict__ bq, T* __restrict__ out,
                                                                        float* __restrict__ lse, AttnGeom g) {
  __shared__ __align__(16) float qs[AT_WARPS][AT_QPW][AT_C];
  __shared__ float ps[AT_WARPS][AT_QPW][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int bh = blockIdx.y, KB = g.Kt + g.Kh + g.Kw;
  const int i0 = (blockIdx.x * AT_WARPS + wid) * AT_QPW;
  if (i0 >= g.Nq) return;
  const T* qb = q + (size_t)bh * g.Nq * AT_C;
  const T* kb = k + (size_t)bh * g.Nk * AT_C;
  const T* vb = v + (size_t)bh * g.Nk * AT_C;
#pragma unroll
  for (int a = 0; a < AT_QPW; ++a)
#pragma unroll
    for (int d = 0; d < AT_V; ++d)
      qs[wid][a][lane + 32 * d] = (i0 + a < g.Nq) ? ldf(qb + (size_t)(i0 + a) * AT_C + lane + 32 * d) * g.scale : 0.f;
  __syncwarp();
  float mx[AT_QPW], l[AT_QPW], acc[AT_QPW][AT_V];
#pragma unroll
  for (int a = 0; a < AT_QPW; ++a) {
    mx[a] = -FLT_MAX, l[a] = 0.f;
#pragma unroll
    for (int d = 0; d < AT_V; ++d) acc[a][d] = 0.f;
  }
  for (int j0 = 0; j0 < g.Nk; j0 += 32) {
    const int j = j0 + lane;
    const bool valid = j < g.Nk;
    float s[AT_QPW];
#pragma unroll
    for (int a = 0; a < AT_QPW; ++a) s[a] = 0.f;
    if (valid) {
      const T* kr = kb + (size_t)j * AT_C;
#pragma unroll 2
      for (int c = 0; c < AT_C; c += 8) {
        float kf[8];
        ld8(kr + c, kf);
#pragma unroll
        for (int a = 0; a < AT_QPW; ++a) {
          const float4 q0 = *reinterpret_cast<const float4*>(&qs[wid][a][c]);
          const float4 q1 = *reinterpret_cast<const float4*>(&qs[wid][a][c + 4]);
          s[a] += kf[0] * q0.x + kf[1] * q0.y + kf[2] * q0.z + kf[3] * q0.w + kf[4] * q1.x + kf[5] * q1.y + kf[6] * q1.z +
                  kf[7] * q1.w;
        }
      }
      if (j > 0) {
#pragma unroll
        for (int a = 0; a < AT_QPW; ++a) {
          const int i = i0 + a;
          if (i > 0 && i < g.Nq) s[a] += attn_bias(bq + ((size_t)bh * (g.Nq - 1) + (i - 1)) * KB, j, g);
        }
      }
    }
#pragma unroll
    for (int a = 0; a < AT_QPW; ++a) {
      const float cm = wmax(valid ? s[a] : -FLT_MAX);
      const float mnew = fmaxf(mx[a], cm);
      const float p = valid ? expf(s[a] - mnew) : 0.f;
      const float corr = expf(mx[a] - mnew);
      l[a] = l[a] * corr + wsum(p);
#pragma unroll
      for (int d = 0; d < AT_V; ++d) acc[a][d] *= corr;
      mx[a] = mnew;
      ps[wid][a][lane] = p;
    }
    __syncwarp();
    const int nj = min(32, g.Nk - j0);
    for (int jj = 0; jj < nj; ++jj) {
      const T* vr = vb + (size_t)(j0 + jj) * AT_C;
      float vf[AT_V];
#pragma unroll
      for (int d = 0; d < AT_V; ++d) vf[d] = ldf(vr + lane + 32 * d);
#pragma unroll
      for (int a = 0; a < AT_QPW; ++a) {
        const float p = ps[wid][a][jj];
#pragma unroll
        for (int d = 0; d < AT_V; ++d) acc[a][d] += p * vf[d];
      }
    }
    __syncwarp();
  }
  const int b = bh / g.heads, h = bh % g.heads;
#pragma unroll
  for (int a = 0; a < AT_QPW; ++a) {
    const int i = i0 + a;
    if (i >= g.Nq) break;
    const float inv = 1.0f / l[a];
    T* orow = out + ((size_t)b * g.Nq + i) * ((size_t)g.heads * AT_C) + (size_t)h * AT_C;
#pragma unroll
    for (int d = 0; d < AT_V; ++d) {
      float o = acc[a][d] * inv;
      if (g.resid && i > 0) o += ldf(qb + (size_t)i * AT_C + lane + 32 * d);
      stf(orow + lane + 32 * d, o);
    }
    if (lane == 0) lse[(size_t)bh * g.Nq + i] = mx[a] + logf(l[a]);
  }
}

// dQ pass (a warp owns AT_QPW queries and walks the keys ONCE):  p = exp(s - lse), dP = dO . v, dS = p (dP - delta) with
// delta = sum_j p dP (= dO . O_attn) -- not known until the walk ends, so the pass accumulates the two halves of
//   dq  = scale * (sum_j p dP k_j  -  delta * sum_j p k_j)  (+ dO for the residual pooling)
//   dbq[i, component(j)] = sum_j p dP - delta * sum_j p      (per component group)
// separately and combines them at the end, all in fp32 (nothing is derived from the rounded forward output).  The bias
// sums are taken in the second phase of every chunk (where p and p dP are read back from shared memory anyway): lane e
// owns components e and e + 32 and adds the keys whose (kt, kh, kw) select it -- the first build did this with
// shared-memory atomics from the key-per-lane phase, where 32 consecutive keys share one kt and ~5 kh: 32-way same-address
// conflicts made this pass 3.4x the forward (2.44 ms vs 0.72 ms on the 1 569 x 393 blocks).
// delta [BH, Nq] is written for the dKV pass; dq has the layout of q; dbq the layout of bq (every row written once).
template <typename T>
__global__ void __launch_bounds__(AT_WARPS * 32) pooled_attn_bwd_q_kernel(
    const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, const float* __restrict__ bq,
    const T* __restrict__ dout, const float* __restrict__ lse, T* __restrict__ dq, float* __restrict__ dbq,
    float* __restrict__ delta, AttnGeom g) {
  __shared__ __align__(16) float qs[AT_WARPS][AT_QPW][AT_C];
  __shared__ __align__(16) float dos[AT_WARPS][AT_QPW][AT_C];
  __shared__ float pss[AT_WARPS][AT_QPW][32];
  __shared__ float pds[AT_WARPS][AT_QPW][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int bh = blockIdx.y, KB = g.Kt + g.Kh + g.Kw;
  const int b = bh / g.heads, h = bh % g.heads;
  const int i0 = (blockIdx.x * AT_WARPS + wid) * AT_QPW;
  if (i0 >= g.Nq) return;
  const T* qb = q + (size_t)bh * g.Nq * AT_C;
  const T* kb = k + (size_t)bh * g.Nk * AT_C;
  const T* vb = v + (size_t)bh * g.Nk * AT_C;
  const size_t orow_pitch = (size_t)g.heads * AT_C;
  float ls[AT_QPW], dpart[AT_QPW];
#pragma unroll
  for (int a = 0; a < AT_QPW; ++a) {
    const int i = i0 + a;
#pragma unroll
    for (int d = 0; d < AT_V; ++d) {
      const int c = lane + 32 * d;
      float qv = 0.f, dv_ = 0.f;
      if (i < g.Nq) {
        qv = ldf(qb + (size_t)i * AT_C + c);
        dv_ = ldf(dout + ((size_t)b * g.Nq + i) * orow_pitch + (size_t)h * AT_C + c);
      }
      qs[wid][a][c] = qv * g.scale, dos[wid][a][c] = dv_;
    }
    ls[a] = (i < g.Nq) ? lse[(size_t)bh * g.Nq + i] : 0.f;
    dpart[a] = 0.f;
  }
  float bp[2][AT_QPW], bd[2][AT_QPW];   // bias components lane and lane + 32: sum of p, sum of p * dP
#pragma unroll
  for (int a = 0; a < AT_QPW; ++a) bp[0][a] = bp[1][a] = bd[0][a] = bd[1][a] = 0.f;
  __syncwarp();
  float acd[AT_QPW][AT_V], acp[AT_QPW][AT_V];
#pragma unroll
  for (int a = 0; a < AT_QPW; ++a)
#pragma unroll
    for (int d = 0; d < AT_V; ++d) acd[a][d] = 0.f, acp[a][d] = 0.f;
  for (int j0 = 0; j0 < g.Nk; j0 += 32) {
    const int j = j0 + lane;
    const bool valid = j < g.Nk;
    float s[AT_QPW], dp[AT_QPW];
#pragma unroll
    for (int a = 0; a < AT_QPW; ++a) s[a] = 0.f, dp[a] = 0.f;
    if (valid) {
      const T* kr = kb + (size_t)j * AT_C;
      const T* vr = vb + (size_t)j * AT_C;
#pragma unroll 2
      for (int c = 0; c < AT_C; c += 8) {
        float kf[8], vf[8];
        ld8(kr + c, kf);
        ld8(vr + c, vf);
#pragma unroll
        for (int a = 0; a < AT_QPW; ++a) {
          const float4 q0 = *reinterpret_cast<const float4*>(&qs[wid][a][c]);
          const float4 q1 = *reinterpret_cast<const float4*>(&qs[wid][a][c + 4]);
          const float4 d0 = *reinterpret_cast<const float4*>(&dos[wid][a][c]);
          const float4 d1 = *reinterpret_cast<const float4*>(&dos[wid][a][c + 4]);
          s[a] += kf[0] * q0.x + kf[1] * q0.y + kf[2] * q0.z + kf[3] * q0.w + kf[4] * q1.x + kf[5] * q1.y + kf[6] * q1.z +
                  kf[7] * q1.w;
          dp[a] += vf[0] * d0.x + vf[1] * d0.y + vf[2] * d0.z + vf[3] * d0.w + vf[4] * d1.x + vf[5] * d1.y + vf[6] * d1.z +
                   vf[7] * d1.w;
        }
      }
    }
#pragma unroll
    for (int a = 0; a < AT_QPW; ++a) {
      const int i = i0 + a;
      float p = 0.f, pd = 0.f;
      if (valid && i < g.Nq) {
        float sc = s[a];
        if (j > 0 && i > 0) sc += attn_bias(bq + ((size_t)bh * (g.Nq - 1) + (i - 1)) * KB, j, g);
        p = expf(sc - ls[a]);
        pd = p * dp[a];
      }
      dpart[a] += pd;
      pss[wid][a][lane] = p, pds[wid][a][lane] = pd;
    }
    __syncwarp();
    const int nj = min(32, g.Nk - j0);
    // grid position of the first non-cls key of this chunk (warp-uniform), advanced key by key below
    const int t0 = max(j0, 1) - 1;
    int ckw = t0 % g.Kw, ckh = (t0 / g.Kw) % g.Kh, ckt = t0 / (g.Kw * g.Kh);
    for (int jj = 0; jj < nj; ++jj) {
      const T* kr = kb + (size_t)(j0 + jj) * AT_C;
      float kf[AT_V];
#pragma unroll
      for (int d = 0; d < AT_V; ++d) kf[d] = ldf(kr + lane + 32 * d);
      float m0 = 0.f, m1 = 0.f;            // does this key select the lane's bias components?
      if (j0 + jj > 0) {
        const int et = ckt, eh = g.Kt + ckh, ew = g.Kt + g.Kh + ckw;
        m0 = (lane == et || lane == eh || lane == ew) ? 1.f : 0.f;
        m1 = (lane + 32 == et || lane + 32 == eh || lane + 32 == ew) ? 1.f : 0.f;
        if (++ckw == g.Kw) {
          ckw = 0;
          if (++ckh == g.Kh) ckh = 0, ++ckt;
        }
      }
#pragma unroll
      for (int a = 0; a < AT_QPW; ++a) {
        const float p = pss[wid][a][jj], pd = pds[wid][a][jj];
#pragma unroll
        for (int d = 0; d < AT_V; ++d) acp[a][d] += p * kf[d], acd[a][d] += pd * kf[d];
        bp[0][a] += m0 * p, bd[0][a] += m0 * pd;
        bp[1][a] += m1 * p, bd[1][a] += m1 * pd;
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int a = 0; a < AT_QPW; ++a) {
    const int i = i0 + a;
    if (i >= g.Nq) break;
    const float dl = wsum(dpart[a]);
    if (lane == 0) delta[(size_t)bh * g.Nq + i] = dl;
#pragma unroll
    for (int d = 0; d < AT_V; ++d) {
      const int c = lane + 32 * d;
      float r = (acd[a][d] - dl * acp[a][d]) * g.scale;
      if (g.resid && i > 0) r += dos[wid][a][c];
      stf(dq + ((size_t)bh * g.Nq + i) * AT_C + c, r);
    }
    if (i > 0) {
      float* brow = dbq + ((size_t)bh * (g.Nq - 1) + (i - 1)) * KB;
      if (lane < KB) brow[lane] = bd[0][a] - dl * bp[0][a];
      if (lane + 32 < KB) brow[lane + 32] = bd[1][a] - dl * bp[1][a];
    }
  }
}

// dK / dV pass (a warp owns AT_QPW keys and walks a slice of the queries, a lane per query): dv_j += sum_i p dO_i,
// dk_j += scale * sum_i dS q_i, accumulated into zero-initialised fp32 buffers (one atomic per warp, slice and element).
template <typename T>
__global__ void __launch_bounds__(AT_WARPS * 32) pooled_attn_bwd_kv_kernel(
    const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, const float* __restrict__ bq,
    const T* __restrict__ dout, const float* __restrict__ lse, const float* __restrict__ delta, float* __restrict__ dk,
    float* __restrict__ dv, AttnGeom g) {
  __shared__ __align__(16) float ks[AT_WARPS][AT_QPW][AT_C];
  __shared__ __align__(16) float vs[AT_WARPS][AT_QPW][AT_C];
  __shared__ float ps[AT_WARPS][AT_QPW][32];
  __shared__ float dss[AT_WARPS][AT_QPW][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int bh = blockIdx.y, KB = g.Kt + g.Kh + g.Kw;
  const int b = bh / g.heads, h = bh % g.heads;
  const int j0 = (blockIdx.x * AT_WARPS + wid) * AT_QPW;
  if (j0 >= g.Nk) return;
  const T* qb = q + (size_t)bh * g.Nq * AT_C;
  const T* kb = k + (size_t)bh * g.Nk * AT_C;
  const T* vb = v + (size_t)bh * g.Nk * AT_C;
  const size_t orow_pitch = (size_t)g.heads * AT_C;
  int kt[AT_QPW], kh[AT_QPW], kw[AT_QPW];
#pragma unroll
  for (int a = 0; a < AT_QPW; ++a) {
    const int j = j0 + a;
#pragma unroll
    for (int d = 0; d < AT_V; ++d) {
      const int c = lane + 32 * d;
      ks[wid][a][c] = (j < g.Nk) ? ldf(kb + (size_t)j * AT_C + c) * g.scale : 0.f;
      vs[wid][a][c] = (j < g.Nk) ? ldf(vb + (size_t)j * AT_C + c) : 0.f;
    }
    const int jj = (j > 0 && j < g.Nk) ? j - 1 : 0;
    kw[a] = jj % g.Kw, kh[a] = (jj / g.Kw) % g.Kh, kt[a] = jj / (g.Kw * g.Kh);
  }
  __syncwarp();
  float adk[AT_QPW][AT_V], adv[AT_QPW][AT_V];
#pragma unroll
  for (int a = 0; a < AT_QPW; ++a)
#pragma unroll
    for (int d = 0; d < AT_V; ++d) adk[a][d] = 0.f, adv[a][d] = 0.f;
  const int per = (((g.Nq + g.qsplit - 1) / g.qsplit) + 31) / 32 * 32;
  const int ibeg = blockIdx.z * per, iend = min(g.Nq, ibeg + per);
  for (int i0 = ibeg; i0 < iend; i0 += 32) {
    const int i = i0 + lane;
    const bool valid = i < iend;
    float s[AT_QPW], dp[AT_QPW];
#pragma unroll
    for (int a = 0; a < AT_QPW; ++a) s[a] = 0.f, dp[a] = 0.f;
    if (valid) {
      const T* qr = qb + (size_t)i * AT_C;
      const T* dr = dout + ((size_t)b * g.Nq + i) * orow_pitch + (size_t)h * AT_C;
#pragma unroll 2
      for (int c = 0; c < AT_C; c += 8) {
        float qf[8], df[8];
        ld8(qr + c, qf);
        ld8(dr + c, df);
#pragma unroll
        for (int a = 0; a < AT_QPW; ++a) {
          const float4 k0 = *reinterpret_cast<const float4*>(&ks[wid][a][c]);
          const float4 k1 = *reinterpret_cast<const float4*>(&ks[wid][a][c + 4]);
          const float4 v0 = *reinterpret_cast<const float4*>(&vs[wid][a][c]);
          const float4 v1 = *reinterpret_cast<const float4*>(&vs[wid][a][c + 4]);
          s[a] += qf[0] * k0.x + qf[1] * k0.y + qf[2] * k0.z + qf[3] * k0.w + qf[4] * k1.x + qf[5] * k1.y + qf[6] * k1.z +
                  qf[7] * k1.w;
          dp[a] += df[0] * v0.x + df[1] * v0.y + df[2] * v0.z + df[3] * v0.w + df[4] * v1.x + df[5] * v1.y + df[6] * v1.z +
                   df[7] * v1.w;
        }
      }
    }
    const float ls = valid ? lse[(size_t)bh * g.Nq + i] : 0.f;
    const float dl = valid ? delta[(size_t)bh * g.Nq + i] : 0.f;
    const float* bqrow = bq + ((size_t)bh * (g.Nq - 1) + (valid && i > 0 ? i - 1 : 0)) * KB;
#pragma unroll
    for (int a = 0; a < AT_QPW; ++a) {
      const int j = j0 + a;
      float p = 0.f, ds = 0.f;
      if (valid && j < g.Nk) {
        float sc = s[a];
        if (i > 0 && j > 0) sc += __ldg(bqrow + kt[a]) + __ldg(bqrow + g.Kt + kh[a]) + __ldg(bqrow + g.Kt + g.Kh + kw[a]);
        p = expf(sc - ls);
        ds = p * (dp[a] - dl);
      }
      ps[wid][a][lane] = p, dss[wid][a][lane] = ds;
    }
    __syncwarp();
    const int ni = min(32, iend - i0);
    for (int ii = 0; ii < ni; ++ii) {
      const T* qr = qb + (size_t)(i0 + ii) * AT_C;
      const T* dr = dout + ((size_t)b * g.Nq + i0 + ii) * orow_pitch + (size_t)h * AT_C;
      float qf[AT_V], df[AT_V];
#pragma unroll
      for (int d = 0; d < AT_V; ++d) qf[d] = ldf(qr + lane + 32 * d), df[d] = ldf(dr + lane + 32 * d);
#pragma unroll
      for (int a = 0; a < AT_QPW; ++a) {
        const float p = ps[wid][a][ii], ds = dss[wid][a][ii];
#pragma unroll
        for (int d = 0; d < AT_V; ++d) adv[a][d] += p * df[d], adk[a][d] += ds * qf[d];
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int a = 0; a < AT_QPW; ++a) {
    const int j = j0 + a;
    if (j >= g.Nk) break;
#pragma unroll
    for (int d = 0; d < AT_V; ++d) {
      const size_t off = ((size_t)bh * g.Nk + j) * AT_C + lane + 32 * d;
      atomicAdd(dk + off, adk[a][d] * g.scale);
      atomicAdd(dv + off, adv[a][d]);
    }
  }
}

inline int grid_for(long long work_items, int per_block) {
  long long blocks = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

int check_pool(const pvrl_pool3d_t* p, const char* who) {
  PVRL_CHECK_ARG(p != nullptr, "%s: null geometry", who);
  PVRL_CHECK_ARG(p->B > 0 && p->heads > 0 && p->C > 0 && p->T > 0 && p->H > 0 && p->W > 0, "%s: empty geometry", who);
  for (int d = 0; d < 3; ++d) {
    PVRL_CHECK_ARG(p->kernel[d] > 0 && p->stride[d] > 0 && p->pad[d] >= 0, "%s: bad kernel / stride / padding", who);
  }
  const int in[3] = {p->T, p->H, p->W};
  for (int d = 0; d < 3; ++d) {
    const int o = (in[d] + 2 * p->pad[d] - p->kernel[d]) / p->stride[d] + 1;
    PVRL_CHECK_ARG(o > 0 && o == p->out[d], "%s: output grid %d along axis %d does not match floor((in + 2p - k) / s) + 1 = %d",
                   who, p->out[d], d, o);
  }
  return 0;
}

PoolGeom to_geom(const pvrl_pool3d_t* p) {
  PoolGeom g;
  g.B = p->B, g.heads = p->heads, g.C = p->C, g.T = p->T, g.H = p->H, g.W = p->W;
  g.KT = p->kernel[0], g.KH = p->kernel[1], g.KW = p->kernel[2];
  g.ST = p->stride[0], g.SH = p->stride[1], g.SW = p->stride[2];
  g.PT = p->pad[0], g.PH = p->pad[1], g.PW = p->pad[2];
  g.OT = p->out[0], g.OH = p->out[1], g.OW = p->out[2];
  g.ld = p->ld;
  return g;
}

}  // namespace
}  // namespace pvrl

namespace pvrl {
// mvit_attn_mma.cu: the forward on mma.sync (bf16 only)
int pooled_attn_fwd_mma_launch(const void* q, const void* k, const void* v, const float* bq, void* out, float* lse, int B,
                               int heads, int Nq, int Nk, int Kt, int Kh, int Kw, float scale, int resid,
                               cudaStream_t stream);
int pooled_attn_bwd_mma_launch(const void* q, const void* k, const void* v, const float* bq, const void* dout, const float* lse,
                               void* dq, float* dk, float* dv, float* dbq, float* delta, int B, int heads, int Nq, int Nk,
                               int Kt, int Kh, int Kw, float scale, int resid, cudaStream_t stream);
}  // namespace pvrl

using namespace pvrl;
#define STREAM static_cast<cudaStream_t>(stream)

// PVRL_MVIT_ATTN_MMA = 1 (default) / 0 selects the mma.sync forward (mvit_attn_mma.cu) or the CUDA-core forward for bf16
// problems (fp32 problems -- the parity mode -- always take the CUDA-core kernel); read on every call so that tests can
// compare the two in one process.
static bool mvit_attn_mma_enabled() {
  const char* e = getenv("PVRL_MVIT_ATTN_MMA");
  return e != nullptr ? atoi(e) != 0 : PVRL_MVIT_ATTN_MMA_DEFAULT != 0;
}

extern "C" int pvrl_ln_any_fwd(const void* x, int32_t x_dtype, const float* w, const float* b, void* y, int32_t y_dtype,
                               float* stats, int32_t M, int32_t D, float eps, void* stream) {
  PVRL_CHECK_ARG(x && w && b && y && stats && M > 0, "pvrl_ln_any_fwd: bad arguments");
  PVRL_CHECK_ARG(D > 0 && D <= 32 * LNA_MAXV, "pvrl_ln_any_fwd: D=%d must be in 1 .. %d", D, 32 * LNA_MAXV);
  const int grid = grid_for(M, 8);
  if (x_dtype == PVRL_F32 && y_dtype == PVRL_F32)
    ln_any_fwd_kernel<float, float><<<grid, 256, 0, STREAM>>>((const float*)x, w, b, (float*)y, stats, M, D, eps);
  else if (x_dtype == PVRL_F32)
    ln_any_fwd_kernel<float, bf16><<<grid, 256, 0, STREAM>>>((const float*)x, w, b, (bf16*)y, stats, M, D, eps);
  else if (y_dtype == PVRL_F32)
    ln_any_fwd_kernel<bf16, float><<<grid, 256, 0, STREAM>>>((const bf16*)x, w, b, (float*)y, stats, M, D, eps);
  else
    ln_any_fwd_kernel<bf16, bf16><<<grid, 256, 0, STREAM>>>((const bf16*)x, w, b, (bf16*)y, stats, M, D, eps);
  return launched("ln_any_fwd_kernel");
}

extern "C" int pvrl_ln_any_bwd(const void* dy, int32_t dy_dtype, const void* x, int32_t x_dtype, const float* w,
                               const float* stats, void* dx, float* dw, float* db, int32_t M, int32_t D, void* stream) {
  PVRL_CHECK_ARG(dy && x && w && stats && dx && M > 0, "pvrl_ln_any_bwd: bad arguments");
  PVRL_CHECK_ARG(D > 0 && D <= 32 * LNA_MAXV, "pvrl_ln_any_bwd: D=%d must be in 1 .. %d", D, 32 * LNA_MAXV);
  int grid = grid_for(M, 8);
  if (grid > num_sms() * 4) grid = num_sms() * 4;   // fewer, longer warps: the column sums end in one atomic per warp and column
  // dx has the dtype of x (the gradient of the tensor that was normalised)
  if (dy_dtype == PVRL_F32 && x_dtype == PVRL_F32)
    ln_any_bwd_kernel<float, float, float>
        <<<grid, 256, 0, STREAM>>>((const float*)dy, (const float*)x, w, stats, (float*)dx, dw, db, M, D);
  else if (dy_dtype == PVRL_BF16 && x_dtype == PVRL_F32)
    ln_any_bwd_kernel<bf16, float, float>
        <<<grid, 256, 0, STREAM>>>((const bf16*)dy, (const float*)x, w, stats, (float*)dx, dw, db, M, D);
  else if (dy_dtype == PVRL_BF16 && x_dtype == PVRL_BF16)
    ln_any_bwd_kernel<bf16, bf16, bf16>
        <<<grid, 256, 0, STREAM>>>((const bf16*)dy, (const bf16*)x, w, stats, (bf16*)dx, dw, db, M, D);
  else
    ln_any_bwd_kernel<float, bf16, bf16>
        <<<grid, 256, 0, STREAM>>>((const float*)dy, (const bf16*)x, w, stats, (bf16*)dx, dw, db, M, D);
  return launched("ln_any_bwd_kernel");
}

extern "C" int pvrl_pool3d_fwd(const void* in, const float* w, void* out, int32_t dtype, const pvrl_pool3d_t* p,
                               void* stream) {
  int rc = check_pool(p, "pvrl_pool3d_fwd");
  if (rc) return rc;
  PVRL_CHECK_ARG(in && out, "pvrl_pool3d_fwd: null buffer");
  PVRL_CHECK_ARG(p->C <= 32 * POOL_MAXV && p->ld >= (long long)p->heads * p->C, "pvrl_pool3d_fwd: C=%d (<= %d), ld=%lld",
                 p->C, 32 * POOL_MAXV, (long long)p->ld);
  if (w == nullptr)
    PVRL_CHECK_ARG(p->out[0] == p->T && p->out[1] == p->H && p->out[2] == p->W, "pvrl_pool3d_fwd: w = NULL is a re-layout only");
  const PoolGeom g = to_geom(p);
  const long long rows = (long long)g.B * g.heads * (1 + g.OT * g.OH * g.OW);
  const int grid = grid_for(rows, 8);
  if (dtype == PVRL_F32)
    pool3d_fwd_kernel<float><<<grid, 256, 0, STREAM>>>((const float*)in, w, (float*)out, g);
  else
    pool3d_fwd_kernel<bf16><<<grid, 256, 0, STREAM>>>((const bf16*)in, w, (bf16*)out, g);
  return launched("pool3d_fwd_kernel");
}

extern "C" int pvrl_pool3d_bwd(const void* dout, const void* in, const float* w, void* din, float* dw, int32_t dtype,
                               const pvrl_pool3d_t* p, void* stream) {
  int rc = check_pool(p, "pvrl_pool3d_bwd");
  if (rc) return rc;
  PVRL_CHECK_ARG(dout && din, "pvrl_pool3d_bwd: null buffer");
  PVRL_CHECK_ARG(p->C <= 32 * POOL_MAXV && p->ld >= (long long)p->heads * p->C, "pvrl_pool3d_bwd: C=%d (<= %d), ld=%lld",
                 p->C, 32 * POOL_MAXV, (long long)p->ld);
  if (w == nullptr)
    PVRL_CHECK_ARG(p->out[0] == p->T && p->out[1] == p->H && p->out[2] == p->W, "pvrl_pool3d_bwd: w = NULL is a re-layout only");
  const PoolGeom g = to_geom(p);
  const long long rows = (long long)g.B * g.heads * (1 + g.T * g.H * g.W);
  const int grid = grid_for(rows, 8);
  if (dtype == PVRL_F32)
    pool3d_bwd_in_kernel<float><<<grid, 256, 0, STREAM>>>((const float*)dout, w, (float*)din, g);
  else
    pool3d_bwd_in_kernel<bf16><<<grid, 256, 0, STREAM>>>((const bf16*)dout, w, (bf16*)din, g);
  rc = launched("pool3d_bwd_in_kernel");
  if (rc || w == nullptr || dw == nullptr) return rc;
  PVRL_CHECK_ARG(in != nullptr, "pvrl_pool3d_bwd: the weight gradient needs the forward input");
  PVRL_CHECK_ARG(p->kernel[0] == 3 && p->kernel[1] == 3 && p->kernel[2] == 3,
                 "pvrl_pool3d_bwd: the weight-gradient kernel is written for 3 x 3 x 3 pooling kernels (MVIT.POOL_KVQ_KERNEL)");
  const long long orows = (long long)g.B * g.heads * g.OT * g.OH * g.OW;
  int gw = grid_for(orows, 8 * 16);
  if (gw > num_sms() * 2) gw = num_sms() * 2;
  // one channel per lane (27 accumulators, 64 registers, 32 resident warps per SM): 1.49 ms vs 2.67 ms for both weight-
  // gradient launches of block 0 (9 clips) on B200 -- the default; PVRL_POOL_DW=1 selects the 254-register kernel
  const char* variant = getenv("PVRL_POOL_DW");
  if (variant == nullptr || atoi(variant) == 2) {
    const dim3 grid2(grid_for(orows, 8 * 8), (g.C + 31) / 32);
    if (dtype == PVRL_F32)
      pool3d_bwd_w_cg_kernel<float><<<grid2, 256, 0, STREAM>>>((const float*)dout, (const float*)in, dw, g);
    else
      pool3d_bwd_w_cg_kernel<bf16><<<grid2, 256, 0, STREAM>>>((const bf16*)dout, (const bf16*)in, dw, g);
    return launched("pool3d_bwd_w_cg_kernel");
  }
  const size_t smem = (size_t)g.C * 27 * sizeof(float);
  if (dtype == PVRL_F32)
    pool3d_bwd_w_kernel<float><<<gw, 256, smem, STREAM>>>((const float*)dout, (const float*)in, dw, g);
  else
    pool3d_bwd_w_kernel<bf16><<<gw, 256, smem, STREAM>>>((const bf16*)dout, (const bf16*)in, dw, g);
  return launched("pool3d_bwd_w_kernel");
}

extern "C" int pvrl_maxpool3d_fwd(const void* x, void* y, int32_t* arg, int32_t dtype, const pvrl_pool3d_t* p,
                                  void* stream) {
  int rc = check_pool(p, "pvrl_maxpool3d_fwd");
  if (rc) return rc;
  PVRL_CHECK_ARG(x && y && arg && p->heads == 1, "pvrl_maxpool3d_fwd: bad arguments (heads must be 1: C is the token width)");
  const PoolGeom g = to_geom(p);
  const long long total = (long long)g.B * (1 + g.OT * g.OH * g.OW) * g.C;
  const int grid = grid_for(total, 256);
  if (dtype == PVRL_F32)
    maxpool3d_fwd_kernel<float><<<grid, 256, 0, STREAM>>>((const float*)x, (float*)y, arg, g);
  else
    maxpool3d_fwd_kernel<bf16><<<grid, 256, 0, STREAM>>>((const bf16*)x, (bf16*)y, arg, g);
  return launched("maxpool3d_fwd_kernel");
}

extern "C" int pvrl_maxpool3d_bwd(const void* dy, const int32_t* arg, float* dx, int32_t dtype, const pvrl_pool3d_t* p,
                                  void* stream) {
  int rc = check_pool(p, "pvrl_maxpool3d_bwd");
  if (rc) return rc;
  PVRL_CHECK_ARG(dy && dx && arg && p->heads == 1, "pvrl_maxpool3d_bwd: bad arguments");
  const PoolGeom g = to_geom(p);
  const long long total = (long long)g.B * (1 + g.OT * g.OH * g.OW) * g.C;
  const int grid = grid_for(total, 256);
  if (dtype == PVRL_F32)
    maxpool3d_bwd_kernel<float><<<grid, 256, 0, STREAM>>>((const float*)dy, arg, dx, g);
  else
    maxpool3d_bwd_kernel<bf16><<<grid, 256, 0, STREAM>>>((const bf16*)dy, arg, dx, g);
  return launched("maxpool3d_bwd_kernel");
}

extern "C" int pvrl_im2col3d(const float* frames, void* out, int32_t out_dtype, int32_t Cin, int32_t Kpad,
                             const pvrl_pool3d_t* p, void* stream) {
  int rc = check_pool(p, "pvrl_im2col3d");
  if (rc) return rc;
  PVRL_CHECK_ARG(frames && out && Cin > 0, "pvrl_im2col3d: bad arguments");
  PVRL_CHECK_ARG(Kpad >= Cin * p->kernel[0] * p->kernel[1] * p->kernel[2], "pvrl_im2col3d: Kpad=%d is shorter than a window", Kpad);
  const PoolGeom g = to_geom(p);
  const long long total = (long long)g.B * g.OT * g.OH * g.OW * Kpad;
  const int grid = grid_for(total, 256 * 4);
  if (out_dtype == PVRL_F32)
    im2col3d_kernel<float><<<grid, 256, 0, STREAM>>>(frames, (float*)out, g, Cin, Kpad);
  else
    im2col3d_kernel<bf16><<<grid, 256, 0, STREAM>>>(frames, (bf16*)out, g, Cin, Kpad);
  return launched("im2col3d_kernel");
}

static int check_attn(const pvrl_pooled_attn_t* a, const char* who) {
  PVRL_CHECK_ARG(a != nullptr, "%s: null descriptor", who);
  PVRL_CHECK_ARG(a->B > 0 && a->heads > 0 && a->Nq > 1 && a->Nk > 1, "%s: empty problem", who);
  PVRL_CHECK_ARG(a->C == AT_C, "%s: head width %d (the kernels are written for %d-wide heads, every MViTv2 stage)", who, a->C, AT_C);
  PVRL_CHECK_ARG(a->Kt > 0 && a->Kh > 0 && a->Kw > 0 && a->Kt * a->Kh * a->Kw == a->Nk - 1,
                 "%s: key grid %d x %d x %d does not hold Nk - 1 = %d keys", who, a->Kt, a->Kh, a->Kw, a->Nk - 1);
  PVRL_CHECK_ARG(a->Kt + a->Kh + a->Kw <= AT_KBMAX, "%s: Kt + Kh + Kw = %d exceeds %d", who, a->Kt + a->Kh + a->Kw, AT_KBMAX);
  PVRL_CHECK_ARG((long long)a->B * a->heads <= 65535, "%s: clips x heads = %lld exceeds the grid limit", who,
                 (long long)a->B * a->heads);
  return 0;
}

static AttnGeom to_attn(const pvrl_pooled_attn_t* a) {
  AttnGeom g;
  g.BH = a->B * a->heads, g.heads = a->heads, g.Nq = a->Nq, g.Nk = a->Nk;
  g.Kt = a->Kt, g.Kh = a->Kh, g.Kw = a->Kw, g.scale = a->scale, g.resid = a->residual_pooling, g.qsplit = 1;
  return g;
}

extern "C" int pvrl_pooled_attn_fwd(const void* q, const void* k, const void* v, const float* bq, void* out, float* lse,
                                    int32_t dtype, const pvrl_pooled_attn_t* a, void* stream) {
  int rc = check_attn(a, "pvrl_pooled_attn_fwd");
  if (rc) return rc;
  PVRL_CHECK_ARG(q && k && v && bq && out && lse, "pvrl_pooled_attn_fwd: null buffer");
  if (dtype == PVRL_BF16 && mvit_attn_mma_enabled())
    return pooled_attn_fwd_mma_launch(q, k, v, bq, out, lse, a->B, a->heads, a->Nq, a->Nk, a->Kt, a->Kh, a->Kw, a->scale,
                                      a->residual_pooling, STREAM);
  const AttnGeom g = to_attn(a);
  const dim3 grid((g.Nq + AT_WARPS * AT_QPW - 1) / (AT_WARPS * AT_QPW), g.BH);
  if (dtype == PVRL_F32)
    pooled_attn_fwd_kernel<float><<<grid, AT_WARPS * 32, 0, STREAM>>>((const float*)q, (const float*)k, (const float*)v, bq,
                                                                      (float*)out, lse, g);
  else
    pooled_attn_fwd_kernel<bf16><<<grid, AT_WARPS * 32, 0, STREAM>>>((const bf16*)q, (const bf16*)k, (const bf16*)v, bq,
                                                                     (bf16*)out, lse, g);
  return launched("pooled_attn_fwd_kernel");
}

extern "C" int pvrl_pooled_attn_bwd(const void* q, const void* k, const void* v, const float* bq, const void* dout,
                                    const float* lse, void* dq, float* dk, float* dv, float* dbq, float* delta,
                                    int32_t dtype, const pvrl_pooled_attn_t* a, void* stream) {
  int rc = check_attn(a, "pvrl_pooled_attn_bwd");
  if (rc) return rc;
  PVRL_CHECK_ARG(q && k && v && bq && dout && lse && dq && dk && dv && dbq && delta, "pvrl_pooled_attn_bwd: null buffer");
  // The mma.sync dQ and dK/dV passes (mvit_attn_mma.cu) for bf16 problems: default since their first hardware run was green
  // (round-1 driver run, XPASS; tests/test_zz_mvit_mma_bwd_gpu.py now holds them strictly).  PVRL_MVIT_ATTN_MMA_BWD=0
  // selects the CUDA-core passes.
  if (dtype == PVRL_BF16) {
    const char* e = getenv("PVRL_MVIT_ATTN_MMA_BWD");
    if (e == nullptr || atoi(e) != 0)
      return pooled_attn_bwd_mma_launch(q, k, v, bq, dout, lse, dq, dk, dv, dbq, delta, a->B, a->heads, a->Nq, a->Nk, a->Kt,
                                        a->Kh, a->Kw, a->scale, a->residual_pooling, STREAM);
  }
  AttnGeom g = to_attn(a);
  const dim3 gq((g.Nq + AT_WARPS * AT_QPW - 1) / (AT_WARPS * AT_QPW), g.BH);
  if (dtype == PVRL_F32)
    pooled_attn_bwd_q_kernel<float><<<gq, AT_WARPS * 32, 0, STREAM>>>((const float*)q, (const float*)k, (const float*)v, bq,
                                                                      (const float*)dout, lse, (float*)dq, dbq, delta, g);
  else
    pooled_attn_bwd_q_kernel<bf16><<<gq, AT_WARPS * 32, 0, STREAM>>>((const bf16*)q, (const bf16*)k, (const bf16*)v, bq,
                                                                     (const bf16*)dout, lse, (bf16*)dq, dbq, delta, g);
  rc = launched("pooled_attn_bwd_q_kernel");
  if (rc) return rc;
  // dK / dV: cut the query range so that the grid fills the chip (few keys, many queries), >= 256 queries per slice
  const int kblocks = (g.Nk + AT_WARPS * AT_QPW - 1) / (AT_WARPS * AT_QPW);
  int split = (num_sms() * 8 + kblocks * g.BH - 1) / (kblocks * g.BH);
  const int max_split = (g.Nq + 255) / 256;
  if (split > max_split) split = max_split;
  if (split < 1) split = 1;
  if (split > 65535) split = 65535;
  g.qsplit = split;
  const dim3 gk(kblocks, g.BH, split);
  if (dtype == PVRL_F32)
    pooled_attn_bwd_kv_kernel<float><<<gk, AT_WARPS * 32, 0, STREAM>>>((const float*)q, (const float*)k, (const float*)v, bq,
                                                                       (const float*)dout, lse, delta, dk, dv, g);
  else
    pooled_attn_bwd_kv_kernel<bf16><<<gk, AT_WARPS * 32, 0, STREAM>>>((const bf16*)q, (const bf16*)k, (const bf16*)v, bq,
                                                                      (const bf16*)dout, lse, delta, dk, dv, g);
  return launched("pooled_attn_bwd_kv_kernel");
}
