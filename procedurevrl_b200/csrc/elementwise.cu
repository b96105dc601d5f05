// HBM-bound kernels around the GEMMs: patch im2col, LayerNorm forward/backward (warp-shuffle, one warp
// per 768-wide row), gather/cast of gradients through the (h w t) <-> (t)(h w) token permutes, bias-gradient
// column sums, weight casts/transposes, the bf16x3 operand split and the embedding gradients.
// All loads/stores are 128-bit and coalesced along the feature dimension.
#include "pvrl_host.h"
#include "pvrl_ptx.cuh"

namespace pvrl {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__device__ __forceinline__ void store4(T* dst, float a, float b, float c, float d);
template <>
__device__ __forceinline__ void store4<float>(float* dst, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(dst) = make_float4(a, b, c, d);
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* dst, float a, float b, float c, float d) {
  uint2 u;
  u.x = pack_bf16x2(a, b);
  u.y = pack_bf16x2(c, d);
  *reinterpret_cast<uint2*>(dst) = u;
}
template <typename T>
__device__ __forceinline__ float4 load4(const T* src);
template <>
__device__ __forceinline__ float4 load4<float>(const float* src) {
  return __ldg(reinterpret_cast<const float4*>(src));
}
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* src) {
  uint2 u = __ldg(reinterpret_cast<const uint2*>(src));
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}

// ---------------------------------------------------------------------------------------- patchify
// out row (b,t,ph,pw), column k = c*P*P + kh*P + kw  == Conv2d(k=s=P) as a GEMM (vit.py:172-179).
template <typename OutT>
__global__ void patchify_kernel(const float* __restrict__ frames, OutT* __restrict__ out, int Bc, int T, int H, int W,
                                int P, long long total_vec) {
  const int nW = W / P, nH = H / P, KP = 3 * P * P, vec_per_row = KP / 4, pv = P / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total_vec;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / vec_per_row;
    const int kv = static_cast<int>(i - row * vec_per_row);
    const int kwv = kv % pv, kh = (kv / pv) % P, c = kv / (pv * P);
    const int pw = static_cast<int>(row % nW);
    const int ph = static_cast<int>((row / nW) % nH);
    const int t = static_cast<int>((row / (nW * nH)) % T);
    const int b = static_cast<int>(row / ((long long)nW * nH * T));
    const float* src = frames + (((long long)(b * 3 + c) * T + t) * H + (ph * P + kh)) * W + pw * P + kwv * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(src));
    store4<OutT>(out + row * KP + kv * 4, v.x, v.y, v.z, v.w);
  }
}

// uint8 frames straight from the decoder (datasets/utils.py:309-326 tensor_normalize happens here instead of on the host:
// x = (u8 / 255 - mean[c]) / std[c]); 4 pixels per thread, 4x less H2D / HBM read traffic than fp32 frames.
template <typename OutT>
__global__ void patchify_u8_kernel(const uint8_t* __restrict__ frames, OutT* __restrict__ out, int Bc, int T, int H, int W,
                                   int P, long long total_vec, float3 scale, float3 shift) {
  const int nW = W / P, nH = H / P, KP = 3 * P * P, vec_per_row = KP / 4, pv = P / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total_vec;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / vec_per_row;
    const int kv = static_cast<int>(i - row * vec_per_row);
    const int kwv = kv % pv, kh = (kv / pv) % P, c = kv / (pv * P);
    const int pw = static_cast<int>(row % nW);
    const int ph = static_cast<int>((row / nW) % nH);
    const int t = static_cast<int>((row / (nW * nH)) % T);
    const int b = static_cast<int>(row / ((long long)nW * nH * T));
    const uint8_t* src = frames + (((long long)(b * 3 + c) * T + t) * H + (ph * P + kh)) * W + pw * P + kwv * 4;
    const uchar4 v = *reinterpret_cast<const uchar4*>(src);
    const float sc = c == 0 ? scale.x : (c == 1 ? scale.y : scale.z), sh = c == 0 ? shift.x : (c == 1 ? shift.y : shift.z);
    store4<OutT>(out + row * KP + kv * 4, fmaf(v.x, sc, sh), fmaf(v.y, sc, sh), fmaf(v.z, sc, sh), fmaf(v.w, sc, sh));
  }
}

__global__ void cls_init_kernel(float* __restrict__ x, const float* __restrict__ cls, const float* __restrict__ pos,
                                int S, int D) {
  const int b = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) x[(long long)b * S * D + d] = cls[d] + pos[d];
}

// ---------------------------------------------------------------------------------------- LayerNorm
constexpr int LN_MAX_VEC = 8;  // D <= 1024
constexpr int LN_WARPS = 8;

__device__ __forceinline__ const float* ln_src_row(const float* x, const float* x_cls, int map, int m, const Geom& g,
                                                   int D) {
  const long long r = map_row(map, m, g);
  if (r >= 0) return x + r * D;
  const long long bt = -r - 1;
  return x_cls + (bt / g.T) * (long long)g.S * D;
}

template <typename OutT>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ x_cls, const float* __restrict__ w,
                     const float* __restrict__ b, OutT* __restrict__ y, float* __restrict__ stats, int M, int D,
                     float eps, int map, Geom g) {
  if (PDL_EARLY_TRIGGER) pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  // Rows are visited from the LAST to the first: the GEMM epilogue that produced x wrote its highest rows last (they are
  // still in the 126 MB L2), and the GEMM that consumes y starts at row 0 -- which this kernel therefore writes last.
  const int m = M - 1 - (blockIdx.x * LN_WARPS + (threadIdx.x >> 5));
  if (m < 0) return;
  const float* xp = ln_src_row(x, x_cls, map, m, g, D);
  const int nvec = D >> 7;
  float4 v[LN_MAX_VEC];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_VEC; ++i)
    if (i < nvec) {
      v[i] = __ldg(reinterpret_cast<const float4*>(xp) + i * 32 + lane);
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  const float mean = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_VEC; ++i)
    if (i < nvec) {
      const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += a * a + bb * bb + c * c + d * d;
    }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / D + eps);
  OutT* yp = y + (long long)m * D;
#pragma unroll
  for (int i = 0; i < LN_MAX_VEC; ++i)
    if (i < nvec) {
      const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + i * 32 + lane);
      const float4 bv = __ldg(reinterpret_cast<const float4*>(b) + i * 32 + lane);
      store4<OutT>(yp + (i * 32 + lane) * 4, (v[i].x - mean) * rstd * ww.x + bv.x, (v[i].y - mean) * rstd * ww.y + bv.y,
                   (v[i].z - mean) * rstd * ww.z + bv.z, (v[i].w - mean) * rstd * ww.w + bv.w);
    }
  if (lane == 0 && stats != nullptr) reinterpret_cast<float2*>(stats)[m] = make_float2(mean, rstd);
}

// dx[src(m)] += rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * w;  dw += dy * xhat;  db += dy.
// One warp per row, NVEC float4 per lane (D = 128 * NVEC); per-lane column sums of dy*xhat / dy live in registers
// across the rows a warp visits and are folded block-wide through shared memory, then one atomicAdd per column.
// What the EMIT variant writes next to dx: the operand the NEXT sub-layer's backward GEMMs read -- exactly what
// pvrl_gather_cast(dx, out, map, rowscale, rs_div, colsum) would produce afterwards -- while the updated row is still in
// registers, so dx is not read a second time (36 gather launches per step disappear).
struct LnEmit {
  void* out;               // act dtype = dy's dtype, [rows of `map`][D]
  const float* rowscale;   // DropPath factors of the next sub-layer (NULL = 1), indexed by emitted row / rs_div
  float* colsum;           // += column sums of the emitted rows (bias gradient of the next Linear), NULL = skip
  int map, rs_div;
  int extra_cls;           // > 0: that many clips' cls rows (untouched by a MAP_SKIPCLS pass) are emitted from dx as they are
};

// emit dx row `xr` (values v, this lane's columns) to every row of em.map that addresses it
// Column-sum accumulation of the LayerNorm backward.  SMEM = false: per-lane registers (folded through shared memory at the
// end of the kernel).  SMEM = true: a warp-private slice of shared memory, read-modify-written as whole float4s (lane-
// contiguous: conflict-free, no atomics) -- 72 fewer live registers at D = 768, which lets two blocks (16 warps) share an
// SM.  (Shared-memory fp32 atomics into one block-wide copy were measured first: 83 us against 71 us with registers.)
template <bool SMEM, int NVEC>
__device__ __forceinline__ void ln_acc(float4 (&reg)[NVEC], float* sm, int i, int lane, float a, float b, float c, float d) {
  if constexpr (SMEM) {
    float4* p = reinterpret_cast<float4*>(sm) + i * 32 + lane;
    float4 v = *p;
    v.x += a, v.y += b, v.z += c, v.w += d;
    *p = v;
  } else {
    reg[i].x += a, reg[i].y += b, reg[i].z += c, reg[i].w += d;
  }
}

template <typename OutT, int NVEC, bool SMEM>
__device__ __forceinline__ void ln_emit_row(const LnEmit& em, const Geom& g, long long xr, const float4 (&v)[NVEC],
                                            float4 (&ec)[NVEC], float* sm_ec, int lane) {
  constexpr int D = NVEC * 128;
  const int xi = static_cast<int>(xr);      // residual-stream rows fit 31 bits (checked by the host: M is an int)
  const int b = xi / g.S, pos = xi - b * g.S;
  long long m2 = xr;
  int reps = 1;
  float ef = 1.0f;
  if (em.map != PVRL_MAP_IDENT) {
    if (pos > 0) {
      const int n = (pos - 1) / g.T, t = (pos - 1) - n * g.T;
      m2 = em.map == PVRL_MAP_SKIPCLS ? xr - b - 1
           : em.map == PVRL_MAP_SPATIAL ? ((long long)b * g.T + t) * (g.HW + 1) + 1 + n
                                        : ((long long)b * g.T + t) * g.HW + n;
    } else if (em.map == PVRL_MAP_SPATIAL) {   // the clip's cls row feeds the cls row of each of its T frames, x 1/T (mean)
      m2 = (long long)b * g.T * (g.HW + 1);
      reps = g.T;
      ef = 1.0f / g.T;
    } else {
      return;                                  // MAP_SKIPCLS / MAP_PATCH have no cls rows
    }
  }
  OutT* out = static_cast<OutT*>(em.out);
  for (int r = 0; r < reps; ++r, m2 += g.HW + 1) {
    const float f = em.rowscale != nullptr ? __ldg(em.rowscale + static_cast<int>(m2) / em.rs_div) * ef : ef;
#pragma unroll
    for (int i = 0; i < NVEC; ++i) {
      const float4 o = make_float4(v[i].x * f, v[i].y * f, v[i].z * f, v[i].w * f);
      store4<OutT>(out + m2 * D + (i * 32 + lane) * 4, o.x, o.y, o.z, o.w);
      ln_acc<SMEM, NVEC>(ec, sm_ec, i, lane, o.x, o.y, o.z, o.w);
    }
  }
}

template <typename InT, int NVEC, bool EMIT, bool SMEM>
__global__ void __launch_bounds__(LN_WARPS * 32, (SMEM || NVEC <= 4) ? 2 : 1)
layernorm_bwd_kernel(const InT* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ x_cls,
                     const float* __restrict__ w, const float* __restrict__ stats, float* __restrict__ dx,
                     float* __restrict__ dw, float* __restrict__ db, int M, int map, Geom g, LnEmit em) {
  if (PDL_EARLY_TRIGGER) pdl_launch_dependents();
  pdl_wait();
  constexpr int D = NVEC * 128;
  constexpr int NACC = EMIT ? 3 : 2;
  // SMEM: [LN_WARPS][NACC][D] warp-private accumulators (dynamic); otherwise one [NACC][D] block-wide copy for the final fold
  extern __shared__ __align__(16) float sdyn[];
  __shared__ float sstat[SMEM ? 1 : NACC * D];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sacc = SMEM ? sdyn + warp * NACC * D : sstat;
  if constexpr (SMEM) {
    for (int i = lane; i < NACC * D / 4; i += 32) reinterpret_cast<float4*>(sacc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
  } else {
    for (int i = threadIdx.x; i < NACC * D; i += blockDim.x) sstat[i] = 0.f;
    __syncthreads();
  }
  float4 aw[NVEC], ab[NVEC], ec[NVEC];   // ec: column sums of the emitted rows (dead code without EMIT)
#pragma unroll
  for (int i = 0; i < NVEC; ++i) aw[i] = make_float4(0.f, 0.f, 0.f, 0.f), ab[i] = aw[i];
#pragma unroll
  for (int i = 0; i < NVEC; ++i) ec[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if constexpr (EMIT) {   // cls rows this pass does not touch: copied out of dx as they are (complete since the previous kernel)
    for (int b = blockIdx.x * LN_WARPS + warp; b < em.extra_cls; b += gridDim.x * LN_WARPS) {
      const long long xr = (long long)b * g.S;
      float4 cur[NVEC];
#pragma unroll
      for (int i = 0; i < NVEC; ++i) cur[i] = *(reinterpret_cast<const float4*>(dx + xr * D) + i * 32 + lane);
      ln_emit_row<InT, NVEC, SMEM>(em, g, xr, cur, ec, sacc + 2 * D, lane);
    }
  }
  // last row first: dy was just written by a GEMM whose final tiles are still in L2, and the gather that follows reads
  // dx from row 0 upwards -- the rows this kernel writes last
  for (int mm = blockIdx.x * LN_WARPS + warp; mm < M; mm += gridDim.x * LN_WARPS) {
    const int m = M - 1 - mm;
    const long long r = map_row(map, m, g);
    const bool cls = r < 0;
    const long long xr = cls ? ((-r - 1) / g.T) * (long long)g.S : r;
    const float* xp = (cls ? x_cls : x) + xr * D;
    const float2 st = reinterpret_cast<const float2*>(stats)[m];
    float4 xh[NVEC], gg[NVEC], cur[NVEC];
    float* dp = dx + xr * D;
#pragma unroll
    for (int i = 0; i < NVEC; ++i) {   // every load of the row -- including the dx it accumulates into -- is in flight
      xh[i] = __ldg(reinterpret_cast<const float4*>(xp) + i * 32 + lane);   // before the first use: one memory round
      gg[i] = load4<InT>(dy + (long long)m * D + (i * 32 + lane) * 4);       // trip per row instead of two
      if (!cls) cur[i] = *(reinterpret_cast<const float4*>(dp) + i * 32 + lane);
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NVEC; ++i) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(w) + i * 32 + lane);
      const float4 d = gg[i];
      xh[i] = make_float4((xh[i].x - st.x) * st.y, (xh[i].y - st.x) * st.y, (xh[i].z - st.x) * st.y,
                          (xh[i].w - st.x) * st.y);
      ln_acc<SMEM, NVEC>(aw, sacc, i, lane, d.x * xh[i].x, d.y * xh[i].y, d.z * xh[i].z, d.w * xh[i].w);
      ln_acc<SMEM, NVEC>(ab, sacc + D, i, lane, d.x, d.y, d.z, d.w);
      gg[i] = make_float4(d.x * wv.x, d.y * wv.y, d.z * wv.z, d.w * wv.w);
      s1 += gg[i].x + gg[i].y + gg[i].z + gg[i].w;
      s2 += gg[i].x * xh[i].x + gg[i].y * xh[i].y + gg[i].z * xh[i].z + gg[i].w * xh[i].w;
    }
    const float c1 = warp_sum(s1) * (1.0f / D), c2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
    for (int i = 0; i < NVEC; ++i) {
      const float4 o = make_float4(st.y * (gg[i].x - c1 - xh[i].x * c2), st.y * (gg[i].y - c1 - xh[i].y * c2),
                                   st.y * (gg[i].z - c1 - xh[i].z * c2), st.y * (gg[i].w - c1 - xh[i].w * c2));
      float* p = dp + (i * 32 + lane) * 4;
      if (cls) {  // T frames of one clip share the cls row (vit.py:138-140)
        atomicAdd(p, o.x), atomicAdd(p + 1, o.y), atomicAdd(p + 2, o.z), atomicAdd(p + 3, o.w);
      } else {
        cur[i] = make_float4(cur[i].x + o.x, cur[i].y + o.y, cur[i].z + o.z, cur[i].w + o.w);
        *reinterpret_cast<float4*>(p) = cur[i];
      }
    }
    if constexpr (EMIT) {
      if (!cls) ln_emit_row<InT, NVEC, SMEM>(em, g, xr, cur, ec, sacc + 2 * D, lane);
    }
  }
  if constexpr (!SMEM) {
#pragma unroll
    for (int i = 0; i < NVEC; ++i) {
      const int c = (i * 32 + lane) * 4;
      atomicAdd(&sacc[c], aw[i].x), atomicAdd(&sacc[c + 1], aw[i].y), atomicAdd(&sacc[c + 2], aw[i].z),
          atomicAdd(&sacc[c + 3], aw[i].w);
      atomicAdd(&sacc[D + c], ab[i].x), atomicAdd(&sacc[D + c + 1], ab[i].y), atomicAdd(&sacc[D + c + 2], ab[i].z),
          atomicAdd(&sacc[D + c + 3], ab[i].w);
    }
    if (EMIT && em.colsum != nullptr) {
#pragma unroll
      for (int i = 0; i < NVEC; ++i) {
        const int c = (i * 32 + lane) * 4;
        atomicAdd(&sacc[2 * D + c], ec[i].x), atomicAdd(&sacc[2 * D + c + 1], ec[i].y);
        atomicAdd(&sacc[2 * D + c + 2], ec[i].z), atomicAdd(&sacc[2 * D + c + 3], ec[i].w);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    float s[3] = {0.f, 0.f, 0.f};
    if constexpr (SMEM) {
      for (int wv = 0; wv < LN_WARPS; ++wv)
#pragma unroll
        for (int k = 0; k < NACC; ++k) s[k] += sdyn[(wv * NACC + k) * D + i];
    } else {
#pragma unroll
      for (int k = 0; k < NACC; ++k) s[k] = sstat[k * D + i];
    }
    if (dw != nullptr) atomicAdd(dw + i, s[0]);
    if (db != nullptr) atomicAdd(db + i, s[1]);
    if (EMIT && em.colsum != nullptr) atomicAdd(em.colsum + i, s[2]);
  }
}

// Default: column sums in warp-private shared memory (128 registers, two 8-warp blocks per SM: 53.7 us per launch at the
// bench shape); PVRL_LN_SMEM_ACC=0 keeps them in registers (254 registers, one block per SM: 71.3 us)
inline bool ln_smem_acc() {
  static const bool on = [] {
    const char* e = getenv("PVRL_LN_SMEM_ACC");
    return e == nullptr || atoi(e) != 0;
  }();
  return on;
}

template <typename InT, int NVEC, bool EMIT, bool SMEM>
int layernorm_bwd_launch3(const InT* d, const float* x, const float* x_cls, const float* w, const float* stats, float* dx,
                          float* dw, float* db, int M, int map, Geom gg, LnEmit em, int grid, cudaStream_t stream) {
  const size_t smem = SMEM ? sizeof(float) * LN_WARPS * (EMIT ? 3 : 2) * NVEC * 128 : 0;
  auto kern = layernorm_bwd_kernel<InT, NVEC, EMIT, SMEM>;
  if (smem > 48 * 1024) {
    static bool configured = false;      // one flag per instantiation
    if (!configured) {
      PVRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      configured = true;
    }
  }
  PVRL_CUDA(launch_pdl(kern, dim3(grid), dim3(LN_WARPS * 32), smem, stream, d, x, x_cls, w, stats, dx, dw, db, M, map, gg, em));
  return launched("layernorm_bwd_kernel");
}

template <typename InT, bool EMIT, bool SMEM>
int layernorm_bwd_launch2(const void* dy, const float* x, const float* x_cls, const float* w, const float* stats,
                          float* dx, float* dw, float* db, int M, int D, int map, Geom gg, LnEmit em, cudaStream_t stream) {
  int grid = (M + LN_WARPS - 1) / LN_WARPS;
  if (grid > 148 * 2) grid = 148 * 2;   // persistent row loop; 1-2 blocks are resident per SM (register-bound)
  const InT* d = static_cast<const InT*>(dy);
  switch (D / 128) {
    case 2: return layernorm_bwd_launch3<InT, 2, EMIT, SMEM>(d, x, x_cls, w, stats, dx, dw, db, M, map, gg, em, grid, stream);
    case 4: return layernorm_bwd_launch3<InT, 4, EMIT, SMEM>(d, x, x_cls, w, stats, dx, dw, db, M, map, gg, em, grid, stream);
    case 6: return layernorm_bwd_launch3<InT, 6, EMIT, SMEM>(d, x, x_cls, w, stats, dx, dw, db, M, map, gg, em, grid, stream);
    case 8: return layernorm_bwd_launch3<InT, 8, EMIT, SMEM>(d, x, x_cls, w, stats, dx, dw, db, M, map, gg, em, grid, stream);
    default: return fail(-1, "pvrl_layernorm_bwd: D=%d not in {256, 512, 768, 1024}", D);
  }
}

template <typename InT, bool EMIT>
int layernorm_bwd_launch(const void* dy, const float* x, const float* x_cls, const float* w, const float* stats,
                         float* dx, float* dw, float* db, int M, int D, int map, Geom gg, LnEmit em, cudaStream_t stream) {
  return ln_smem_acc() ? layernorm_bwd_launch2<InT, EMIT, true>(dy, x, x_cls, w, stats, dx, dw, db, M, D, map, gg, em, stream)
                       : layernorm_bwd_launch2<InT, EMIT, false>(dy, x, x_cls, w, stats, dx, dw, db, M, D, map, gg, em, stream);
}

// ---------------------------------------------------------------------------------------- gather + cast
// One warp per row (NVEC float4 per lane, the whole row in flight), persistent over the rows like the LayerNorm kernels
// -- the access pattern that streams HBM at ~5 TB/s here.  Because a lane keeps its columns across the rows its warp
// visits, the bias gradient colsum(out) comes for free: per-lane partial sums -> shared memory -> one fp32 atomic per
// column and block, and the separate pvrl_colsum pass over dY disappears.
template <typename OutT, int NVEC>
__global__ void __launch_bounds__(LN_WARPS * 32)
gather_cast_kernel(const float* __restrict__ src, OutT* __restrict__ out, const float* __restrict__ rowscale, int rs_div,
                   int M, int map, Geom g, float* __restrict__ colsum) {
  if (PDL_EARLY_TRIGGER) pdl_launch_dependents();
  pdl_wait();
  constexpr int D = NVEC * 128;
  __shared__ float sacc[D];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 acc[NVEC];
#pragma unroll
  for (int i = 0; i < NVEC; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int m = blockIdx.x * LN_WARPS + warp; m < M; m += gridDim.x * LN_WARPS) {
    long long r = map_row(map, m, g);
    float f = rowscale != nullptr ? __ldg(rowscale + m / rs_div) : 1.0f;
    if (r < 0) {  // cls row of a spatial sequence: d(mean over T) = 1/T
      r = ((-r - 1) / g.T) * (long long)g.S;
      f *= 1.0f / g.T;
    }
    float4 v[NVEC];
#pragma unroll
    for (int i = 0; i < NVEC; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(src + r * D) + i * 32 + lane);
#pragma unroll
    for (int i = 0; i < NVEC; ++i) {
      const float4 o = make_float4(v[i].x * f, v[i].y * f, v[i].z * f, v[i].w * f);
      store4<OutT>(out + (long long)m * D + (i * 32 + lane) * 4, o.x, o.y, o.z, o.w);
      acc[i].x += o.x, acc[i].y += o.y, acc[i].z += o.z, acc[i].w += o.w;
    }
  }
  if (colsum != nullptr) {
    for (int i = threadIdx.x; i < D; i += blockDim.x) sacc[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NVEC; ++i) {
      const int c = (i * 32 + lane) * 4;
      atomicAdd(&sacc[c], acc[i].x), atomicAdd(&sacc[c + 1], acc[i].y), atomicAdd(&sacc[c + 2], acc[i].z),
          atomicAdd(&sacc[c + 3], acc[i].w);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += blockDim.x) atomicAdd(colsum + i, sacc[i]);
  }
}

template <typename OutT>
int gather_cast_launch(const float* src, void* out, const float* rowscale, int rs_div, int M, int D, int map, Geom gg,
                       float* colsum, cudaStream_t stream) {
  int grid = (M + LN_WARPS - 1) / LN_WARPS;
  if (grid > 148 * 6) grid = 148 * 6;
  OutT* o = static_cast<OutT*>(out);
  switch (D / 128) {
    case 2: launch_pdl(gather_cast_kernel<OutT, 2>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, src, o, rowscale, rs_div, M, map, gg, colsum); break;
    case 4: launch_pdl(gather_cast_kernel<OutT, 4>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, src, o, rowscale, rs_div, M, map, gg, colsum); break;
    case 6: launch_pdl(gather_cast_kernel<OutT, 6>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, src, o, rowscale, rs_div, M, map, gg, colsum); break;
    case 8: launch_pdl(gather_cast_kernel<OutT, 8>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, src, o, rowscale, rs_div, M, map, gg, colsum); break;
    default: return fail(-1, "pvrl_gather_cast: D=%d not in {256, 512, 768, 1024}", D);
  }
  return launched("gather_cast_kernel");
}

__global__ void cls_merge_kernel(const float* __restrict__ x0, const float* __restrict__ side, float* __restrict__ x2,
                                 int T, int S, int D) {
  if (PDL_EARLY_TRIGGER) pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += side[((long long)b * T + t) * D + d];
    x2[(long long)b * S * D + d] = x0[(long long)b * S * D + d] + s / T;
  }
}

// ---------------------------------------------------------------------------------------- column sums
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ a, long long lda, float* __restrict__ out, int M, int N,
                              int rows_per_block) {
  if (PDL_EARLY_TRIGGER) pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8][128];
  const int col = blockIdx.x * 128 + threadIdx.x * 4;  // blockDim = (32, 8)
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < N)
    for (int r = r0 + threadIdx.y; r < r1; r += 8) {
      const float4 v = load4<T>(a + (long long)r * lda + col);
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
  float* rr = &red[threadIdx.y][threadIdx.x * 4];
  rr[0] = acc.x, rr[1] = acc.y, rr[2] = acc.z, rr[3] = acc.w;
  __syncthreads();
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if (tid < 128 && blockIdx.x * 128 + tid < N) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][tid];
    atomicAdd(out + blockIdx.x * 128 + tid, s);
  }
}

// ---------------------------------------------------------------------------------------- weight cast
template <typename OutT>
__global__ void cast_weight_kernel(const float* __restrict__ w, OutT* __restrict__ o, OutT* __restrict__ oT, int rows,
                                   int cols) {
  __shared__ float tile[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = blockIdx.y * 32 + j;
    float v = 0.f;
    if (r < rows && c < cols) {
      v = w[(long long)r * cols + c];
      if (o != nullptr) o[(long long)r * cols + c] = static_cast<OutT>(v);
    }
    tile[j][threadIdx.x] = v;
  }
  __syncthreads();
  if (oT != nullptr) {
    const int r = blockIdx.y * 32 + threadIdx.x;  // original row -> fast index of the transposed output
    for (int j = threadIdx.y; j < 32; j += 8) {
      const int cc = blockIdx.x * 32 + j;
      if (r < rows && cc < cols) oT[(long long)cc * rows + r] = static_cast<OutT>(tile[threadIdx.x][j]);
    }
  }
}

// All weights of the model in ONE launch (85 cast_weight launches per optimizer step otherwise): the descriptor table
// lives in device memory, a block finds its matrix by binary search over the tile offsets.
template <typename OutT>
__device__ __forceinline__ void st_pair(OutT* dst, float a, float b);
template <>
__device__ __forceinline__ void st_pair<float>(float* dst, float a, float b) {
  *reinterpret_cast<float2*>(dst) = make_float2(a, b);
}
template <>
__device__ __forceinline__ void st_pair<__nv_bfloat16>(__nv_bfloat16* dst, float a, float b) {
  *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(a, b);
}

// 64 x 64 tiles, two elements per thread: every global access of the copy AND of the transposed copy is a full
// 128-byte (bf16) row segment per warp.  rows and cols must be even (all Linear weights here are multiples of 64).
template <typename OutT>
__global__ void __launch_bounds__(256)
cast_weight_multi_kernel(const pvrl_cast_desc_t* __restrict__ descs, int n) {
  __shared__ float tile[64][65];
  int lo = 0, hi = n - 1;
  while (lo < hi) {   // last descriptor with tile0 <= blockIdx.x
    const int mid = (lo + hi + 1) >> 1;
    if (descs[mid].tile0 <= static_cast<int>(blockIdx.x)) lo = mid;
    else hi = mid - 1;
  }
  const pvrl_cast_desc_t d = descs[lo];
  const int t = blockIdx.x - d.tile0;
  const int bx = t % d.tiles_x, by = t / d.tiles_x;
  const float* w = d.w;
  OutT* o = static_cast<OutT*>(d.out);
  OutT* oT = static_cast<OutT*>(d.outT);
  const int rows = d.rows, cols = d.cols;
  const int c = bx * 64 + threadIdx.x * 2;
  for (int j = threadIdx.y; j < 64; j += 8) {
    const int r = by * 64 + j;
    float2 v = make_float2(0.f, 0.f);
    if (r < rows && c < cols) {
      v = *reinterpret_cast<const float2*>(w + (long long)r * cols + c);
      if (o != nullptr) st_pair<OutT>(o + (long long)r * cols + c, v.x, v.y);
    }
    tile[j][threadIdx.x * 2] = v.x, tile[j][threadIdx.x * 2 + 1] = v.y;
  }
  __syncthreads();
  if (oT != nullptr) {
    const int r = by * 64 + threadIdx.x * 2;   // original rows (2 per thread) -> fast index of the transposed output
    for (int j = threadIdx.y; j < 64; j += 8) {
      const int cc = bx * 64 + j;
      if (r < rows && cc < cols)
        st_pair<OutT>(oT + (long long)cc * rows + r, tile[threadIdx.x * 2][j], tile[threadIdx.x * 2 + 1][j]);
    }
  }
}

// a = hi + lo (+ O(2^-17 |a|)); the three K-concatenated segments pair up as hi*hi + hi*lo + lo*hi.
__global__ void split3_kernel(const float* __restrict__ a, __nv_bfloat16* __restrict__ out, int M, int K, int pattern,
                              int along) {
  const long long total = (long long)M * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const float v = a[i];
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    const __nv_bfloat16 s1 = pattern == 0 ? hi : lo, s2 = pattern == 0 ? lo : hi;
    if (along == 1) {
      const long long m = i / K, k = i - m * K;
      __nv_bfloat16* o = out + m * 3 * K + k;
      o[0] = hi, o[K] = s1, o[2 * (long long)K] = s2;
    } else {
      out[i] = hi, out[total + i] = s1, out[2 * total + i] = s2;
    }
  }
}

// ---------------------------------------------------------------------------------------- embed gradients
__global__ void embed_bwd_kernel(const float* __restrict__ dx, float* __restrict__ dcls, float* __restrict__ dpos,
                                 float* __restrict__ dtime, int Bc, int D, Geom g) {
  const int n = blockIdx.x;  // 0..HW-1 : patch index; block HW handles the cls row
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    if (n == g.HW) {
      float s = 0.f;
      for (int b = 0; b < Bc; ++b) s += dx[(long long)b * g.S * D + d];
      if (dcls != nullptr) dcls[d] += s;
      if (dpos != nullptr) dpos[d] += s;
      continue;
    }
    float pos = 0.f;
    for (int t = 0; t < g.T; ++t) {
      float s = 0.f;
      for (int b = 0; b < Bc; ++b) s += dx[((long long)b * g.S + 1 + (long long)n * g.T + t) * D + d];
      pos += s;
      if (dtime != nullptr) atomicAdd(dtime + (long long)t * D + d, s);
    }
    if (dpos != nullptr) dpos[(long long)(1 + n) * D + d] += pos;
  }
}

inline int grid_for(long long work, int block, int cap_per_sm = 16) {
  long long g = (work + block - 1) / block;
  const long long cap = 148LL * cap_per_sm;
  return static_cast<int>(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace
}  // namespace pvrl

using namespace pvrl;
#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int pvrl_patchify(const float* frames, void* out, int32_t out_dtype, int32_t Bc, int32_t T, int32_t H,
                             int32_t W, int32_t patch, void* stream) {
  PVRL_CHECK_ARG(frames && out && Bc > 0 && T > 0, "pvrl_patchify: bad arguments");
  PVRL_CHECK_ARG(patch % 4 == 0 && H % patch == 0 && W % patch == 0, "pvrl_patchify: H=%d W=%d patch=%d", H, W, patch);
  const long long total_vec = (long long)Bc * T * (H / patch) * (W / patch) * 3 * patch * patch / 4;
  const int grid = grid_for(total_vec, 256);
  if (out_dtype == PVRL_F32)
    patchify_kernel<float><<<grid, 256, 0, STREAM>>>(frames, static_cast<float*>(out), Bc, T, H, W, patch, total_vec);
  else
    patchify_kernel<__nv_bfloat16>
        <<<grid, 256, 0, STREAM>>>(frames, static_cast<__nv_bfloat16*>(out), Bc, T, H, W, patch, total_vec);
  return launched("patchify_kernel");
}

extern "C" int pvrl_patchify_u8(const uint8_t* frames, void* out, int32_t out_dtype, int32_t Bc, int32_t T, int32_t H,
                                int32_t W, int32_t patch, const float* mean3, const float* std3, void* stream) {
  PVRL_CHECK_ARG(frames && out && mean3 && std3 && Bc > 0 && T > 0, "pvrl_patchify_u8: bad arguments");
  PVRL_CHECK_ARG(patch % 4 == 0 && H % patch == 0 && W % patch == 0, "pvrl_patchify_u8: H=%d W=%d patch=%d", H, W, patch);
  PVRL_CHECK_ARG(std3[0] != 0.f && std3[1] != 0.f && std3[2] != 0.f, "pvrl_patchify_u8: zero std");
  const long long total_vec = (long long)Bc * T * (H / patch) * (W / patch) * 3 * patch * patch / 4;
  const int grid = grid_for(total_vec, 256);
  const float3 scale = make_float3(1.0f / (255.0f * std3[0]), 1.0f / (255.0f * std3[1]), 1.0f / (255.0f * std3[2]));
  const float3 shift = make_float3(-mean3[0] / std3[0], -mean3[1] / std3[1], -mean3[2] / std3[2]);
  if (out_dtype == PVRL_F32)
    patchify_u8_kernel<float><<<grid, 256, 0, STREAM>>>(frames, static_cast<float*>(out), Bc, T, H, W, patch, total_vec,
                                                        scale, shift);
  else
    patchify_u8_kernel<__nv_bfloat16><<<grid, 256, 0, STREAM>>>(frames, static_cast<__nv_bfloat16*>(out), Bc, T, H, W, patch,
                                                                total_vec, scale, shift);
  return launched("patchify_u8_kernel");
}

extern "C" int pvrl_cls_init(float* x, const float* cls_token, const float* pos_embed, int32_t Bc, int32_t S,
                             int32_t D, void* stream) {
  PVRL_CHECK_ARG(x && cls_token && pos_embed && Bc > 0, "pvrl_cls_init: bad arguments");
  cls_init_kernel<<<Bc, 256, 0, STREAM>>>(x, cls_token, pos_embed, S, D);
  return launched("cls_init_kernel");
}

extern "C" int pvrl_layernorm_fwd(const float* x, const float* x_cls, const float* w, const float* b, void* y,
                                  int32_t y_dtype, float* stats, int32_t M, int32_t D, float eps, int32_t map,
                                  pvrl_geom_t g, void* stream) {
  PVRL_CHECK_ARG(x && w && b && y && M > 0, "pvrl_layernorm_fwd: bad arguments");
  PVRL_CHECK_ARG(D % 128 == 0 && D <= 128 * LN_MAX_VEC, "pvrl_layernorm_fwd: D=%d must be a multiple of 128, <= 1024", D);
  PVRL_CHECK_ARG(map != PVRL_MAP_SPATIAL || x_cls != nullptr, "pvrl_layernorm_fwd: MAP_SPATIAL needs x_cls");
  const Geom gg(g.T > 0 ? g.T : 1, g.HW > 0 ? g.HW : 1);
  const int grid = (M + LN_WARPS - 1) / LN_WARPS;
  if (y_dtype == PVRL_F32)
    launch_pdl(layernorm_fwd_kernel<float>, dim3(grid), dim3(LN_WARPS * 32), 0, STREAM, x, x_cls, w, b,
               static_cast<float*>(y), stats, M, D, eps, map, gg);
  else
    launch_pdl(layernorm_fwd_kernel<__nv_bfloat16>, dim3(grid), dim3(LN_WARPS * 32), 0, STREAM, x, x_cls, w, b,
               static_cast<__nv_bfloat16*>(y), stats, M, D, eps, map, gg);
  return launched("layernorm_fwd_kernel");
}

extern "C" int pvrl_layernorm_bwd(const void* dy, int32_t dy_dtype, const float* x, const float* x_cls, const float* w,
                                  const float* stats, float* dx, float* dw, float* db, int32_t M, int32_t D,
                                  int32_t map, pvrl_geom_t g, void* stream) {
  PVRL_CHECK_ARG(dy && x && w && stats && dx && M > 0, "pvrl_layernorm_bwd: bad arguments");
  PVRL_CHECK_ARG(D % 128 == 0 && D <= 128 * LN_MAX_VEC, "pvrl_layernorm_bwd: D=%d must be a multiple of 128, <= 1024", D);
  PVRL_CHECK_ARG(map != PVRL_MAP_SPATIAL || x_cls != nullptr, "pvrl_layernorm_bwd: MAP_SPATIAL needs x_cls");
  const Geom gg(g.T > 0 ? g.T : 1, g.HW > 0 ? g.HW : 1);
  const LnEmit none = {};
  return dy_dtype == PVRL_F32
             ? layernorm_bwd_launch<float, false>(dy, x, x_cls, w, stats, dx, dw, db, M, D, map, gg, none, STREAM)
             : layernorm_bwd_launch<__nv_bfloat16, false>(dy, x, x_cls, w, stats, dx, dw, db, M, D, map, gg, none, STREAM);
}

extern "C" int pvrl_layernorm_bwd_emit(const void* dy, int32_t dy_dtype, const float* x, const float* x_cls,
                                       const float* w, const float* stats, float* dx, float* dw, float* db, int32_t M,
                                       int32_t D, int32_t map, pvrl_geom_t g, void* emit_out, int32_t emit_map,
                                       const float* emit_rowscale, int32_t emit_rs_div, float* emit_colsum,
                                       void* stream) {
  PVRL_CHECK_ARG(dy && x && w && stats && dx && M > 0 && emit_out, "pvrl_layernorm_bwd_emit: bad arguments");
  PVRL_CHECK_ARG(D % 128 == 0 && D <= 128 * LN_MAX_VEC, "pvrl_layernorm_bwd_emit: D=%d must be a multiple of 128, <= 1024", D);
  PVRL_CHECK_ARG(map != PVRL_MAP_SPATIAL || x_cls != nullptr, "pvrl_layernorm_bwd_emit: MAP_SPATIAL needs x_cls");
  PVRL_CHECK_ARG(emit_rowscale == nullptr || emit_rs_div > 0, "pvrl_layernorm_bwd_emit: rowscale needs rs_div > 0");
  const Geom gg(g.T > 0 ? g.T : 1, g.HW > 0 ? g.HW : 1);
  // the pairs for which "every row of emit_map is addressed by a row this pass completes" holds (vit.py:130-157 backward)
  const bool ok = (map == PVRL_MAP_IDENT && emit_map == PVRL_MAP_SPATIAL) ||
                  (map == PVRL_MAP_SPATIAL && emit_map == PVRL_MAP_SKIPCLS) ||
                  (map == PVRL_MAP_SKIPCLS && (emit_map == PVRL_MAP_IDENT || emit_map == PVRL_MAP_PATCH));
  PVRL_CHECK_ARG(ok, "pvrl_layernorm_bwd_emit: map %d cannot emit map %d", map, emit_map);
  const int rows_per_clip = map == PVRL_MAP_IDENT ? gg.S : map == PVRL_MAP_SPATIAL ? gg.T * (gg.HW + 1) : gg.L;
  PVRL_CHECK_ARG(M % rows_per_clip == 0, "pvrl_layernorm_bwd_emit: M=%d is not a whole number of clips", M);
  LnEmit em;
  em.out = emit_out, em.rowscale = emit_rowscale, em.colsum = emit_colsum, em.map = emit_map;
  em.rs_div = emit_rs_div > 0 ? emit_rs_div : 1;
  em.extra_cls = (map == PVRL_MAP_SKIPCLS && emit_map == PVRL_MAP_IDENT) ? M / rows_per_clip : 0;
  return dy_dtype == PVRL_F32
             ? layernorm_bwd_launch<float, true>(dy, x, x_cls, w, stats, dx, dw, db, M, D, map, gg, em, STREAM)
             : layernorm_bwd_launch<__nv_bfloat16, true>(dy, x, x_cls, w, stats, dx, dw, db, M, D, map, gg, em, STREAM);
}

extern "C" int pvrl_gather_cast(const float* src, void* out, int32_t out_dtype, const float* rowscale, int32_t rs_div,
                                int32_t M, int32_t D, int32_t map, pvrl_geom_t g, float* colsum, void* stream) {
  PVRL_CHECK_ARG(src && out && M > 0 && D % 4 == 0 && D <= 4096, "pvrl_gather_cast: bad arguments");
  PVRL_CHECK_ARG(rowscale == nullptr || rs_div > 0, "pvrl_gather_cast: rowscale needs rs_div > 0");
  const Geom gg(g.T > 0 ? g.T : 1, g.HW > 0 ? g.HW : 1);
  const int rd = rs_div > 0 ? rs_div : 1;
  return out_dtype == PVRL_F32 ? gather_cast_launch<float>(src, out, rowscale, rd, M, D, map, gg, colsum, STREAM)
                               : gather_cast_launch<__nv_bfloat16>(src, out, rowscale, rd, M, D, map, gg, colsum, STREAM);
}

extern "C" int pvrl_cls_merge(const float* x0, const float* side, float* x2, int32_t Bc, int32_t T, int32_t S,
                              int32_t D, void* stream) {
  PVRL_CHECK_ARG(x0 && side && x2 && Bc > 0 && T > 0, "pvrl_cls_merge: bad arguments");
  launch_pdl(cls_merge_kernel, dim3(Bc), dim3(256), 0, STREAM, x0, side, x2, T, S, D);
  return launched("cls_merge_kernel");
}

extern "C" int pvrl_colsum(const void* a, int32_t a_dtype, int64_t lda, float* out, int32_t M, int32_t N,
                           void* stream) {
  PVRL_CHECK_ARG(a && out && M > 0 && N > 0 && N % 4 == 0, "pvrl_colsum: bad arguments");
  const int rows_per_block = 256;
  dim3 grid((N + 127) / 128, (M + rows_per_block - 1) / rows_per_block), block(32, 8);
  if (a_dtype == PVRL_F32)
    launch_pdl(colsum_kernel<float>, grid, block, 0, STREAM, static_cast<const float*>(a), (long long)lda, out, M, N,
               rows_per_block);
  else
    launch_pdl(colsum_kernel<__nv_bfloat16>, grid, block, 0, STREAM, static_cast<const __nv_bfloat16*>(a), (long long)lda,
               out, M, N, rows_per_block);
  return launched("colsum_kernel");
}

extern "C" int pvrl_cast_weight(const float* w, void* w_out, void* wT_out, int32_t out_dtype, int32_t rows,
                                int32_t cols, void* stream) {
  PVRL_CHECK_ARG(w && (w_out || wT_out) && rows > 0 && cols > 0, "pvrl_cast_weight: bad arguments");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  if (out_dtype == PVRL_F32)
    cast_weight_kernel<float>
        <<<grid, block, 0, STREAM>>>(w, static_cast<float*>(w_out), static_cast<float*>(wT_out), rows, cols);
  else
    cast_weight_kernel<__nv_bfloat16><<<grid, block, 0, STREAM>>>(w, static_cast<__nv_bfloat16*>(w_out),
                                                                   static_cast<__nv_bfloat16*>(wT_out), rows, cols);
  return launched("cast_weight_kernel");
}

extern "C" int pvrl_cast_weight_multi(const pvrl_cast_desc_t* descs_dev, int32_t n, int32_t total_tiles,
                                      int32_t out_dtype, void* stream) {
  PVRL_CHECK_ARG(descs_dev && n > 0 && total_tiles > 0, "pvrl_cast_weight_multi: bad arguments");   // 64 x 64 tiles
  dim3 block(32, 8);
  if (out_dtype == PVRL_F32)
    cast_weight_multi_kernel<float><<<total_tiles, block, 0, STREAM>>>(descs_dev, n);
  else
    cast_weight_multi_kernel<__nv_bfloat16><<<total_tiles, block, 0, STREAM>>>(descs_dev, n);
  return launched("cast_weight_multi_kernel");
}

extern "C" int pvrl_split3(const float* a, void* out, int32_t M, int32_t K, int32_t pattern, int32_t along,
                           void* stream) {
  PVRL_CHECK_ARG(a && out && M > 0 && K > 0, "pvrl_split3: bad arguments");
  PVRL_CHECK_ARG((pattern == 0 || pattern == 1) && (along == 0 || along == 1), "pvrl_split3: bad pattern/along");
  const int grid = grid_for((long long)M * K, 256);
  split3_kernel<<<grid, 256, 0, STREAM>>>(a, static_cast<__nv_bfloat16*>(out), M, K, pattern, along);
  return launched("split3_kernel");
}

extern "C" int pvrl_embed_bwd(const float* dx, float* dcls, float* dpos, float* dtime, int32_t Bc, int32_t D,
                              pvrl_geom_t g, void* stream) {
  PVRL_CHECK_ARG(dx && Bc > 0 && D > 0 && g.T > 0 && g.HW > 0, "pvrl_embed_bwd: bad arguments");
  const Geom gg(g.T, g.HW);
  embed_bwd_kernel<<<g.HW + 1, 256, 0, STREAM>>>(dx, dcls, dpos, dtime, Bc, D, gg);
  return launched("embed_bwd_kernel");
}
