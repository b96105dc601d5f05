// Exact fp32 softmax attention on CUDA cores for head_dim 64 and any sequence length (forward) /
// seq <= 208 (backward).  Used for the temporal axis (T = 4..32 tokens per sequence: the op is HBM-bound,
// one thread per query row, K/V broadcast from shared memory), for the parity ("bf16x3") mode and as the
// on-device cross-check of the tcgen05 spatial kernel.   Attention.forward, vit.py:84-88.
#include "pvrl_host.h"
#include "pvrl_ptx.cuh"

namespace pvrl {
namespace {

constexpr int HD = 64;        // head dim
constexpr int ROW = HD + 4;   // padded smem row (floats): conflict-free 128-bit row-per-lane access
constexpr int KEY_BLOCK = 256;

template <typename T>
__device__ __forceinline__ void load_row64(const T* src, float* dst);  // 64 contiguous elements -> fp32 smem/regs
template <>
__device__ __forceinline__ void load_row64<float>(const float* src, float* dst) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
    dst[4 * i] = v.x, dst[4 * i + 1] = v.y, dst[4 * i + 2] = v.z, dst[4 * i + 3] = v.w;
  }
}
template <>
__device__ __forceinline__ void load_row64<__nv_bfloat16>(const __nv_bfloat16* src, float* dst) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src) + i);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    dst[8 * i] = a.x, dst[8 * i + 1] = a.y, dst[8 * i + 2] = b.x, dst[8 * i + 3] = b.y;
    dst[8 * i + 4] = c.x, dst[8 * i + 5] = c.y, dst[8 * i + 6] = d.x, dst[8 * i + 7] = d.y;
  }
}
template <typename T>
__device__ __forceinline__ void store_row64(T* dst, const float* src);
template <>
__device__ __forceinline__ void store_row64<float>(float* dst, const float* src) {
#pragma unroll
  for (int i = 0; i < 16; ++i)
    reinterpret_cast<float4*>(dst)[i] = make_float4(src[4 * i], src[4 * i + 1], src[4 * i + 2], src[4 * i + 3]);
}
template <>
__device__ __forceinline__ void store_row64<__nv_bfloat16>(__nv_bfloat16* dst, const float* src) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint4 u;
    u.x = pack_bf16x2(src[8 * i], src[8 * i + 1]);
    u.y = pack_bf16x2(src[8 * i + 2], src[8 * i + 3]);
    u.z = pack_bf16x2(src[8 * i + 4], src[8 * i + 5]);
    u.w = pack_bf16x2(src[8 * i + 6], src[8 * i + 7]);
    reinterpret_cast<uint4*>(dst)[i] = u;
  }
}

// cooperative copy of `rows` rows of 64 elements (global row pitch `pitch` elements) into padded fp32 smem
template <typename T>
__device__ __forceinline__ void stage_rows(const T* src, long long pitch, int rows, float* dst) {
  // 8 elements per thread-iteration: 8 threads cover one row
  for (int i = threadIdx.x; i < rows * 8; i += blockDim.x) {
    const int r = i >> 3, c = (i & 7) * 8;
    const T* s = src + (long long)r * pitch + c;
    float* d = dst + r * ROW + c;
    if (sizeof(T) == 2) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(s));
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), dd = unpack_bf16x2(u.w);
      d[0] = a.x, d[1] = a.y, d[2] = b.x, d[3] = b.y, d[4] = cc.x, d[5] = cc.y, d[6] = dd.x, d[7] = dd.y;
    } else {
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(s)), v1 = __ldg(reinterpret_cast<const float4*>(s) + 1);
      d[0] = v0.x, d[1] = v0.y, d[2] = v0.z, d[3] = v0.w, d[4] = v1.x, d[5] = v1.y, d[6] = v1.z, d[7] = v1.w;
    }
  }
}

__device__ __forceinline__ float dot64(const float* __restrict__ a_reg, const float* __restrict__ b_smem) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float4 k = *reinterpret_cast<const float4*>(b_smem + 4 * i);
    s0 = fmaf(a_reg[4 * i], k.x, s0), s1 = fmaf(a_reg[4 * i + 1], k.y, s1);
    s2 = fmaf(a_reg[4 * i + 2], k.z, s2), s3 = fmaf(a_reg[4 * i + 3], k.w, s3);
  }
  return (s0 + s1) + (s2 + s3);
}
__device__ __forceinline__ float dot64_ss(const float* __restrict__ a_smem, const float* __restrict__ b_smem) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float4 a = *reinterpret_cast<const float4*>(a_smem + 4 * i);
    const float4 k = *reinterpret_cast<const float4*>(b_smem + 4 * i);
    s0 = fmaf(a.x, k.x, s0), s1 = fmaf(a.y, k.y, s1), s2 = fmaf(a.z, k.z, s2), s3 = fmaf(a.w, k.w, s3);
  }
  return (s0 + s1) + (s2 + s3);
}
__device__ __forceinline__ void axpy64(float* __restrict__ acc_reg, float a, const float* __restrict__ x_smem) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float4 v = *reinterpret_cast<const float4*>(x_smem + 4 * i);
    acc_reg[4 * i] = fmaf(a, v.x, acc_reg[4 * i]), acc_reg[4 * i + 1] = fmaf(a, v.y, acc_reg[4 * i + 1]);
    acc_reg[4 * i + 2] = fmaf(a, v.z, acc_reg[4 * i + 2]), acc_reg[4 * i + 3] = fmaf(a, v.w, acc_reg[4 * i + 3]);
  }
}

// grid.x = ceil(n_seq*H / pairs), grid.y = query tiles; block = pairs * q_per_pair threads.
// A "pair" is one (sequence, head).  Thread (pair p, query i) keeps q, the running max/sum and o in registers.
template <typename T>
__global__ void __launch_bounds__(256)
attn_fwd_kernel(const T* __restrict__ qkv, T* __restrict__ out, float* __restrict__ lse, int n_seq, int seq, int H,
                float scale, int pairs, int q_per_pair) {
  extern __shared__ float smem[];
  const int kb = seq < KEY_BLOCK ? seq : KEY_BLOCK;
  float* sK = smem;                         // [pairs][kb][ROW]
  float* sV = smem + (size_t)pairs * kb * ROW;
  const int p = threadIdx.x / q_per_pair;
  const int qi = blockIdx.y * q_per_pair + threadIdx.x % q_per_pair;
  const long long pair0 = (long long)blockIdx.x * pairs;
  const long long pair = pair0 + p;
  const long long n_pairs = (long long)n_seq * H;
  const int C = H * HD;
  const long long pitch = 3LL * C;
  const bool active = pair < n_pairs && qi < seq;
  const int s_idx = static_cast<int>(pair / H), h = static_cast<int>(pair % H);

  float q[HD], o[HD];
  float mx = -INFINITY, l = 0.f;
#pragma unroll
  for (int d = 0; d < HD; ++d) o[d] = 0.f;
  if (active) {
    load_row64<T>(qkv + ((long long)s_idx * seq + qi) * pitch + h * HD, q);
#pragma unroll
    for (int d = 0; d < HD; ++d) q[d] *= scale;
  }
  for (int k0 = 0; k0 < seq; k0 += kb) {
    const int nk = min(kb, seq - k0);
    __syncthreads();
    for (int pp = 0; pp < pairs; ++pp) {
      const long long pr = pair0 + pp;
      if (pr >= n_pairs) break;
      const int ss = static_cast<int>(pr / H), hh = static_cast<int>(pr % H);
      const T* base = qkv + ((long long)ss * seq + k0) * pitch + hh * HD;
      stage_rows<T>(base + C, pitch, nk, sK + (size_t)pp * kb * ROW);
      stage_rows<T>(base + 2 * C, pitch, nk, sV + (size_t)pp * kb * ROW);
    }
    __syncthreads();
    if (active) {
      const float* kp = sK + (size_t)p * kb * ROW;
      const float* vp = sV + (size_t)p * kb * ROW;
      for (int j = 0; j < nk; ++j) {
        const float s = dot64(q, kp + j * ROW);
        if (s > mx) {
          const float corr = __expf(mx - s);  // exp(-inf) = 0 on the first key
          l *= corr;
#pragma unroll
          for (int d = 0; d < HD; ++d) o[d] *= corr;
          mx = s;
        }
        const float pj = __expf(s - mx);
        l += pj;
        axpy64(o, pj, vp + j * ROW);
      }
    }
  }
  if (active) {
    const float inv = 1.0f / l;
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] *= inv;
    store_row64<T>(out + ((long long)s_idx * seq + qi) * C + h * HD, o);
    if (lse != nullptr) lse[pair * seq + qi] = mx + __logf(l);
  }
}

// One CTA per group of `pairs` (sequence, head) pairs, whole sequence resident (seq <= 208).
// Phase 1: thread = query  -> dq.   Phase 2: thread = key -> dk, dv.   No atomics, deterministic.
template <typename T>
__global__ void __launch_bounds__(256)
attn_bwd_kernel(const T* __restrict__ qkv, const T* __restrict__ out, const T* __restrict__ dout,
                const float* __restrict__ lse, T* __restrict__ dqkv, int n_seq, int seq, int H, float scale,
                int pairs) {
  extern __shared__ float smem[];
  const size_t blk = (size_t)pairs * seq * ROW;
  float* sQ = smem;
  float* sK = sQ + blk;
  float* sV = sK + blk;
  float* sdO = sV + blk;
  float* sLse = sdO + blk;                 // [pairs*seq]
  float* sDelta = sLse + (size_t)pairs * seq;
  const long long pair0 = (long long)blockIdx.x * pairs;
  const long long n_pairs = (long long)n_seq * H;
  const int C = H * HD;
  const long long pitch = 3LL * C;

  for (int pp = 0; pp < pairs; ++pp) {
    const long long pr = pair0 + pp;
    if (pr >= n_pairs) break;
    const int ss = static_cast<int>(pr / H), hh = static_cast<int>(pr % H);
    const T* base = qkv + (long long)ss * seq * pitch + hh * HD;
    stage_rows<T>(base, pitch, seq, sQ + (size_t)pp * seq * ROW);
    stage_rows<T>(base + C, pitch, seq, sK + (size_t)pp * seq * ROW);
    stage_rows<T>(base + 2 * C, pitch, seq, sV + (size_t)pp * seq * ROW);
    stage_rows<T>(dout + (long long)ss * seq * C + hh * HD, C, seq, sdO + (size_t)pp * seq * ROW);
  }
  const int p = threadIdx.x / seq, i = threadIdx.x % seq;
  const long long pair = pair0 + p;
  const bool active = p < pairs && pair < n_pairs;
  const int s_idx = static_cast<int>(pair / H), h = static_cast<int>(pair % H);
  if (active) {
    // delta_i = dO_i . O_i  (O read straight from global; each thread its own row)
    float orow[HD], dorow[HD];
    load_row64<T>(out + ((long long)s_idx * seq + i) * C + h * HD, orow);
    load_row64<T>(dout + ((long long)s_idx * seq + i) * C + h * HD, dorow);
    float dsum = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) dsum = fmaf(orow[d], dorow[d], dsum);
    sDelta[p * seq + i] = dsum;
    sLse[p * seq + i] = lse[pair * seq + i];
  }
  __syncthreads();
  if (active) {
    const float* Qp = sQ + (size_t)p * seq * ROW;
    const float* Kp = sK + (size_t)p * seq * ROW;
    const float* Vp = sV + (size_t)p * seq * ROW;
    const float* dOp = sdO + (size_t)p * seq * ROW;
    const float* Lp = sLse + p * seq;
    const float* Dp = sDelta + p * seq;
    float acc[HD], acc2[HD];
    // ---- phase 1: dq_i = scale * sum_j ds_ij k_j
    {
      const float li = Lp[i], di = Dp[i];
#pragma unroll
      for (int d = 0; d < HD; ++d) acc[d] = 0.f;
      for (int j = 0; j < seq; ++j) {
        const float s = dot64_ss(Qp + i * ROW, Kp + j * ROW) * scale;
        const float pij = __expf(s - li);
        const float dp = dot64_ss(dOp + i * ROW, Vp + j * ROW);
        axpy64(acc, pij * (dp - di) * scale, Kp + j * ROW);
      }
      store_row64<T>(dqkv + ((long long)s_idx * seq + i) * pitch + h * HD, acc);
    }
    // ---- phase 2: this thread is key j = i:  dv_j = sum_i p_ij dO_i ;  dk_j = scale * sum_i ds_ij q_i
    {
      const int j = i;
#pragma unroll
      for (int d = 0; d < HD; ++d) acc[d] = 0.f, acc2[d] = 0.f;
      for (int qi = 0; qi < seq; ++qi) {
        const float s = dot64_ss(Qp + qi * ROW, Kp + j * ROW) * scale;
        const float pij = __expf(s - Lp[qi]);
        const float dp = dot64_ss(dOp + qi * ROW, Vp + j * ROW);
        axpy64(acc2, pij, dOp + qi * ROW);
        axpy64(acc, pij * (dp - Dp[qi]) * scale, Qp + qi * ROW);
      }
      store_row64<T>(dqkv + ((long long)s_idx * seq + j) * pitch + C + h * HD, acc);
      store_row64<T>(dqkv + ((long long)s_idx * seq + j) * pitch + 2 * C + h * HD, acc2);
    }
  }
}


// ---- long sequences (joint space-time attention: 1 + HW*T = 1569 tokens, vit.py:124-127) --------------------------------
// Flash-style backward on CUDA cores: nothing of size seq x seq is materialised.  K_dq: thread = query (q, dO, dq in
// registers), keys / values stream through shared memory in blocks of 64.  K_dkv: thread = key (k, then k + v, in
// registers), queries / dO / lse / delta stream through shared memory; two passes over the queries (dv, then dk) keep
// the live registers below 200.  fp32 arithmetic, deterministic, no atomics.  This path exists for coverage of
// TIMESFORMER.ATTENTION_TYPE joint_space_time (no shipped config uses it); it is not tuned.
constexpr int LONG_BLK = 64;      // streamed rows per shared-memory block
constexpr int LONG_THREADS = 128;

template <typename T>
__global__ void __launch_bounds__(LONG_THREADS)
attn_bwd_dq_long_kernel(const T* __restrict__ qkv, const T* __restrict__ out, const T* __restrict__ dout,
                        const float* __restrict__ lse, T* __restrict__ dqkv, int seq, int H, float scale) {
  __shared__ __align__(16) float sK[LONG_BLK * ROW], sV[LONG_BLK * ROW];
  const long long pair = blockIdx.x;
  const int s_idx = static_cast<int>(pair / H), h = static_cast<int>(pair % H);
  const int C = H * HD;
  const long long pitch = 3LL * C;
  const int i = blockIdx.y * LONG_THREADS + threadIdx.x;
  const bool active = i < seq;
  float q[HD], go[HD], acc[HD];
  float li = 0.f, di = 0.f;
#pragma unroll
  for (int d = 0; d < HD; ++d) q[d] = 0.f, go[d] = 0.f, acc[d] = 0.f;
  if (active) {
    load_row64<T>(qkv + ((long long)s_idx * seq + i) * pitch + h * HD, q);
    load_row64<T>(dout + ((long long)s_idx * seq + i) * C + h * HD, go);
    load_row64<T>(out + ((long long)s_idx * seq + i) * C + h * HD, acc);   // acc temporarily holds O_i
#pragma unroll
    for (int d = 0; d < HD; ++d) di = fmaf(acc[d], go[d], di), acc[d] = 0.f, q[d] *= scale;
    li = lse[pair * seq + i];
  }
  const T* base = qkv + (long long)s_idx * seq * pitch + h * HD;
  for (int k0 = 0; k0 < seq; k0 += LONG_BLK) {
    const int nk = min(LONG_BLK, seq - k0);
    __syncthreads();
    stage_rows<T>(base + (long long)k0 * pitch + C, pitch, nk, sK);
    stage_rows<T>(base + (long long)k0 * pitch + 2 * C, pitch, nk, sV);
    __syncthreads();
    if (active)
      for (int j = 0; j < nk; ++j) {
        const float p = __expf(dot64(q, sK + j * ROW) - li);
        const float dp = dot64(go, sV + j * ROW);
        axpy64(acc, p * (dp - di) * scale, sK + j * ROW);
      }
  }
  if (active) store_row64<T>(dqkv + ((long long)s_idx * seq + i) * pitch + h * HD, acc);
}

template <typename T>
__global__ void __launch_bounds__(LONG_THREADS)
attn_bwd_dkv_long_kernel(const T* __restrict__ qkv, const T* __restrict__ out, const T* __restrict__ dout,
                         const float* __restrict__ lse, T* __restrict__ dqkv, int seq, int H, float scale) {
  __shared__ __align__(16) float sQ[LONG_BLK * ROW], sdO[LONG_BLK * ROW];
  __shared__ float sL[LONG_BLK], sD[LONG_BLK];
  const long long pair = blockIdx.x;
  const int s_idx = static_cast<int>(pair / H), h = static_cast<int>(pair % H);
  const int C = H * HD;
  const long long pitch = 3LL * C;
  const int j = blockIdx.y * LONG_THREADS + threadIdx.x;
  const bool active = j < seq;
  const T* qbase = qkv + (long long)s_idx * seq * pitch + h * HD;
  const T* obase = out + (long long)s_idx * seq * C + h * HD;
  const T* gbase = dout + (long long)s_idx * seq * C + h * HD;
  float k[HD], v[HD], acc[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) k[d] = 0.f, v[d] = 0.f;
  if (active) {
    load_row64<T>(qbase + (long long)j * pitch + C, k);
#pragma unroll
    for (int d = 0; d < HD; ++d) k[d] *= scale;      // s_ij = q_i . (scale k_j)
  }
  for (int pass = 0; pass < 2; ++pass) {             // pass 0: dv_j = sum_i p_ij dO_i ; pass 1: dk_j = scale sum_i ds_ij q_i
    if (pass == 1 && active) load_row64<T>(qbase + (long long)j * pitch + 2 * C, v);
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] = 0.f;
    for (int q0 = 0; q0 < seq; q0 += LONG_BLK) {
      const int nq = min(LONG_BLK, seq - q0);
      __syncthreads();
      stage_rows<T>(qbase + (long long)q0 * pitch, pitch, nq, sQ);
      stage_rows<T>(gbase + (long long)q0 * C, C, nq, sdO);
      if (threadIdx.x < nq) {
        sL[threadIdx.x] = lse[pair * seq + q0 + threadIdx.x];
        if (pass == 1) {                             // delta_i = dO_i . O_i, recomputed per block (O is read from global)
          float orow[HD], grow[HD];
          load_row64<T>(obase + (long long)(q0 + threadIdx.x) * C, orow);
          load_row64<T>(gbase + (long long)(q0 + threadIdx.x) * C, grow);
          float dsum = 0.f;
#pragma unroll
          for (int d = 0; d < HD; ++d) dsum = fmaf(orow[d], grow[d], dsum);
          sD[threadIdx.x] = dsum;
        }
      }
      __syncthreads();
      if (active)
        for (int i = 0; i < nq; ++i) {
          const float p = __expf(dot64(k, sQ + i * ROW) - sL[i]);
          if (pass == 0) {
            axpy64(acc, p, sdO + i * ROW);
          } else {
            const float dp = dot64(v, sdO + i * ROW);
            axpy64(acc, p * (dp - sD[i]) * scale, sQ + i * ROW);
          }
        }
    }
    if (active) store_row64<T>(dqkv + ((long long)s_idx * seq + j) * pitch + (pass == 0 ? 2 * C : C) + h * HD, acc);
  }
}

}  // namespace
}  // namespace pvrl

namespace pvrl {   // attention_small.cu: register-only kernels for seq <= 32 (the temporal axis)
template <typename T>
int attn_small_fwd_launch(const void* qkv, void* out, float* lse, int n_seq, int seq, int H, float scale,
                          cudaStream_t stream);
template <typename T>
int attn_small_bwd_launch(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int n_seq,
                          int seq, int H, float scale, cudaStream_t stream);
// attention_t8.cu: mma.sync kernels for seq <= 8, bf16
int attn_t8_fwd_launch(const void* qkv, void* out, float* lse, int n_seq, int seq, int H, float scale,
                       cudaStream_t stream);
int attn_t8_bwd_launch(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int n_seq,
                       int seq, int H, float scale, cudaStream_t stream);
// attention_t32.cu: mma.sync kernels for 9..32 frames (bf16)
int attn_t32_fwd_launch(const void* qkv, void* out, float* lse, int n_seq, int seq, int H, float scale,
                        cudaStream_t stream);
int attn_t32_bwd_launch(const void* qkv, const void* dout, const float* lse, void* dqkv, int n_seq, int seq, int H,
                        float scale, cudaStream_t stream);
}  // namespace pvrl

using namespace pvrl;

template <typename T>
static int attn_fwd_launch(const void* qkv, void* out, float* lse, int n_seq, int seq, int H, float scale,
                           cudaStream_t stream) {
  int q_per_pair = seq >= 224 ? 256 : ((seq + 31) / 32) * 32;
  if (seq < 32) q_per_pair = seq;  // pack several short sequences per CTA
  int pairs = 1;
  if (seq < 128) pairs = 128 / seq > 0 ? 128 / seq : 1;
  if (pairs * q_per_pair > 256) pairs = 256 / q_per_pair;
  const int kb = seq < KEY_BLOCK ? seq : KEY_BLOCK;
  const size_t smem = 2 * (size_t)pairs * kb * ROW * sizeof(float);
  auto kern = attn_fwd_kernel<T>;
  static bool configured = false;   // per instantiation; not repeated so launches stay CUDA-graph capturable
  if (!configured) {
    PVRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  const long long n_pairs = (long long)n_seq * H;
  dim3 grid(static_cast<unsigned>((n_pairs + pairs - 1) / pairs), (seq + q_per_pair - 1) / q_per_pair);
  kern<<<grid, pairs * q_per_pair, smem, stream>>>(static_cast<const T*>(qkv), static_cast<T*>(out), lse, n_seq, seq, H,
                                                   scale, pairs, q_per_pair);
  return launched("attn_fwd_kernel");
}

template <typename T>
static int attn_bwd_launch(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int n_seq,
                           int seq, int H, float scale, cudaStream_t stream) {
  int pairs = 128 / seq > 0 ? 128 / seq : 1;
  const int threads = ((pairs * seq + 31) / 32) * 32;
  const size_t smem = (4 * (size_t)pairs * seq * ROW + 2 * (size_t)pairs * seq) * sizeof(float);
  auto kern = attn_bwd_kernel<T>;
  static bool configured = false;
  if (!configured) {
    PVRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const long long n_pairs = (long long)n_seq * H;
  kern<<<static_cast<unsigned>((n_pairs + pairs - 1) / pairs), threads, smem, stream>>>(
      static_cast<const T*>(qkv), static_cast<const T*>(out), static_cast<const T*>(dout), lse, static_cast<T*>(dqkv),
      n_seq, seq, H, scale, pairs);
  return launched("attn_bwd_kernel");
}

extern "C" int pvrl_attn_fwd(const void* qkv, void* out, float* lse, int32_t dtype, int32_t n_seq, int32_t seq,
                             int32_t H, float scale, void* stream) {
  PVRL_CHECK_ARG(qkv && out && n_seq > 0 && seq > 0 && H > 0, "pvrl_attn_fwd: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (seq <= 8 && dtype == PVRL_BF16) return attn_t8_fwd_launch(qkv, out, lse, n_seq, seq, H, scale, st);
  if (seq <= 32 && dtype == PVRL_BF16) return attn_t32_fwd_launch(qkv, out, lse, n_seq, seq, H, scale, st);
  if (seq <= 32)
    return dtype == PVRL_F32 ? attn_small_fwd_launch<float>(qkv, out, lse, n_seq, seq, H, scale, st)
                             : attn_small_fwd_launch<__nv_bfloat16>(qkv, out, lse, n_seq, seq, H, scale, st);
  return dtype == PVRL_F32 ? attn_fwd_launch<float>(qkv, out, lse, n_seq, seq, H, scale, st)
                           : attn_fwd_launch<__nv_bfloat16>(qkv, out, lse, n_seq, seq, H, scale, st);
}

template <typename T>
static int attn_bwd_long_launch(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                                int n_seq, int seq, int H, float scale, cudaStream_t stream) {
  dim3 grid(static_cast<unsigned>(n_seq) * H, (seq + LONG_THREADS - 1) / LONG_THREADS);
  attn_bwd_dq_long_kernel<T><<<grid, LONG_THREADS, 0, stream>>>(static_cast<const T*>(qkv), static_cast<const T*>(out),
                                                               static_cast<const T*>(dout), lse, static_cast<T*>(dqkv),
                                                               seq, H, scale);
  int rc = launched("attn_bwd_dq_long_kernel");
  if (rc) return rc;
  attn_bwd_dkv_long_kernel<T><<<grid, LONG_THREADS, 0, stream>>>(static_cast<const T*>(qkv), static_cast<const T*>(out),
                                                                static_cast<const T*>(dout), lse, static_cast<T*>(dqkv),
                                                                seq, H, scale);
  return launched("attn_bwd_dkv_long_kernel");
}

extern "C" int pvrl_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                             int32_t dtype, int32_t n_seq, int32_t seq, int32_t H, float scale, void* stream) {
  PVRL_CHECK_ARG(qkv && out && dout && lse && dqkv && n_seq > 0 && seq > 0 && H > 0, "pvrl_attn_bwd: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (seq > 208)
    return dtype == PVRL_F32 ? attn_bwd_long_launch<float>(qkv, out, dout, lse, dqkv, n_seq, seq, H, scale, st)
                             : attn_bwd_long_launch<__nv_bfloat16>(qkv, out, dout, lse, dqkv, n_seq, seq, H, scale, st);
  if (seq <= 8 && dtype == PVRL_BF16) return attn_t8_bwd_launch(qkv, out, dout, lse, dqkv, n_seq, seq, H, scale, st);
  if (seq <= 32 && dtype == PVRL_BF16) return attn_t32_bwd_launch(qkv, dout, lse, dqkv, n_seq, seq, H, scale, st);
  if (seq <= 32)
    return dtype == PVRL_F32
               ? attn_small_bwd_launch<float>(qkv, out, dout, lse, dqkv, n_seq, seq, H, scale, st)
               : attn_small_bwd_launch<__nv_bfloat16>(qkv, out, dout, lse, dqkv, n_seq, seq, H, scale, st);
  return dtype == PVRL_F32 ? attn_bwd_launch<float>(qkv, out, dout, lse, dqkv, n_seq, seq, H, scale, st)
                           : attn_bwd_launch<__nv_bfloat16>(qkv, out, dout, lse, dqkv, n_seq, seq, H, scale, st);
}
