// The video<->step-text matching head and the pre-training loss in fp32 (latency-sized problems:
// M = 18..26 rows, 512-d embeddings, K = 778 / 9871 step phrases).  vit.py:300-307, train_net.py:153-162.
#include "pvrl_host.h"
#include "pvrl_ptx.cuh"

namespace pvrl {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide reductions over 256 threads (8 warps)
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) s += red[i];
  return s;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = -INFINITY;
  for (int i = 0; i < (blockDim.x >> 5); ++i) s = fmaxf(s, red[i]);
  return s;
}

// ---- small linear: one warp per output feature j, loops over the M rows --------------------------------
__global__ void linear_small_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                        const float* __restrict__ b, float* __restrict__ y, int M, int K, int N) {
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (j >= N) return;
  for (int m = 0; m < M; ++m) {
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s = fmaf(x[(long long)m * K + k], w[(long long)j * K + k], s);
    s = warp_sum(s);
    if (lane == 0) y[(long long)m * N + j] = s + (b != nullptr ? b[j] : 0.f);
  }
}
__global__ void linear_small_dw_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                       float* __restrict__ dw, float* __restrict__ db, int M, int K, int N) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)N * K) return;
  const int j = static_cast<int>(i / K), k = static_cast<int>(i - (long long)j * K);
  float s = 0.f, sb = 0.f;
  for (int m = 0; m < M; ++m) {
    const float d = dy[(long long)m * N + j];
    s = fmaf(d, x[(long long)m * K + k], s);
    sb += d;
  }
  dw[i] += s;
  if (k == 0 && db != nullptr) db[j] += sb;
}
__global__ void linear_small_dx_kernel(const float* __restrict__ w, const float* __restrict__ dy,
                                       float* __restrict__ dx, int M, int K, int N) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)M * K) return;
  const int m = static_cast<int>(i / K), k = static_cast<int>(i - (long long)m * K);
  float s = 0.f;
  for (int j = 0; j < N; ++j) s = fmaf(dy[(long long)m * N + j], w[(long long)j * K + k], s);
  dx[i] = s;
}

// ---- L2 normalisation: one warp per row -----------------------------------------------------------------
__global__ void l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ norms, int M,
                                  int C) {
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= M) return;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float v = x[(long long)m * C + c];
    s = fmaf(v, v, s);
  }
  const float nrm = sqrtf(warp_sum(s));
  for (int c = lane; c < C; c += 32) y[(long long)m * C + c] = x[(long long)m * C + c] / nrm;
  if (lane == 0 && norms != nullptr) norms[m] = nrm;
}
__global__ void l2norm_bwd_kernel(const float* __restrict__ y, const float* __restrict__ norms,
                                  const float* __restrict__ dy, float* __restrict__ dx, int M, int C) {
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= M) return;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s = fmaf(y[(long long)m * C + c], dy[(long long)m * C + c], s);
  s = warp_sum(s);
  const float inv = 1.0f / norms[m];
  for (int c = lane; c < C; c += 32)
    dx[(long long)m * C + c] = (dy[(long long)m * C + c] - y[(long long)m * C + c] * s) * inv;
}

// ---- similarity logits: one warp per class (label row in registers), loops over the M embeddings -------
constexpr int SIM_MAX_VEC = 8;  // K <= 1024
__global__ void sim_logits_fwd_kernel(const float* __restrict__ emb, const float* __restrict__ label,
                                      float* __restrict__ logits, int M, int C, int K, float inv_temp) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  const int nvec = K >> 7;
  float4 lab[SIM_MAX_VEC];
#pragma unroll
  for (int i = 0; i < SIM_MAX_VEC; ++i)
    if (i < nvec) lab[i] = __ldg(reinterpret_cast<const float4*>(label + (long long)c * K) + i * 32 + lane);
  for (int m = 0; m < M; ++m) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < SIM_MAX_VEC; ++i)
      if (i < nvec) {
        const float4 e = __ldg(reinterpret_cast<const float4*>(emb + (long long)m * K) + i * 32 + lane);
        s += e.x * lab[i].x + e.y * lab[i].y + e.z * lab[i].z + e.w * lab[i].w;
      }
    s = warp_sum(s);
    if (lane == 0) logits[(long long)m * C + c] = s * inv_temp;
  }
}
// demb[m,k] += inv_temp * sum_c dlogits[m,c] label[c,k];  block = class chunk, thread = k, 32 rows per pass
constexpr int SIMB_CH = 64, SIMB_MT = 32;
__global__ void sim_logits_bwd_kernel(const float* __restrict__ dlogits, const float* __restrict__ label,
                                      float* __restrict__ demb, int M, int C, int K, float inv_temp) {
  __shared__ float sdl[SIMB_MT][SIMB_CH + 1];
  const int c0 = blockIdx.x * SIMB_CH;
  const int nc = min(SIMB_CH, C - c0);
  for (int m0 = 0; m0 < M; m0 += SIMB_MT) {
    const int nm = min(SIMB_MT, M - m0);
    __syncthreads();
    for (int i = threadIdx.x; i < SIMB_MT * SIMB_CH; i += blockDim.x) {
      const int mm = i / SIMB_CH, cc = i % SIMB_CH;
      sdl[mm][cc] = (mm < nm && cc < nc) ? dlogits[(long long)(m0 + mm) * C + c0 + cc] : 0.f;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      float acc[SIMB_MT];
#pragma unroll
      for (int mm = 0; mm < SIMB_MT; ++mm) acc[mm] = 0.f;
      for (int cc = 0; cc < nc; cc += 8) {   // 8 bank rows in flight per thread: the loop is bound by L2 latency, not FMAs
        float l[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) l[u] = cc + u < nc ? __ldg(label + (long long)(c0 + cc + u) * K + k) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int cu = min(cc + u, SIMB_CH - 1);
#pragma unroll
          for (int mm = 0; mm < SIMB_MT; ++mm) acc[mm] = fmaf(sdl[mm][cu], l[u], acc[mm]);
        }
      }
#pragma unroll
      for (int mm = 0; mm < SIMB_MT; ++mm)
        if (mm < nm) atomicAdd(demb + (long long)(m0 + mm) * K + k, acc[mm] * inv_temp);
    }
  }
}

// ---- KL(top-k teacher || softmax(pred)) per row: one 256-thread block per row --------------------------
constexpr int TOPK_MAX = 8;
__global__ void __launch_bounds__(256)
kl_topk_loss_kernel(const float* __restrict__ pred, const float* __restrict__ teacher, float* __restrict__ row_loss,
                    float* __restrict__ dpred, float* __restrict__ teacher_out, int M, int K, int topk, float gscale) {
  extern __shared__ float sm[];
  float* sp = sm;          // pred row
  float* st = sm + K;      // teacher row -> teacher probabilities
  __shared__ float red[8];
  __shared__ float cand[256 * TOPK_MAX];
  __shared__ float tv[TOPK_MAX];
  const int m = blockIdx.x;
  const float* pr = pred + (long long)m * K;
  const float* tr = teacher + (long long)m * K;
  float pmax = -INFINITY, tmax = -INFINITY;
  for (int c = threadIdx.x; c < K; c += blockDim.x) {
    const float a = pr[c], b = tr[c];
    sp[c] = a, st[c] = b;
    pmax = fmaxf(pmax, a), tmax = fmaxf(tmax, b);
  }
  pmax = block_max(pmax, red);
  tmax = block_max(tmax, red);
  float psum = 0.f, tsum = 0.f;
  for (int c = threadIdx.x; c < K; c += blockDim.x) {
    psum += expf(sp[c] - pmax);
    const float e = expf(st[c] - tmax);
    st[c] = e;
    tsum += e;
  }
  psum = block_sum(psum, red);
  tsum = block_sum(tsum, red);
  const float plse = pmax + logf(psum);
  const float tinv = 1.0f / tsum;
  // teacher softmax + per-thread top-k values (train_net.py:153-155)
  float best[TOPK_MAX];
#pragma unroll
  for (int i = 0; i < TOPK_MAX; ++i) best[i] = -1.f;
  for (int c = threadIdx.x; c < K; c += blockDim.x) {
    const float t = st[c] * tinv;
    st[c] = t;
    if (topk > 0 && t > best[topk - 1]) {
      float v = t;
#pragma unroll
      for (int i = 0; i < TOPK_MAX; ++i)
        if (i < topk && v > best[i]) {
          const float tmp = best[i];
          best[i] = v;
          v = tmp;
        }
    }
  }
  float wsum = 1.0f;
  if (topk > 0) {
#pragma unroll
    for (int i = 0; i < TOPK_MAX; ++i)
      if (i < topk) cand[threadIdx.x * TOPK_MAX + i] = best[i];
    __syncthreads();
    if (threadIdx.x < 32) {  // warp 0 extracts the k largest candidates, one per round
      for (int r = 0; r < topk; ++r) {
        float bv = -2.f;
        int bi = -1;
        for (int i = threadIdx.x; i < 256 * TOPK_MAX; i += 32) {
          if ((i % TOPK_MAX) >= topk) continue;
          const float v = cand[i];
          if (v > bv) bv = v, bi = i;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ov > bv || (ov == bv && oi >= 0 && (bi < 0 || oi < bi))) bv = ov, bi = oi;
        }
        if (threadIdx.x == 0) {
          tv[r] = bv;
          cand[bi] = -3.f;
        }
        __syncwarp();
      }
    }
    __syncthreads();
    // keep entries equal to a top-k value, once per matching slot (train_net.py:156-157), renormalise (:158)
    float ws = 0.f;
    for (int c = threadIdx.x; c < K; c += blockDim.x) {
      const float t = st[c];
      int cnt = 0;
#pragma unroll
      for (int i = 0; i < TOPK_MAX; ++i)
        if (i < topk && t == tv[i]) ++cnt;
      const float v = t * cnt;
      st[c] = v;
      ws += v;
    }
    wsum = block_sum(ws, red);
  }
  const float winv = 1.0f / wsum;
  const float gs = gscale / M;
  float loss = 0.f;
  for (int c = threadIdx.x; c < K; c += blockDim.x) {
    const float t = st[c] * winv;
    const float logp = sp[c] - plse;
    if (t > 0.f) loss += t * (logf(t) - logp);  // KLDivLoss pointwise, xlogy semantics at t == 0
    if (dpred != nullptr) dpred[(long long)m * K + c] = (expf(logp) - t) * gs;
    if (teacher_out != nullptr) teacher_out[(long long)m * K + c] = t;
  }
  loss = block_sum(loss, red);
  if (threadIdx.x == 0 && row_loss != nullptr) row_loss[m] = loss / M;
}

__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ x, float* __restrict__ y, int K) {
  __shared__ float red[8];
  const float* xr = x + (long long)blockIdx.x * K;
  float* yr = y + (long long)blockIdx.x * K;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < K; c += blockDim.x) mx = fmaxf(mx, xr[c]);
  mx = block_max(mx, red);
  float s = 0.f;
  for (int c = threadIdx.x; c < K; c += blockDim.x) s += expf(xr[c] - mx);
  s = block_sum(s, red);
  const float inv = 1.0f / s;
  for (int c = threadIdx.x; c < K; c += blockDim.x) yr[c] = expf(xr[c] - mx) * inv;
}

}  // namespace
}  // namespace pvrl

using namespace pvrl;
#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int pvrl_linear_small_fwd(const float* x, const float* w, const float* b, float* y, int32_t M, int32_t K,
                                     int32_t N, void* stream) {
  PVRL_CHECK_ARG(x && w && y && M > 0 && K > 0 && N > 0, "pvrl_linear_small_fwd: bad arguments");
  linear_small_fwd_kernel<<<(N + 7) / 8, 256, 0, STREAM>>>(x, w, b, y, M, K, N);
  return launched("linear_small_fwd_kernel");
}

extern "C" int pvrl_linear_small_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* db,
                                     int32_t M, int32_t K, int32_t N, void* stream) {
  PVRL_CHECK_ARG(x && w && dy && M > 0 && K > 0 && N > 0, "pvrl_linear_small_bwd: bad arguments");
  if (dw != nullptr) {
    linear_small_dw_kernel<<<static_cast<unsigned>(((long long)N * K + 255) / 256), 256, 0, STREAM>>>(x, dy, dw, db, M,
                                                                                                       K, N);
    int rc = launched("linear_small_dw_kernel");
    if (rc) return rc;
  }
  if (dx != nullptr) {
    linear_small_dx_kernel<<<static_cast<unsigned>(((long long)M * K + 255) / 256), 256, 0, STREAM>>>(w, dy, dx, M, K,
                                                                                                       N);
    return launched("linear_small_dx_kernel");
  }
  return 0;
}

extern "C" int pvrl_l2norm_fwd(const float* x, float* y, float* norms, int32_t M, int32_t C, void* stream) {
  PVRL_CHECK_ARG(x && y && M > 0 && C > 0, "pvrl_l2norm_fwd: bad arguments");
  l2norm_fwd_kernel<<<(M + 7) / 8, 256, 0, STREAM>>>(x, y, norms, M, C);
  return launched("l2norm_fwd_kernel");
}

extern "C" int pvrl_l2norm_bwd(const float* y, const float* norms, const float* dy, float* dx, int32_t M, int32_t C,
                               void* stream) {
  PVRL_CHECK_ARG(y && norms && dy && dx && M > 0 && C > 0, "pvrl_l2norm_bwd: bad arguments");
  l2norm_bwd_kernel<<<(M + 7) / 8, 256, 0, STREAM>>>(y, norms, dy, dx, M, C);
  return launched("l2norm_bwd_kernel");
}

extern "C" int pvrl_sim_logits_fwd(const float* emb, const float* label, float* logits, int32_t M, int32_t C,
                                   int32_t K, float inv_temp, void* stream) {
  PVRL_CHECK_ARG(emb && label && logits && M > 0 && C > 0, "pvrl_sim_logits_fwd: bad arguments");
  PVRL_CHECK_ARG(K % 128 == 0 && K <= 128 * SIM_MAX_VEC, "pvrl_sim_logits_fwd: K=%d must be a multiple of 128, <= 1024", K);
  sim_logits_fwd_kernel<<<(C + 7) / 8, 256, 0, STREAM>>>(emb, label, logits, M, C, K, inv_temp);
  return launched("sim_logits_fwd_kernel");
}

extern "C" int pvrl_sim_logits_bwd(const float* dlogits, const float* label, float* demb, int32_t M, int32_t C,
                                   int32_t K, float inv_temp, void* stream) {
  PVRL_CHECK_ARG(dlogits && label && demb && M > 0 && C > 0 && K > 0, "pvrl_sim_logits_bwd: bad arguments");
  sim_logits_bwd_kernel<<<(C + SIMB_CH - 1) / SIMB_CH, 256, 0, STREAM>>>(dlogits, label, demb, M, C, K, inv_temp);
  return launched("sim_logits_bwd_kernel");
}

extern "C" int pvrl_kl_topk_loss(const float* pred, const float* teacher_logits, float* row_loss, float* dpred,
                                 float* teacher_out, int32_t M, int32_t K, int32_t topk, float gscale, void* stream) {
  PVRL_CHECK_ARG(pred && teacher_logits && M > 0 && K > 0, "pvrl_kl_topk_loss: bad arguments");
  PVRL_CHECK_ARG(topk >= 0 && topk <= TOPK_MAX, "pvrl_kl_topk_loss: topk=%d out of [0, %d]", topk, TOPK_MAX);
  const size_t smem = 2 * (size_t)K * sizeof(float);
  PVRL_CHECK_ARG(smem <= 200 * 1024, "pvrl_kl_topk_loss: K=%d too large for the single-block-per-row kernel", K);
  static bool configured = false;
  if (!configured) {
    PVRL_CUDA(cudaFuncSetAttribute(kl_topk_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  kl_topk_loss_kernel<<<M, 256, smem, STREAM>>>(pred, teacher_logits, row_loss, dpred, teacher_out, M, K, topk, gscale);
  return launched("kl_topk_loss_kernel");
}

extern "C" int pvrl_softmax_rows(const float* x, float* y, int32_t M, int32_t K, void* stream) {
  PVRL_CHECK_ARG(x && y && M > 0 && K > 0, "pvrl_softmax_rows: bad arguments");
  softmax_rows_kernel<<<M, 256, 0, STREAM>>>(x, y, K);
  return launched("softmax_rows_kernel");
}
