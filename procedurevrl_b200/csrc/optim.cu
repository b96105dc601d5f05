// Optimizer step over the FLAT parameter / gradient buffers: torch.optim.AdamW / Adam / SGD as built by
// lib/models/optimizer.py:90-114 and driven by tools/train_net.py:176-192 (optimizer.step(); optimizer.zero_grad()).
// One launch per parameter group: every element is read once (p, g, m, v = 16 B) and written once (p, m, v, and g = 0 for
// the next backward's accumulating dW kernels = 16 B) -- pure HBM streaming, 128-bit accesses, 4 independent 16-byte
// loads per array in flight per thread.  Learning rate and step count live in DEVICE memory so that a CUDA graph of the
// whole training step can be replayed while the host changes the rate between replays (lr_policy.py schedules).
#include "pvrl_host.h"

namespace pvrl {
namespace {

constexpr int THREADS = 256;
constexpr int UNROLL = 4;

__global__ void optim_tick_kernel(float* step) { *step += 1.f; }

struct AdamCoef {
  float lr_wd;       // lr * weight_decay
  float step_size;   // lr / (1 - beta1^t)
  float inv_sqrt_bc2;
};

// omb1 / omb2 = 1 - beta, rounded from the double difference (what torch's kernels use), not from 1.f - float(beta)
__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, const AdamCoef& c, float omb1,
                                          float omb2, float eps, float wd, bool decoupled) {
  if (decoupled)
    p -= c.lr_wd * p;           // AdamW: p *= 1 - lr * wd
  else
    g = fmaf(wd, p, g);         // Adam: L2 term joins the gradient
  m = fmaf(omb1, g - m, m);
  v = fmaf(omb2, g * g - v, v);
  const float denom = sqrtf(v) * c.inv_sqrt_bc2 + eps;
  p -= c.step_size * (m / denom);
}

template <bool ZERO>
__global__ void __launch_bounds__(THREADS)
adam_flat_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 long long n, const float* __restrict__ lr_dev, const float* __restrict__ step_dev, float lr_mult,
                 double beta1d, double beta2d, float eps, float wd, int decoupled, float grad_scale) {
  __shared__ AdamCoef sc;
  if (threadIdx.x == 0) {
    const double t = static_cast<double>(*step_dev);
    const float lr = *lr_dev * lr_mult;
    sc.lr_wd = lr * wd;
    sc.step_size = static_cast<float>(static_cast<double>(lr) / (1.0 - pow(beta1d, t)));
    sc.inv_sqrt_bc2 = static_cast<float>(1.0 / sqrt(1.0 - pow(beta2d, t)));
  }
  __syncthreads();
  const AdamCoef c = sc;
  const float omb1 = static_cast<float>(1.0 - beta1d), omb2 = static_cast<float>(1.0 - beta2d);
  const bool dec = decoupled != 0;
  // scalar head up to the first 16-byte boundary (all four arrays share their misalignment: checked by the host)
  const long long head = min(n, static_cast<long long>(((16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15) >> 2));
  const long long nvec = (n - head) >> 2;
  const long long tail0 = head + (nvec << 2);
  if (blockIdx.x == 0) {
    for (long long i = threadIdx.x; i < head + (n - tail0); i += THREADS) {
      const long long j = i < head ? i : tail0 + (i - head);
      float pp = p[j], mm = m[j], vv = v[j];
      adam_elem(pp, g[j] * grad_scale, mm, vv, c, omb1, omb2, eps, wd, dec);
      p[j] = pp, m[j] = mm, v[j] = vv;
      if (ZERO) g[j] = 0.f;
    }
  }
  float4* p4 = reinterpret_cast<float4*>(p + head);
  float4* g4 = reinterpret_cast<float4*>(g + head);
  float4* m4 = reinterpret_cast<float4*>(m + head);
  float4* v4 = reinterpret_cast<float4*>(v + head);
  const long long stride = static_cast<long long>(gridDim.x) * THREADS;
  for (long long i0 = static_cast<long long>(blockIdx.x) * THREADS + threadIdx.x; i0 < nvec; i0 += stride * UNROLL) {
    float4 pp[UNROLL], gg[UNROLL], mm[UNROLL], vv[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long i = i0 + u * stride;
      if (i < nvec) pp[u] = p4[i], gg[u] = g4[i], mm[u] = m4[i], vv[u] = v4[i];
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long i = i0 + u * stride;
      if (i < nvec) {
        adam_elem(pp[u].x, gg[u].x * grad_scale, mm[u].x, vv[u].x, c, omb1, omb2, eps, wd, dec);
        adam_elem(pp[u].y, gg[u].y * grad_scale, mm[u].y, vv[u].y, c, omb1, omb2, eps, wd, dec);
        adam_elem(pp[u].z, gg[u].z * grad_scale, mm[u].z, vv[u].z, c, omb1, omb2, eps, wd, dec);
        adam_elem(pp[u].w, gg[u].w * grad_scale, mm[u].w, vv[u].w, c, omb1, omb2, eps, wd, dec);
        p4[i] = pp[u], m4[i] = mm[u], v4[i] = vv[u];
        if (ZERO) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
}

__device__ __forceinline__ void sgd_elem(float& p, float g, float& buf, float lr, float momentum, float dampening,
                                         bool nesterov, float wd, bool first) {
  g = fmaf(wd, p, g);
  if (momentum != 0.f) {
    buf = first ? g : fmaf(momentum, buf, (1.f - dampening) * g);
    g = nesterov ? fmaf(momentum, buf, g) : buf;
  }
  p -= lr * g;
}

template <bool ZERO>
__global__ void __launch_bounds__(THREADS)
sgd_flat_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ buf, long long n,
                const float* __restrict__ lr_dev, const float* __restrict__ step_dev, float lr_mult, float momentum,
                float dampening, int nesterov, float wd, float grad_scale) {
  const float lr = __ldg(lr_dev) * lr_mult;
  const bool first = __ldg(step_dev) <= 1.f;      // torch.optim.SGD: the momentum buffer starts as a copy of the gradient
  const bool nes = nesterov != 0;
  const long long head = min(n, static_cast<long long>(((16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15) >> 2));
  const long long nvec = (n - head) >> 2;
  const long long tail0 = head + (nvec << 2);
  if (blockIdx.x == 0) {
    for (long long i = threadIdx.x; i < head + (n - tail0); i += THREADS) {
      const long long j = i < head ? i : tail0 + (i - head);
      float pp = p[j], bb = buf[j];
      sgd_elem(pp, g[j] * grad_scale, bb, lr, momentum, dampening, nes, wd, first);
      p[j] = pp, buf[j] = bb;
      if (ZERO) g[j] = 0.f;
    }
  }
  float4* p4 = reinterpret_cast<float4*>(p + head);
  float4* g4 = reinterpret_cast<float4*>(g + head);
  float4* b4 = reinterpret_cast<float4*>(buf + head);
  const long long stride = static_cast<long long>(gridDim.x) * THREADS;
  for (long long i0 = static_cast<long long>(blockIdx.x) * THREADS + threadIdx.x; i0 < nvec; i0 += stride * UNROLL) {
    float4 pp[UNROLL], gg[UNROLL], bb[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long i = i0 + u * stride;
      if (i < nvec) pp[u] = p4[i], gg[u] = g4[i], bb[u] = b4[i];
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long i = i0 + u * stride;
      if (i < nvec) {
        sgd_elem(pp[u].x, gg[u].x * grad_scale, bb[u].x, lr, momentum, dampening, nes, wd, first);
        sgd_elem(pp[u].y, gg[u].y * grad_scale, bb[u].y, lr, momentum, dampening, nes, wd, first);
        sgd_elem(pp[u].z, gg[u].z * grad_scale, bb[u].z, lr, momentum, dampening, nes, wd, first);
        sgd_elem(pp[u].w, gg[u].w * grad_scale, bb[u].w, lr, momentum, dampening, nes, wd, first);
        p4[i] = pp[u], b4[i] = bb[u];
        if (ZERO) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
}

inline bool same_misalignment(const void* a, const void* b) {
  return ((reinterpret_cast<uintptr_t>(a) ^ reinterpret_cast<uintptr_t>(b)) & 15) == 0;
}

inline int stream_grid(long long n) {
  const long long per_block = static_cast<long long>(THREADS) * 4 * UNROLL;
  const long long want = (n + per_block - 1) / per_block;
  const long long cap = static_cast<long long>(num_sms()) * 2;     // ~90 registers per thread: 2 resident CTAs per SM
  return static_cast<int>(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace
}  // namespace pvrl

using namespace pvrl;
#define STREAM static_cast<cudaStream_t>(stream)

extern "C" int pvrl_optim_tick(float* step_dev, void* stream) {
  PVRL_CHECK_ARG(step_dev, "pvrl_optim_tick: null step counter");
  optim_tick_kernel<<<1, 1, 0, STREAM>>>(step_dev);
  return launched("pvrl_optim_tick");
}

extern "C" int pvrl_adam_flat(float* p, float* g, float* m, float* v, int64_t n, const float* lr_dev,
                              const float* step_dev, float lr_mult, double beta1, double beta2, float eps,
                              float weight_decay, int32_t decoupled, float grad_scale, int32_t zero_grad, void* stream) {
  PVRL_CHECK_ARG(p && g && m && v && lr_dev && step_dev && n > 0, "pvrl_adam_flat: bad arguments");
  PVRL_CHECK_ARG((reinterpret_cast<uintptr_t>(p) & 3) == 0 && same_misalignment(p, g) && same_misalignment(p, m) &&
                     same_misalignment(p, v),
                 "pvrl_adam_flat: p, g, m, v must share their offset from a 16-byte boundary");
  const int grid = stream_grid(n);
  if (zero_grad)
    adam_flat_kernel<true><<<grid, THREADS, 0, STREAM>>>(p, g, m, v, n, lr_dev, step_dev, lr_mult, beta1, beta2, eps,
                                                         weight_decay, decoupled, grad_scale);
  else
    adam_flat_kernel<false><<<grid, THREADS, 0, STREAM>>>(p, g, m, v, n, lr_dev, step_dev, lr_mult, beta1, beta2, eps,
                                                          weight_decay, decoupled, grad_scale);
  return launched("pvrl_adam_flat");
}

// dst[i] = (dst[i] + sum_s src[s * stride + i]) * scale: the local reduction of the copy-engine gradient exchange
// (procedurevrl_b200/grad_exchange.py): `n_src` peers' copies of this rank's gradient chunk sit in a staging buffer.
template <bool VEC>
__global__ void __launch_bounds__(THREADS)
reduce_chunks_kernel(float* __restrict__ dst, const float* __restrict__ src, int n_src, long long stride, long long n,
                     float scale) {
  const long long nvec = VEC ? n >> 2 : 0;
  const long long step = static_cast<long long>(gridDim.x) * THREADS;
  for (long long i = static_cast<long long>(blockIdx.x) * THREADS + threadIdx.x; i < nvec; i += step) {
    float4 a = reinterpret_cast<const float4*>(dst)[i];
    for (int s = 0; s < n_src; ++s) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(src + s * stride) + i);
      a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
    }
    a.x *= scale, a.y *= scale, a.z *= scale, a.w *= scale;
    reinterpret_cast<float4*>(dst)[i] = a;
  }
  // tail of the vector path (block 0), or everything when the range does not start on a 16-byte boundary (grid-strided)
  const long long i0 = VEC ? (nvec << 2) + threadIdx.x : static_cast<long long>(blockIdx.x) * THREADS + threadIdx.x;
  if (VEC && blockIdx.x != 0) return;
  for (long long i = i0; i < n; i += VEC ? THREADS : step) {
    float a = dst[i];
    for (int s = 0; s < n_src; ++s) a += src[s * stride + i];
    dst[i] = a * scale;
  }
}

extern "C" int pvrl_reduce_chunks(float* dst, const float* src, int32_t n_src, int64_t stride, int64_t n, float scale,
                                  void* stream) {
  PVRL_CHECK_ARG(dst && (src || n_src == 0) && n > 0 && n_src >= 0, "pvrl_reduce_chunks: bad arguments");
  PVRL_CHECK_ARG((reinterpret_cast<uintptr_t>(dst) & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 3) == 0,
                 "pvrl_reduce_chunks: dst, src must be fp32-aligned");
  const bool vec = (reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && stride % 4 == 0;
  if (vec)
    reduce_chunks_kernel<true><<<stream_grid(n), THREADS, 0, STREAM>>>(dst, src, n_src, stride, n, scale);
  else   // a range that does not start on a 16-byte boundary (never the case for the model's parameter layout)
    reduce_chunks_kernel<false><<<stream_grid(n), THREADS, 0, STREAM>>>(dst, src, n_src, stride, n, scale);
  return launched("pvrl_reduce_chunks");
}

extern "C" int pvrl_sgd_flat(float* p, float* g, float* buf, int64_t n, const float* lr_dev, const float* step_dev,
                             float lr_mult, float momentum, float dampening, int32_t nesterov, float weight_decay,
                             float grad_scale, int32_t zero_grad, void* stream) {
  PVRL_CHECK_ARG(p && g && buf && lr_dev && step_dev && n > 0, "pvrl_sgd_flat: bad arguments");
  PVRL_CHECK_ARG((reinterpret_cast<uintptr_t>(p) & 3) == 0 && same_misalignment(p, g) && same_misalignment(p, buf),
                 "pvrl_sgd_flat: p, g, buf must share their offset from a 16-byte boundary");
  PVRL_CHECK_ARG(!(nesterov && (momentum <= 0.f || dampening != 0.f)),
                 "pvrl_sgd_flat: Nesterov momentum requires a momentum and zero dampening");
  const int grid = stream_grid(n);
  if (zero_grad)
    sgd_flat_kernel<true><<<grid, THREADS, 0, STREAM>>>(p, g, buf, n, lr_dev, step_dev, lr_mult, momentum, dampening,
                                                        nesterov, weight_decay, grad_scale);
  else
    sgd_flat_kernel<false><<<grid, THREADS, 0, STREAM>>>(p, g, buf, n, lr_dev, step_dev, lr_mult, momentum, dampening,
                                                         nesterov, weight_decay, grad_scale);
  return launched("pvrl_sgd_flat");
}
