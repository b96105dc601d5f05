// Temporal attention: sequences of T = 2..32 frames per spatial token, head_dim 64 (vit.py:130-133 -> :84-88).
// 3528 x 12 tiny (T x T) problems per 18 clips: the op is purely HBM-bound (read qkv once, write o once).
// Mapping: an octet of 8 lanes owns one (sequence, head, row); lane c holds dims [8c, 8c+8) of every vector, so
// each row access is one coalesced 128-byte (bf16) request per octet, dot products finish with three xor-shuffles,
// nothing is staged in shared memory and register use stays ~50/thread (full occupancy to cover HBM latency).
// The backward exchanges delta / lse of a sequence's rows through 1 KB of shared memory.
#include "pvrl_host.h"
#include "pvrl_ptx.cuh"

namespace pvrl {
namespace {

constexpr int HD = 64;
constexpr int THREADS = 256;            // 32 rows per block
constexpr int ROWS = THREADS / 8;

template <typename T>
__device__ __forceinline__ void ld8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void ld8<float>(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
}
template <>
__device__ __forceinline__ void ld8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x, v[1] = a.y, v[2] = b.x, v[3] = b.y, v[4] = c.x, v[5] = c.y, v[6] = d.x, v[7] = d.y;
}
template <typename T>
__device__ __forceinline__ void st8(T* p, const float (&v)[8]);
template <>
__device__ __forceinline__ void st8<float>(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void st8<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]), u.y = pack_bf16x2(v[2], v[3]), u.z = pack_bf16x2(v[4], v[5]), u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ float dot8(const float (&a)[8], const float (&b)[8]) {
  float s = a[0] * b[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) s = fmaf(a[i], b[i], s);
  return s;
}
__device__ __forceinline__ float octet_sum(float v) {   // all 8 lanes of the octet get the total
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

template <typename T, int TMAX>
__global__ void __launch_bounds__(THREADS)
attn_small_fwd_kernel(const T* __restrict__ qkv, T* __restrict__ out, float* __restrict__ lse, long long n_rows,
                      int seq, int H, float scale) {
  const long long row = blockIdx.x * (long long)ROWS + (threadIdx.x >> 3);    // (pair, i) flattened
  const int c = (threadIdx.x & 7) * 8;
  const bool active = row < n_rows;
  const long long pair = active ? row / seq : 0;
  const int i = active ? static_cast<int>(row - pair * seq) : 0;
  const int s_idx = static_cast<int>(pair / H), h = static_cast<int>(pair % H);
  const int C = H * HD;
  const long long pitch = 3LL * C;
  const T* base = qkv + (long long)s_idx * seq * pitch + h * HD + c;
  float q[8], kv[8], s[TMAX];
  ld8<T>(base + i * pitch, q);
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < TMAX; ++j)
    if (j < seq) {
      ld8<T>(base + j * pitch + C, kv);
      s[j] = octet_sum(dot8(q, kv)) * scale;
      mx = fmaxf(mx, s[j]);
    }
  float l = 0.f;
#pragma unroll
  for (int j = 0; j < TMAX; ++j)
    if (j < seq) {
      s[j] = __expf(s[j] - mx);
      l += s[j];
    }
  const float inv = 1.0f / l;
  float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < TMAX; ++j)
    if (j < seq) {
      ld8<T>(base + j * pitch + 2 * C, kv);
      const float w = s[j] * inv;
#pragma unroll
      for (int d = 0; d < 8; ++d) o[d] = fmaf(w, kv[d], o[d]);
    }
  if (active) {
    st8<T>(out + ((long long)s_idx * seq + i) * C + h * HD + c, o);
    if (lse != nullptr && c == 0) lse[row] = mx + __logf(l);
  }
}

// Block = ROWS rows = whole sequences only (rows_used = (ROWS / seq) * seq, the rest of the block idles).
template <typename T, int TMAX>
__global__ void __launch_bounds__(THREADS)
attn_small_bwd_kernel(const T* __restrict__ qkv, const T* __restrict__ out, const T* __restrict__ dout,
                      const float* __restrict__ lse, T* __restrict__ dqkv, long long n_pairs, int seq, int H,
                      float scale, int pairs_per_block) {
  __shared__ float s_delta[ROWS], s_lse[ROWS];
  const int lr = threadIdx.x >> 3;                  // local row
  const int c = (threadIdx.x & 7) * 8;
  const int lp = lr / seq, i = lr % seq;
  const long long pair = blockIdx.x * (long long)pairs_per_block + lp;
  const bool active = lp < pairs_per_block && pair < n_pairs;
  const long long pr = active ? pair : 0;
  const int s_idx = static_cast<int>(pr / H), h = static_cast<int>(pr % H);
  const int C = H * HD;
  const long long pitch = 3LL * C;
  const T* qb = qkv + (long long)s_idx * seq * pitch + h * HD + c;
  const T* ob = out + (long long)s_idx * seq * C + h * HD + c;
  const T* dob = dout + (long long)s_idx * seq * C + h * HD + c;
  T* dqb = dqkv + (long long)s_idx * seq * pitch + h * HD + c;

  float dOi[8], t8[8], u8[8];
  ld8<T>(dob + (long long)i * C, dOi);
  ld8<T>(ob + (long long)i * C, t8);
  const float di = octet_sum(dot8(dOi, t8));        // delta_i = dO_i . O_i
  const float li = lse[pr * seq + i];
  if (c == 0) s_delta[lr] = di, s_lse[lr] = li;
  __syncthreads();
  // ---- phase 1: this octet = query i:  dq_i = scale * sum_j p_ij (dO_i . v_j - delta_i) k_j
  float qi8[8], acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  ld8<T>(qb + i * pitch, qi8);
#pragma unroll
  for (int j = 0; j < TMAX; ++j)
    if (j < seq) {
      ld8<T>(qb + j * pitch + C, t8);               // k_j
      ld8<T>(qb + j * pitch + 2 * C, u8);           // v_j
      const float p = __expf(octet_sum(dot8(qi8, t8)) * scale - li);
      const float ds = p * (octet_sum(dot8(dOi, u8)) - di) * scale;
#pragma unroll
      for (int d = 0; d < 8; ++d) acc[d] = fmaf(ds, t8[d], acc[d]);
    }
  if (active) st8<T>(dqb + i * pitch, acc);
  // ---- phase 2: this octet = key j = i:  dv_j = sum_q p_qj dO_q ;  dk_j = scale * sum_q ds_qj q_q
  float kj[8], vj[8], dk[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, dv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  ld8<T>(qb + i * pitch + C, kj);
  ld8<T>(qb + i * pitch + 2 * C, vj);
#pragma unroll
  for (int qq = 0; qq < TMAX; ++qq)
    if (qq < seq) {
      ld8<T>(qb + qq * pitch, t8);                  // q_q
      ld8<T>(dob + (long long)qq * C, u8);          // dO_q
      const int sr = active ? lp * seq + qq : 0;   // idle octets (rows past the last whole sequence) stay in bounds
      const float p = __expf(octet_sum(dot8(t8, kj)) * scale - s_lse[sr]);
      const float ds = p * (octet_sum(dot8(u8, vj)) - s_delta[sr]) * scale;
#pragma unroll
      for (int d = 0; d < 8; ++d) {
        dv[d] = fmaf(p, u8[d], dv[d]);
        dk[d] = fmaf(ds, t8[d], dk[d]);
      }
    }
  if (active) {
    st8<T>(dqb + i * pitch + C, dk);
    st8<T>(dqb + i * pitch + 2 * C, dv);
  }
}

}  // namespace

// Called from pvrl_attn_fwd / pvrl_attn_bwd (attention_simt.cu) when seq <= 32.
template <typename T>
int attn_small_fwd_launch(const void* qkv, void* out, float* lse, int n_seq, int seq, int H, float scale,
                          cudaStream_t stream) {
  const long long n_rows = (long long)n_seq * H * seq;
  const unsigned grid = static_cast<unsigned>((n_rows + ROWS - 1) / ROWS);
  if (seq <= 8)
    attn_small_fwd_kernel<T, 8><<<grid, THREADS, 0, stream>>>(static_cast<const T*>(qkv), static_cast<T*>(out), lse,
                                                              n_rows, seq, H, scale);
  else
    attn_small_fwd_kernel<T, 32><<<grid, THREADS, 0, stream>>>(static_cast<const T*>(qkv), static_cast<T*>(out), lse,
                                                               n_rows, seq, H, scale);
  return launched("attn_small_fwd_kernel");
}

template <typename T>
int attn_small_bwd_launch(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int n_seq,
                          int seq, int H, float scale, cudaStream_t stream) {
  const long long n_pairs = (long long)n_seq * H;
  const int ppb = ROWS / seq;
  const unsigned grid = static_cast<unsigned>((n_pairs + ppb - 1) / ppb);
  if (seq <= 8)
    attn_small_bwd_kernel<T, 8><<<grid, THREADS, 0, stream>>>(
        static_cast<const T*>(qkv), static_cast<const T*>(out), static_cast<const T*>(dout), lse, static_cast<T*>(dqkv),
        n_pairs, seq, H, scale, ppb);
  else
    attn_small_bwd_kernel<T, 32><<<grid, THREADS, 0, stream>>>(
        static_cast<const T*>(qkv), static_cast<const T*>(out), static_cast<const T*>(dout), lse, static_cast<T*>(dqkv),
        n_pairs, seq, H, scale, ppb);
  return launched("attn_small_bwd_kernel");
}

template int attn_small_fwd_launch<float>(const void*, void*, float*, int, int, int, float, cudaStream_t);
template int attn_small_fwd_launch<__nv_bfloat16>(const void*, void*, float*, int, int, int, float, cudaStream_t);
template int attn_small_bwd_launch<float>(const void*, const void*, const void*, const float*, void*, int, int, int,
                                          float, cudaStream_t);
template int attn_small_bwd_launch<__nv_bfloat16>(const void*, const void*, const void*, const float*, void*, int, int,
                                                  int, float, cudaStream_t);

}  // namespace pvrl
