// Persistent warp-specialised bf16 GEMM for sm_100a: TMA (128B-swizzled tiles) -> 4-stage smem ring ->
// tcgen05.mma (128x256x16, fp32 accumulators double-buffered in TMEM) -> fused epilogues straight from
// TMEM to global memory.  One CTA per SM; warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM
// allocator, warps 4..11 = epilogue (each owns a 32-lane TMEM quarter and one 128-column half).
//
// Covers every dense contraction of the TimeSformer block (SURVEY.md 2a K1,K4,K6,K8,K10 and their
// backward K15): "NT" for y = x W^T (+ dX through pre-transposed weights) and "TN" (MN-major operands,
// split-K + fp32 red.add) for dW = dY^T X.
#include <mutex>
#include <unordered_map>

#include "pvrl_host.h"
#include "pvrl_ptx.cuh"

namespace pvrl {
namespace {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;          // 16 KB
constexpr int B_BYTES = BN * BK * 2;          // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int CHUNK_BYTES = 64 * BK * 2;      // one 64-wide MN chunk of a TN tile (8 KB)
constexpr int NUM_THREADS = 384;
constexpr int NUM_EPI_WARPS = 8;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 128 /*barriers*/;

struct GemmArgs {
  int M, N, K;
  int k_splits, kb_per_split;
  void* out;
  long long ldo;
  void* out2;
  const float* bias;
  const float* rowscale;
  int rs_div;
  int map;
  const void* aux;
  long long ld_aux;
  const float* resid;
  const float* add_pos;
  const float* add_time;
  Geom g;
};

// ---- epilogue helpers: one thread owns one output row and 32 consecutive columns -------------------
template <typename OutT>
__device__ __forceinline__ void store_row32(OutT* dst, const float (&v)[32]);
template <>
__device__ __forceinline__ void store_row32<float>(float* dst, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    reinterpret_cast<float4*>(dst)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
template <>
__device__ __forceinline__ void store_row32<__nv_bfloat16>(__nv_bfloat16* dst, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 u;
    u.x = pack_bf16x2(v[8 * j], v[8 * j + 1]);
    u.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
    u.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
    u.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
    reinterpret_cast<uint4*>(dst)[j] = u;
  }
}
template <typename T>
__device__ __forceinline__ void load_row32(const T* src, float (&v)[32]);
template <>
__device__ __forceinline__ void load_row32<float>(const float* src, float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float4 f = __ldg(reinterpret_cast<const float4*>(src) + j);
    v[4 * j] = f.x, v[4 * j + 1] = f.y, v[4 * j + 2] = f.z, v[4 * j + 3] = f.w;
  }
}
template <>
__device__ __forceinline__ void load_row32<__nv_bfloat16>(const __nv_bfloat16* src, float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 u = __ldg(reinterpret_cast<const uint4*>(src) + j);
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    v[8 * j] = a.x, v[8 * j + 1] = a.y, v[8 * j + 2] = b.x, v[8 * j + 3] = b.y;
    v[8 * j + 4] = c.x, v[8 * j + 5] = c.y, v[8 * j + 6] = d.x, v[8 * j + 7] = d.y;
  }
}

template <int EPI, typename OutT>
__device__ __forceinline__ void epilogue_row(const GemmArgs& p, int m, int n0, float (&acc)[32]) {
  if (EPI == PVRL_EPI_ATOMIC) {
    float* dst = reinterpret_cast<float*>(p.out) + (long long)m * p.ldo + n0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * j), "f"(acc[4 * j]),
                   "f"(acc[4 * j + 1]), "f"(acc[4 * j + 2]), "f"(acc[4 * j + 3])
                   : "memory");
    return;
  }
  if (p.bias != nullptr) {
    float b[32];
    load_row32<float>(p.bias + n0, b);
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] += b[j];
  }
  if (EPI == PVRL_EPI_GELU) {
    store_row32<OutT>(reinterpret_cast<OutT*>(p.out) + (long long)m * p.ldo + n0, acc);
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = gelu_erf(acc[j]);
    store_row32<OutT>(reinterpret_cast<OutT*>(p.out2) + (long long)m * p.ldo + n0, acc);
    return;
  }
  if (EPI == PVRL_EPI_DGELU) {
    float a[32];
    load_row32<OutT>(reinterpret_cast<const OutT*>(p.aux) + (long long)m * p.ld_aux + n0, a);
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] *= gelu_erf_grad(a[j]);
  }
  if (p.rowscale != nullptr) {
    const float s = __ldg(p.rowscale + m / p.rs_div);
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] *= s;
  }
  const long long orow = map_row(p.map, m, p.g);
  if (EPI == PVRL_EPI_RESID) {
    float* out = reinterpret_cast<float*>(p.out);
    if (orow < 0) {  // cls row of a spatial sequence: park it for the mean over frames (vit.py:147-149)
      store_row32<float>(reinterpret_cast<float*>(p.out2) + (-orow - 1) * p.ldo + n0, acc);
      return;
    }
    if (p.resid != nullptr) {
      float r[32];
      load_row32<float>(p.resid + orow * p.ldo + n0, r);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] += r[j];
    }
    if (p.add_pos != nullptr) {  // MAP_PATCH: + pos_embed[1+n] + time_embed[t]  (vit.py:373-404)
      const int bt = m / p.g.HW, n = m - bt * p.g.HW, t = bt % p.g.T;
      float r[32];
      load_row32<float>(p.add_pos + (long long)(1 + n) * p.N + n0, r);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] += r[j];
      load_row32<float>(p.add_time + (long long)t * p.N + n0, r);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] += r[j];
    }
    store_row32<float>(out + orow * p.ldo + n0, acc);
    return;
  }
  // STORE / DGELU
  store_row32<OutT>(reinterpret_cast<OutT*>(p.out) + orow * p.ldo + n0, acc);
}

// tile index -> (m block, n block, k split).  NT: n fastest, so the CTAs of a wave share A tiles through L2 and the
// (small) weight matrix stays resident.  TN (dW, split-K): the k split is the SLOWEST index, so at any time the
// resident CTAs cover all output tiles of one or two contraction slices and every slice of dY / X is fetched from
// HBM once instead of once per output tile.
template <bool TN>
__device__ __forceinline__ void decode_tile(int tile, int m_tiles, int n_tiles, int k_splits, int& m_blk, int& n_blk,
                                            int& ks) {
  if (TN) {
    const int per = m_tiles * n_tiles;
    ks = tile / per;
    const int rest = tile - ks * per;
    n_blk = rest % n_tiles, m_blk = rest / n_tiles;
  } else {
    ks = tile % k_splits;
    const int rest = tile / k_splits;
    n_blk = rest % n_tiles, m_blk = rest / n_tiles;
  }
}

template <int EPI, typename OutT, bool TN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t tiles_addr = (raw_addr + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024 B alignment
  uint8_t* smem = smem_raw + (tiles_addr - raw_addr);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  const uint32_t bars_addr = tiles_addr + STAGES * STAGE_BYTES;
  // barrier slots: full[0..3], empty[4..7], tmem_full[8..9], tmem_empty[10..11], tmem ptr at slot 12
  auto full_bar = [&](int s) { return bars_addr + 8u * s; };
  auto empty_bar = [&](int s) { return bars_addr + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bars_addr + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bars_addr + 8u * (2 * STAGES + 2 + s); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (p.M + BM - 1) / BM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int num_kb = (p.K + BK - 1) / BK;
  const int total_tiles = m_tiles * n_tiles * p.k_splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), NUM_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int m_blk, n_blk, ks;
        decode_tile<TN>(tile, m_tiles, n_tiles, p.k_splits, m_blk, n_blk, ks);
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(num_kb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), STAGE_BYTES);
          const uint32_t sA = tiles_addr + stage * STAGE_BYTES;
          const uint32_t sB = sA + A_BYTES;
          if (!TN) {
            tma_load_2d(sA, &tmA, full_bar(stage), kb * BK, m_blk * BM);
            tma_load_2d(sB, &tmB, full_bar(stage), kb * BK, n_blk * BN);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c)
              tma_load_2d(sA + c * CHUNK_BYTES, &tmA, full_bar(stage), m_blk * BM + c * 64, kb * BK);
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_2d(sB + c * CHUNK_BYTES, &tmB, full_bar(stage), n_blk * BN + c * 64, kb * BK);
          }
          if (++stage == STAGES) stage = 0, phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, TN ? 1 : 0, TN ? 1 : 0);
      // descriptor advance per UMMA_K = 16: K-major +32 B inside the swizzle atom, MN-major +16 k-rows
      constexpr uint32_t kstep = TN ? 16u * 128u : 32u;
      constexpr uint32_t lbo = TN ? CHUNK_BYTES : 16u;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int m_blk, n_blk, ks;
        decode_tile<TN>(tile, m_tiles, n_tiles, p.k_splits, m_blk, n_blk, ks);
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(num_kb, kb0 + p.kb_per_split);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sA = tiles_addr + stage * STAGE_BYTES;
          const uint32_t sB = sA + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = make_smem_desc(sA + k * kstep, lbo, 1024);
            const uint64_t bdesc = make_smem_desc(sB + k * kstep, lbo, 1024);
            umma_bf16(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));  // smem slot reusable once these MMAs have read it
          if (++stage == STAGES) stage = 0, phase ^= 1u;
        }
        umma_commit(tfull_bar(acc));      // accumulator complete -> epilogue
        if (++acc == 2) acc = 0, acc_phase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: TMEM -> regs -> global
    const int quarter = warp & 3;          // TMEM lanes [32*quarter, +32) are the only ones this warp may read
    const int half = (warp - 4) >> 2;      // column half of the 256-wide accumulator
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int m_blk, n_blk, ks;
      decode_tile<TN>(tile, m_tiles, n_tiles, p.k_splits, m_blk, n_blk, ks);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int m = m_blk * BM + quarter * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int col0 = half * 128 + c * 32;
        const int n0 = n_blk * BN + col0;
        if (n0 >= p.N) break;  // warp-uniform
        uint32_t raw[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + col0, raw);
        tmem_ld_wait();
        if (m < p.M) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
          epilogue_row<EPI, OutT>(p, m, n0, v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) acc = 0, acc_phase ^= 1u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

struct MapKey {
  const void* ptr;
  uint64_t inner, outer, ld;
  uint32_t box_inner, box_outer;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && ld == o.ld && box_inner == o.box_inner &&
           box_outer == o.box_outer;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h = h * 1000003u ^ k.inner;
    h = h * 1000003u ^ k.outer;
    h = h * 1000003u ^ k.ld;
    h = h * 1000003u ^ (k.box_inner << 16 | k.box_outer);
    return h;
  }
};

}  // namespace

// ---- host: TMA descriptor cache ---------------------------------------------------------------------
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

// 2-D bf16 tensor map over a row-major [outer, inner] matrix with leading dimension ld (elements),
// box = [box_outer, box_inner], 128-byte swizzle, zero fill out of bounds.
int make_tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                      uint32_t box_inner, uint32_t box_outer) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{ptr, inner, outer, ld, box_inner, box_outer};
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) return fail(-2, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15u) != 0 || (ld * 2) % 16 != 0)
    return fail(-1, "TMA operand must be 16-byte aligned with a leading dimension multiple of 8 elements");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(-3, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  std::lock_guard<std::mutex> lock(mu);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return 0;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

namespace {

template <int EPI, typename OutT, bool TN>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& a, cudaStream_t stream) {
  auto kern = gemm_bf16_kernel<EPI, OutT, TN>;
  static bool configured = false;
  if (!configured) {
    PVRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  const int m_tiles = (a.M + BM - 1) / BM, n_tiles = (a.N + BN - 1) / BN;
  const int total = m_tiles * n_tiles * a.k_splits;
  const int grid = total < num_sms() ? total : num_sms();
  kern<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb, a);
  return launched("gemm_bf16_kernel");
}

}  // namespace
}  // namespace pvrl

extern "C" int pvrl_gemm_bf16(const pvrl_gemm_t* d, void* stream_) {
  using namespace pvrl;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PVRL_CHECK_ARG(d != nullptr, "pvrl_gemm_bf16: null descriptor");
  PVRL_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0, "pvrl_gemm_bf16: empty problem M=%d N=%d K=%d", d->M, d->N, d->K);
  PVRL_CHECK_ARG(d->N % 32 == 0, "pvrl_gemm_bf16: N=%d must be a multiple of 32", d->N);
  PVRL_CHECK_ARG(d->A && d->B && d->out, "pvrl_gemm_bf16: null operand");
  PVRL_CHECK_ARG(d->epilogue >= PVRL_EPI_STORE && d->epilogue <= PVRL_EPI_ATOMIC, "pvrl_gemm_bf16: bad epilogue %d",
                 d->epilogue);
  if (d->trans == 1) PVRL_CHECK_ARG(d->epilogue == PVRL_EPI_ATOMIC, "pvrl_gemm_bf16: TN form supports ATOMIC only");
  if (d->epilogue == PVRL_EPI_GELU) PVRL_CHECK_ARG(d->out2 != nullptr, "pvrl_gemm_bf16: GELU needs out2");
  if (d->epilogue == PVRL_EPI_DGELU) PVRL_CHECK_ARG(d->aux != nullptr, "pvrl_gemm_bf16: DGELU needs aux");
  if (d->rowscale) PVRL_CHECK_ARG(d->rs_div > 0, "pvrl_gemm_bf16: rowscale needs rs_div > 0");
  if (d->map == PVRL_MAP_SPATIAL)
    PVRL_CHECK_ARG(d->epilogue == PVRL_EPI_RESID && d->out2 != nullptr,
                   "pvrl_gemm_bf16: MAP_SPATIAL needs the RESID epilogue and a cls side buffer");

  GemmArgs a;
  a.M = d->M, a.N = d->N, a.K = d->K;
  a.out = d->out, a.ldo = d->ldo, a.out2 = d->out2;
  a.bias = d->bias, a.rowscale = d->rowscale, a.rs_div = d->rs_div > 0 ? d->rs_div : 1;
  a.map = d->map, a.aux = d->aux, a.ld_aux = d->ld_aux, a.resid = d->resid;
  a.add_pos = d->add_pos, a.add_time = d->add_time;
  a.g = Geom(d->g.T > 0 ? d->g.T : 1, d->g.HW > 0 ? d->g.HW : 1);

  const int num_kb = (d->K + BK - 1) / BK;
  int splits = 1;
  if (d->epilogue == PVRL_EPI_ATOMIC) {
    splits = d->k_splits;
    if (splits <= 0) {  // pick the split count that fills whole waves of SMs
      const int tiles = ((d->M + BM - 1) / BM) * ((d->N + BN - 1) / BN);
      const int sms = num_sms();
      double best = -1.0;
      splits = 1;
      for (int s = 1; s <= 64 && s <= num_kb; ++s) {
        if ((num_kb + s - 1) / s < 4 && s > 1) break;
        const int work = tiles * s;
        const double util = static_cast<double>(work) / (((work + sms - 1) / sms) * sms);
        if (util > best + 0.02) best = util, splits = s;
      }
    }
    if (splits > num_kb) splits = num_kb;
  }
  a.kb_per_split = (num_kb + splits - 1) / splits;
  a.k_splits = (num_kb + a.kb_per_split - 1) / a.kb_per_split;  // no empty split

  CUtensorMap ta, tb;
  int rc;
  if (d->trans == 0) {
    if ((rc = make_tmap_2d_bf16(&ta, d->A, d->K, d->M, d->lda, BK, BM))) return rc;
    if ((rc = make_tmap_2d_bf16(&tb, d->B, d->K, d->N, d->ldb, BK, BN))) return rc;
  } else {
    if ((rc = make_tmap_2d_bf16(&ta, d->A, d->M, d->K, d->lda, 64, BK))) return rc;
    if ((rc = make_tmap_2d_bf16(&tb, d->B, d->N, d->K, d->ldb, 64, BK))) return rc;
  }

  const bool f32 = d->out_dtype == PVRL_F32;
  switch (d->epilogue) {
    case PVRL_EPI_STORE:
      return f32 ? launch_gemm<PVRL_EPI_STORE, float, false>(ta, tb, a, stream)
                 : launch_gemm<PVRL_EPI_STORE, __nv_bfloat16, false>(ta, tb, a, stream);
    case PVRL_EPI_GELU:
      return f32 ? launch_gemm<PVRL_EPI_GELU, float, false>(ta, tb, a, stream)
                 : launch_gemm<PVRL_EPI_GELU, __nv_bfloat16, false>(ta, tb, a, stream);
    case PVRL_EPI_DGELU:
      return f32 ? launch_gemm<PVRL_EPI_DGELU, float, false>(ta, tb, a, stream)
                 : launch_gemm<PVRL_EPI_DGELU, __nv_bfloat16, false>(ta, tb, a, stream);
    case PVRL_EPI_RESID:
      return launch_gemm<PVRL_EPI_RESID, float, false>(ta, tb, a, stream);
    default:
      return d->trans ? launch_gemm<PVRL_EPI_ATOMIC, float, true>(ta, tb, a, stream)
                      : launch_gemm<PVRL_EPI_ATOMIC, float, false>(ta, tb, a, stream);
  }
}
