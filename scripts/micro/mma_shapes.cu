// Micro-benchmark (development aid): cycles for a burst of tcgen05.mma instructions of one shape issued by one thread,
// SS (A in smem) vs TS (A in TMEM), K-major vs MN-major B.  nvcc -gencode arch=compute_100a,code=sm_100a -o mma_shapes
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../procedurevrl_b200/csrc/pvrl_ptx.cuh"
using namespace pvrl;

__global__ void __launch_bounds__(128) k(long long* out, int N, int count, int ts, int b_mn, int reps) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar;
  const uint32_t bar_a = smem_u32(&bar);
  if (threadIdx.x == 0) { mbar_init(bar_a, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(smem_u32(&slot));
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, 0, b_mn);
    const uint32_t sA = base, sB = base + 64 * 1024;
    for (int r = 0; r < reps; ++r) {
      long long t0 = clock64();
      for (int i = 0; i < count; ++i) {
        const int kk = i & 3;
        uint64_t bd = b_mn ? make_smem_desc(sB + (i % 16) * 2048, 8192, 1024) : make_smem_desc(sB + kk * 32, 16, 1024);
        if (ts) umma_bf16_ts(tmem + 256, tmem + (i % 16) * 8, bd, idesc, i > 0);
        else umma_bf16(tmem + 256, make_smem_desc(sA + kk * 32, 16, 1024), bd, idesc, i > 0);
      }
      long long t1 = clock64();
      umma_commit(bar_a);
      mbar_wait(bar_a, r & 1);
      long long t2 = clock64();
      out[2 * r] = t1 - t0; out[2 * r + 1] = t2 - t0;
    }
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

int main() {
  long long* d; cudaMalloc(&d, 64 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int Ns[] = {64, 128, 208, 256};
  int counts[] = {1, 4, 13, 26};
  for (int ts = 0; ts < 2; ++ts) for (int bmn = 0; bmn < 2; ++bmn) for (int N : Ns) for (int c : counts) {
    if (bmn && N > 128) continue;
    k<<<1, 128, 200 * 1024>>>(d, N, c, ts, bmn, 4);
    long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s B=%s N=%3d count=%2d: issue %5lld total %5lld cyc (%.1f cyc/mma)  %s\n", ts ? "TS" : "SS", bmn ? "MN" : "K ", N, c, h[6], h[7],
           (double)h[7] / c, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
