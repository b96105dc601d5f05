// Micro-benchmark (development aid): TMEM -> register (tcgen05.ld) and register -> TMEM (tcgen05.st) throughput per SM for
// 1..16 warps.  Warps w and w + 4 share a TMEM lane quarter.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../procedurevrl_b200/csrc/pvrl_ptx.cuh"
using namespace pvrl;

template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, long long* cyc, int iters, int n_warps) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t tb = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 64;
  uint32_t a[32], b[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) a[i] = b[i] = threadIdx.x + i;
  tmem_st16(tb, reinterpret_cast<uint32_t(&)[16]>(a));
  tmem_st_wait();
  __syncthreads();
  long long t0 = clock64();
  if (warp < n_warps) {
    for (int it = 0; it < iters; ++it) {
      if (MODE == 0) {            // two x32 loads in flight, one wait
        tmem_ld32(tb, a);
        tmem_ld32(tb + 32, b);
        tmem_ld_wait_on(a);
        tmem_ld_wait_on(b);
      } else if (MODE == 1) {     // x16 loads, one wait each (the kernel's pattern)
        tmem_ld16(tb, reinterpret_cast<uint32_t(&)[16]>(a));
        tmem_ld_wait_on(reinterpret_cast<uint32_t(&)[16]>(a));
        tmem_ld16(tb + 16, reinterpret_cast<uint32_t(&)[16]>(b));
        tmem_ld_wait_on(reinterpret_cast<uint32_t(&)[16]>(b));
      } else {                    // x16 stores
        tmem_st16(tb, reinterpret_cast<uint32_t(&)[16]>(a));
        tmem_st16(tb + 16, reinterpret_cast<uint32_t(&)[16]>(b));
        tmem_st_wait();
      }
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += a[i] ^ b[i];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int MODE>
void run(const char* name, int bytes_per_iter_per_warp) {
  float* o; long long* c; cudaMalloc(&o, 4096); cudaMalloc(&c, 8);
  for (int w : {1, 2, 4, 8, 16}) {
    const int iters = 2000;
    k<MODE><<<1, 512>>>(o, c, iters, w);
    k<MODE><<<1, 512>>>(o, c, iters, w);
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-28s warps=%2d: %7.1f cycles/iter  %7.1f B/clk per SM\n", name, w, (double)h / iters,
           (double)bytes_per_iter_per_warp * w * iters / h);
  }
}

int main() {
  run<0>("tcgen05.ld 2 x .x32", 2 * 32 * 32 * 4);
  run<1>("tcgen05.ld .x16 + wait, x2", 2 * 16 * 32 * 4);
  run<2>("tcgen05.st 2 x .x16", 2 * 16 * 32 * 4);
  return 0;
}
