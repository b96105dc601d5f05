// Micro-benchmark (development aid): per-SMSP throughput of MUFU.EX2, FFMA, FFMA2 (packed fp32x2), F2FP and the softmax
// inner-loop mix, for 1 / 2 / 4 warps per scheduler.  One CTA, warps w with w % 4 == 0 all land on SMSP 0.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../procedurevrl_b200/csrc/pvrl_ptx.cuh"
using namespace pvrl;

template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, long long* cyc, int iters, int warps_per_smsp) {
  const int warp = threadIdx.x >> 5;
  const bool active = (warp & 3) == 0 && (warp >> 2) < warps_per_smsp;
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = -0.001f * (threadIdx.x + i);
  float acc = 0.f;
  uint64_t acc2 = 0ull;
  __syncthreads();
  long long t0 = clock64();
  if (active) {
    for (int it = 0; it < iters; ++it) {
      if (MODE == 0) {          // 16 MUFU.EX2
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = ex2_approx(x[i]) - 1.0009765625f;
      } else if (MODE == 1) {   // 16 FFMA (3-register form)
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], x[(i + 1) & 15], x[(i + 5) & 15]);
      } else if (MODE == 2) {   // 8 FFMA2
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          float a, b;
          f2_unpack(f2_fma(f2_pack(x[i], x[i + 1]), f2_pack(x[(i + 2) & 15], x[(i + 3) & 15]), f2_pack(x[(i + 6) & 15], x[(i + 7) & 15])), a, b);
          x[i] = a; x[i + 1] = b;
        }
      } else if (MODE == 3) {   // softmax mix, scalar: 16 x (FFMA, MUFU, FADD) + 8 F2FP
        uint32_t pk = 0;
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          float p0 = ex2_approx(fmaf(x[i], 1.01f, -0.5f)), p1 = ex2_approx(fmaf(x[i + 1], 1.01f, -0.5f));
          acc += p0; acc += p1;
          pk ^= pack_bf16x2(p0, p1);
          x[i] = p0 - 1.25f; x[i + 1] = p1 - 1.25f;
        }
        acc += __uint_as_float(pk & 0x3f800000);
      } else if (MODE == 4) {   // softmax mix, packed: 8 x (FFMA2, 2 MUFU, FADD2, F2FP)
        uint32_t pk = 0;
        const uint64_t s2 = f2_pack(1.01f, 1.01f), m2 = f2_pack(-0.5f, -0.5f);
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          float a, b;
          f2_unpack(f2_fma(f2_pack(x[i], x[i + 1]), s2, m2), a, b);
          float p0 = ex2_approx(a), p1 = ex2_approx(b);
          acc2 = f2_add(acc2, f2_pack(p0, p1));
          pk ^= pack_bf16x2(p0, p1);
          x[i] = p0 - 1.25f; x[i + 1] = p1 - 1.25f;
        }
        acc += __uint_as_float(pk & 0x3f800000);
      } else if (MODE == 5) {   // 16 FMNMX
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fmaxf(x[i], x[(i + 3) & 15] + 1.f);
      }
    }
  }
  long long t1 = clock64();
  float s = acc;
  float a, b; f2_unpack(acc2, a, b); s += a + b;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter) {
  float* o; long long* c; cudaMalloc(&o, 4096); cudaMalloc(&c, 8);
  for (int w : {1, 2, 4}) {
    const int iters = 2000;
    k<MODE><<<1, 512>>>(o, c, iters, w);
    k<MODE><<<1, 512>>>(o, c, iters, w);
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-44s warps/SMSP=%d: %.2f cycles per warp-instruction (pipe, aggregated over the scheduler's warps)\n", name, w,
           (double)h / ((double)iters * per_iter * w));
  }
}

int main() {
  run<0>("MUFU.EX2 (+FADD)", 16);
  run<1>("FFMA 3-reg", 16);
  run<2>("FFMA2 packed (per packed instr)", 8);
  run<5>("FADD+FMNMX", 16);
  run<3>("softmax mix scalar (per element)", 16);
  run<4>("softmax mix packed (per element)", 16);
  return 0;
}
