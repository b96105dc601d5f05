// Host-side helpers shared by the .cu translation units: error reporting and launch accounting.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <utility>

#include "../../include/pvrl.h"

namespace pvrl {

char* last_error_buf();              // thread-local, 512 bytes
std::atomic<int64_t>& launch_counter();

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

// Call after every kernel launch: counts it and converts a launch error into a return code.
inline int launched(const char* what) {
  launch_counter().fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(static_cast<int>(e), "%s: %s", what, cudaGetErrorString(e));
  return 0;
}

#define PVRL_CHECK_ARG(cond, ...) \
  do {                            \
    if (!(cond)) return ::pvrl::fail(-1, __VA_ARGS__); \
  } while (0)

#define PVRL_CUDA(expr)                                                                         \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess) return ::pvrl::fail(static_cast<int>(e__), "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

// Programmatic dependent launch: the kernel may be scheduled while its predecessor in the stream is still draining (its
// CTAs take SMs as the predecessor's CTAs exit, run their prologue -- barrier init, TMEM allocation, descriptor prefetch --
// and then block in griddepcontrol.wait until the predecessor has completed and flushed).  Every kernel launched through
// this helper calls pdl_wait() before its first dependent global access.
// Measured on the 18-clip step (profiles/README.md): with every kernel releasing its dependents at its start the step got
// 6 % SLOWER (35.0 vs 32.9 ms), with the implicit trigger at grid completion it is neutral (33.1 vs 33.1 ms) -- the
// kernels here are long enough that launch latency is already hidden by the CUDA graph.  So it is opt-in: PVRL_PDL=1.
inline bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("PVRL_PDL");
    return e != nullptr && atoi(e) != 0;
  }();
  return on;
}
template <typename Kern, typename... Args>
inline cudaError_t launch_pdl(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn();   // cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda)
int num_sms();
// cached 2-D bf16 tensor map over a row-major [outer, inner] matrix (gemm_sm100.cu)
int make_tmap_2d_bf16(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                      uint32_t box_outer);

struct Geom {
  int T, HW, L, S;
  __host__ __device__ Geom() : T(1), HW(1), L(1), S(2) {}
  __host__ __device__ Geom(int t, int hw) : T(t), HW(hw), L(t * hw), S(1 + t * hw) {}
};

// Row maps of include/pvrl.h.  Returns the residual-stream row of logical row m; for the cls rows of
// MAP_SPATIAL returns -(bt + 1) (the caller decides what a cls row means for it).
__host__ __device__ inline long long map_row(int map, int m, const Geom& g) {
  switch (map) {
    case PVRL_MAP_SKIPCLS:
      return (long long)m + m / g.L + 1;
    case PVRL_MAP_SPATIAL: {
      int bt = m / (g.HW + 1), n = m - bt * (g.HW + 1);
      if (n == 0) return -(long long)(bt + 1);
      int b = bt / g.T, t = bt - b * g.T;
      return (long long)b * g.S + 1 + (long long)(n - 1) * g.T + t;
    }
    case PVRL_MAP_PATCH: {
      int bt = m / g.HW, n = m - bt * g.HW;
      int b = bt / g.T, t = bt - b * g.T;
      return (long long)b * g.S + 1 + (long long)n * g.T + t;
    }
    case PVRL_MAP_CLS:
      return (long long)m * g.S;
    default:
      return m;
  }
}

}  // namespace pvrl
