import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from procedurevrl_b200 import ops
L = ops.lib()
M, N, K = 28242, 768, 3072
A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(N, K, device="cuda").bfloat16(); out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
def t(n=20):
    for _ in range(3): ops.gemm(A, B, out, M=M, N=N, K=K)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): ops.gemm(A, B, out, M=M, N=N, K=K)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
us = t()
L.pvrl_debug_gemm2_max_pairs.restype = ctypes.c_int
print("pairs override", os.environ.get("PVRL_GEMM2_PAIRS"), "max active pairs", L.pvrl_debug_gemm2_max_pairs(), f"{us:.1f} us {2*M*N*K/us/1e6:.0f} TF/s")
