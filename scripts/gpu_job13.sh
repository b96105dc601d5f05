set -x
mkdir -p gpurun_out
python scripts/gemm2_probe.py 2>&1 | tail -2
for p in 16 32 37 48 64 72 74; do PVRL_GEMM2_PAIRS=$p python scripts/gemm2_probe.py 2>&1 | tail -1; done
