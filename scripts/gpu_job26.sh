set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 200 python scripts/op_bench.py --only "gemm" 2>&1 | grep -E "GELU"
timeout 200 python scripts/op_bench.py --only "colsum" 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
