set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 200 python scripts/op_bench.py --only gemm --json gpurun_out/opbench25.json 2>&1 | tail -14
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
