set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "gather" 2>&1 | tail -3
timeout 300 python scripts/op_bench.py --cold --iters 10 --json gpurun_out/opbench28_cold.json 2>&1 | tail -32
