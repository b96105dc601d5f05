set -x
for w in 1 2 3; do echo "WAVES $w"; PVRL_GEMM2_WAVES=$w timeout 200 python scripts/op_bench.py --only "dW" 2>&1 | tail -4; done
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "gemm" 2>&1 | tail -3
