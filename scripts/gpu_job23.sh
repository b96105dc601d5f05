set -x
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "layernorm" 2>&1 | tail -4
timeout 200 python scripts/op_bench.py --only layernorm 2>&1 | tail -6
