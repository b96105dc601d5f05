set -x
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "attn_tensor" 2>&1 | tail -8
timeout 200 python scripts/op_bench.py --only attn_tc 2>&1 | tail -2
timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
