set -x
python scripts/gemm2_probe.py 2>&1 | tail -1
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "gemm or attn_tensor" 2>&1 | tail -5
timeout 200 python scripts/op_bench.py --only gemm 2>&1 | tail -14
timeout 200 python scripts/op_bench.py --only attn_tc 2>&1 | tail -3
