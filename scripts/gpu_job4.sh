set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/t4_ops.log
tail -n 5 gpurun_out/t4_ops.log
timeout 300 python scripts/op_bench.py --json gpurun_out/opbench4_auto.json > gpurun_out/opbench4_auto.log 2>&1
PVRL_GEMM_BN=256 timeout 300 python scripts/op_bench.py --only gemm --json gpurun_out/opbench4_bn256.json > gpurun_out/opbench4_bn256.log 2>&1
PVRL_GEMM_BN=192 timeout 300 python scripts/op_bench.py --only gemm --json gpurun_out/opbench4_bn192.json > gpurun_out/opbench4_bn192.log 2>&1
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/t4_model.log
tail -n 5 gpurun_out/t4_model.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench5.log 2>&1
tail -n 2 gpurun_out/bench5.log
timeout 900 ncu --set full --clock-control none --csv --page raw --log-file gpurun_out/ncu4_ops_raw.csv python scripts/op_bench.py --iters 1 --warm 0 > gpurun_out/ncu4_ops.log 2>&1
cat gpurun_out/opbench4_auto.log
