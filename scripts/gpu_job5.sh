set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/t5_ops.log
tail -n 3 gpurun_out/t5_ops.log
timeout 200 python scripts/op_bench.py --only gemm --json gpurun_out/opbench5_auto.json > gpurun_out/opbench5_auto.log 2>&1
PVRL_GEMM_BN=256 timeout 200 python scripts/op_bench.py --only gemm --json gpurun_out/opbench5_bn256.json > gpurun_out/opbench5_bn256.log 2>&1
PVRL_GEMM_BN=192 timeout 200 python scripts/op_bench.py --only gemm --json gpurun_out/opbench5_bn192.json > gpurun_out/opbench5_bn192.log 2>&1
timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/t5_model.log
tail -n 3 gpurun_out/t5_model.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench6.log 2>&1
tail -n 1 gpurun_out/bench6.log
cat gpurun_out/opbench5_auto.log
