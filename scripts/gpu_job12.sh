set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "gemm" 2>&1 | tail -40 > gpurun_out/t12_gemm.log
tail -n 25 gpurun_out/t12_gemm.log
if grep -q "failed\|error\|Error" gpurun_out/t12_gemm.log; then echo "GEMM2 FAILED"; exit 0; fi
timeout 200 python scripts/op_bench.py --only gemm --json gpurun_out/opbench12_2cta.json > gpurun_out/opbench12_2cta.log 2>&1
cat gpurun_out/opbench12_2cta.log
PVRL_GEMM_2CTA=0 timeout 200 python scripts/op_bench.py --only gemm --json gpurun_out/opbench12_1cta.json > gpurun_out/opbench12_1cta.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/t12.log
tail -n 6 gpurun_out/t12.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench12.log 2>&1
tail -n 1 gpurun_out/bench12.log
