set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/t6_ops.log
tail -n 8 gpurun_out/t6_ops.log
timeout 200 python scripts/op_bench.py --json gpurun_out/opbench6.json > gpurun_out/opbench6.log 2>&1
timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/t6_model.log
tail -n 8 gpurun_out/t6_model.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench7.log 2>&1
tail -n 1 gpurun_out/bench7.log
cat gpurun_out/opbench6.log
