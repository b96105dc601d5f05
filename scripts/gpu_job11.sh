set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/t11.log
tail -n 12 gpurun_out/t11.log
timeout 200 python scripts/order_bench.py > gpurun_out/order_bench2.log 2>&1
head -4 gpurun_out/order_bench2.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench11.log 2>&1
tail -n 1 gpurun_out/bench11.log
