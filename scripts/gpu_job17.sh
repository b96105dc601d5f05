set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/t17.log
tail -n 20 gpurun_out/t17.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench17.log 2>&1
tail -n 1 gpurun_out/bench17.log
