set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "ot_ or order" 2>&1 | tail -40 > gpurun_out/t7_ot.log
tail -n 30 gpurun_out/t7_ot.log
timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/t7_model.log
tail -n 8 gpurun_out/t7_model.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench8.log 2>&1
tail -n 1 gpurun_out/bench8.log
