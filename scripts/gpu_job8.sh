set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "ot_ or order" 2>&1 | tail -40 > gpurun_out/t8_ot.log
tail -n 30 gpurun_out/t8_ot.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench9.log 2>&1
tail -n 1 gpurun_out/bench9.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches9.csv python bench.py --profile --steps 1 > gpurun_out/ncu_launch9.log 2>&1
