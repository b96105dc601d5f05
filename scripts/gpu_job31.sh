set -x
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "attn" 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches31.csv python bench.py --profile --steps 1 > gpurun_out/ncu_launch31.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
