set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/t3.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench4_graph.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches3.csv python bench.py --profile --steps 1 > gpurun_out/ncu_launch3.log 2>&1
tail -n 8 gpurun_out/t3.log; tail -n 3 gpurun_out/bench4_graph.log
