set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/t_all2.log
python bench.py --steps 10 --warmup 3 --no-graph > gpurun_out/bench3_eager.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench3_graph.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches2.csv python bench.py --profile --steps 1 > gpurun_out/ncu_launch2.log 2>&1
tail -n 5 gpurun_out/t_all2.log; tail -n 3 gpurun_out/bench3_eager.log; tail -n 3 gpurun_out/bench3_graph.log
