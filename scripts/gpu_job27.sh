set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches27.csv python bench.py --profile --steps 1 > gpurun_out/ncu_launch27.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --frames 32 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
