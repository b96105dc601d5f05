set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/t20.log
tail -n 8 gpurun_out/t20.log
timeout 200 python scripts/op_bench.py --json gpurun_out/opbench20.json > gpurun_out/opbench20.log 2>&1
grep -E "GELU|gather|colsum|cast" gpurun_out/opbench20.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench20.log 2>&1
tail -n 1 gpurun_out/bench20.log
