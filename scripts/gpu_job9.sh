set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "ot_ or order or attn_simt" 2>&1 | tail -40 > gpurun_out/t9.log
tail -n 30 gpurun_out/t9.log
timeout 200 python scripts/op_bench.py --only attn --json gpurun_out/opbench9.json > gpurun_out/opbench9.log 2>&1
cat gpurun_out/opbench9.log
timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/t9_model.log
tail -n 8 gpurun_out/t9_model.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench10.log 2>&1
tail -n 1 gpurun_out/bench10.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches10.csv python bench.py --profile --steps 1 > gpurun_out/ncu_launch10.log 2>&1
