set -x
mkdir -p gpurun_out
timeout 200 compute-sanitizer --tool memcheck python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "attn_simt and 3-7-5" 2>&1 | grep -v "^$" | head -60 > gpurun_out/t10_sanitizer.log
head -40 gpurun_out/t10_sanitizer.log
timeout 200 python scripts/order_bench.py > gpurun_out/order_bench.log 2>&1
cat gpurun_out/order_bench.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_tc -o gpurun_out/prof_attn_tc python scripts/op_bench.py --only attn_tc --iters 1 --warm 1 > gpurun_out/ncu_attn_tc.log 2>&1
ls -la gpurun_out/*.ncu-rep
