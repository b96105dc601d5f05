set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 400 python bench.py > gpurun_out/bench29.log 2>&1
tail -n 1 gpurun_out/bench29.log
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench29_ref.log 2>&1
tail -n 1 gpurun_out/bench29_ref.log | cut -c1-300
timeout 600 ncu --set full --clock-control none -k regex:gemm -s 300 -c 14 --csv --page raw --log-file gpurun_out/prof_gemm29_raw.csv python bench.py --profile --steps 1 > gpurun_out/ncu_gemm29.log 2>&1
timeout 400 ncu --set full --clock-control none -k "regex:attn_t8|layernorm|gather_cast|attn_tc" -s 200 -c 12 --csv --page raw --log-file gpurun_out/prof_misc29_raw.csv python bench.py --profile --steps 1 > gpurun_out/ncu_misc29.log 2>&1
timeout 200 python scripts/op_bench.py --json gpurun_out/opbench29.json > gpurun_out/opbench29.log 2>&1
