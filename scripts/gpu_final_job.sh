set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/bench_final.log 2>&1; tail -n 1 gpurun_out/bench_final.log | cut -c1-300
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_final.log 2>&1; tail -n 1 gpurun_out/bench_ref_final.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches_final.csv python bench.py --profile --steps 1 > gpurun_out/ncu_launch_final.log 2>&1
timeout 600 ncu --set full --clock-control none -k "regex:layernorm_bwd|ot_linear_dw" -s 38 -c 5 --csv --page raw --log-file gpurun_out/prof_ln_final_raw.csv python bench.py --profile --steps 1 > gpurun_out/ncu_ln_final.log 2>&1
timeout 300 python bench.py --frames 32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_t32_final.log 2>&1; tail -n 1 gpurun_out/bench_t32_final.log | cut -c1-200
python scripts/summarize_ncu_raw.py gpurun_out/prof_ln_final_raw.csv | cut -c1-200
