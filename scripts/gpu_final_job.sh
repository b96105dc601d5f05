#!/bin/bash
# Round-end validation on one B200 (run through gpurun): full GPU test suite, smoke, the default bench line, the ncu launch
# list of the steady-state step and full-set captures of the attention / residual-GEMM kernels.  Everything lands in gpurun_out/.
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_tests.log 2>&1; tail -4 gpurun_out/r02_final_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_final_bench.log 2>&1; grep "^{" gpurun_out/r02_final_bench.log | tail -1 > gpurun_out/r02_bench_n1_final.json; cut -c1-400 gpurun_out/r02_bench_n1_final.json
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | grep "^{" | tail -1 > gpurun_out/r02_bench_reference_arm_final.json; cut -c1-300 gpurun_out/r02_bench_reference_arm_final.json
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --profile --steps 1 > gpurun_out/ncu_launch_final.log 2>&1
python scripts/summarize_launches.py gpurun_out/r02_launches_final.csv | head -30
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:attn_sp_fwd|attn_tc_bwd" -s 12 -c 2 -o gpurun_out/r02_prof_attn_final -f python bench.py --profile --steps 1 > gpurun_out/ncu_attn_final.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:gemm -s 300 -c 14 --csv --page raw --log-file gpurun_out/r02_prof_gemm_final_raw.csv python bench.py --profile --steps 1 > gpurun_out/ncu_gemm_final.log 2>&1
python scripts/summarize_ncu_raw.py gpurun_out/r02_prof_gemm_final_raw.csv | cut -c1-200
