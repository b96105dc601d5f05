set -x
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/t_all.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --profile --steps 1 > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 150 -c 6 -o gpurun_out/prof_gemm python bench.py --profile --steps 1 > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 12 -c 2 -o gpurun_out/prof_attn_tc python bench.py --profile --steps 1 > gpurun_out/ncu_attn.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:layernorm|attn_fwd_kernel|attn_bwd_kernel|colsum|gather_cast" -s 40 -c 8 -o gpurun_out/prof_misc python bench.py --profile --steps 1 > gpurun_out/ncu_misc.log 2>&1
tail -n 5 gpurun_out/t_all.log; tail -n 2 gpurun_out/bench2.log; ls -la gpurun_out
