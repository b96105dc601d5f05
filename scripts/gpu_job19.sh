set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches19.csv python bench.py --profile --steps 1 > gpurun_out/ncu_launch19.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm -s 300 -c 14 -o gpurun_out/prof_gemm19 python bench.py --profile --steps 1 > gpurun_out/ncu_gemm19.log 2>&1
ncu -i gpurun_out/prof_gemm19.ncu-rep --page raw --csv > gpurun_out/prof_gemm19_raw.csv 2>/dev/null
timeout 400 ncu --set full --clock-control none -k "regex:attn_t8|layernorm|gather_cast|attn_tc" -s 20 -c 10 --csv --page raw --log-file gpurun_out/prof_misc19_raw.csv python bench.py --profile --steps 1 > gpurun_out/ncu_misc19.log 2>&1
ls -la gpurun_out | tail -8
