# What the round-end driver runs on a B200 box, plus the profiling captures kept under profiles/ (see profiles/README.md).
#   gpurun --timeout 2400 -- 'bash scripts/gpu_check.sh'
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/bench.log 2>&1; tail -n 1 gpurun_out/bench.log
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2>&1; tail -n 1 gpurun_out/bench_ref.log
# launch list of the steady-state step (durations only; absolute times under ncu are serialised, the SHARES are what count)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches.csv python bench.py --profile --steps 1 > gpurun_out/ncu_launch.log 2>&1
# full-set captures inside the real step: GEMMs; the HBM-bound kernels added last (LayerNorm backward + emit, flat AdamW,
# weight cast); the 9..32-frame temporal attention at T = 32
timeout 600 ncu --set full --clock-control none -k regex:gemm -s 300 -c 14 --csv --page raw --log-file gpurun_out/prof_gemm_raw.csv python bench.py --profile --steps 1 > gpurun_out/ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none -k "regex:layernorm_bwd|adam_flat|cast_weight_multi" -s 38 -c 6 --csv --page raw --log-file gpurun_out/prof_hbm_raw.csv python bench.py --profile --steps 1 > gpurun_out/ncu_hbm.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:attn_t32 -s 10 -c 4 --csv --page raw --log-file gpurun_out/prof_t32_raw.csv python bench.py --frames 32 --profile --steps 1 > gpurun_out/ncu_t32.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches.csv | head -30
python scripts/summarize_ncu_raw.py gpurun_out/prof_gemm_raw.csv | cut -c1-200
python scripts/summarize_ncu_raw.py gpurun_out/prof_hbm_raw.csv | cut -c1-200
python scripts/summarize_ncu_raw.py gpurun_out/prof_t32_raw.csv | cut -c1-200
# MViTv2-S 16x224 (config 5): 9-clip forward + backward (+ flat AdamW), launch list of one step (give ncu ~3 min: ~1500 launches
# per step incl. torch glue), and the mma.sync backward's first hardware run with its per-launch times
timeout 120 python scripts/mvit_bench.py --steps 3 --warmup 2 > gpurun_out/mvit_bench.log 2>&1; tail -n 1 gpurun_out/mvit_bench.log
timeout 120 python scripts/mvit_bench.py --steps 3 --warmup 2 --optimizer > gpurun_out/mvit_bench_opt.log 2>&1; tail -n 1 gpurun_out/mvit_bench_opt.log
PVRL_MVIT_ATTN_MMA_BWD=1 timeout 120 python scripts/mvit_bench.py --steps 3 --warmup 2 > gpurun_out/mvit_bench_mma_bwd.log 2>&1; tail -n 1 gpurun_out/mvit_bench_mma_bwd.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_mvit.csv python scripts/mvit_bench.py --steps 1 --warmup 0 > gpurun_out/ncu_mvit.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_mvit.csv | head -30
timeout 300 python -m pytest tests/test_zz_mvit_mma_bwd_gpu.py -m gpu -q -s -rxX 2>&1 | grep "mma bwd\|passed\|failed\|xfail\|xpass"
timeout 200 python scripts/mvit_attn_bench.py > gpurun_out/mvit_attn_bench.log 2>&1; cat gpurun_out/mvit_attn_bench.log
