set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for m in 2 1 0; do PVRL_GEMM_2CTA=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('mode $m', j['value'], j['ms_per_step'], j['roofline']['achieved'], j['clocks'])"; done
