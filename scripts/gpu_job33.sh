set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for p in 1 0 1 0; do PVRL_PDL=$p timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('PDL $p', j['value'], j['ms_per_step'], j['config']['loss'], j['clocks']['sm_mhz'])"; done
