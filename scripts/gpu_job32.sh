set -x
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "patchify" 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --input-u8 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('u8', j['value'], j['ms_per_step'], j['e2e'], j['config']['loss'])"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('f32', j['value'], j['ms_per_step'], j['e2e'], j['config']['loss'])"
