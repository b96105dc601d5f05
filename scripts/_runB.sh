cd /root/repo
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "gemm" 2>&1 | tail -3
for ew in 16 8; do echo "== PVRL_GEMM2_EW=$ew"; PVRL_GEMM2_EW=$ew timeout 200 python scripts/op_bench.py --only "GELU" --cold --iters 20 2>&1 | grep -v Warn; done
for lib in new old new old; do echo "== bench $lib"; if [ $lib = old ]; then export PVRL_LIB=/root/repo/scripts/micro/lib_old.so; else unset PVRL_LIB; fi; timeout 200 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline 2>&1 | grep "^{" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'], d['config']['loss'])
"; done
