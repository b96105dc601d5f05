set -x
mkdir -p gpurun_out
for b in 3 0; do
PVRL_AR_BLOCKS_PER_BUCKET=$b timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$b bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_b$b.log 2>&1
echo "exit code $?"
grep '"metric"' gpurun_out/bench_n2_b$b.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('buckets $b', j['value'], j['ms_per_step'], j['config'].get('loss'), j['clocks'])"
done
