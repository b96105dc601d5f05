set -x
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_final.log 2>&1
echo "exit code $?"
grep '"metric"' gpurun_out/bench_n2_final.log | cut -c1-330
