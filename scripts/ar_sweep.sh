# Gradient-exchange variants at N GPUs (run under gpurun --gpus N):  bash scripts/ar_sweep.sh N
# default = one all-reduce of the flat buffer after the backward; PVRL_AR_BLOCKS_PER_BUCKET=n = buckets of n encoder blocks
# exchanged on a side stream during the backward; NCCL_MAX_CTAS caps the SMs NCCL takes from the persistent GEMMs.
N=${1:-2}
run() {
  echo "== $*"
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 20 --warmup 5 --no-extras 2>&1 | grep -E '^\{|rror|did not return' | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('   value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'clk', d['clocks']['sm_mhz'])
    else: print('   ', l.strip()[:200])
"
}
run PVRL_AR_BLOCKS_PER_BUCKET=0
run PVRL_AR_BLOCKS_PER_BUCKET=0 NCCL_MAX_CTAS=8
run PVRL_AR_BLOCKS_PER_BUCKET=4 NCCL_MAX_CTAS=8
run PVRL_AR_BLOCKS_PER_BUCKET=4 NCCL_MAX_CTAS=4
run PVRL_AR_BLOCKS_PER_BUCKET=2 NCCL_MAX_CTAS=4
run PVRL_AR_BLOCKS_PER_BUCKET=6 NCCL_MAX_CTAS=16
