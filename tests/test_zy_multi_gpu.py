"""Two-GPU checks (skipped on a one-GPU box): the data-parallel step under torch.distributed.run with every gradient
exchange variant -- copy engines (default, with and without the split optimizer pass) and NCCL -- ends with the same
parameters on every rank and across variants, and every process exits normally (lib/models/build.py:49-53
is the reference's boundary: replicas + gradient averaging)."""
import json
import os
import random
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, steps=3):
    env = dict(os.environ)
    env.update(env_extra)
    port = 29600 + random.randrange(300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "scripts", "dp_step_check.py"), str(steps)]
    r = subprocess.run(cmd, env=env, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-4000:]
    return json.loads(lines[0])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_exchange_variants_agree_and_exit_normally():
    ce = _run({"PVRL_GRAD_EXCHANGE": "ce"})
    ce_unsplit = _run({"PVRL_GRAD_EXCHANGE": "ce", "PVRL_SPLIT_UPDATE": "0"})
    ce_one = _run({"PVRL_GRAD_EXCHANGE": "ce", "PVRL_AR_BLOCKS_PER_BUCKET": "0"})
    nccl = _run({"PVRL_GRAD_EXCHANGE": "nccl"})
    assert ce["exchange"] == "ce" and ce["split_update"] and ce["blocks_per_bucket"] == 1 and ce["mode"] == "graph"
    assert nccl["exchange"] == "nccl" and ce_one["blocks_per_bucket"] == 0
    for run in (ce, ce_unsplit, ce_one, nccl):
        assert run["world"] == 2 and run["ranks_agree"], run
    # two ranks: (a + b) / 2 is the same number whichever engine adds it; what differs from run to run is the order of the
    # split-K fp32 atomics inside the dW GEMMs, i.e. the last bits of the gradients
    for other in (ce_unsplit, ce_one, nccl):
        assert other["losses"] == pytest.approx(ce["losses"], rel=2e-3), (ce, other)
        # (AdamW's first steps move a parameter by ~lr * sign(g): a gradient whose last bits straddle zero flips 2 * lr)
        for k in ("param_abs_sum", "param_sq_sum"):
            assert other[k] == pytest.approx(ce[k], rel=1e-5), (k, ce, other)
        assert other["param_sum"] == pytest.approx(ce["param_sum"], abs=1e-6 * ce["param_abs_sum"]), (ce, other)
