"""First measurement of the MViTv2-S 16 x 224 path (BASELINE config 5): forward + loss + backward of one video (9 clips)
per step on one GPU, eager dispatch (no CUDA graph yet), CUDA-event timing.  NOT the contract bench line (bench.py times
config 2); the JSON line this prints is kept under profiles/ as the starting point for the tcgen05 attention work.

    python scripts/mvit_bench.py --steps 3 --warmup 2            # timing
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_mvit.csv \
        python scripts/mvit_bench.py --steps 1 --warmup 1        # launch list (shares only)
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from procedurevrl_b200 import ops  # noqa: E402
from procedurevrl_b200.lib.config import get_cfg  # noqa: E402
from procedurevrl_b200.lib.models import MODEL_REGISTRY  # noqa: E402

TRAIN_FLOPS_PER_CLIP = 3 * 128.45e9          # MViTv2-S 16 x 224 forward = 128.45 GFLOP, step = 3x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--clips", type=int, default=9)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3"])
    ap.add_argument("--optimizer", action="store_true", help="include the flat AdamW step (construct_optimizer) in the timed step")
    a = ap.parse_args()
    with open(os.path.join(ROOT, "tests", "golden", "mvit_full_geometry.json")) as f:
        fg = json.load(f)
    c = get_cfg()
    c.merge_from_list(["DEV.MATCH_LANG_EMB", True, "DEV.TEST_LANG_EMB", os.path.join(ROOT, "tests", "golden", "clip_step_emb_coin.pt"),
                       "TRAIN.LABEL_EMB", "", "MODEL.TEXT_MODEL", "", "MODEL.MODEL_NAME", "MViT", "MODEL.NUM_CLASSES", 778,
                       "MODEL.PRETRAINED", False, "DATA.NUM_FRAMES", 16, "DATA.TRAIN_CROP_SIZE", 224, "DATA.INPUT_CHANNEL_NUM", [3],
                       "B200.PRECISION", a.precision])
    for k, v in fg["mvit"].items():
        c.MVIT[k] = v
    torch.manual_seed(0)
    m = MODEL_REGISTRY.get("MViT")(c).cuda().train()
    for p in m.parameters():
        p.requires_grad_(True)
    g = torch.Generator().manual_seed(1)
    u8 = torch.randint(0, 256, (a.clips, 3, 16, 224, 224), generator=g, dtype=torch.uint8)
    x = ((u8.float() / 255.0 - 0.45) / 0.225).cuda()
    labels = torch.arange(a.clips, device="cuda") % 778

    opt = None
    if a.optimizer:
        from procedurevrl_b200.lib.models import optimizer as optim
        c.merge_from_list(["SOLVER.OPTIMIZING_METHOD", "adamw", "SOLVER.BASE_LR", 5e-5, "SOLVER.WEIGHT_DECAY", 0.05])
        opt = optim.construct_optimizer(m, c)

    def step():
        if opt is None:
            for p in m.parameters():
                p.grad = None
        else:
            opt.zero_grad()
        loss = torch.nn.functional.cross_entropy(m(x), labels)
        loss.backward()
        if opt is not None:
            opt.step()
        return loss

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    n0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    cps = a.clips / ms * 1e3
    print(json.dumps({"metric": "clips/sec MViTv2-S 16x224 forward+backward" + (" + flat AdamW step" if opt is not None else " (no optimizer step)") + ", 1 GPU", "value": round(cps, 2),
                      "unit": "clips/s", "ms_per_step": round(ms, 2), "steps": a.steps, "warmup": a.warmup, "n_gpus": 1,
                      "dtype": a.precision, "data": "synthetic", "loss": round(loss.item(), 4),
                      "gpu_launches_per_step": (ops.launch_count() - n0) // a.steps,
                      "achieved_tflops": round(cps * TRAIN_FLOPS_PER_CLIP / 1e12, 1),
                      "config": {"workload": "MViTv2-S 16x224, 9 clips (1 video) per step, MATCH_LANG_EMB head, eager dispatch",
                                 "clips": a.clips}, "peaks_file": bool(peaks)}))


if __name__ == "__main__":
    main()
