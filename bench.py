#!/usr/bin/env python
"""Headline benchmark: clips/sec of one TimeSformer-B 8x224 HowTo100M stage-2 pretrain step (forward + KL/MSE
loss + backward + gradient all-reduce + AdamW) on N B200s, synthetic clips (BASELINE.json configs[1]/[2]).

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference algorithm (oracle port) on the host CPU cores

Prints ONE JSON line on rank 0 (contract in the task statement): value = whole-job clips/s with inputs resident
in HBM, e2e = same metric with the H2D copy of the frames and the D2H read of the loss inside the timed region,
roofline = achieved TFLOP/s of the tcgen05 GEMM kernel family (CUDA events around every launch in the timed
region) against the measured sustained bf16 peak, cpu_baseline = the oracle on the host cores."""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "clips/sec TimeSformer-B 8x224 pretrain step"
UNIT = "clips/s"
N_STEP_PHRASES, EMB = 9871, 512           # data/clip_step_emb_ht100m_vbphrase.pth is fp32 [9871, 512]


def fwd_flops_per_clip(T, depth=12, D=768, N=196, h=12, d=64):
    """SURVEY.md 8d closed form (2 FLOP/MAC)."""
    L, Ls = N * T, (N + 1) * T
    per_block = (2 * L * D * 3 * D + 2 * L * D * D * 2 + 4 * N * h * T * T * d + 2 * Ls * D * 3 * D + 2 * Ls * D * D
                 + 4 * T * h * (N + 1) ** 2 * d + 4 * (L + 1) * D * 4 * D)
    return depth * per_block + 2 * L * 768 * D


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            j = json.load(f)
        return j.get("bf16_tflops_sustained", 1400.0), j.get("hbm_gbs", 6650.0), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.th.join(timeout=2)
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def pretrain_cfg(T, depth, label_path, precision):
    from procedurevrl_b200.lib.config import get_cfg
    c = get_cfg()
    # configs/HowTo100M/procedurevrl_adamw.yaml, model-relevant keys
    c.merge_from_list(["DEV.ENABLE", True, "DEV.MATCH_LANG_EMB", True, "DEV.ORDER_PRETRAIN_ENABLED", True,
                       "DEV.ORDER_TFM_LAYERS", 4, "TRAIN.DATASET", "howto100m_develop", "TRAIN.BATCH_SIZE", 16,
                       "TRAIN.LABEL_EMB", label_path, "TRAIN.TOPK", 5, "DATA.NUM_FRAMES", T, "DATA.TRAIN_CROP_SIZE", 224,
                       "DATA.TEST_CROP_SIZE", 224, "MODEL.MODEL_NAME", "vit_base_patch16_224_develop",
                       "MODEL.NUM_CLASSES", N_STEP_PHRASES, "MODEL.ARCH", "vit", "MODEL.LOSS_FUNC", "kldiv",
                       "MODEL.TEXT_MODEL", "clip_vit_b_16", "MODEL.PRETRAINED", False, "MODEL.DROP_PATH", 0.1,
                       "TIMESFORMER.DEPTH", depth, "B200.PRECISION", precision])
    return c


def synthetic_label_bank(path):
    """Synthetic stand-in for the shipped CLIP step-phrase bank (same shape / scale: std 0.40)."""
    if not os.path.exists(path):
        g = torch.Generator().manual_seed(1234)
        torch.save(0.4 * torch.randn(N_STEP_PHRASES, EMB, generator=g), path)
    return path


def synthetic_batch(Bv, T, seed, as_u8=False):
    """uint8 frames -> /255 -> (x - 0.45) / 0.225; 0.4 * randn text / CLIP-visual embeddings."""
    g = torch.Generator().manual_seed(seed)
    u8 = torch.randint(0, 256, (Bv, 9, 3, T, 224, 224), generator=g, dtype=torch.uint8)
    frames = u8 if as_u8 else (u8.float() / 255.0 - 0.45) / 0.225
    meta = {"clip_text_emb": 0.4 * torch.randn(Bv * 9, EMB, generator=g),
            "clip_vis_feat": 0.4 * torch.randn(Bv * 9, EMB, generator=g)}
    return frames, meta


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch.distributed as dist
    from procedurevrl_b200 import functional as PF
    from procedurevrl_b200 import ops
    from procedurevrl_b200.lib.models import build_model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"

    T, Bv = args.frames, args.videos_per_gpu
    import tempfile
    bank = synthetic_label_bank(os.path.join(tempfile.gettempdir(), f"pvrl_synthetic_step_bank_{os.getpid()}.pt"))
    torch.manual_seed(0)
    cfg = pretrain_cfg(T, args.depth, bank, args.precision)
    cfg.NUM_GPUS = world if args.ddp else 1        # the flat-gradient trainer does its own all-reduce (no DDP wrapper)
    model = build_model(cfg)
    inner = (model.module if hasattr(model, "module") else model).model
    with torch.no_grad():                       # a fresh reference init has all-zero temporal_fc / time_embed
        for blk in inner.blocks:
            torch.nn.init.trunc_normal_(blk.temporal_fc.weight, std=0.02)
        torch.nn.init.trunc_normal_(inner.time_embed, std=0.02)
    model.train()

    frames_h, meta_h = synthetic_batch(Bv, T, seed=cfg.RNG_SEED + rank, as_u8=args.input_u8)
    frames_pin = frames_h.pin_memory()
    frames = frames_h.to(dev)
    meta = {k: v.to(dev) for k, v in meta_h.items()}
    clips_per_step = Bv * 9 * world

    if args.ddp:        # reference-style driver loop: DistributedDataParallel wrapper + per-op dispatch (train_net.py:146-191)
        params = [p for p in model.parameters() if p.requires_grad]
        opt = torch.optim.AdamW(params, lr=5e-5, weight_decay=1e-4, fused=True)     # procedurevrl_adamw.yaml SOLVER

        def step(fr):
            pred, teacher, mse = model([fr, meta])
            loss, _, _ = PF.pretrain_loss(pred, teacher, mse, topk=cfg.TRAIN.TOPK)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            return loss
        mode = "ddp-eager"
    else:               # flat-gradient step, single NCCL all-reduce, optionally replayed from a CUDA graph
        from procedurevrl_b200.trainer import PretrainStep
        trainer = PretrainStep(model, cfg, lr=5e-5, weight_decay=1e-4,
                               process_group=dist.group.WORLD if world > 1 else None, use_graph=not args.no_graph)
        if not args.no_graph and not args.profile:
            trainer.capture(frames, meta, warmup=2)

        def step(fr):
            return trainer(fr, meta)
        mode = "graph" if trainer.graph is not None else "eager"

    # CUDA events around every GEMM launch (the dominant kernel family) for the roofline
    gemm_events, real_gemm = [], ops.gemm

    def timed_gemm(A, B, out, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = real_gemm(A, B, out, **kw)
        e1.record()
        gemm_events.append((e0, e1, 2.0 * kw["M"] * kw["N"] * kw["K"]))
        return r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(frames)
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    graphed = mode == "graph"
    if not graphed:
        ops.gemm = timed_gemm
    launches0 = ops.launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for _ in range(args.steps):
        loss = step(frames)
    t1.record()
    barrier()
    launches = ops.launch_count() - launches0
    ms = t0.elapsed_time(t1)
    clocks = sampler.stop() if rank == 0 else None
    if graphed:
        # graph replays launch the recorded kernels without going through the C ABI: count them from one eager step,
        # which also carries the CUDA events around every GEMM launch (events cannot time nodes inside a graph)
        ops.gemm = timed_gemm
        launches0 = ops.launch_count()
        te0, te1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # The eager step is host-bound (one Python / ctypes call per launch): on an idle GPU the event recorded BEFORE a
        # launch is stamped ~10 us before the kernel starts, and that host gap would be billed to the kernel.  Park the
        # GPU on a spin kernel first, long enough for the host to enqueue the whole step, so that events and kernels are
        # consumed back to back and an event pair brackets the kernel's execution only.
        torch.cuda._sleep(int(0.12 * 1.9e9))
        te0.record()
        trainer._eager(frames, meta)
        te1.record()
        barrier()
        launches = (ops.launch_count() - launches0) * args.steps
        eager_ms = te0.elapsed_time(te1)
    ops.gemm = real_gemm
    gemm_ms = sum(a.elapsed_time(b) for a, b, _ in gemm_events)
    gemm_flops = sum(f for _, _, f in gemm_events)
    n_gemm = len(gemm_events)

    # end to end: pinned host frames -> device every step, loss read back (a host sync) every step.  As in the reference's
    # loop (non_blocking .cuda() copies of a pinned DataLoader batch, train_net.py:103-107) the H2D copy of step i+1 is
    # issued on a copy stream while step i computes; each step's input still crosses PCIe inside the timed region
    # (exactly `steps` copies), and the step consumes it through a device-to-device copy into the graph's static input.
    barrier()
    n_e2e = 0 if args.profile else args.steps
    copy_stream = torch.cuda.Stream()
    stage = torch.empty_like(frames)
    staged, consumed = torch.cuda.Event(), torch.cuda.Event()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()

    def issue_h2d():
        copy_stream.wait_event(consumed)                     # the previous step has read the staging buffer
        with torch.cuda.stream(copy_stream):
            stage.copy_(frames_pin, non_blocking=True)
            staged.record(copy_stream)

    consumed.record()
    if n_e2e:
        issue_h2d()
    for i in range(n_e2e):
        torch.cuda.current_stream().wait_event(staged)
        frames.copy_(stage, non_blocking=True)
        consumed.record()
        if i + 1 < n_e2e:
            issue_h2d()
        last = step(frames).item()
    if args.profile:
        last = loss.item()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)

    if world > 1:
        t = torch.tensor([ms, e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = t.tolist()
    ms_step = ms / args.steps
    value = clips_per_step / (ms_step / 1e3)
    e2e = clips_per_step / (e2e_ms / args.steps / 1e3) if not args.profile else 0.0

    if rank == 0:
        peak, _, peak_src = measured_peaks()
        achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None
        traffic = None
        tp = os.path.join(ROOT, "profiles", "gemm_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        step_flops = 3 * fwd_flops_per_clip(T, args.depth) * Bv * 9
        out = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "bf16x3", "data": "synthetic",
            "config": {"workload": f"TimeSformer-B {T}x224 HowTo100M stage-2 pretrain step (fwd+KL/MSE loss+bwd+"
                                   f"allreduce+AdamW), {Bv} videos x 9 clips per GPU, DROP_PATH 0.1",
                       "depth": args.depth, "clips_per_gpu": Bv * 9, "parallelism": f"dp{world}", "dispatch": mode,
                       "l2": "per-step working set ~14 GB >> 126 MB L2 (no flush needed)",
                       "loss": round(float(last), 4)},
            "e2e": {"value": round(e2e, 2), "unit": UNIT, "h2d_bytes_per_step": frames_pin.numel() * frames_pin.element_size(),
                    "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05, all epilogues)",
                         "achieved": round(achieved, 1) if achieved else None, "peak": peak, "unit": "TFLOP/s",
                         "frac": round(achieved / peak, 4) if achieved else None, "traffic": traffic,
                         "peak_source": f"bf16_tflops_sustained, {peak_src}",
                         "launches": n_gemm, "gemm_share_of_step": round(gemm_ms / ms_step, 4),
                         "timed_in": ("CUDA events around every GEMM launch of one eager step (GPU parked on a spin kernel while the "
                                      "host enqueues it, so no host gap is billed to a kernel) run after the graph-replayed timed "
                                      "region (events cannot time nodes inside a graph); share = that GEMM time / ms_per_step")
                         if graphed else "timed region",
                         "step_mfu": round(step_flops / (ms_step / 1e3) / 1e12 / peak, 4)},
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(T, args.depth, budget_s=25.0)
        print(json.dumps(out), flush=True)
    if world > 1:
        # Leave without tearing NCCL down: the step's CUDA graph holds the communicator's all-reduce, and
        # ncclCommDestroy behind destroy_process_group() hung the 2-GPU run after the result line was printed.
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


# ------------------------------------------------------------------------------------------------ CPU arm (oracle)
def _oracle_step_fn(T, depth, n_clips, full_pretrain):
    """One optimisation step of the reference algorithm on the host cores (oracle/timesformer_oracle.py)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import timesformer_oracle as O
    p = O.seeded_state(depth=depth, frames=T, seed=0, with_order=True)
    for v in p.values():
        v.requires_grad_(True)
    opt = torch.optim.AdamW(list(p.values()), lr=5e-5, weight_decay=1e-4)
    g = torch.Generator().manual_seed(0)
    label = torch.nn.functional.normalize(0.4 * torch.randn(N_STEP_PHRASES, EMB, generator=g), dim=1)
    if full_pretrain:
        Bv = n_clips // 9
        frames = O.synthetic_clips(Bv, 9, 3, T, 224, 224, seed=1)
        text, vis = 0.4 * torch.randn(Bv * 9, EMB, generator=g), 0.4 * torch.randn(Bv * 9, EMB, generator=g)
        draws = O.synthetic_draws(Bv)

        def step():
            pred, teach, mse = O.pretrain_forward(p, frames, text, vis, label, draws, depth=depth)
            loss, _, _ = O.pretrain_loss(pred, teach, mse)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
        return step
    frames = O.synthetic_clips(n_clips, 3, T, 224, 224, seed=1)
    teach = O.pseudo_labels(0.4 * torch.randn(n_clips, EMB, generator=g), 0.4 * torch.randn(n_clips, EMB, generator=g), label)

    def step():
        logits = O.match_lang_forward(p, frames, label, depth=depth)
        loss, _, _ = O.pretrain_loss(logits, teach, [torch.zeros(1), torch.zeros(1)])
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
    return step


def cpu_baseline(T, depth, budget_s):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = _oracle_step_fn(T, depth, 1, False)
    t = time.perf_counter()
    step()                                   # warm-up / calibration on one clip
    t1 = time.perf_counter() - t
    n = int(max(1, min(9, budget_s / max(t1, 1e-3) / 2)))
    step = _oracle_step_fn(T, depth, n, False)
    t = time.perf_counter()
    step()
    dt = time.perf_counter() - t
    return {"value": round(n / dt, 4), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle (torch fp32 CPU restatement of the reference) fwd+KL loss+bwd+AdamW on {n} clip(s) "
                      f"of the same {T}x224 workload, 1 timed step after a 1-clip warm-up"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    T, depth = args.frames, args.depth
    cal = _oracle_step_fn(T, depth, 1, False)
    t = time.perf_counter()
    cal()
    t1 = time.perf_counter() - t
    total = args.steps + args.warmup
    full = 9 * t1 * total <= 240.0
    n = 9 if full else int(max(1, min(9, 240.0 / total / max(t1, 1e-3))))
    step = _oracle_step_fn(T, depth, n, full)
    for _ in range(args.warmup):
        step()
    t = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t
    value = n * args.steps / dt
    sample = (f"1 video x 9 clips full pretrain step (oracle.pretrain_forward + loss + backward + AdamW)" if full else
              f"{n} clip(s): encoder + head + KL top-k loss + backward + AdamW (order transformer omitted to bound time)")
    out = {"impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 1),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"TimeSformer-B {T}x224 HowTo100M stage-2 pretrain step on the host CPU", "depth": depth},
           "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--depth", type=int, default=12)
    ap.add_argument("--videos-per-gpu", type=int, default=2)        # TRAIN.BATCH_SIZE 16 / NUM_GPUS 8
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="dispatch every kernel from Python instead of replaying a CUDA graph")
    ap.add_argument("--input-u8", action="store_true",
                    help="feed uint8 frames (normalisation fused into the patch im2col kernel): 4x less H2D traffic in e2e")
    ap.add_argument("--ddp", action="store_true", help="reference-style DistributedDataParallel wrapper (eager)")
    ap.add_argument("--profile", action="store_true",
                    help="short run for ncu: 1 warm-up + --steps timed steps, no e2e / cpu legs (never a bench value)")
    args = ap.parse_args()
    if args.impl == "b200":
        args.warmup = 1 if args.profile else max(args.warmup, 3)
    if args.profile:
        args.no_cpu_baseline = True
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
