"""N > 1 path on CPU: two gloo ranks, each with a replica of the model mirror behind the data-parallel wrapper of
lib/models/build.py, shard a batch of clips; the all-reduced gradients must equal the single-process gradients
of the whole batch.  The C-ABI ops are replaced by their torch restatements (tests/shadow_ops.py) -- this checks
the host logic of the data-parallel boundary (one autograd node producing every encoder gradient, bucketed
all-reduce, no unused-parameter walk), not the kernels."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _small_cfg(bank):
    from procedurevrl_b200.lib.config import get_cfg
    c = get_cfg()
    c.merge_from_list(["DEV.MATCH_LANG_EMB", True, "DEV.TEST_LANG_EMB", bank, "MODEL.MODEL_NAME",
                       "vit_base_patch16_224_develop", "MODEL.PRETRAINED", False, "MODEL.DROP_PATH", 0.0,
                       "TIMESFORMER.DEPTH", 1, "MODEL.NUM_CLASSES", 778, "DATA.NUM_FRAMES", 2, "DATA.TRAIN_CROP_SIZE", 32,
                       "B200.PRECISION", "bf16x3", "B200.GRAD_BUCKET_MB", 1, "NUM_GPUS", 2])
    return c


def _setup():
    for p in (ROOT, HERE, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import shadow_ops
    from procedurevrl_b200 import ops
    from procedurevrl_b200.lib.models.vit import VisionTransformer
    for n in shadow_ops.ALL:
        setattr(ops, n, getattr(shadow_ops, n))
    VisionTransformer._require_cuda = False


def _model(bank):
    from procedurevrl_b200.lib.models import MODEL_REGISTRY
    torch.manual_seed(3)
    m = MODEL_REGISTRY.get("vit_base_patch16_224_develop")(_small_cfg(bank))
    with torch.no_grad():
        for blk in m.model.blocks:
            torch.nn.init.normal_(blk.temporal_fc.weight, std=0.02)
        torch.nn.init.normal_(m.model.time_embed, std=0.02)
    for p in m.parameters():
        p.requires_grad_(True)
    return m.train()


def _batch():
    g = torch.Generator().manual_seed(5)
    return torch.randn(4, 3, 2, 32, 32, generator=g), torch.tensor([1, 50, 300, 700])


def _worker(rank, world, port, bank, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    _setup()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from procedurevrl_b200.lib.models.build import wrap_data_parallel
    m = wrap_data_parallel(_model(bank), _small_cfg(bank))
    x, y = _batch()
    shard = slice(rank * 2, rank * 2 + 2)
    torch.nn.functional.cross_entropy(m(x[shard]), y[shard]).backward()
    if rank == 0:
        torch.save({k: p.grad.clone() for k, p in m.module.named_parameters()}, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradients_match_full_batch(gold_dir, tmp_path):
    bank = os.path.join(gold_dir, "clip_step_emb_coin.pt")
    out_path = str(tmp_path / "grads.pt")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, bank, out_path), nprocs=2, join=True)
    ddp = torch.load(out_path)
    _setup()
    try:
        m = _model(bank)
        x, y = _batch()
        torch.nn.functional.cross_entropy(m(x), y).backward()
        ref = {k: p.grad for k, p in m.named_parameters()}
        assert set(ref) == set(ddp)
        for k in ref:
            # bf16x3 products are exact to ~2^-17; shard-vs-full differs only by summation order
            assert (ddp[k] - ref[k]).abs().max().item() <= 1e-4 * ref[k].abs().max().item() + 1e-7, k
    finally:
        import importlib
        from procedurevrl_b200 import ops
        from procedurevrl_b200.lib.models.vit import VisionTransformer
        importlib.reload(ops)
        VisionTransformer._require_cuda = True


def test_allreduce_bucket_ranges_tile_the_flat_buffer():
    """Host logic of the overlapped gradient exchange (trainer.bucket_ranges): block buckets + front + tail cover every
    element of the flat gradient buffer exactly once, for every bucket size."""
    from procedurevrl_b200.trainer import bucket_ranges
    names = [("model.cls_token", 768), ("model.pos_embed", 197 * 768), ("model.patch_embed.proj.weight", 768 * 768)]
    for i in range(12):
        names += [(f"model.blocks.{i}.norm1.weight", 768), (f"model.blocks.{i}.attn.qkv.weight", 2304 * 768),
                  (f"model.blocks.{i}.mlp.fc2.bias", 768)]
    names += [("model.norm.weight", 768), ("model.head.weight", 512 * 768),
              ("model.order_tfm.temporalModelling.resblocks.0.ln_1.weight", 512)]
    total = sum(s for _, s in names)
    for bpb in (1, 2, 3, 4, 5, 12, 20):
        ranges, front_end, tail_start = bucket_ranges(names, 12, bpb)
        cover = [(0, front_end)] + sorted(ranges.values()) + [(tail_start, total)]
        assert cover[0][0] == 0 and cover[-1][1] == total
        assert all(cover[k][1] == cover[k + 1][0] for k in range(len(cover) - 1)), (bpb, cover)
        assert all(lo > 0 for lo in ranges)          # block 0's bucket always rides with the final exchange


def _linear_worker(rank, world, port, bank, out_path):
    """the reference's fine-tuning flow under DDP -- build_model wraps FIRST, construct_optimizer freezes the
    encoder for TRAIN.LINEAR afterwards (optimizer.py:25-31), and with MODEL.NUM_SEG > 0 `order_tfm.pad_embedding` never
    receives a gradient.  With find_unused_parameters=False the second iteration dies in 'Expected to have finished
    reduction'; with the reference's True (build.py:49-53) it trains."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    _setup()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from procedurevrl_b200.lib.models import MODEL_REGISTRY
    from procedurevrl_b200.lib.models.build import wrap_data_parallel
    from procedurevrl_b200.lib.models.optimizer import parameter_groups
    cfg = _small_cfg(bank)
    cfg.merge_from_list(["DEV.MATCH_LANG_EMB", False, "MODEL.NUM_SEG", 2, "TRAIN.LINEAR", True, "MODEL.DROP_E", 0.0])
    torch.manual_seed(3)
    m = MODEL_REGISTRY.get("vit_base_patch16_224_develop")(cfg).train()
    ddp = wrap_data_parallel(m, cfg)
    groups = parameter_groups(ddp, cfg)                      # freezes the encoder AFTER the wrap, as train_net.py does
    opt = torch.optim.SGD([{"params": [p for p in g["params"] if p.requires_grad], "lr": 0.1} for g in groups])
    g = torch.Generator().manual_seed(5 + rank)
    losses = []
    for it in range(3):
        x = torch.randn(2, 3, 2 * 2, 32, 32, generator=g)    # [B, 3, NUM_SEG * T, H, W]
        y = torch.tensor([1 + it, 50 + rank])
        loss = torch.nn.functional.cross_entropy(ddp(x), y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    if rank == 0:
        torch.save({"losses": losses, "pad_grad_none": m.model.order_tfm.pad_embedding.weight.grad is None,
                    "enc_frozen": not m.model.blocks[0].attn.qkv.weight.requires_grad,
                    "head_cls": m.model.head_cls.weight.detach().clone()}, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_linear_finetune_with_forecast_under_ddp(gold_dir, tmp_path):
    bank = os.path.join(gold_dir, "clip_step_emb_coin.pt")
    out_path = str(tmp_path / "linear.pt")
    port = 31500 + os.getpid() % 2000
    mp.spawn(_linear_worker, args=(2, port, bank, out_path), nprocs=2, join=True)
    r = torch.load(out_path)
    assert len(r["losses"]) == 3 and all(l == l for l in r["losses"])
    assert r["enc_frozen"]
    try:
        pass
    finally:
        import importlib
        from procedurevrl_b200 import ops
        from procedurevrl_b200.lib.models.vit import VisionTransformer
        importlib.reload(ops)
        VisionTransformer._require_cuda = True


def test_peer_exchange_chunks_tile_every_range():
    """Host logic of the copy-engine gradient exchange (grad_exchange.PeerGradExchange.chunk_bounds): the per-rank chunks of
    any range are disjoint, ordered, 16-byte granular and cover it exactly, for every world size."""
    from procedurevrl_b200.grad_exchange import PeerGradExchange
    for world in (2, 3, 4, 8):
        for a, b in ((0, 134_600_000), (12345, 13_000_001), (5, 8), (0, 1), (7, 7 + 4 * world), (100, 100)):
            ch = PeerGradExchange.chunk_bounds(a, b, world)
            assert len(ch) == world and ch[0][0] == a and ch[-1][1] == b
            assert all(ch[i][1] == ch[i + 1][0] for i in range(world - 1))
            assert all(0 <= e - s for s, e in ch) and all((s - a) % 4 == 0 for s, e in ch if e > s)
