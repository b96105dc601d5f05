"""autograd.Function wrappers of the head / similarity / loss kernels (include/pvrl.h, head_loss.cu).
They own save-for-backward; all arithmetic happens in the C-ABI library."""
import torch

from . import ops


def _f32(t):
    return t.contiguous().float() if (t.dtype != torch.float32 or not t.is_contiguous()) else t


def _few_rows_ok(x, w):
    """shapes the `pvrl_ot_linear_*` few-row kernels take (they run the 18-26 row head ~10x faster than the generic
    small-linear kernels): contraction a multiple of 128 up to 2048, output width a multiple of 4."""
    return x.shape[1] % 128 == 0 and x.shape[1] <= 2048 and w.shape[0] % 4 == 0


class _LinearSmall(torch.autograd.Function):
    """y = x @ w.t() + b for a handful of rows: `head` 768->512 (vit.py:301), `head_cls` (vit.py:321)."""

    @staticmethod
    def forward(ctx, x, w, b):
        x, w = _f32(x), _f32(w)
        y = torch.empty(x.shape[0], w.shape[0], device=x.device, dtype=torch.float32)
        if _few_rows_ok(x, w):
            ops.ot_linear_fwd(x, w, None if b is None else _f32(b), y)
        else:
            ops.linear_small_fwd(x, w, b, y)
        ctx.save_for_backward(x, w)
        ctx.has_b = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = _f32(dy)
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dx = torch.empty_like(x) if need_x else None
        dw = torch.zeros_like(w) if need_w else None
        db = torch.zeros(w.shape[0], device=w.device) if (need_w and ctx.has_b) else None
        if _few_rows_ok(x, w):
            if need_x:
                ops.ot_linear_dx(dy, w, dx)
            if need_w:
                ops.ot_linear_dw(dy, x, dw, db)
        else:
            ops.linear_small_bwd(x, w, dy, dx, dw, db)
        if not (ctx.has_b and ctx.needs_input_grad[2]):
            db = None
        return dx, dw, db


class _L2Norm(torch.autograd.Function):
    """x / x.norm(dim=1, keepdim=True)  (vit.py:302,306,311,316,333,340,431)."""

    @staticmethod
    def forward(ctx, x):
        x = _f32(x)
        y, n = torch.empty_like(x), torch.empty(x.shape[0], device=x.device, dtype=torch.float32)
        ops.l2norm_fwd(x, y, n)
        ctx.save_for_backward(y, n)
        return y

    @staticmethod
    def backward(ctx, dy):
        y, n = ctx.saved_tensors
        return ops.l2norm_bwd(y, n, _f32(dy), torch.empty_like(y))


class _SimLogits(torch.autograd.Function):
    """emb @ label_emb.t() / temp  (vit.py:307,334,341,432); label_emb is a constant bank."""

    @staticmethod
    def forward(ctx, emb, label, inv_temp):
        emb = _f32(emb)
        out = torch.empty(emb.shape[0], label.shape[0], device=emb.device, dtype=torch.float32)
        ops.sim_logits_fwd(emb, label, out, inv_temp)
        ctx.label, ctx.inv_temp, ctx.shape = label, inv_temp, emb.shape
        return out

    @staticmethod
    def backward(ctx, dlogits):
        demb = torch.zeros(ctx.shape, device=dlogits.device, dtype=torch.float32)
        ops.sim_logits_bwd(_f32(dlogits), ctx.label, demb, ctx.inv_temp)
        return demb, None, None


class _KLTopKLoss(torch.autograd.Function):
    """train_net.py:153-162: KLDivLoss(batchmean)(log_softmax(pred), renorm(top-k(softmax(teacher))))."""

    @staticmethod
    def forward(ctx, pred, teacher_logits, topk):
        pred, teacher_logits = _f32(pred), _f32(teacher_logits.detach())
        row = torch.empty(pred.shape[0], device=pred.device, dtype=torch.float32)
        dpred = torch.empty_like(pred)
        ops.kl_topk_loss(pred, teacher_logits, row, dpred, None, topk, 1.0)
        ctx.save_for_backward(dpred)
        return row.sum()

    @staticmethod
    def backward(ctx, g):
        (dpred,) = ctx.saved_tensors
        return dpred * g, None, None


def linear_small(x, w, b=None):
    return _LinearSmall.apply(x, w, b)


def l2_normalize(x):
    return _L2Norm.apply(x)


def similarity_logits(emb, label_emb_n, temp):
    return _SimLogits.apply(emb, label_emb_n, 1.0 / temp)


def kl_topk_loss(pred, teacher_logits, topk=5):
    return _KLTopKLoss.apply(pred, teacher_logits, topk)


def softmax_rows(x):
    x = _f32(x)
    return ops.softmax_rows(x, torch.empty_like(x))


def pretrain_loss(pred, teacher_logits, mse_pair, topk=5):
    """loss of tools/train_net.py:152-162 for the order-pretraining branch: KL(top-k teacher) + MSE."""
    loss1 = kl_topk_loss(pred, teacher_logits, topk)
    loss2 = torch.nn.functional.mse_loss(mse_pair[0], mse_pair[1], reduction="mean")
    return loss1 + loss2, loss1, loss2
