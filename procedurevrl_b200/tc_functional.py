"""nn.Linear / Mlp as autograd Functions over the tcgen05 GEMM of the C ABI (`pvrl_gemm_bf16`), for host code that is
scheduled by autograd instead of by an explicit engine -- the MViTv2 encoder mirror (lib/models/mvit.py; reference
lib/models/slowfast_mvit/attention.py:223-230,386-391, common.py:7-34), whose Linear layers are 75 % of its FLOPs with
widths 96 ... 3072 and row counts 393 ... 25 089 per clip.

  tc_linear(x, weight, bias)            y = x W^T + b          dX = dY W,  dW += dY^T X (split-K, fp32 atomics),  db = colsum dY
  tc_mlp(x, w1, b1, w2, b2)             y = gelu_erf(x W1^T + b1) W2^T + b2 with the GELU applied in fc1's epilogue, gelu'
                                        saved (bf16) and multiplied in the epilogue of fc2's dX GEMM (EPI_GELU / EPI_DGELU)

precision "bf16": operands rounded to bf16, fp32 accumulate, bf16 activations out (throughput mode);
precision "bf16x3": the error-compensated three-term split on the same kernel, fp32 activations (parity mode).
Weights keep their fp32 masters; bf16 operand copies ([N, K] for the forward, [K, N] for dX) are cached per parameter and
re-cast when the parameter's version or storage changes.  There is no fallback: every path ends in `ops.gemm`."""
import weakref

import torch

from . import ops

_WCACHE = {}      # id(weight) -> (weakref to the weight, version key, W operand [N, K'], W^T operand [K, N'])


def _weight_operands(w, x3):
    """bf16 operand copies of an fp32 master: W [N, K] for the forward, W^T [K, N] for dX (three-term splits in the parity
    mode).  Cached per parameter OBJECT: `id()` alone is not an identity -- CPython hands a dead tensor's id to the next one
    and the caching allocator hands out the same address -- so an entry also holds a weak reference that must still point
    at `w`, and the key carries the shape."""
    key = (w._version, w.data_ptr(), tuple(w.shape), x3)
    hit = _WCACHE.get(id(w))
    if hit is not None and hit[0]() is not w:
        hit = None                                  # the id was recycled: a different tensor lives there now
    if hit is not None and hit[1] == key:
        return hit[2], hit[3]
    w2 = w.detach().reshape(w.shape[0], -1).contiguous()
    N, K = w2.shape
    mul = 3 if x3 else 1
    if hit is not None and hit[2].shape == (N, mul * K) and hit[2].device == w2.device:
        wb, wt = hit[2], hit[3]                     # refresh in place: operand addresses stay stable (CUDA graphs)
    else:
        wb = torch.empty(N, mul * K, device=w2.device, dtype=torch.bfloat16)
        wt = torch.empty(K, mul * N, device=w2.device, dtype=torch.bfloat16)
    if not x3:
        ops.cast_weight(w2, wb, wt)
    else:
        wt32 = torch.empty(K, N, device=w2.device, dtype=torch.float32)
        ops.cast_weight(w2, None, wt32)
        ops.split3(w2, wb, N, K, 1, 1)
        ops.split3(wt32, wt, K, N, 1, 1)
    if len(_WCACHE) > 4096:                         # entries of dead parameters (their weak references are gone)
        for k in [k for k, v in _WCACHE.items() if v[0]() is None]:
            del _WCACHE[k]
    _WCACHE[id(w)] = (weakref.ref(w), key, wb, wt)
    return wb, wt


def _a_operand(a, M, K, x3):
    """activation [M, K] -> GEMM A operand: bf16 as is, or the [hi | hi | lo] split of an fp32 activation."""
    if not x3:
        return (a if a.dtype == torch.bfloat16 else a.to(torch.bfloat16)).contiguous(), K
    a3 = torch.empty(M, 3 * K, device=a.device, dtype=torch.bfloat16)
    ops.split3(a.float().contiguous(), a3, M, K, 0, 1)
    return a3, 3 * K


def _dw(dy, x, gw, M, N, K, x3):
    """gw[N, K] += dy[M, N]^T x[M, K] (TN GEMM, split-K partial sums meet in fp32 atomics)."""
    if not x3:
        ops.gemm(dy.contiguous(), x.contiguous(), gw, M=N, N=K, K=M, trans=1, epilogue=ops.EPI_ATOMIC, ldo=K)
    else:
        a3 = torch.empty(3 * M, N, device=dy.device, dtype=torch.bfloat16)
        b3 = torch.empty(3 * M, K, device=dy.device, dtype=torch.bfloat16)
        ops.split3(dy.float().contiguous(), a3, M, N, 0, 0)
        ops.split3(x.float().contiguous(), b3, M, K, 1, 0)
        ops.gemm(a3, b3, gw, M=N, N=K, K=3 * M, trans=1, epilogue=ops.EPI_ATOMIC, ldo=K)


def _check_dims(K, N):
    if K % 8 or N % 8:
        raise ValueError(f"tc_linear: in / out features must be multiples of 8 for the TMA operand maps (got {K} -> {N})")


class _TCLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, x3):
        K, N = x.shape[-1], w.shape[0]
        _check_dims(K, N)
        lead = x.shape[:-1]
        x2 = x.reshape(-1, K)
        M = x2.shape[0]
        act = torch.float32 if x3 else torch.bfloat16
        A, Keff = _a_operand(x2, M, K, x3)
        wb, _ = _weight_operands(w, x3)
        y = torch.empty(M, N, device=x.device, dtype=act)
        ops.gemm(A, wb, y, M=M, N=N, K=Keff, bias=None if b is None else b.detach().float().contiguous())
        ctx.save_for_backward(x2 if x3 else A, w)
        ctx.x3, ctx.has_b, ctx.lead, ctx.in_dtype = x3, b is not None, lead, x.dtype
        return y.reshape(*lead, N)

    @staticmethod
    def backward(ctx, dy):
        xs, w = ctx.saved_tensors
        x3 = ctx.x3
        N, K = w.shape[0], xs.shape[1]
        dy2 = dy.reshape(-1, N)
        dy2 = (dy2.float() if x3 else dy2.to(torch.bfloat16)).contiguous()
        M = dy2.shape[0]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            A, Keff = _a_operand(dy2, M, N, x3)
            _, wt = _weight_operands(w, x3)
            dx2 = torch.empty(M, K, device=dy.device, dtype=torch.float32 if x3 else torch.bfloat16)
            ops.gemm(A, wt, dx2, M=M, N=K, K=Keff)
            dx = dx2.reshape(*ctx.lead, K).to(ctx.in_dtype)
        if ctx.needs_input_grad[1]:
            dw = torch.zeros(N, K, device=dy.device, dtype=torch.float32)
            _dw(dy2, xs, dw, M, N, K, x3)
            dw = dw.view_as(w)
        if ctx.has_b and ctx.needs_input_grad[2]:
            db = torch.zeros(N, device=dy.device, dtype=torch.float32)
            ops.colsum(dy2, db, M, N)
        return dx, dw, db, None


class _TCMlp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, x3):
        K, Hd, N = x.shape[-1], w1.shape[0], w2.shape[0]
        _check_dims(K, Hd)
        _check_dims(Hd, N)
        lead = x.shape[:-1]
        x2 = x.reshape(-1, K)
        M = x2.shape[0]
        act = torch.float32 if x3 else torch.bfloat16
        A, Keff = _a_operand(x2, M, K, x3)
        h = torch.empty(M, Hd, device=x.device, dtype=act)            # gelu(z): operand of fc2 and of its dW
        dact = torch.empty(M, Hd, device=x.device, dtype=act)         # gelu'(z): multiplied in fc2's dX epilogue
        ops.gemm(A, _weight_operands(w1, x3)[0], dact, M=M, N=Hd, K=Keff, bias=b1.detach().float().contiguous(),
                 epilogue=ops.EPI_GELU, out2=h)
        H, Heff = _a_operand(h, M, Hd, x3)
        y = torch.empty(M, N, device=x.device, dtype=act)
        ops.gemm(H, _weight_operands(w2, x3)[0], y, M=M, N=N, K=Heff, bias=b2.detach().float().contiguous())
        ctx.save_for_backward(x2 if x3 else A, h if x3 else H, dact, w1, w2)
        ctx.x3, ctx.lead, ctx.in_dtype = x3, lead, x.dtype
        return y.reshape(*lead, N)

    @staticmethod
    def backward(ctx, dy):
        xs, hs, dact, w1, w2 = ctx.saved_tensors
        x3 = ctx.x3
        K, Hd, N = xs.shape[1], w1.shape[0], w2.shape[0]
        dy2 = dy.reshape(-1, N)
        dy2 = (dy2.float() if x3 else dy2.to(torch.bfloat16)).contiguous()
        M = dy2.shape[0]
        dev = dy.device
        act = torch.float32 if x3 else torch.bfloat16
        db2 = torch.zeros(N, device=dev)
        ops.colsum(dy2, db2, M, N)
        dw2 = torch.zeros(N, Hd, device=dev)
        _dw(dy2, hs, dw2, M, N, Hd, x3)
        A, Keff = _a_operand(dy2, M, N, x3)
        d_pre = torch.empty(M, Hd, device=dev, dtype=act)
        db1 = torch.zeros(Hd, device=dev)
        ops.gemm(A, _weight_operands(w2, x3)[1], d_pre, M=M, N=Hd, K=Keff, epilogue=ops.EPI_DGELU, aux=dact, colsum=db1)
        dw1 = torch.zeros(Hd, K, device=dev)
        _dw(d_pre, xs, dw1, M, Hd, K, x3)
        dx = None
        if ctx.needs_input_grad[0]:
            P, Peff = _a_operand(d_pre, M, Hd, x3)
            dx2 = torch.empty(M, K, device=dev, dtype=act)
            ops.gemm(P, _weight_operands(w1, x3)[1], dx2, M=M, N=K, K=Peff)
            dx = dx2.reshape(*ctx.lead, K).to(ctx.in_dtype)
        return dx, dw1.view_as(w1), db1, dw2.view_as(w2), db2, None


def tc_linear(x, weight, bias=None, precision="bf16"):
    """y = x @ weight.T + bias on the tcgen05 GEMM; x [..., K], weight [N, K] fp32 master, returns [..., N] (bf16, or
    fp32 in the "bf16x3" parity mode)."""
    assert precision in ("bf16", "bf16x3")
    return _TCLinear.apply(x, weight, bias, precision == "bf16x3")


def tc_mlp(x, fc1_weight, fc1_bias, fc2_weight, fc2_bias, precision="bf16"):
    """fc2(gelu_erf(fc1(x))) with the GELU and its derivative fused into the GEMM epilogues."""
    assert precision in ("bf16", "bf16x3")
    return _TCMlp.apply(x, fc1_weight, fc1_bias, fc2_weight, fc2_bias, precision == "bf16x3")
