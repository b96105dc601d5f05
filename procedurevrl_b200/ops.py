"""ctypes binding of libpvrl_sm100.so (include/pvrl.h): the only way the host code reaches the GPU math.

There is no fallback: if the shared library is missing or a call returns non-zero this raises.  Every
function takes torch CUDA tensors purely as (pointer, shape) carriers and enqueues on torch's current
stream; nothing here computes."""
import ctypes
import os

import torch

from . import build as _build

BF16, F32 = 0, 1
MAP_IDENT, MAP_SKIPCLS, MAP_SPATIAL, MAP_PATCH, MAP_CLS = 0, 1, 2, 3, 4
EPI_STORE, EPI_GELU, EPI_DGELU, EPI_RESID, EPI_ATOMIC = 0, 1, 2, 3, 4

_c_void_p, _c_int, _c_i64, _c_float = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float


class Geom(ctypes.Structure):
    _fields_ = [("T", _c_int), ("HW", _c_int)]


class GemmDesc(ctypes.Structure):
    _fields_ = [
        ("M", _c_int), ("N", _c_int), ("K", _c_int), ("trans", _c_int),
        ("A", _c_void_p), ("lda", _c_i64), ("B", _c_void_p), ("ldb", _c_i64),
        ("epilogue", _c_int), ("out_dtype", _c_int),
        ("out", _c_void_p), ("ldo", _c_i64), ("out2", _c_void_p),
        ("bias", _c_void_p), ("rowscale", _c_void_p), ("rs_div", _c_int), ("map", _c_int),
        ("aux", _c_void_p), ("ld_aux", _c_i64),
        ("resid", _c_void_p), ("add_pos", _c_void_p), ("add_time", _c_void_p),
        ("g", Geom), ("k_splits", _c_int), ("colsum", _c_void_p),
    ]


class Pool3dDesc(ctypes.Structure):
    _fields_ = [("B", _c_int), ("heads", _c_int), ("C", _c_int), ("T", _c_int), ("H", _c_int), ("W", _c_int),
                ("kernel", _c_int * 3), ("stride", _c_int * 3), ("pad", _c_int * 3), ("out", _c_int * 3), ("ld", _c_i64)]


class PooledAttnDesc(ctypes.Structure):
    _fields_ = [("B", _c_int), ("heads", _c_int), ("Nq", _c_int), ("Nk", _c_int), ("C", _c_int),
                ("Kt", _c_int), ("Kh", _c_int), ("Kw", _c_int), ("scale", _c_float), ("residual_pooling", _c_int)]


_SIGS = {
    "pvrl_gemm_bf16": [ctypes.POINTER(GemmDesc), _c_void_p],
    "pvrl_patchify": [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p],
    "pvrl_patchify_u8": [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, ctypes.POINTER(_c_float),
                         ctypes.POINTER(_c_float), _c_void_p],
    "pvrl_cls_init": [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p],
    "pvrl_layernorm_fwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_int,
                           _c_float, _c_int, Geom, _c_void_p],
    "pvrl_layernorm_bwd": [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                           _c_void_p, _c_int, _c_int, _c_int, Geom, _c_void_p],
    "pvrl_layernorm_bwd_emit": [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                _c_void_p, _c_int, _c_int, _c_int, Geom, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p,
                                _c_void_p],
    "pvrl_gather_cast": [_c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_int, _c_int, Geom, _c_void_p,
                         _c_void_p],
    "pvrl_cls_merge": [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p],
    "pvrl_colsum": [_c_void_p, _c_int, _c_i64, _c_void_p, _c_int, _c_int, _c_void_p],
    "pvrl_cast_weight": [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p],
    "pvrl_cast_weight_multi": [_c_void_p, _c_int, _c_int, _c_int, _c_void_p],
    "pvrl_split3": [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p],
    "pvrl_embed_bwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, Geom, _c_void_p],
    "pvrl_attn_fwd": [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_float, _c_void_p],
    "pvrl_attn_bwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_float,
                      _c_void_p],
    "pvrl_attn_tc_fwd": [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_float, _c_void_p],
    "pvrl_attn_tc_bwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_float,
                         _c_void_p],
    "pvrl_debug_sp_trace": [_c_void_p],
    "pvrl_reduce_chunks": [_c_void_p, _c_void_p, _c_int, ctypes.c_int64, ctypes.c_int64, _c_float, _c_void_p],
    "pvrl_linear_small_fwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p],
    "pvrl_linear_small_bwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int,
                              _c_void_p],
    "pvrl_l2norm_fwd": [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p],
    "pvrl_l2norm_bwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p],
    "pvrl_sim_logits_fwd": [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_float, _c_void_p],
    "pvrl_sim_logits_bwd": [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_float, _c_void_p],
    "pvrl_kl_topk_loss": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_float,
                          _c_void_p],
    "pvrl_softmax_rows": [_c_void_p, _c_void_p, _c_int, _c_int, _c_void_p],
    "pvrl_ot_linear_fwd": [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_float, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                           _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p],
    "pvrl_ot_linear_dx": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p],
    "pvrl_ot_linear_dw": [_c_void_p, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int,
                          _c_void_p],
    "pvrl_ot_ln_bwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int,
                       _c_void_p],
    "pvrl_ot_attn_fwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p],
    "pvrl_ot_attn_bwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p],
    "pvrl_ot_embed_fwd": [_c_void_p, _c_void_p, _c_void_p, _c_float, _c_float, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                          _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p],
    "pvrl_ot_embed_bwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int,
                          _c_int, _c_int, _c_void_p],
    "pvrl_ln_any_fwd": [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_float,
                        _c_void_p],
    "pvrl_ln_any_bwd": [_c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int,
                        _c_int, _c_void_p],
    "pvrl_pool3d_fwd": [_c_void_p, _c_void_p, _c_void_p, _c_int, ctypes.POINTER(Pool3dDesc), _c_void_p],
    "pvrl_pool3d_bwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, ctypes.POINTER(Pool3dDesc), _c_void_p],
    "pvrl_maxpool3d_fwd": [_c_void_p, _c_void_p, _c_void_p, _c_int, ctypes.POINTER(Pool3dDesc), _c_void_p],
    "pvrl_maxpool3d_bwd": [_c_void_p, _c_void_p, _c_void_p, _c_int, ctypes.POINTER(Pool3dDesc), _c_void_p],
    "pvrl_im2col3d": [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, ctypes.POINTER(Pool3dDesc), _c_void_p],
    "pvrl_pooled_attn_fwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int,
                             ctypes.POINTER(PooledAttnDesc), _c_void_p],
    "pvrl_pooled_attn_bwd": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                             _c_void_p, _c_void_p, _c_void_p, _c_int, ctypes.POINTER(PooledAttnDesc), _c_void_p],
    "pvrl_optim_tick": [_c_void_p, _c_void_p],
    "pvrl_adam_flat": [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_i64, _c_void_p, _c_void_p, _c_float,
                       ctypes.c_double, ctypes.c_double, _c_float, _c_float, _c_int, _c_float, _c_int, _c_void_p],
    "pvrl_sgd_flat": [_c_void_p, _c_void_p, _c_void_p, _c_i64, _c_void_p, _c_void_p, _c_float, _c_float, _c_float,
                      _c_int, _c_float, _c_float, _c_int, _c_void_p],
}

_lib = None


def lib():
    """The loaded C-ABI library.  Fails loudly when it has not been built (python -m procedurevrl_b200.build)."""
    global _lib
    if _lib is None:
        path = os.environ.get("PVRL_LIB") or _build.LIB_PATH        # PVRL_LIB: A/B runs of an alternative build of the same ABI
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `python -m procedurevrl_b200.build` (there is no fallback path)")
        L = ctypes.CDLL(path)
        for name, args in _SIGS.items():
            fn = getattr(L, name)
            fn.argtypes, fn.restype = args, ctypes.c_int
        L.pvrl_last_error.restype = ctypes.c_char_p
        L.pvrl_abi_version.restype = ctypes.c_int
        L.pvrl_launch_count.restype = ctypes.c_int64
        if L.pvrl_abi_version() != 1:
            raise RuntimeError("libpvrl_sm100.so ABI version mismatch")
        _lib = L
    return _lib


def exported_symbols():
    return list(_SIGS) + ["pvrl_last_error", "pvrl_abi_version", "pvrl_launch_count"]


def launch_count():
    return int(lib().pvrl_launch_count())


def _check(rc, name):
    if rc != 0:
        raise RuntimeError(f"{name} failed (rc={rc}): {lib().pvrl_last_error().decode()}")


def _p(t):
    if t is None:
        return None
    assert t.is_cuda, "pvrl ops take CUDA tensors only (no CPU fallback)"
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _dt(t):
    if t.dtype == torch.bfloat16:
        return BF16
    assert t.dtype == torch.float32, t.dtype
    return F32


def _geom(T=1, HW=1):
    return Geom(int(T), int(HW))


# ---------------------------------------------------------------------------------------------- GEMM
def gemm(A, B, out, *, M, N, K, trans=0, epilogue=EPI_STORE, out2=None, bias=None, rowscale=None, rs_div=0,
         map=MAP_IDENT, aux=None, resid=None, add_pos=None, add_time=None, T=1, HW=1, k_splits=0,
         lda=None, ldb=None, ldo=None, colsum=None):
    """out = epilogue(A @ B^T) -- see pvrl_gemm_t in include/pvrl.h.  A, B bf16; contiguous 2-D tensors."""
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    d = GemmDesc()
    d.M, d.N, d.K, d.trans = M, N, K, trans
    d.A, d.lda = _p(A), lda if lda is not None else A.stride(0)
    d.B, d.ldb = _p(B), ldb if ldb is not None else B.stride(0)
    d.epilogue = epilogue
    d.out_dtype = F32 if epilogue in (EPI_RESID, EPI_ATOMIC) else _dt(out)
    d.out, d.ldo = _p(out), ldo if ldo is not None else out.stride(-2)
    d.out2 = _p(out2)
    d.bias, d.rowscale, d.rs_div, d.map = _p(bias), _p(rowscale), rs_div, map
    d.aux, d.ld_aux = _p(aux), (aux.stride(0) if aux is not None else 0)
    d.resid, d.add_pos, d.add_time = _p(resid), _p(add_pos), _p(add_time)
    d.g = _geom(T, HW)
    d.k_splits = k_splits
    d.colsum = _p(colsum)
    _check(lib().pvrl_gemm_bf16(ctypes.byref(d), _stream()), "pvrl_gemm_bf16")
    return out


# ---------------------------------------------------------------------------------------------- elementwise
def patchify(frames, out, patch=16):
    Bc, C, T, H, W = frames.shape
    assert C == 3 and frames.dtype == torch.float32 and frames.is_contiguous()
    _check(lib().pvrl_patchify(_p(frames), _p(out), _dt(out), Bc, T, H, W, patch, _stream()), "pvrl_patchify")
    return out


def patchify_u8(frames, out, patch, mean, std):
    """uint8 frames [Bc, 3, T, H, W] -> normalised im2col rows (the host-side tensor_normalize fused into the kernel)."""
    Bc, C, T, H, W = frames.shape
    assert C == 3 and frames.dtype == torch.uint8 and frames.is_contiguous()
    m3, s3 = (_c_float * 3)(*[float(v) for v in mean]), (_c_float * 3)(*[float(v) for v in std])
    _check(lib().pvrl_patchify_u8(_p(frames), _p(out), _dt(out), Bc, T, H, W, patch, m3, s3, _stream()), "pvrl_patchify_u8")
    return out


def cls_init(x, cls_token, pos_embed):
    Bc, S, D = x.shape
    _check(lib().pvrl_cls_init(_p(x), _p(cls_token), _p(pos_embed), Bc, S, D, _stream()), "pvrl_cls_init")


def layernorm_fwd(x, w, b, y, stats, M, D, eps, map=MAP_IDENT, x_cls=None, T=1, HW=1):
    _check(lib().pvrl_layernorm_fwd(_p(x), _p(x_cls), _p(w), _p(b), _p(y), _dt(y), _p(stats), M, D, eps, map,
                                    _geom(T, HW), _stream()), "pvrl_layernorm_fwd")
    return y


def layernorm_bwd(dy, x, w, stats, dx, dw, db, M, D, map=MAP_IDENT, x_cls=None, T=1, HW=1, emit=None):
    """emit = (out, map, rowscale, rs_div, colsum): also write what gather_cast(dx, out, ..., map, rowscale, rs_div,
    colsum) would produce after this call (see pvrl_layernorm_bwd_emit for the supported map pairs)."""
    if emit is None:
        _check(lib().pvrl_layernorm_bwd(_p(dy), _dt(dy), _p(x), _p(x_cls), _p(w), _p(stats), _p(dx), _p(dw), _p(db), M, D,
                                        map, _geom(T, HW), _stream()), "pvrl_layernorm_bwd")
        return
    out, emap, rowscale, rs_div, colsum = emit
    assert out.dtype == dy.dtype, "the emitted operand has the dtype of dy"
    _check(lib().pvrl_layernorm_bwd_emit(_p(dy), _dt(dy), _p(x), _p(x_cls), _p(w), _p(stats), _p(dx), _p(dw), _p(db), M, D,
                                         map, _geom(T, HW), _p(out), emap, _p(rowscale), rs_div or 0, _p(colsum),
                                         _stream()), "pvrl_layernorm_bwd_emit")


def gather_cast(src, out, M, D, map=MAP_IDENT, rowscale=None, rs_div=0, T=1, HW=1, colsum=None):
    _check(lib().pvrl_gather_cast(_p(src), _p(out), _dt(out), _p(rowscale), rs_div, M, D, map, _geom(T, HW), _p(colsum),
                                  _stream()), "pvrl_gather_cast")
    return out


def cls_merge(x0, side, x2, Bc, T, S, D):
    _check(lib().pvrl_cls_merge(_p(x0), _p(side), _p(x2), Bc, T, S, D, _stream()), "pvrl_cls_merge")


def colsum(a, out, M, N):
    _check(lib().pvrl_colsum(_p(a), _dt(a), a.stride(0), _p(out), M, N, _stream()), "pvrl_colsum")


def cast_weight(w, w_out, wT_out):
    rows, cols = w.shape
    ref = w_out if w_out is not None else wT_out
    _check(lib().pvrl_cast_weight(_p(w), _p(w_out), _p(wT_out), _dt(ref), rows, cols, _stream()), "pvrl_cast_weight")


class CastDesc(ctypes.Structure):
    _fields_ = [("w", _c_void_p), ("out", _c_void_p), ("outT", _c_void_p), ("rows", _c_int), ("cols", _c_int),
                ("tile0", _c_int), ("tiles_x", _c_int)]


def cast_weight_table(triples):
    """Device descriptor table for cast_weight_multi: triples of (w fp32 [rows, cols], out, outT) -> (table, n, tiles, dtype)."""
    arr = (CastDesc * len(triples))()
    t0 = 0
    for i, (w, o, oT) in enumerate(triples):
        rows, cols = w.shape
        assert rows % 2 == 0 and cols % 2 == 0, "cast_weight_multi moves element pairs"
        tx = (cols + 63) // 64
        arr[i] = CastDesc(_p(w), _p(o), _p(oT), rows, cols, t0, tx)
        t0 += tx * ((rows + 63) // 64)
    raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(triples[0][0].device)
    return raw, len(triples), t0, _dt(triples[0][1] if triples[0][1] is not None else triples[0][2])


def cast_weight_multi(table):
    raw, n, tiles, dt = table
    _check(lib().pvrl_cast_weight_multi(_p(raw), n, tiles, dt, _stream()), "pvrl_cast_weight_multi")


def split3(a, out, M, K, pattern, along):
    assert a.dtype == torch.float32 and out.dtype == torch.bfloat16
    _check(lib().pvrl_split3(_p(a), _p(out), M, K, pattern, along, _stream()), "pvrl_split3")
    return out


def embed_bwd(dx, dcls, dpos, dtime, Bc, D, T, HW):
    _check(lib().pvrl_embed_bwd(_p(dx), _p(dcls), _p(dpos), _p(dtime), Bc, D, _geom(T, HW), _stream()),
           "pvrl_embed_bwd")


# ---------------------------------------------------------------------------------------------- attention
def attn_fwd(qkv, out, lse, n_seq, seq, H, scale):
    _check(lib().pvrl_attn_fwd(_p(qkv), _p(out), _p(lse), _dt(qkv), n_seq, seq, H, scale, _stream()), "pvrl_attn_fwd")
    return out


def attn_bwd(qkv, out, dout, lse, dqkv, n_seq, seq, H, scale):
    _check(lib().pvrl_attn_bwd(_p(qkv), _p(out), _p(dout), _p(lse), _p(dqkv), _dt(qkv), n_seq, seq, H, scale,
                               _stream()), "pvrl_attn_bwd")
    return dqkv


def attn_tc_fwd(qkv, out, lse, n_seq, seq, H, scale):
    assert qkv.dtype == torch.bfloat16
    _check(lib().pvrl_attn_tc_fwd(_p(qkv), _p(out), _p(lse), n_seq, seq, H, scale, _stream()), "pvrl_attn_tc_fwd")
    return out


def attn_tc_bwd(qkv, out, dout, lse, dqkv, n_seq, seq, H, scale):
    assert qkv.dtype == torch.bfloat16
    _check(lib().pvrl_attn_tc_bwd(_p(qkv), _p(out), _p(dout), _p(lse), _p(dqkv), n_seq, seq, H, scale, _stream()),
           "pvrl_attn_tc_bwd")
    return dqkv


def reduce_chunks(dst, src, n_src, stride, n, scale):
    """dst[:n] = (dst[:n] + sum_s src[s * stride : s * stride + n]) * scale (fp32): local half of the copy-engine gradient exchange."""
    assert dst.dtype == torch.float32 and (n_src == 0 or src.dtype == torch.float32)
    _check(lib().pvrl_reduce_chunks(_p(dst), _p(src) if n_src else None, n_src, stride, n, scale, _stream()), "pvrl_reduce_chunks")
    return dst


def debug_sp_trace():
    """PVRL_SP_TRACE=1: clock64 stamps of CTA 0 of the last persistent spatial-attention launch, [20, 16, 8] int64."""
    import numpy as np
    buf = np.zeros((20, 16, 8), dtype=np.int64)
    torch.cuda.synchronize()
    n = lib().pvrl_debug_sp_trace(buf.ctypes.data)
    return buf if n else None


# ---------------------------------------------------------------------------------------------- head / loss
def linear_small_fwd(x, w, b, y):
    M, K = x.shape
    N = w.shape[0]
    _check(lib().pvrl_linear_small_fwd(_p(x), _p(w), _p(b), _p(y), M, K, N, _stream()), "pvrl_linear_small_fwd")
    return y


def linear_small_bwd(x, w, dy, dx, dw, db):
    M, K = x.shape
    N = w.shape[0]
    _check(lib().pvrl_linear_small_bwd(_p(x), _p(w), _p(dy), _p(dx), _p(dw), _p(db), M, K, N, _stream()),
           "pvrl_linear_small_bwd")


def l2norm_fwd(x, y, norms):
    M, C = x.shape
    _check(lib().pvrl_l2norm_fwd(_p(x), _p(y), _p(norms), M, C, _stream()), "pvrl_l2norm_fwd")
    return y


def l2norm_bwd(y, norms, dy, dx):
    M, C = y.shape
    _check(lib().pvrl_l2norm_bwd(_p(y), _p(norms), _p(dy), _p(dx), M, C, _stream()), "pvrl_l2norm_bwd")
    return dx


def sim_logits_fwd(emb, label, logits, inv_temp):
    M, K = emb.shape
    C = label.shape[0]
    _check(lib().pvrl_sim_logits_fwd(_p(emb), _p(label), _p(logits), M, C, K, inv_temp, _stream()),
           "pvrl_sim_logits_fwd")
    return logits


def sim_logits_bwd(dlogits, label, demb, inv_temp):
    M, C = dlogits.shape
    K = label.shape[1]
    _check(lib().pvrl_sim_logits_bwd(_p(dlogits), _p(label), _p(demb), M, C, K, inv_temp, _stream()),
           "pvrl_sim_logits_bwd")
    return demb


def kl_topk_loss(pred, teacher_logits, row_loss, dpred, teacher_out, topk, gscale=1.0):
    M, K = pred.shape
    _check(lib().pvrl_kl_topk_loss(_p(pred), _p(teacher_logits), _p(row_loss), _p(dpred), _p(teacher_out), M, K, topk,
                                   gscale, _stream()), "pvrl_kl_topk_loss")


def softmax_rows(x, y):
    M, K = x.shape
    _check(lib().pvrl_softmax_rows(_p(x), _p(y), M, K, _stream()), "pvrl_softmax_rows")
    return y


# ---------------------------------------------------------------------------------------------- order transformer
OT_X_PLAIN, OT_X_LN, OT_X_QGELU = 0, 1, 2


def _f32c(*ts):
    for t in ts:
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous()), "this argument must be a contiguous fp32 tensor"


def ot_linear_fwd(x, W, bias, y, x_mode=OT_X_PLAIN, ln_w=None, ln_b=None, eps=1e-5, xhat=None, rstd=None, resid=None,
                  act_out=None):
    M, K = x.shape
    N = W.shape[0]
    _f32c(x, W, bias, y, ln_w, ln_b, xhat, rstd, resid, act_out)
    _check(lib().pvrl_ot_linear_fwd(_p(x), x_mode, _p(ln_w), _p(ln_b), eps, _p(xhat), _p(rstd), _p(W), _p(bias), _p(resid),
                                    _p(y), _p(act_out), M, N, K, _stream()), "pvrl_ot_linear_fwd")
    return y


def ot_linear_dx(dY, W, dA, pre=None):
    M, N = dY.shape
    K = W.shape[1]
    _f32c(dY, W, dA, pre)
    _check(lib().pvrl_ot_linear_dx(_p(dY), _p(W), _p(pre), _p(dA), M, N, K, _stream()), "pvrl_ot_linear_dx")
    return dA


def ot_linear_dw(dY, A, dW, db, a_mode=OT_X_PLAIN, ln_w=None, ln_b=None):
    M, N = dY.shape
    K = A.shape[1]
    _f32c(dY, A, dW, db, ln_w, ln_b)
    _check(lib().pvrl_ot_linear_dw(_p(dY), _p(A), a_mode, _p(ln_w), _p(ln_b), _p(dW), _p(db), M, N, K, _stream()),
           "pvrl_ot_linear_dw")


def ot_ln_bwd(dA, xhat, rstd, w, dh, dw, db):
    M, C = dA.shape
    _f32c(dA, xhat, rstd, w, dh, dw, db)
    _check(lib().pvrl_ot_ln_bwd(_p(dA), _p(xhat), _p(rstd), _p(w), _p(dh), _p(dw), _p(db), M, C, _stream()),
           "pvrl_ot_ln_bwd")


def ot_attn_fwd(qkv, pad_start, probs, o, B, S, H):
    _f32c(qkv, probs, o)
    assert pad_start is None or pad_start.dtype == torch.int64
    _check(lib().pvrl_ot_attn_fwd(_p(qkv), _p(pad_start), _p(probs), _p(o), B, S, H, _stream()), "pvrl_ot_attn_fwd")
    return o


def ot_attn_bwd(qkv, probs, dO, dqkv, B, S, H):
    _f32c(qkv, probs, dO, dqkv)
    _check(lib().pvrl_ot_attn_bwd(_p(qkv), _p(probs), _p(dO), _p(dqkv), B, S, H, _stream()), "pvrl_ot_attn_bwd")
    return dqkv


def ot_embed_fwd(video, src, noise, ca, cb, mask_inds, pad_start, type_w, pos_w, pad_w, tvec, h, B, S):
    C = video.shape[1]
    _f32c(video, src, noise, type_w, pos_w, pad_w, tvec, h)
    assert mask_inds.dtype == torch.int64 and pad_start.dtype == torch.int64
    _check(lib().pvrl_ot_embed_fwd(_p(video), _p(src), _p(noise), ca, cb, _p(mask_inds), _p(pad_start), _p(type_w),
                                   _p(pos_w), _p(pad_w), _p(tvec), _p(h), B, S, C, _stream()), "pvrl_ot_embed_fwd")
    return h


def ot_embed_bwd(dh, mask_inds, pad_start, dvideo, dtype, dpos, dpad, dtvec, B, S):
    C = dh.shape[1]
    _f32c(dh, dvideo, dtype, dpos, dpad, dtvec)
    _check(lib().pvrl_ot_embed_bwd(_p(dh), _p(mask_inds), _p(pad_start), _p(dvideo), _p(dtype), _p(dpos), _p(dpad),
                                   _p(dtvec), B, S, C, _stream()), "pvrl_ot_embed_bwd")


# ---------------------------------------------------------------------------------------------- optimizer (flat buffers)
def optim_tick(step):
    _f32c(step)
    _check(lib().pvrl_optim_tick(_p(step), _stream()), "pvrl_optim_tick")


def adam_flat(p, g, m, v, lr, step, *, lr_mult=1.0, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, decoupled=True,
              grad_scale=1.0, zero_grad=True):
    """One parameter group of torch.optim.AdamW (decoupled) / Adam on flat fp32 slices; lr / step are device scalars."""
    _f32c(p, g, m, v, lr, step)
    assert p.numel() == g.numel() == m.numel() == v.numel()
    _check(lib().pvrl_adam_flat(_p(p), _p(g), _p(m), _p(v), p.numel(), _p(lr), _p(step), lr_mult, beta1, beta2, eps,
                                weight_decay, int(decoupled), grad_scale, int(zero_grad), _stream()), "pvrl_adam_flat")


def sgd_flat(p, g, buf, lr, step, *, lr_mult=1.0, momentum=0.0, dampening=0.0, nesterov=False, weight_decay=0.0,
             grad_scale=1.0, zero_grad=True):
    """One parameter group of torch.optim.SGD on flat fp32 slices; lr / step are device scalars."""
    _f32c(p, g, buf, lr, step)
    assert p.numel() == g.numel() == buf.numel()
    _check(lib().pvrl_sgd_flat(_p(p), _p(g), _p(buf), p.numel(), _p(lr), _p(step), lr_mult, momentum, dampening,
                               int(nesterov), weight_decay, grad_scale, int(zero_grad), _stream()), "pvrl_sgd_flat")


# ---------------------------------------------------------------------------------------------- MViTv2 encoder ops
def pool_out_grid(grid, kernel, stride, pad):
    """floor((in + 2 pad - kernel) / stride) + 1 per axis (nn.Conv3d / nn.MaxPool3d without ceil_mode)."""
    return [(g + 2 * p - k) // s + 1 for g, k, s, p in zip(grid, kernel, stride, pad)]


def _pool_desc(B, heads, C, grid, kernel, stride, pad, ld):
    d = Pool3dDesc()
    d.B, d.heads, d.C = int(B), int(heads), int(C)
    d.T, d.H, d.W = (int(g) for g in grid)
    d.kernel, d.stride, d.pad = (_c_int * 3)(*kernel), (_c_int * 3)(*stride), (_c_int * 3)(*pad)
    d.out = (_c_int * 3)(*pool_out_grid(grid, kernel, stride, pad))
    d.ld = int(ld)
    return d


def _col_ptr(t, col0):
    """address of column `col0` of the first row of a contiguous [.., ld] CUDA tensor"""
    assert t.is_cuda and t.is_contiguous(), "pvrl ops take contiguous CUDA tensors only (no CPU fallback)"
    return t.data_ptr() + int(col0) * t.element_size()


def ln_any_fwd(x, w, b, y, stats, M, D, eps):
    """y = LayerNorm(x) over the last axis of [M, D], any D <= 1024; stats [M, 2] = (mean, rstd)."""
    _f32c(w, b, stats)
    assert x.is_contiguous() and y.is_contiguous()
    _check(lib().pvrl_ln_any_fwd(_p(x), _dt(x), _p(w), _p(b), _p(y), _dt(y), _p(stats), M, D, eps, _stream()),
           "pvrl_ln_any_fwd")
    return y


def ln_any_bwd(dy, x, w, stats, dx, dw, db, M, D):
    """dx (dtype of x) written; dw / db (fp32) accumulated."""
    _f32c(w, stats, dw, db)
    assert dy.is_contiguous() and x.is_contiguous() and dx.is_contiguous() and dx.dtype == x.dtype
    _check(lib().pvrl_ln_any_bwd(_p(dy), _dt(dy), _p(x), _dt(x), _p(w), _p(stats), _p(dx), _p(dw), _p(db), M, D, _stream()),
           "pvrl_ln_any_bwd")


def pool3d_fwd(src, col0, w, out, heads, C, grid, kernel, stride, pad):
    """attention_pool of one of Q / K / V: src [B, 1 + T*H*W, ld] (columns [col0, col0 + heads*C) are this tensor),
    w [C, kt*kh*kw] fp32 or None (re-layout only) -> out [B, heads, 1 + To*Ho*Wo, C]."""
    B, _, ld = src.shape
    assert out.dtype == src.dtype and out.is_contiguous()
    if w is not None:
        _f32c(w)
    d = _pool_desc(B, heads, C, grid, kernel, stride, pad, ld)
    _check(lib().pvrl_pool3d_fwd(_col_ptr(src, col0), _p(w), _p(out), _dt(src), ctypes.byref(d), _stream()), "pvrl_pool3d_fwd")
    return out


def pool3d_bwd(dout, src, col0, w, din, dw, heads, C, grid, kernel, stride, pad):
    """din [B, 1 + T*H*W, ld]: columns [col0, col0 + heads*C) are written; dw [C, 27] (or None) accumulated."""
    B, _, ld = din.shape
    assert dout.dtype == din.dtype and dout.is_contiguous() and (src is None or src.shape == din.shape)
    if w is not None:
        _f32c(w, dw)
    d = _pool_desc(B, heads, C, grid, kernel, stride, pad, ld)
    _check(lib().pvrl_pool3d_bwd(_p(dout), None if src is None else _col_ptr(src, col0), _p(w), _col_ptr(din, col0), _p(dw),
                                 _dt(din), ctypes.byref(d), _stream()), "pvrl_pool3d_bwd")


def maxpool3d_fwd(x, y, arg, grid, kernel, stride, pad):
    """x [B, 1 + T*H*W, D] -> y [B, 1 + To*Ho*Wo, D], arg int32 [B, To*Ho*Wo, D] (winning input token)."""
    B, _, D = x.shape
    assert x.is_contiguous() and y.is_contiguous() and y.dtype == x.dtype and arg.dtype == torch.int32
    d = _pool_desc(B, 1, D, grid, kernel, stride, pad, D)
    _check(lib().pvrl_maxpool3d_fwd(_p(x), _p(y), _p(arg), _dt(x), ctypes.byref(d), _stream()), "pvrl_maxpool3d_fwd")
    return y


def maxpool3d_bwd(dy, arg, dx, grid, kernel, stride, pad):
    """dx fp32 [B, 1 + T*H*W, D], zero-initialised by the caller, += dy scattered through arg."""
    B, _, D = dx.shape
    _f32c(dx)
    assert dy.is_contiguous() and arg.dtype == torch.int32
    d = _pool_desc(B, 1, D, grid, kernel, stride, pad, D)
    _check(lib().pvrl_maxpool3d_bwd(_p(dy), _p(arg), _p(dx), _dt(dy), ctypes.byref(d), _stream()), "pvrl_maxpool3d_bwd")


def im2col3d(frames, out, kernel, stride, pad):
    """frames fp32 [B, Cin, T, H, W] -> out [B * To*Ho*Wo, Kpad] rows of the Conv3d stem (zero-padded columns)."""
    B, Cin, T, H, W = frames.shape
    _f32c(frames)
    assert out.is_contiguous()
    d = _pool_desc(B, 1, Cin, (T, H, W), kernel, stride, pad, 0)
    _check(lib().pvrl_im2col3d(_p(frames), _p(out), _dt(out), Cin, out.shape[1], ctypes.byref(d), _stream()), "pvrl_im2col3d")
    return out


def _attn_desc(q, k, kgrid, scale, resid):
    B, heads, Nq, C = q.shape
    d = PooledAttnDesc()
    d.B, d.heads, d.Nq, d.Nk, d.C = B, heads, Nq, k.shape[2], C
    d.Kt, d.Kh, d.Kw = (int(g) for g in kgrid)
    d.scale, d.residual_pooling = float(scale), int(bool(resid))
    return d


def pooled_attn_fwd(q, k, v, bq, out, lse, kgrid, scale, resid):
    """q [B, heads, Nq, 96], k / v [B, heads, Nk, 96], bq fp32 [B, heads, Nq - 1, Kt + Kh + Kw] -> out [B, Nq, heads * 96],
    lse fp32 [B, heads, Nq] (see pvrl_pooled_attn_fwd in include/pvrl.h)."""
    _f32c(bq, lse)
    assert q.is_contiguous() and k.is_contiguous() and v.is_contiguous() and out.is_contiguous()
    assert q.dtype == k.dtype == v.dtype == out.dtype
    d = _attn_desc(q, k, kgrid, scale, resid)
    _check(lib().pvrl_pooled_attn_fwd(_p(q), _p(k), _p(v), _p(bq), _p(out), _p(lse), _dt(q), ctypes.byref(d), _stream()),
           "pvrl_pooled_attn_fwd")
    return out


def pooled_attn_bwd(q, k, v, bq, dout, lse, dq, dk, dv, dbq, delta, kgrid, scale, resid):
    """dq (like q), dbq (like bq), delta written; dk / dv fp32, zero-initialised by the caller, accumulated."""
    _f32c(bq, lse, dk, dv, dbq, delta)
    assert dout.is_contiguous() and dout.dtype == q.dtype and dq.dtype == q.dtype and dq.is_contiguous()
    d = _attn_desc(q, k, kgrid, scale, resid)
    _check(lib().pvrl_pooled_attn_bwd(_p(q), _p(k), _p(v), _p(bq), _p(dout), _p(lse), _p(dq), _p(dk), _p(dv),
                                      _p(dbq), _p(delta), _dt(q), ctypes.byref(d), _stream()), "pvrl_pooled_attn_bwd")
