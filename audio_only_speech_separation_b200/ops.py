"""Python-level operators over the C-ABI (one function per reference call they replace).

Each function validates its arguments the way the reference module would fail (TypeError / ValueError), requires
CUDA tensors, launches on the current torch stream and returns torch tensors.  None of them has a CPU path.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, lib, ptr, require_cuda, stream_ptr

__all__ = [
    "split_feature",
    "merge_feature",
    "segment_channels_last",
    "overlap_add_channels_last",
    "linear",
    "linear_wgrad",
    "split_bf16",
    "LstmPack",
    "bilstm_forward",
    "bilstm_backward",
    "groupnorm_residual",
    "attention",
    "attention_backward",
    "attention_tensor_cores",
    "add_layernorm",
    "layernorm_backward",
    "split_rows",
    "linear_planes",
    "linear_wgrad_planes",
    "attention_planes",
]


def _prec(precision) -> int:
    if precision in (0, 1):
        return precision
    p = str(precision).lower()
    if p in ("fp32", "float32"):
        return _lib.PREC_FP32
    if p in ("bf16", "bfloat16"):
        return _lib.PREC_BF16
    raise ValueError(f"precision must be 'fp32' or 'bf16', got {precision!r}")


def split_feature(x: torch.Tensor, block_size: int):
    """``split_feature`` of look2hear/models/utils/gc3_basics.py:79-91: ``[B,N,L] -> ([B,N,K,S], rest)``, bit-exact."""
    if x.ndim != 3:
        raise ValueError(f"expected [B, N, L], got {tuple(x.shape)}")
    require_cuda(x, "input")
    if x.dtype != torch.float32:
        raise TypeError("split_feature: float32 only")
    B, N, L = x.shape
    rest, S = _lib.seg_geometry(L, block_size)
    y = torch.empty(B, N, block_size, S, device=x.device, dtype=x.dtype)
    check(lib().dp_segment_f32(ptr(x), ptr(y), B, N, L, block_size, stream_ptr()), "dp_segment_f32")
    return y, rest


def merge_feature(y: torch.Tensor, rest: int) -> torch.Tensor:
    """``merge_feature`` of gc3_basics.py:94-109: ``[B,N,K,S] -> [B,N,L]``, bit-exact."""
    if y.ndim != 4:
        raise ValueError(f"expected [B, N, K, S], got {tuple(y.shape)}")
    require_cuda(y, "input")
    if y.dtype != torch.float32:
        raise TypeError("merge_feature: float32 only")
    B, N, K, S = y.shape
    L = (S // 2) * K - K // 2 - rest
    if L <= 0:
        raise ValueError(f"rest={rest} is inconsistent with K={K}, S={S}")
    x = torch.empty(B, N, L, device=y.device, dtype=y.dtype)
    check(lib().dp_overlap_add_f32(ptr(y), ptr(x), B, N, K, S, L, stream_ptr()), "dp_overlap_add_f32")
    return x


def segment_channels_last(f: torch.Tensor, block_size: int) -> torch.Tensor:
    """Same map as :func:`split_feature` on ``[B,L,C] -> [B,S,K,C]`` (the engine's internal layout)."""
    require_cuda(f, "input")
    B, L, Cc = f.shape
    _, S = _lib.seg_geometry(L, block_size)
    x = torch.empty(B, S, block_size, Cc, device=f.device, dtype=torch.float32)
    check(lib().dp_segment_cl_f32(ptr(f), ptr(x), B, L, block_size, Cc, stream_ptr()), "dp_segment_cl_f32")
    return x


def overlap_add_channels_last(x: torch.Tensor, L: int) -> torch.Tensor:
    require_cuda(x, "input")
    B, S, K, Cc = x.shape
    f = torch.empty(B, L, Cc, device=x.device, dtype=torch.float32)
    check(lib().dp_overlap_add_cl_f32(ptr(x), ptr(f), B, L, K, Cc, stream_ptr()), "dp_overlap_add_cl_f32")
    return f


def split_bf16(w: torch.Tensor):
    """fp32 -> (bf16 hi, bf16 lo) with ``w ~= hi + lo``."""
    require_cuda(w, "w")
    hi = torch.empty(w.shape, device=w.device, dtype=torch.bfloat16)
    lo = torch.empty(w.shape, device=w.device, dtype=torch.bfloat16)
    check(lib().dp_split_bf16(ptr(w), ptr(hi), ptr(lo), w.numel(), stream_ptr()), "dp_split_bf16")
    return hi, lo


def linear(a, weight, bias=None, *, w_kn=False, relu=False, out=None, accumulate=False, stats=None, rows_per_group=0,
           bias_scale=1.0, precision="fp32", lda=None, rows=None):
    """``y = a @ W^T + bias`` (``nn.Linear`` / 1x1 conv).  ``weight`` is ``[N,K]`` (or ``[K,N]`` with ``w_kn``)."""
    require_cuda(a, "a")
    require_cuda(weight, "weight")
    hi, lo = split_bf16(weight.float().contiguous())
    if w_kn:
        K, N = weight.shape
    else:
        N, K = weight.shape
    M = rows if rows is not None else a.shape[0]
    lda = lda if lda is not None else a.stride(0)
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.float32)
    check(
        lib().dp_linear_f32(ptr(a), lda, ptr(hi), ptr(lo), weight.shape[1], int(w_kn), ptr(bias), float(bias_scale), ptr(out),
                            out.stride(0), M, N, K, int(relu), int(accumulate), ptr(stats), int(rows_per_group), _prec(precision),
                            stream_ptr()),
        "dp_linear_f32",
    )
    return out


def linear_wgrad(a, b, out, *, scale=1.0, precision="fp32"):
    """``out[Mo,No] += scale * a[P,Mo]^T @ b[P,No]`` (weight gradient of :func:`linear`)."""
    for t, n in ((a, "a"), (b, "b"), (out, "out")):
        require_cuda(t, n)
    P, Mo = a.shape
    No = b.shape[1]
    check(
        lib().dp_linear_wgrad_f32(ptr(a), a.stride(0), ptr(b), b.stride(0), ptr(out), out.stride(0), P, Mo, No, float(scale),
                                  _prec(precision), stream_ptr()),
        "dp_linear_wgrad_f32",
    )
    return out


class LstmPack:
    """Packed weights of one bidirectional ``nn.LSTM(64, 128)`` (gc3_basics.py:16)."""

    def __init__(self, lstm: torch.nn.LSTM):
        names = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"]
        ws = [getattr(lstm, n).detach().float().contiguous() for n in names] + [
            getattr(lstm, n + "_reverse").detach().float().contiguous() for n in names
        ]
        if tuple(ws[0].shape) != (512, 64) or tuple(ws[1].shape) != (512, 128):
            raise ValueError("the recurrent kernel is specialised for input 64 / hidden 128")
        for w in ws:
            require_cuda(w, "lstm weight")
        self.buf = torch.empty(lib().dp_lstm_pack_bytes(), device=ws[0].device, dtype=torch.uint8)
        check(lib().dp_lstm_pack(*[ptr(w) for w in ws], ptr(self.buf), stream_ptr()), "dp_lstm_pack")
        self._keep = ws


def _seq_map(layout: str, B: int, S: int, K: int):
    if layout == "intra":  # sequences (b,s) along k
        return B * S, K, 1 << 30, 0, K, 1
    if layout == "inter":  # sequences (b,k) along s
        return B * K, S, K, S * K, 1, K
    raise ValueError("layout must be 'intra' or 'inter'")


def bilstm_forward(pack: LstmPack, x: torch.Tensor, layout: str, *, save=False, precision="fp32"):
    """BiLSTM over a channels-last dual-path tensor ``x[B,S,K,64]`` along K (``intra``) or S (``inter``).

    Returns ``(H[B,S,K,256], G, Cst)``; ``G``/``Cst`` hold what :func:`bilstm_backward` needs when ``save``.
    """
    require_cuda(x, "x")
    B, S, K, N = x.shape
    P = B * S * K
    G = torch.empty(P, 1024, device=x.device, dtype=torch.float32)
    H = torch.empty(B, S, K, 256, device=x.device, dtype=torch.float32)
    Cst = torch.empty(P, 256, device=x.device, dtype=torch.float32) if save else None
    nseq, ln, qdiv, s_hi, s_lo, s_t = _seq_map(layout, B, S, K)
    check(
        lib().dp_bilstm_forward_f32(ptr(pack.buf), ptr(x), ptr(G), ptr(H), ptr(Cst), P, nseq, ln, qdiv, s_hi, s_lo, s_t, int(save),
                                    _prec(precision), stream_ptr()),
        "dp_bilstm_forward_f32",
    )
    return H, G, Cst


def bilstm_backward(pack: LstmPack, G, Cst, dH, shape, layout: str, *, precision="fp32"):
    """BPTT: ``dH[B,S,K,256]`` -> ``(dx[B,S,K,64], dbias[1024] packed)``; ``G`` becomes d(pre-activations) (packed columns)."""
    B, S, K = shape
    P = B * S * K
    require_cuda(dH, "dH")
    dx = torch.empty(B, S, K, 64, device=dH.device, dtype=torch.float32)
    dbias = torch.zeros(1024, device=dH.device, dtype=torch.float32)
    nseq, ln, qdiv, s_hi, s_lo, s_t = _seq_map(layout, B, S, K)
    check(
        lib().dp_bilstm_backward_f32(ptr(pack.buf), ptr(G), ptr(Cst), ptr(dH), ptr(dx), 0, ptr(dbias), P, nseq, ln, qdiv, s_hi, s_lo,
                                     s_t, _prec(precision), stream_ptr()),
        "dp_bilstm_backward_f32",
    )
    return dx, dbias


def groupnorm_residual(y, res, gamma, beta, stats, rows_per_group, eps, *, concat=None):
    """``res + GroupNorm(1,C)(y)`` on channels-last rows, statistics given as fp64 (sum, sumsq) per group."""
    require_cuda(y, "y")
    rows = y.numel() // y.shape[-1]
    Cc = y.shape[-1]
    groups = rows // rows_per_group
    mr = torch.empty(groups, 2, device=y.device, dtype=torch.float32)
    check(lib().dp_groupnorm_finalize(ptr(stats), ptr(mr), groups, float(rows_per_group * Cc), float(eps), stream_ptr()),
          "dp_groupnorm_finalize")
    out = torch.empty_like(y)
    cw, cb, slope = concat if concat is not None else (None, None, None)
    check(
        lib().dp_groupnorm_residual_f32(ptr(y), ptr(res), ptr(out), ptr(mr), ptr(gamma), ptr(beta), rows, rows_per_group, Cc, ptr(cw),
                                        ptr(cb), ptr(slope), stream_ptr()),
        "dp_groupnorm_residual_f32",
    )
    return out


def attention(qkv: torch.Tensor, heads: int, layout: str, *, save=False):
    """Self-attention core of ``nn.MultiheadAttention`` on a channels-last dual-path tensor ``qkv[B,S,K,3E]`` (the
    in_proj output ``[q|k|v]``), along K (``intra``) or S (``inter``).  Returns ``(o[B,S,K,E], lse or None)``."""
    require_cuda(qkv, "qkv")
    B, S, K, E3 = qkv.shape
    E = E3 // 3
    o = torch.empty(B, S, K, E, device=qkv.device, dtype=torch.float32)
    lse = torch.empty(B * S * K, heads, device=qkv.device, dtype=torch.float32) if save else None
    nseq, ln, qdiv, s_hi, s_lo, s_t = _seq_map(layout, B, S, K)
    check(lib().dp_attention_forward_f32(ptr(qkv), ptr(o), ptr(lse), E, heads, nseq, ln, qdiv, s_hi, s_lo, s_t, stream_ptr()),
          "dp_attention_forward_f32")
    return o, lse


def attention_tensor_cores(qkv: torch.Tensor, heads: int, layout: str, *, precision="fp32", save=False):
    """:func:`attention` on the warp-level tensor cores (online softmax; sequences <= 320): returns ``(o, lse)``."""
    require_cuda(qkv, "qkv")
    B, S, K, E3 = qkv.shape
    E = E3 // 3
    o = torch.empty(B, S, K, E, device=qkv.device, dtype=torch.float32)
    lse = torch.empty(B * S * K, heads, device=qkv.device, dtype=torch.float32) if save else None
    nseq, ln, qdiv, s_hi, s_lo, s_t = _seq_map(layout, B, S, K)
    check(lib().dp_attention_forward_tc_f32(ptr(qkv), ptr(o), None, None, ptr(lse), E, heads, nseq, ln, qdiv, s_hi, s_lo, s_t, _prec(precision),
                                            stream_ptr()), "dp_attention_forward_tc_f32")
    return o, lse


def attention_backward(qkv, o, lse, d_o, heads: int, layout: str, *, tensor_cores=False, precision="fp32"):
    """Backward of :func:`attention` from the saved log-sum-exp.  ``tensor_cores``: the mma.sync kernel the engines use
    (sequences <= 256; bf16x3 products in fp32 mode), otherwise the exact fp32 CUDA-core kernel."""
    require_cuda(d_o, "d_o")
    B, S, K, E3 = qkv.shape
    d_qkv = torch.empty_like(qkv)
    nseq, ln, qdiv, s_hi, s_lo, s_t = _seq_map(layout, B, S, K)
    if tensor_cores:
        check(lib().dp_attention_backward_tc_f32(ptr(qkv), ptr(o), ptr(lse), ptr(d_o), ptr(d_qkv), E3 // 3, heads, nseq, ln, qdiv, s_hi, s_lo,
                                                 s_t, _prec(precision), stream_ptr()), "dp_attention_backward_tc_f32")
        return d_qkv
    check(lib().dp_attention_backward_f32(ptr(qkv), ptr(o), ptr(lse), ptr(d_o), ptr(d_qkv), E3 // 3, heads, nseq, ln, qdiv, s_hi, s_lo, s_t,
                                          stream_ptr()), "dp_attention_backward_f32")
    return d_qkv


def add_layernorm(a, b, gamma, beta, eps, *, res=None, save_z=False):
    """``res + LayerNorm(a + b)`` on rows of ``E`` channels (``nn.LayerNorm`` after a residual add)."""
    require_cuda(a, "a")
    E = a.shape[-1]
    rows = a.numel() // E
    out = torch.empty_like(a)
    z = torch.empty_like(a) if save_z else None
    check(lib().dp_add_layernorm_f32(ptr(a), ptr(b), ptr(z), ptr(out), ptr(res), ptr(gamma), ptr(beta), rows, E, float(eps), stream_ptr()),
          "dp_add_layernorm_f32")
    return out, z


def layernorm_backward(dy, z, gamma, eps):
    require_cuda(dy, "dy")
    E = dy.shape[-1]
    rows = dy.numel() // E
    dz = torch.empty_like(dy)
    dgamma = torch.zeros(E, device=dy.device, dtype=torch.float32)
    dbeta = torch.zeros(E, device=dy.device, dtype=torch.float32)
    check(lib().dp_layernorm_backward_f32(ptr(dy), ptr(z), ptr(dz), None, ptr(gamma), rows, E, float(eps), ptr(dgamma), ptr(dbeta),
                                          stream_ptr()), "dp_layernorm_backward_f32")
    return dz, dgamma, dbeta


def split_rows(x: torch.Tensor, *, relu=False, lo=True):
    """fp32 rows ``[rows, C]`` -> bf16 (hi, lo) planes, the operand format of the TMA-fed tcgen05 GEMMs."""
    require_cuda(x, "x")
    rows, Cc = x.shape
    hi = torch.empty(rows, Cc, device=x.device, dtype=torch.bfloat16)
    lo_t = torch.empty(rows, Cc, device=x.device, dtype=torch.bfloat16) if lo else None
    check(lib().dp_split_rows_f32(ptr(x), x.stride(0), ptr(hi), ptr(lo_t), rows, Cc, int(relu), stream_ptr()), "dp_split_rows_f32")
    return hi, lo_t


def linear_planes(a_hi, a_lo, w_hi, w_lo, bias=None, *, act=0, out=None, accumulate=False, planes_out=False, bias_scale=1.0,
                  precision="fp32"):
    """``act(a @ W^T + bias)`` on pre-split operands (TMA-fed tcgen05 kernel).  Returns ``(C fp32, (C_hi, C_lo) or None)``."""
    M, K = a_hi.shape
    N = w_hi.shape[0]
    if out is None:
        out = torch.empty(M, N, device=a_hi.device, dtype=torch.float32)
    ch = cl = None
    if planes_out:
        ch = torch.empty(M, N, device=a_hi.device, dtype=torch.bfloat16)
        cl = torch.empty(M, N, device=a_hi.device, dtype=torch.bfloat16)
    check(lib().dp_linear_planes_f32(ptr(a_hi), ptr(a_lo), a_hi.stride(0), ptr(w_hi), ptr(w_lo), w_hi.stride(0), ptr(bias), float(bias_scale),
                                     ptr(out), out.stride(0), ptr(ch), ptr(cl), N, M, N, K, int(act), int(accumulate), _prec(precision),
                                     stream_ptr()), "dp_linear_planes_f32")
    return out, ((ch, cl) if planes_out else None)


def linear_wgrad_planes(a, b0, out0, b1=None, out1=None, *, tr0=False, tr1=False, scale=1.0, precision="fp32"):
    """``out0[Mo,nb0] += a^T b0`` and (same pass over ``a``) ``out1[Mo,nb1] += a^T b1``; ``a``/``b*`` are (hi, lo) plane pairs
    of ``[P, cols]`` tensors (column slices allowed)."""
    P, Mo = a[0].shape
    z = (None, None)
    b1 = b1 if b1 is not None else z
    check(lib().dp_linear_wgrad_planes_f32(ptr(a[0]), ptr(a[1]), a[0].stride(0), Mo, ptr(b0[0]), ptr(b0[1]), b0[0].stride(0), b0[0].shape[1],
                                           ptr(b1[0]), ptr(b1[1]), b1[0].stride(0) if b1[0] is not None else 0,
                                           b1[0].shape[1] if b1[0] is not None else 0, ptr(out0), out0.stride(0), int(tr0), ptr(out1),
                                           out1.stride(0) if out1 is not None else 0, int(tr1), P, float(scale), _prec(precision),
                                           stream_ptr()), "dp_linear_wgrad_planes_f32")
    return out0, out1


def attention_planes(qkv: torch.Tensor, heads: int, layout: str, *, precision="fp32", save=False):
    """Self-attention core on the tcgen05 kernel: ``qkv[B,S,K,3E]`` is split into bf16 hi/lo planes (what the QKV GEMM writes in
    the engines) and streamed by TMA.  Returns ``(o fp32 [B,S,K,E], (o_hi, o_lo), lse or None)``."""
    require_cuda(qkv, "qkv")
    B, S, K, E3 = qkv.shape
    E = E3 // 3
    hi, lo = split_rows(qkv.reshape(-1, E3).contiguous())
    o = torch.empty(B, S, K, E, device=qkv.device, dtype=torch.float32)
    oh = torch.empty(B * S * K, E, device=qkv.device, dtype=torch.bfloat16)
    ol = torch.empty(B * S * K, E, device=qkv.device, dtype=torch.bfloat16)
    lse = torch.empty(B * S * K, heads, device=qkv.device, dtype=torch.float32) if save else None
    check(lib().dp_attention_forward_planes_f32(ptr(hi), ptr(lo), ptr(o), ptr(oh), ptr(ol), ptr(lse), E, heads, int(layout == "inter"), B, S, K,
                                                _prec(precision), stream_ptr()), "dp_attention_forward_planes_f32")
    return o, (oh, ol), lse
