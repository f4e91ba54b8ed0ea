"""CPU oracle for the look2hear dual-path separation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the shipped
product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker (or as the timed CPU reference arm), never as a fallback for the CUDA
path.

It is a functional restatement (plain ``torch`` CPU tensors, no ``nn.Module``)
of the reference algorithm, written from the reference's behaviour and citing
the file:line each function follows (paths relative to the reference repo):

* ``pad_input`` / framing         look2hear/models/gc3_network.py:108-131
* encoder / bottleneck / mask / decoder
                                  look2hear/models/gc3_network.py:133-184
* ``split_feature``               look2hear/models/utils/gc3_basics.py:63-91
* ``merge_feature``               look2hear/models/utils/gc3_basics.py:94-109
* ``ProjRNN`` (BiLSTM + Linear)   look2hear/models/utils/gc3_basics.py:7-24
* ``DPRNN`` stack                 look2hear/models/utils/dprnn.py:53-88
* ``DPTNet`` stack                look2hear/models/utils/dptnet.py:26-162
* ``DP_Wrapper``                  look2hear/models/utils/groupcomm.py:100-114
* ``PairwiseNegSDR``              look2hear/losses/matrix.py:13-57
* ``PITLossWrapper``              look2hear/losses/pit_wrapper.py:30-131

The heavy arithmetic of the reference lives in PyTorch itself (``nn.LSTM``,
``nn.GroupNorm``, ``nn.Conv1d`` ... pinned by the reference at torch==1.11.0,
``env.yaml:24``; semantics unchanged in the torch 2.11 of this image).  The
restatement spells those primitives out (gate order i,f,g,o; biased variance;
...) so the algorithm is visible; ``lstm_impl="aten"`` swaps the explicit time
loop for ``torch.lstm`` (the very op ``nn.LSTM`` dispatches to) so the oracle
can be timed as the reference's CPU path.

Parity pinning: the reference has no tests or golden vectors of its own
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself, imported in the build container: ``tests/golden/make_golden.py``
generates ``tests/golden/*.npz`` from ``/root/reference`` and asserts this file
agrees with it; ``tests/test_oracle_golden.py`` re-checks the oracle against the
committed vectors on every run.

Everything is differentiable through torch autograd, so the same functions serve
as the gradient oracle.
"""
from __future__ import annotations

import itertools
import math
from typing import Callable, Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]

# ----------------------------------------------------------------------------
# geometry (integer index work; also restated in C in oracle/seg_index.c)
# ----------------------------------------------------------------------------


def wave_rest(T: int, win: int) -> int:
    """gc3_network.py:123 -- samples appended so frames tile the waveform."""
    stride = win // 2
    return win - (stride + T % win) % win


def num_frames(T: int, win: int) -> int:
    """L = (T + rest + 2*stride - win)/stride + 1."""
    stride = win // 2
    return (T + wave_rest(T, win) + 2 * stride - win) // stride + 1


def seg_rest(L: int, K: int) -> int:
    """gc3_basics.py:68 -- frames appended before chunking."""
    P = K // 2
    return K - (P + L % K) % K


def num_chunks(L: int, K: int) -> int:
    """gc3_basics.py:84-89 -- S = 2 * (Lp - P) / K with Lp = L + rest + 2P."""
    P = K // 2
    Lp = L + seg_rest(L, K) + 2 * P
    return 2 * ((Lp - P) // K)


def split_feature(x: Tensor, K: int) -> Tuple[Tensor, int]:
    """50%-overlap chunking ``[B,N,L] -> [B,N,K,S]`` (gc3_basics.py:63-91).

    Closed form: ``out[b,n,k,s] = x[b,n,(s-1)*P + k]`` when that index lies in
    ``[0,L)`` and 0 otherwise (chunk ``s`` starts at padded offset ``s*P`` and
    the padded signal has ``P`` leading zeros).
    """
    B, N, L = x.shape
    P = K // 2
    rest = seg_rest(L, K)
    S = num_chunks(L, K)
    k = torch.arange(K, device=x.device).view(K, 1)
    s = torch.arange(S, device=x.device).view(1, S)
    src = (s - 1) * P + k  # [K,S]
    valid = (src >= 0) & (src < L)
    gathered = x[:, :, src.clamp(0, max(L - 1, 0)).reshape(-1)].reshape(B, N, K, S)
    out = torch.where(valid.view(1, 1, K, S), gathered, torch.zeros((), dtype=x.dtype, device=x.device))
    return out.contiguous(), rest


def merge_feature(y: Tensor, rest: int) -> Tensor:
    """Overlap-add ``[B,N,K,S] -> [B,N,L]`` (gc3_basics.py:94-109).

    ``out[t] = y[(t+P)%K, 2*floor((t+P)/K)] + y[t%K, 2*floor(t/K)+1]`` for
    ``t in [0, (S/2)*K - P - rest)``.
    """
    B, N, K, S = y.shape
    P = K // 2
    L = (S // 2) * K - P - rest
    t = torch.arange(L, device=y.device)
    k1, s1 = (t + P) % K, 2 * ((t + P) // K)
    k2, s2 = t % K, 2 * (t // K) + 1
    flat = y.reshape(B, N, K * S)
    return (flat[:, :, k1 * S + s1] + flat[:, :, k2 * S + s2]).contiguous()


# ----------------------------------------------------------------------------
# primitives the reference takes from torch.nn
# ----------------------------------------------------------------------------

MatMul = Callable[[Tensor, Tensor], Tensor]


def _mm(a: Tensor, b: Tensor) -> Tensor:
    return a @ b


def group_norm1(x: Tensor, weight: Tensor, bias: Tensor, eps: float) -> Tensor:
    """``nn.GroupNorm(1, C, eps)``: per-sample mean / biased variance over all of
    ``C x spatial``; per-channel affine."""
    B, C = x.shape[:2]
    flat = x.reshape(B, -1)
    mean = flat.mean(dim=1, keepdim=True)
    var = ((flat - mean) ** 2).mean(dim=1, keepdim=True)
    xhat = ((flat - mean) / torch.sqrt(var + eps)).reshape(x.shape)
    shape = [1, C] + [1] * (x.ndim - 2)
    return xhat * weight.view(shape) + bias.view(shape)


def layer_norm(x: Tensor, weight: Tensor, bias: Tensor, eps: float) -> Tensor:
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps) * weight + bias


def lstm_direction(
    x: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor, reverse: bool, mm: MatMul = _mm
) -> Tensor:
    """One direction of a 1-layer ``nn.LSTM`` with zero initial state.

    ``x`` is ``[Nb, L, I]`` (batch_first).  Gate rows are ordered i, f, g, o;
    ``gates = x_t W_ih^T + b_ih + h_{t-1} W_hh^T + b_hh``.
    """
    Nb, L, _ = x.shape
    H = w_hh.shape[1]
    xg = mm(x.reshape(Nb * L, -1), w_ih.t()).reshape(Nb, L, 4 * H) + (b_ih + b_hh)
    h = x.new_zeros(Nb, H)
    c = x.new_zeros(Nb, H)
    outs = [None] * L
    steps = range(L - 1, -1, -1) if reverse else range(L)
    for t in steps:
        gates = xg[:, t] + mm(h, w_hh.t())
        i, f, g, o = gates.split(H, dim=1)
        i, f, o = torch.sigmoid(i), torch.sigmoid(f), torch.sigmoid(o)
        g = torch.tanh(g)
        c = f * c + i * g
        h = o * torch.tanh(c)
        outs[t] = h
    return torch.stack(outs, dim=1)


def bilstm(x: Tensor, sd: StateDict, prefix: str, impl: str = "aten", mm: MatMul = _mm, batch_first: bool = True) -> Tensor:
    """Bidirectional 1-layer LSTM, output ``[.., 2H]`` = ``[h_fwd | h_bwd]``."""
    names = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"]
    fw = [sd[prefix + n] for n in names]
    bw = [sd[prefix + n + "_reverse"] for n in names]
    if impl == "aten":
        H = fw[1].shape[1]
        nb = x.shape[0] if batch_first else x.shape[1]
        zeros = x.new_zeros(2, nb, H)
        # the "train" flag only selects cuDNN's training kernels (needed for backward on CUDA); dropout is 0 on this path
        train = torch.is_grad_enabled() and any(t.requires_grad for t in fw + bw + [x])
        out, _, _ = torch.lstm(x, (zeros, zeros), fw + bw, True, 1, 0.0, train, True, batch_first)
        return out
    xb = x if batch_first else x.transpose(0, 1)
    out = torch.cat([lstm_direction(xb, *fw, reverse=False, mm=mm), lstm_direction(xb, *bw, reverse=True, mm=mm)], dim=2)
    return out if batch_first else out.transpose(0, 1)


def proj_rnn(x: Tensor, sd: StateDict, prefix: str, impl: str, mm: MatMul) -> Tensor:
    """``ProjRNN.forward`` (gc3_basics.py:19-24): BiLSTM then Linear(2H -> I)."""
    h = bilstm(x, sd, prefix + "rnn.", impl, mm)
    y = mm(h.reshape(-1, h.shape[2]), sd[prefix + "proj.weight"].t()) + sd[prefix + "proj.bias"]
    return y.reshape(x.shape)


def prelu(x: Tensor, slope: Tensor) -> Tensor:
    return torch.where(x >= 0, x, slope * x)


# ----------------------------------------------------------------------------
# dual-path stacks
# ----------------------------------------------------------------------------


def dprnn_stack(x: Tensor, sd: StateDict, prefix: str, layers: int, unfold: bool, impl: str, mm: MatMul) -> Tensor:
    """``DPRNN.forward`` for ``num_group == 1`` (dprnn.py:53-88).  ``x``: [B,N,K,S]."""
    B, N, K, S = x.shape
    out = x
    for i in range(layers):
        # intra-chunk: sequences along K, one per (b, s)          dprnn.py:67-73
        row_in = out.permute(0, 3, 2, 1).reshape(B * S, K, N)
        row = proj_rnn(row_in, sd, f"{prefix}row_rnn.{i}.", impl, mm)
        row = row.reshape(B, S, K, N).permute(0, 3, 2, 1)
        row = group_norm1(row, sd[f"{prefix}row_norm.{i}.weight"], sd[f"{prefix}row_norm.{i}.bias"], 1e-8)
        out = out + row
        # inter-chunk: sequences along S, one per (b, k)          dprnn.py:76-82
        col_in = out.permute(0, 2, 3, 1).reshape(B * K, S, N)
        col = proj_rnn(col_in, sd, f"{prefix}col_rnn.{i}.", impl, mm)
        col = col.reshape(B, K, S, N).permute(0, 3, 1, 2)
        col = group_norm1(col, sd[f"{prefix}col_norm.{i}.weight"], sd[f"{prefix}col_norm.{i}.bias"], 1e-8)
        out = out + col
        if unfold:  # depthwise 1x1 conv + PReLU                    dprnn.py:31-34,82
            w = sd[f"{prefix}concat_block.0.weight"].view(1, N, 1, 1)
            b = sd[f"{prefix}concat_block.0.bias"].view(1, N, 1, 1)
            out = prelu(out * w + b, sd[f"{prefix}concat_block.1.weight"])
    # final 1x1 Conv2d                                               dprnn.py:85
    w = sd[f"{prefix}output.weight"].reshape(-1, N)
    y = mm(out.permute(0, 2, 3, 1).reshape(-1, N), w.t()) + sd[f"{prefix}output.bias"]
    return y.reshape(B, K, S, -1).permute(0, 3, 1, 2)


def mha_self(x: Tensor, sd: StateDict, prefix: str, heads: int, mm: MatMul) -> Tensor:
    """``nn.MultiheadAttention`` self-attention, seq-first ``[L,Nb,E]``, no mask."""
    L, Nb, E = x.shape
    d = E // heads
    qkv = mm(x.reshape(L * Nb, E), sd[prefix + "in_proj_weight"].t()) + sd[prefix + "in_proj_bias"]
    q, k, v = qkv.reshape(L, Nb, 3, heads, d).permute(2, 1, 3, 0, 4)  # each [Nb,h,L,d]
    att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(d), dim=-1)
    o = (att @ v).permute(2, 0, 1, 3).reshape(L * Nb, E)
    o = mm(o, sd[prefix + "out_proj.weight"].t()) + sd[prefix + "out_proj.bias"]
    return o.reshape(L, Nb, E)


def dptnet_layer(x: Tensor, sd: StateDict, prefix: str, impl: str, mm: MatMul) -> Tensor:
    """``TransformerEncoderLayer.forward`` (dptnet.py:66-82), seq-first input."""
    src = x + mha_self(x, sd, prefix + "self_attn.", 4, mm)
    src = layer_norm(src, sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"], 1e-5)
    h = torch.relu(bilstm(src, sd, prefix + "linear1.", impl, mm, batch_first=False))
    L, Nb, _ = h.shape
    src2 = (mm(h.reshape(L * Nb, -1), sd[prefix + "linear2.weight"].t()) + sd[prefix + "linear2.bias"]).reshape(L, Nb, -1)
    src = src + src2
    return layer_norm(src, sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"], 1e-5)


def dptnet_stack(x: Tensor, sd: StateDict, prefix: str, layers: int, unfold: bool, impl: str, mm: MatMul) -> Tensor:
    """``DPTNet.forward`` for ``num_group == 1`` (dptnet.py:133-162)."""
    B, N, K, S = x.shape
    out = x
    for i in range(layers):
        row_in = out.permute(0, 3, 2, 1).reshape(B * S, K, N)
        row = dptnet_layer(row_in.permute(1, 0, 2), sd, f"{prefix}row_xfmr.{i}.transformer.", impl, mm).permute(1, 0, 2)
        out = out + row.reshape(B, S, K, N).permute(0, 3, 2, 1)
        col_in = out.permute(0, 2, 3, 1).reshape(B * K, S, N)
        col = dptnet_layer(col_in.permute(1, 0, 2), sd, f"{prefix}col_xfmr.{i}.transformer.", impl, mm).permute(1, 0, 2)
        out = out + col.reshape(B, K, S, N).permute(0, 3, 1, 2)
        if unfold:
            w = sd[f"{prefix}concat_block.0.weight"].view(1, N, 1, 1)
            b = sd[f"{prefix}concat_block.0.bias"].view(1, N, 1, 1)
            out = prelu(out * w + b, sd[f"{prefix}concat_block.1.weight"])
    w = sd[f"{prefix}output.weight"].reshape(-1, N)
    y = mm(out.permute(0, 2, 3, 1).reshape(-1, N), w.t()) + sd[f"{prefix}output.bias"]
    return y.reshape(B, K, S, -1).permute(0, 3, 1, 2)


# ----------------------------------------------------------------------------
# TasNet (gc3_network.py) for group_size == 1
# ----------------------------------------------------------------------------


def tasnet_forward(
    sd: StateDict,
    mixture: Tensor,
    *,
    enc_dim: int = 64,
    bn_dim: int = 64,
    win: int = 16,
    layer: int = 6,
    num_spk: int = 2,
    module: str = "DPRNN",
    block_size: int = 100,
    unfold: bool = False,
    lstm_impl: str = "aten",
    mm: MatMul = _mm,
    taps: Optional[dict] = None,
) -> Tensor:
    """``TasNet.forward`` (gc3_network.py:133-184), ``group_size == 1``.

    ``taps`` (optional dict) receives intermediate tensors for kernel-level
    parity tests: ``enc_output``, ``enc_feature``, ``blocks``, ``dp_out``,
    ``feature_map``, ``mask``.
    """
    was_one_d = mixture.ndim == 1
    x = mixture.unsqueeze(0) if was_one_d else mixture
    if x.ndim == 3:
        x = x.squeeze(1)
    B, T = x.shape
    stride = win // 2
    rest = wave_rest(T, win)
    xp = F.pad(x, (stride, rest + stride))  # gc3_network.py:124-129
    # encoder: Conv1d(1, enc_dim, win, stride, bias=False), no nonlinearity   :140
    frames = xp.unfold(1, win, stride)  # [B, L, win]
    L = frames.shape[1]
    w_enc = sd["encoder.weight"].reshape(enc_dim, win)
    enc = mm(frames.reshape(B * L, win), w_enc.t()).reshape(B, L, enc_dim).permute(0, 2, 1)  # [B,N,L]
    # bottleneck: GroupNorm(1, N, eps=finfo.eps) + 1x1 conv without bias       :53-56,142
    g = group_norm1(enc, sd["bottleneck.0.weight"], sd["bottleneck.0.bias"], float(torch.finfo(torch.float32).eps))
    w_bn = sd["bottleneck.1.weight"].reshape(bn_dim, enc_dim)
    feat = mm(g.permute(0, 2, 1).reshape(B * L, enc_dim), w_bn.t()).reshape(B, L, bn_dim).permute(0, 2, 1)
    # DP_Wrapper (groupcomm.py:100-114), num_spk=1 inside the wrapper
    blocks, srest = split_feature(feat, block_size)
    pfx = "seq_model.seq_model."
    stack = dprnn_stack if module == "DPRNN" else dptnet_stack
    dp = stack(blocks, sd, pfx, layer, unfold, lstm_impl, mm)
    fmap = merge_feature(dp, srest)  # [B, bn_dim, L]
    # mask: Conv1d(bn_dim, enc_dim*num_spk, 1) + ReLU; applied to the RAW encoder output  :169-174
    w_m = sd["mask.0.weight"].reshape(enc_dim * num_spk, bn_dim)
    m = torch.relu(mm(fmap.permute(0, 2, 1).reshape(B * L, bn_dim), w_m.t()) + sd["mask.0.bias"])
    m = m.reshape(B, L, num_spk, enc_dim).permute(0, 2, 3, 1)  # [B,C,N,L]
    masked = m * enc.unsqueeze(1)
    # decoder: ConvTranspose1d(enc_dim, 1, win, stride, bias=False) = per-frame matvec + stride OLA  :177
    w_dec = sd["decoder.weight"].reshape(enc_dim, win)
    fr = mm(masked.permute(0, 1, 3, 2).reshape(B * num_spk * L, enc_dim), w_dec).reshape(B * num_spk, L, win)
    wav = fr.new_zeros(B * num_spk, (L - 1) * stride + win)
    half = fr.reshape(B * num_spk, L, 2, stride)
    wav[:, : L * stride] += half[:, :, 0].reshape(B * num_spk, L * stride)
    wav[:, stride:] += half[:, :, 1].reshape(B * num_spk, L * stride)
    out = wav[:, stride : wav.shape[1] - (rest + stride)].reshape(B, num_spk, T)  # :178-179
    if taps is not None:
        taps.update(enc_output=enc, enc_feature=feat, blocks=blocks, dp_out=dp, feature_map=fmap, mask=m)
    return out.squeeze(0) if was_one_d else out


# ----------------------------------------------------------------------------
# losses
# ----------------------------------------------------------------------------


def pairwise_neg_sdr(ests: Tensor, targets: Tensor, sdr_type: str = "sisdr", zero_mean: bool = True, eps: float = 1e-8) -> Tensor:
    """``PairwiseNegSDR.forward`` (losses/matrix.py:22-57) -> ``[B, n_est, n_tgt]``."""
    if targets.size() != ests.size() or targets.ndim != 3:
        raise TypeError(f"Inputs must be of shape [batch, n_src, time], got {targets.size()} and {ests.size()} instead")
    if zero_mean:
        targets = targets - targets.mean(dim=2, keepdim=True)
        ests = ests - ests.mean(dim=2, keepdim=True)
    t = targets.unsqueeze(1)  # [B,1,n,T]
    e = ests.unsqueeze(2)  # [B,n,1,T]
    if sdr_type in ("sisdr", "sdsdr"):
        dot = (e * t).sum(dim=3, keepdim=True)
        energy = (t**2).sum(dim=3, keepdim=True) + eps
        proj = dot * t / energy
    else:
        proj = t.expand(-1, ests.shape[1], -1, -1)
    noise = e - t if sdr_type in ("sdsdr", "snr") else e - proj
    sdr = (proj**2).sum(dim=3) / ((noise**2).sum(dim=3) + eps)
    return -10.0 * torch.log10(sdr + eps)


def pit_loss(
    ests: Tensor, targets: Tensor, sdr_type: str = "sisdr", threshold_byloss: bool = True, return_ests: bool = False
):
    """``PITLossWrapper(pairwise, pit_from="pw_mtx")`` (pit_wrapper.py:30-67,96-131).

    Pair matrix is ``[b, est, tgt]``; permutation ``p`` assigns estimate ``p[i]``
    to target ``i``; ties resolve to the first permutation in
    ``itertools.permutations`` order.  Returns ``loss`` or ``(loss, reordered,
    perm_indices)``.
    """
    pw = pairwise_neg_sdr(ests, targets, sdr_type)
    n = pw.shape[-1]
    perms = list(itertools.permutations(range(n)))
    pwl = pw.transpose(-1, -2)  # [b, tgt, est]
    loss_set = torch.stack([sum(pwl[:, i, p[i]] for i in range(n)) / n for p in perms], dim=1)
    min_loss, idx = torch.min(loss_set, dim=1)
    kept = min_loss
    if threshold_byloss:
        sel = min_loss > -30
        if bool(sel.any()):
            kept = min_loss[sel]
    loss = kept.mean()
    if not return_ests:
        return loss
    perm_idx = torch.tensor(perms, dtype=torch.long)[idx]  # [B,n]
    reordered = torch.stack([e[p] for e, p in zip(ests, perm_idx)])
    return loss, reordered, perm_idx


# ----------------------------------------------------------------------------
# reference training-step semantics (audio_litmodule.py:73-88, audio_train.py:48,128)
# ----------------------------------------------------------------------------


def adam_clip_step(params, grads, exp_avg, exp_avg_sq, step: int, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, max_norm=5.0):
    """``clip_grad_norm_(max_norm)`` followed by ``torch.optim.Adam`` (wd 0), in place.

    Restates torch semantics: ``clip_coef = min(1, max_norm / (total_norm + 1e-6))``;
    ``denom = sqrt(v) / sqrt(1 - b2^t) + eps``; ``p -= lr / (1 - b1^t) * m / denom``.
    """
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).item()
    coef = min(1.0, max_norm / (total + 1e-6))
    b1, b2 = betas
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        g = g * coef
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = v.sqrt() / math.sqrt(1 - b2**step) + eps
        p.addcdiv_(m, denom, value=-lr / (1 - b1**step))
    return total
