"""CPU oracle for the GroupComm path of ``TasNet`` (``group_size > 1``, ``module="DPRNN"`` or ``"DPTNet"``).

TEST INFRASTRUCTURE ONLY (same rule as ``oracle/dualpath_oracle.py``: only ``tests/``, ``smoke()`` and the bench's CPU legs may import
it, as the checker).  Functional restatement on plain ``torch`` CPU tensors, citing the reference lines each function follows:

* ``TAC`` (transform - average - concatenate)     look2hear/models/utils/gc3_basics.py:28-60
* ``GC_RNN`` (TAC + ProjRNN + GroupNorm per group) look2hear/models/utils/groupcomm.py:10-45
* grouped ``DPRNN`` stack                          look2hear/models/utils/dprnn.py:53-88
* grouped ``DPTNet`` stack                         look2hear/models/utils/dptnet.py:133-162
* context encoder / decoder, grouped mask          look2hear/models/gc3_network.py:59-61,145-175

Pinned by ``tests/golden/make_golden_groupcomm.py`` against the reference itself (``TasNet(..., group_size=16)``, the configuration
of the reference's ``unit_tests.py:66-86``); the golden outputs are committed under ``tests/golden/``.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .dualpath_oracle import dptnet_layer, group_norm1, merge_feature, prelu, proj_rnn, split_feature, wave_rest, _mm

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]


def tac(x: Tensor, sd: StateDict, prefix: str) -> Tensor:
    """``TAC.forward`` (gc3_basics.py:38-60).  ``x``: [B, G, n, T]."""
    B, G, n, T = x.shape
    rows = x.permute(0, 3, 1, 2).reshape(B * T * G, n)
    y = prelu(rows @ sd[prefix + "TAC_input.0.weight"].t() + sd[prefix + "TAC_input.0.bias"], sd[prefix + "TAC_input.1.weight"])
    y = y.reshape(B * T, G, -1)
    m = prelu(y.mean(1) @ sd[prefix + "TAC_mean.0.weight"].t() + sd[prefix + "TAC_mean.0.bias"], sd[prefix + "TAC_mean.1.weight"])
    cat = torch.cat([y, m.unsqueeze(1).expand_as(y)], 2).reshape(B * T * G, -1)
    o = prelu(cat @ sd[prefix + "TAC_output.0.weight"].t() + sd[prefix + "TAC_output.0.bias"], sd[prefix + "TAC_output.1.weight"])
    o = o.reshape(B, T, G, n).permute(0, 2, 3, 1).reshape(B * G, n, T)
    o = group_norm1(o, sd[prefix + "TAC_norm.weight"], sd[prefix + "TAC_norm.bias"], 1e-5)   # nn.GroupNorm default eps
    return x + o.reshape(B, G, n, T)


def gc_rnn(x: Tensor, sd: StateDict, prefix: str, G: int, layers: int, impl: str = "aten") -> Tensor:
    """``GC_RNN.forward`` (groupcomm.py:26-45).  ``x``: [B, dim, L]."""
    B, dim, L = x.shape
    out = x.reshape(B, G, dim // G, L)
    for i in range(layers):
        out = tac(out, sd, f"{prefix}TAC.{i}.").transpose(2, 3).reshape(B * G, L, dim // G)
        r = proj_rnn(out, sd, f"{prefix}rnn.{i}.", impl, _mm)
        nrm = group_norm1(r.transpose(1, 2), sd[f"{prefix}LN.{i}.weight"], sd[f"{prefix}LN.{i}.bias"], 1e-5)
        out = (out + nrm.transpose(1, 2)).reshape(B, G, L, dim // G).transpose(2, 3)
    return out.reshape(B, dim, L)


def dprnn_group_stack(x: Tensor, sd: StateDict, prefix: str, G: int, layers: int, impl: str = "aten", unfold: bool = False) -> Tensor:
    """``DPRNN.forward`` with ``num_group > 1`` (dprnn.py:53-88; the wrapper's ``num_spk`` is 1).  ``x``: [B, N, d1, d2] -> same."""
    B, N, d1, d2 = x.shape
    n = N // G
    out = x.reshape(B, G, n, d1, d2)
    for i in range(layers):
        out = tac(out.reshape(B, G, n, d1 * d2), sd, f"{prefix}TAC.{i}.").reshape(B * G, n, d1, d2)
        row_in = out.permute(0, 3, 2, 1).reshape(B * G * d2, d1, n)
        row = proj_rnn(row_in, sd, f"{prefix}row_rnn.{i}.", impl, _mm).reshape(B * G, d2, d1, n).permute(0, 3, 2, 1)
        out = out + group_norm1(row, sd[f"{prefix}row_norm.{i}.weight"], sd[f"{prefix}row_norm.{i}.bias"], 1e-8)
        col_in = out.permute(0, 2, 3, 1).reshape(B * G * d1, d2, n)
        col = proj_rnn(col_in, sd, f"{prefix}col_rnn.{i}.", impl, _mm).reshape(B * G, d1, d2, n).permute(0, 3, 1, 2)
        out = out + group_norm1(col, sd[f"{prefix}col_norm.{i}.weight"], sd[f"{prefix}col_norm.{i}.bias"], 1e-8)
        if unfold:  # depthwise 1x1 conv + PReLU at group width; the RNNs / norms / this block are one shared instance (dprnn.py:26-34,82)
            cw = sd[f"{prefix}concat_block.0.weight"].view(1, n, 1, 1)
            cb = sd[f"{prefix}concat_block.0.bias"].view(1, n, 1, 1)
            out = prelu(out * cw + cb, sd[f"{prefix}concat_block.1.weight"])
    w = sd[f"{prefix}output.weight"].reshape(-1, n)
    y = out.permute(0, 2, 3, 1).reshape(-1, n) @ w.t() + sd[f"{prefix}output.bias"]
    y = y.reshape(B, G, d1, d2, -1).permute(0, 1, 4, 2, 3)     # [B, G, n_out, d1, d2]; num_spk == 1: the transpose(1, 2) is a no-op
    return y.reshape(B, -1, d1, d2)


def dptnet_group_stack(x: Tensor, sd: StateDict, prefix: str, G: int, layers: int, impl: str = "aten", unfold: bool = False) -> Tensor:
    """``DPTNet.forward`` with ``num_group > 1`` (dptnet.py:133-162): TAC, then a transformer layer (4 heads at width n, a BiLSTM as the
    feed-forward) along the chunk and across the chunks of every group.  ``x``: [B, N, d1, d2] -> same."""
    B, N, d1, d2 = x.shape
    n = N // G
    out = x.reshape(B, G, n, d1, d2)
    for i in range(layers):
        out = tac(out.reshape(B, G, n, d1 * d2), sd, f"{prefix}TAC.{i}.").reshape(B * G, n, d1, d2)
        row_in = out.permute(0, 3, 2, 1).reshape(B * G * d2, d1, n)
        row = dptnet_layer(row_in.permute(1, 0, 2), sd, f"{prefix}row_xfmr.{i}.transformer.", impl, _mm).permute(1, 0, 2)
        out = out + row.reshape(B * G, d2, d1, n).permute(0, 3, 2, 1)
        col_in = out.permute(0, 2, 3, 1).reshape(B * G * d1, d2, n)
        col = dptnet_layer(col_in.permute(1, 0, 2), sd, f"{prefix}col_xfmr.{i}.transformer.", impl, _mm).permute(1, 0, 2)
        out = out + col.reshape(B * G, d1, d2, n).permute(0, 3, 1, 2)
        if unfold:
            cw = sd[f"{prefix}concat_block.0.weight"].view(1, n, 1, 1)
            cb = sd[f"{prefix}concat_block.0.bias"].view(1, n, 1, 1)
            out = prelu(out * cw + cb, sd[f"{prefix}concat_block.1.weight"])
    w = sd[f"{prefix}output.weight"].reshape(-1, n)
    y = out.permute(0, 2, 3, 1).reshape(-1, n) @ w.t() + sd[f"{prefix}output.bias"]
    return y.reshape(B, G, d1, d2, -1).permute(0, 1, 4, 2, 3).reshape(B, -1, d1, d2)


def tasnet_gc_forward(sd: StateDict, mixture: Tensor, *, enc_dim=64, bn_dim=64, win=16, layer=6, num_spk=2, context_size=24,
                      group_size=16, block_size=100, unfold=False, module="DPRNN", lstm_impl="aten", taps: Optional[dict] = None) -> Tensor:
    """``TasNet.forward`` with ``group_size > 1`` (gc3_network.py:133-184)."""
    was_one_d = mixture.ndim == 1
    x = mixture.unsqueeze(0) if was_one_d else mixture
    if x.ndim == 3:
        x = x.squeeze(1)
    B, T = x.shape
    G, stride = group_size, win // 2
    rest = wave_rest(T, win)
    xp = F.pad(x, (stride, rest + stride))
    frames = xp.unfold(1, win, stride)
    L = frames.shape[1]
    enc = (frames.reshape(B * L, win) @ sd["encoder.weight"].reshape(enc_dim, win).t()).reshape(B, L, enc_dim).permute(0, 2, 1)
    g = group_norm1(enc, sd["bottleneck.0.weight"], sd["bottleneck.0.bias"], float(torch.finfo(torch.float32).eps))
    feat = (g.permute(0, 2, 1).reshape(B * L, enc_dim) @ sd["bottleneck.1.weight"].reshape(bn_dim, enc_dim).t())
    feat = feat.reshape(B, L, bn_dim).permute(0, 2, 1)
    # context encoding (gc3_network.py:145-151)
    blk, crest = split_feature(feat, context_size)            # [B, N, ctx, Lc]
    Lc = blk.shape[-1]
    sq_in = blk.permute(0, 3, 1, 2).reshape(B * Lc, bn_dim, context_size)
    sq = gc_rnn(sq_in, sd, "context_enc.", G, 2, lstm_impl)
    sq_mean = sq.mean(2).reshape(B, Lc, bn_dim).transpose(1, 2)
    # sequence modelling: DP_Wrapper (groupcomm.py:100-114)
    dblk, drest = split_feature(sq_mean, block_size)
    stack = dprnn_group_stack if module == "DPRNN" else dptnet_group_stack
    dp = stack(dblk, sd, "seq_model.seq_model.", G, layer, lstm_impl, unfold)
    fmap = merge_feature(dp, drest).reshape(B, -1, Lc)
    # context decoding (gc3_network.py:160-166)
    fm = fmap.unsqueeze(2) + blk
    fm = fm.permute(0, 3, 1, 2).reshape(B * Lc, bn_dim, context_size)
    un = gc_rnn(fm, sd, "context_dec.", G, 2, lstm_impl).reshape(B, Lc, bn_dim, context_size).permute(0, 2, 3, 1)
    un = merge_feature(un, crest)                              # [B, N, L]
    # grouped mask (gc3_network.py:169-175)
    n, eg = bn_dim // G, enc_dim // G
    w_m = sd["mask.0.weight"].reshape(eg * num_spk, n)
    m = torch.relu(un.reshape(B * G, n, L).permute(0, 2, 1).reshape(-1, n) @ w_m.t() + sd["mask.0.bias"])
    m = m.reshape(B, G, L, num_spk, eg).permute(0, 3, 1, 4, 2).reshape(B, num_spk, enc_dim, L)
    masked = m * enc.unsqueeze(1)
    w_dec = sd["decoder.weight"].reshape(enc_dim, win)
    fr = (masked.permute(0, 1, 3, 2).reshape(B * num_spk * L, enc_dim) @ w_dec).reshape(B * num_spk, L, win)
    wav = fr.new_zeros(B * num_spk, (L - 1) * stride + win)
    half = fr.reshape(B * num_spk, L, 2, stride)
    wav[:, : L * stride] += half[:, :, 0].reshape(B * num_spk, L * stride)
    wav[:, stride:] += half[:, :, 1].reshape(B * num_spk, L * stride)
    out = wav[:, stride : wav.shape[1] - (rest + stride)].reshape(B, num_spk, T)
    if taps is not None:
        taps.update(enc_output=enc, enc_feature=feat, squeeze_mean=sq_mean, dp_out=dp, feature_map=fmap, unsqueeze_output=un, mask=m)
    return out.squeeze(0) if was_one_d else out
