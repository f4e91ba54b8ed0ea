"""CPU oracle for the SepFormer dual-path transformer (look2hear/models/sepformer.py).

TEST INFRASTRUCTURE ONLY (same rules as ``dualpath_oracle.py``: imported by ``tests/``, ``__graft_entry__.smoke()`` and
the CPU legs of ``bench.py`` only, never by the product path).

Functional restatement with plain ``torch`` tensors, citing the reference file:line each function follows (paths
relative to the reference repo).  The primitives the reference takes from ``torch.nn`` (``nn.MultiheadAttention``,
``nn.LayerNorm``, ``nn.GroupNorm``, ``nn.Linear``, ``nn.Conv1d`` ...) are spelled out so the algorithm is visible.

Parity pinning: the reference has no tests or golden vectors; ``tests/golden/make_golden.py``
imports the real ``look2hear.models.Sepformer`` in the build container, asserts this file reproduces it (rel-L2 <= 2e-6)
and commits the vectors; ``tests/test_oracle_golden.py`` re-checks on every run.

Covered configuration space: everything ``configs/sepformer_base.yml`` uses (``norm_before`` True or False,
positional encoding on/off, gLN, non-causal, eval mode / dropout 0).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from .dualpath_oracle import group_norm1, layer_norm, merge_feature, prelu, split_feature

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]


def positional_encoding(L: int, E: int, dtype=torch.float32) -> Tensor:
    """``PositionalEncoding`` (sepformer.py:61-80): pe[pos,2i] = sin(pos*w_i), pe[pos,2i+1] = cos(pos*w_i).

    Only used when the state dict carries no ``pe`` buffer; the model path always reads the registered buffer.
    """
    pe = torch.zeros(L, E)
    pos = torch.arange(0, L).unsqueeze(1).float()
    den = torch.exp(torch.arange(0, E, 2).float() * -(math.log(10000.0) / E))
    pe[:, 0::2] = torch.sin(pos * den)
    pe[:, 1::2] = torch.cos(pos * den)
    return pe.to(dtype)


def mha(x: Tensor, sd: StateDict, prefix: str, heads: int, drop=None) -> Tensor:
    """``nn.MultiheadAttention`` self-attention on batch-first ``[Nb, L, E]`` (the wrapper at sepformer.py:83-215
    permutes to seq-first and back), no mask; the averaged weights it also returns are discarded (sepformer.py:554)."""
    Nb, L, E = x.shape
    d = E // heads
    qkv = x.reshape(Nb * L, E) @ sd[prefix + "in_proj_weight"].t() + sd[prefix + "in_proj_bias"]
    q, k, v = qkv.reshape(Nb, L, 3, heads, d).permute(2, 0, 3, 1, 4)  # each [Nb, h, L, d]
    att = torch.softmax((q * (1.0 / math.sqrt(d))) @ k.transpose(-1, -2), dim=-1)
    if drop is not None:
        att = drop(0, att)  # nn.MultiheadAttention(dropout=p): dropout on the probabilities (sepformer.py:124-128)
    o = (att @ v).permute(0, 2, 1, 3).reshape(Nb * L, E)
    o = o @ sd[prefix + "out_proj.weight"].t() + sd[prefix + "out_proj.bias"]
    return o.reshape(Nb, L, E)


def encoder_layer(x: Tensor, sd: StateDict, prefix: str, heads: int, norm_before: bool, drop=None) -> Tensor:
    """``TransformerEncoderLayer.forward`` (sepformer.py:320-370).  ``drop(site, tensor)`` (optional) applies the training-time
    dropout of site 0 (attention probabilities), 1 (``dropout1``, :355), 2 (FFN hidden, :261), 3 (``dropout2``, :366) with an
    explicit mask; ``None`` = eval."""
    ident = (lambda site, t: t) if drop is None else drop
    src1 = layer_norm(x, sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"], 1e-6) if norm_before else x
    x = x + ident(1, mha(src1, sd, prefix + "self_att.att.", heads, drop))
    if not norm_before:
        x = layer_norm(x, sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"], 1e-6)
    src1 = layer_norm(x, sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"], 1e-6) if norm_before else x
    h = ident(2, torch.relu(src1 @ sd[prefix + "pos_ffn.ffn.0.weight"].t() + sd[prefix + "pos_ffn.ffn.0.bias"]))  # sepformer.py:258-263
    x = x + ident(3, h @ sd[prefix + "pos_ffn.ffn.3.weight"].t() + sd[prefix + "pos_ffn.ffn.3.bias"])
    if not norm_before:
        x = layer_norm(x, sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"], 1e-6)
    return x


def transformer_block(x: Tensor, sd: StateDict, prefix: str, layers: int, heads: int, norm_before: bool, use_pe: bool, drop=None) -> Tensor:
    """``TransformerBlock.forward`` (sepformer.py:541-556) + ``TransformerEncoder.forward`` (:438-467): PE added once,
    ``layers`` encoder layers, final LayerNorm."""
    if use_pe:
        x = x + sd[prefix + "pos_enc.pe"][:, : x.shape[1]]
    for l in range(layers):
        x = encoder_layer(x, sd, f"{prefix}mdl.layers.{l}.", heads, norm_before, None if drop is None else (lambda site, t, l=l: drop(l, site, t)))
    return layer_norm(x, sd[prefix + "mdl.norm.weight"], sd[prefix + "mdl.norm.bias"], 1e-6)


def global_ln(x: Tensor, gamma: Tensor, beta: Tensor, eps: float = 1e-8) -> Tensor:
    """``GlobalLN`` (models/utils/normalizations.py:17-47): statistics over all non-batch dims, per-channel affine."""
    return group_norm1(x, gamma, beta, eps)


def dual_block(x: Tensor, sd: StateDict, prefix: str, cfg: dict, drop=None) -> Tensor:
    """``Dual_Computation_Block.forward`` (sepformer.py:600-642).  ``x``: [B, N, K, S].  ``drop(path, layer, site, tensor)``: see
    :func:`encoder_layer` (path 0 = intra, rows ``(b, s)`` along ``k``; 1 = inter, rows ``(b, k)`` along ``s``)."""
    B, N, K, S = x.shape
    intra = x.permute(0, 3, 2, 1).reshape(B * S, K, N)
    intra = transformer_block(intra, sd, prefix + "intra_mdl.", cfg["intra_numlayers"], cfg["intra_nhead"], cfg["intra_norm_before"],
                              cfg["intra_use_positional"], None if drop is None else (lambda l, site, t: drop(0, l, site, t)))
    intra = intra.reshape(B, S, K, N).permute(0, 3, 2, 1)
    intra = global_ln(intra, sd[prefix + "intra_norm.gamma"], sd[prefix + "intra_norm.beta"]) + x
    inter = intra.permute(0, 2, 3, 1).reshape(B * K, S, N)
    inter = transformer_block(inter, sd, prefix + "inter_mdl.", cfg["inter_numlayers"], cfg["inter_nhead"], cfg["inter_norm_before"],
                              cfg["inter_use_positional"], None if drop is None else (lambda l, site, t: drop(1, l, site, t)))
    inter = inter.reshape(B, K, S, N).permute(0, 3, 1, 2)
    return global_ln(inter, sd[prefix + "inter_norm.gamma"], sd[prefix + "inter_norm.beta"]) + intra


DEFAULTS = dict(encoder_kernel_size=16, encoder_in_nchannels=1, encoder_out_nchannels=256, masknet_chunksize=250, masknet_numlayers=2,
                masknet_norm="gLN", masknet_numspks=2, intra_numlayers=8, inter_numlayers=8, intra_nhead=8, inter_nhead=8, intra_dffn=1024,
                inter_dffn=1024, intra_use_positional=True, inter_use_positional=True, intra_norm_before=True, inter_norm_before=True,
                intra_causal=False, inter_causal=False)


def sepformer_forward(sd: StateDict, mix: Tensor, taps: Optional[dict] = None, dropout=None, **config) -> Tensor:
    """``Sepformer.forward`` (sepformer.py:986-1016) with ``Dual_Path_Model.forward`` (:706-760).

    ``dropout(block, path, layer, site, tensor)`` (optional): training-time dropout with explicit masks, see :func:`encoder_layer`."""
    cfg = dict(DEFAULTS, **config)
    assert cfg["masknet_norm"] == "gLN" and not cfg["intra_causal"] and not cfg["inter_causal"] and cfg["encoder_in_nchannels"] == 1
    was_one_d = mix.ndim == 1
    x = mix.unsqueeze(0) if was_one_d else mix
    if x.ndim == 3:
        x = x.squeeze(1)
    B, T = x.shape
    win, N, spk, K = cfg["encoder_kernel_size"], cfg["encoder_out_nchannels"], cfg["masknet_numspks"], cfg["masknet_chunksize"]
    stride = win // 2
    # Encoder: Conv1d(1, N, win, stride, bias=False) + ReLU, no padding            sepformer.py:23-40
    frames = x.unfold(1, win, stride)  # [B, L, win]
    L = frames.shape[1]
    mix_w = torch.relu(frames.reshape(B * L, win) @ sd["encoder.conv1d.weight"].reshape(N, win).t()).reshape(B, L, N).permute(0, 2, 1)
    # masknet                                                                      sepformer.py:725-731
    h = group_norm1(mix_w, sd["masknet.norm.weight"], sd["masknet.norm.bias"], 1e-8)
    h = (h.permute(0, 2, 1).reshape(B * L, N) @ sd["masknet.conv1d.weight"].reshape(N, N).t()).reshape(B, L, N).permute(0, 2, 1)
    blocks, gap = split_feature(h, K)  # _Segmentation == split_feature (verified bit-equal)
    for j in range(cfg["masknet_numlayers"]):
        blocks = dual_block(blocks, sd, f"masknet.dual_mdl.{j}.", cfg, None if dropout is None else (lambda path, l, site, t, j=j: dropout(j, path, l, site, t)))
    y = prelu(blocks, sd["masknet.prelu.weight"])                                   # :736
    Kc, S = y.shape[2], y.shape[3]
    y = y.permute(0, 2, 3, 1).reshape(-1, N) @ sd["masknet.conv2d.weight"].reshape(N * spk, N).t() + sd["masknet.conv2d.bias"]  # :739
    y = y.reshape(B, Kc, S, spk, N).permute(0, 3, 4, 1, 2).reshape(B * spk, N, Kc, S)   # :743
    y = merge_feature(y, gap)  # _over_add == merge_feature                          :746
    yl = y.permute(0, 2, 1).reshape(B * spk * L, N)
    out = torch.tanh(yl @ sd["masknet.output.0.weight"].reshape(N, N).t() + sd["masknet.output.0.bias"])
    gate = torch.sigmoid(yl @ sd["masknet.output_gate.0.weight"].reshape(N, N).t() + sd["masknet.output_gate.0.bias"])
    m = torch.relu((out * gate) @ sd["masknet.end_conv1x1.weight"].reshape(N, N).t())   # :747-755
    est_mask = m.reshape(B, spk, L, N).permute(1, 0, 3, 2)  # [spk, B, N, L]          :754-758
    sep_h = mix_w.unsqueeze(0) * est_mask                                               # :997-998
    # decoder ConvTranspose1d(N, 1, win, stride, bias=False) on rows ordered (spk, b), then the reference's
    # reshape(B, spks, -1) of those rows (a (spk,batch) scramble for B > 1, reproduced on purpose)
    rows = sep_h.reshape(spk * B, N, L)
    fr = (rows.permute(0, 2, 1).reshape(spk * B * L, N) @ sd["decoder.weight"].reshape(N, win)).reshape(spk * B, L, win)
    wav = fr.new_zeros(spk * B, (L - 1) * stride + win)
    half = fr.reshape(spk * B, L, 2, stride)
    wav[:, : L * stride] += half[:, :, 0].reshape(spk * B, L * stride)
    wav[:, stride:] += half[:, :, 1].reshape(spk * B, L * stride)
    est = wav.reshape(B, spk, -1)                                                       # :1004
    T_est = est.shape[2]
    if T > T_est:
        est = torch.nn.functional.pad(est, (0, T - T_est))                              # :1007-1012
    else:
        est = est[:, :, :T]
    if taps is not None:
        taps.update(mix_w=mix_w, blocks=blocks, est_mask=est_mask)
    return est.squeeze(0) if was_one_d else est
