"""Drop-in ``Sepformer`` (look2hear/models/sepformer.py:849-1020) whose forward runs in the CUDA engine.

Constructor signature, ``forward(mix) -> est_source`` contract, ``model_name`` and every ``state_dict`` key / shape
(including the ``pos_enc.pe`` buffers and the gLN ``gamma`` / ``beta``) are those of the reference, so
``getattr(models, "Sepformer")(sample_rate=..., **audionet_config)`` and ``from_pretrain`` keep working.  The
sub-modules are *parameter containers only*, nested and registered in the reference's order and re-initialised by the
same recursive reset, so the default initialisation is reproduced bit for bit under the same seed; none of their
``forward`` methods is ever called.  All parameters and the positional-encoding tables live as views of ONE flat fp32
buffer that the engine reads through an offset table (include/dualpath_b200.h).

The engine covers inference (``model.eval()`` / ``torch.no_grad()``, TMA-fed tcgen05 GEMMs and attention, fp32-parity and bf16
modes) and training (pre-norm layers on both engines, post-norm layers on the TMA engine): forward + backward are two engine calls behind
one autograd node, gradients are
checked against autograd through the reference algorithm.  The reference trains with dropout 0.1 in four places per layer
(attention probabilities, attention output, FFN hidden, FFN output); ``model.dropout`` (default 0.1, like the
reference) applies them in ``train()`` mode on the TMA engine (enc_dim 128 / 256): masks are a counter-based function of a seed drawn
from torch's generator per forward, regenerated -- not stored -- by the backward; smaller layers, which run on the mma.sync engine,
raise unless ``model.dropout = 0.0``.
Reference quirk kept on purpose: for batch > 1 the output rows are the reference's ``reshape`` of (spk, batch)-ordered
decoder rows (sepformer.py:1004), identity only for batch 1 (the YAML's ``batch_size: 1``).
"""
from __future__ import annotations

import copy
import ctypes as C
import math
import os
from typing import List

import torch
import torch.nn as nn

from .. import _lib
from .._lib import check, lib, ptr, stream_ptr
from .base_model import BaseModel


class _Encoder(nn.Module):  # sepformer.py:23-33
    def __init__(self, kernel_size, out_channels, in_channels):
        super().__init__()
        self.conv1d = nn.Conv1d(in_channels, out_channels, kernel_size, stride=kernel_size // 2, groups=1, bias=False)


class _PositionalEncoding(nn.Module):  # sepformer.py:61-71
    def __init__(self, input_size, max_len=2500):
        super().__init__()
        pe = torch.zeros(max_len, input_size, requires_grad=False)
        positions = torch.arange(0, max_len).unsqueeze(1).float()
        denominator = torch.exp(torch.arange(0, input_size, 2).float() * -(math.log(10000.0) / input_size))
        pe[:, 0::2] = torch.sin(positions * denominator)
        pe[:, 1::2] = torch.cos(positions * denominator)
        self.register_buffer("pe", pe.unsqueeze(0))


class _MHA(nn.Module):  # sepformer.py:112-133
    def __init__(self, nhead, d_model, dropout):
        super().__init__()
        self.att = nn.MultiheadAttention(embed_dim=d_model, num_heads=nhead, dropout=dropout, bias=True, add_bias_kv=False,
                                         add_zero_attn=False, kdim=None, vdim=None)


class _FFN(nn.Module):  # sepformer.py:243-263
    def __init__(self, d_ffn, input_size, dropout):
        super().__init__()
        self.ffn = nn.Sequential(nn.Linear(input_size, d_ffn), nn.ReLU(), nn.Dropout(dropout), nn.Linear(d_ffn, input_size))


class _EncoderLayer(nn.Module):  # sepformer.py:299-319
    def __init__(self, d_ffn, nhead, d_model, dropout):
        super().__init__()
        self.self_att = _MHA(nhead, d_model, dropout)
        self.pos_ffn = _FFN(d_ffn, d_model, dropout)
        self.norm1 = nn.LayerNorm(d_model, eps=1e-6, elementwise_affine=True)
        self.norm2 = nn.LayerNorm(d_model, eps=1e-6, elementwise_affine=True)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)


class _TransformerEncoder(nn.Module):  # sepformer.py:404-436
    def __init__(self, num_layers, nhead, d_ffn, d_model, dropout):
        super().__init__()
        self.layers = nn.ModuleList([_EncoderLayer(d_ffn, nhead, d_model, dropout) for _ in range(num_layers)])
        self.norm = nn.LayerNorm(d_model, eps=1e-6, elementwise_affine=True)


class _TransformerBlock(nn.Module):  # sepformer.py:497-539
    def __init__(self, num_layers, d_model, nhead, d_ffn, use_positional_encoding, dropout=0.1):
        super().__init__()
        self.mdl = _TransformerEncoder(num_layers, nhead, d_ffn, d_model, dropout)
        if use_positional_encoding:
            self.pos_enc = _PositionalEncoding(input_size=d_model)


class _GlobalLN(nn.Module):  # models/utils/normalizations.py:30-47 (beta initialised to ONES)
    def __init__(self, channel_size):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(channel_size), requires_grad=True)
        self.beta = nn.Parameter(torch.ones(channel_size), requires_grad=True)


class _DualBlock(nn.Module):  # sepformer.py:590-598
    def __init__(self, intra_mdl, inter_mdl, out_channels):
        super().__init__()
        self.intra_mdl = intra_mdl
        self.inter_mdl = inter_mdl
        self.intra_norm = _GlobalLN(out_channels)
        self.inter_norm = _GlobalLN(out_channels)


class _DualPathModel(nn.Module):  # sepformer.py:672-704
    def __init__(self, in_channels, out_channels, intra_model, inter_model, num_layers, K, num_spks):
        super().__init__()
        self.norm = nn.GroupNorm(1, in_channels, eps=1e-8)
        self.conv1d = nn.Conv1d(in_channels, out_channels, 1, bias=False)
        self.dual_mdl = nn.ModuleList([])
        for _ in range(num_layers):
            self.dual_mdl.append(copy.deepcopy(_DualBlock(intra_model, inter_model, out_channels)))
        self.conv2d = nn.Conv2d(out_channels, out_channels * num_spks, kernel_size=1)
        self.end_conv1x1 = nn.Conv1d(out_channels, in_channels, 1, bias=False)
        self.prelu = nn.PReLU()
        self.activation = nn.ReLU()
        self.output = nn.Sequential(nn.Conv1d(out_channels, out_channels, 1), nn.Tanh())
        self.output_gate = nn.Sequential(nn.Conv1d(out_channels, out_channels, 1), nn.Sigmoid())


def _reset_layer_recursively(layer):
    """Same traversal as sepformer.py:977-984: reset this layer, then recurse into EVERY descendant (so deep modules are
    reset once per ancestor chain); the random stream consumed is what pins the default initialisation."""
    if hasattr(layer, "reset_parameters"):
        layer.reset_parameters()
    for child in layer.modules():
        if layer != child:
            _reset_layer_recursively(child)


class _SepformerFunction(torch.autograd.Function):
    """Autograd node for the whole network: forward and backward are single engine calls."""

    @staticmethod
    def forward(ctx, model, mixture, *params):
        # dropout masks are a function of (seed, layer, site, element): draw one seed per forward from torch's generator and hand the
        # same (p, seed) to the backward, which regenerates the masks
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("Sepformer: the engine has no gradient with respect to the input waveform; detach the mixture")
        p, seed = model._draw_dropout()
        check(lib().dp_sepformer_set_dropout(model._handle, p, seed), "dp_sepformer_set_dropout")
        est, ws = model._engine_forward_train(mixture)
        ctx.model, ctx.ws, ctx.dims, ctx.drop = model, ws, mixture.shape, (p, seed)
        return est

    @staticmethod
    def backward(ctx, d_est):
        model = ctx.model
        if ctx.ws is None:
            raise RuntimeError("Sepformer backward: the saved workspace was already consumed")
        gflat = torch.zeros_like(model._flat)
        B, T = ctx.dims
        check(lib().dp_sepformer_set_dropout(model._handle, ctx.drop[0], ctx.drop[1]), "dp_sepformer_set_dropout")
        check(lib().dp_sepformer_backward(model._handle, ptr(model._flat), ptr(model._pack), ptr(d_est.contiguous().float()), ptr(gflat),
                                          ptr(ctx.ws), B, T, model._prec(), stream_ptr()), "dp_sepformer_backward")
        ctx.ws = None
        grads = []
        for e, o in zip(model._views, model._view_off):
            t = model._tensor_of(e)
            grads.append(None if isinstance(e, tuple) else gflat[o : o + t.numel()].view(t.shape))
        return (None, None, *[g for g in grads if g is not None])


class Sepformer(BaseModel):
    def __init__(
        self,
        encoder_kernel_size=16,
        encoder_in_nchannels=1,
        encoder_out_nchannels=256,
        masknet_chunksize=250,
        masknet_numlayers=2,
        masknet_norm="gLN",
        masknet_numspks=2,
        intra_numlayers=8,
        inter_numlayers=8,
        intra_nhead=8,
        inter_nhead=8,
        intra_dffn=1024,
        inter_dffn=1024,
        intra_use_positional=True,
        inter_use_positional=True,
        intra_norm_before=True,
        inter_norm_before=True,
        intra_causal=False,
        inter_causal=False,
        sample_rate=8000,
    ):
        super().__init__(sample_rate=sample_rate)
        if masknet_norm != "gLN":
            raise NotImplementedError(f"masknet_norm={masknet_norm!r}: the accelerated path implements gLN (configs/sepformer_base.yml)")
        if intra_causal or inter_causal:
            raise NotImplementedError("causal SepFormer is not on the accelerated path (configs/sepformer_base.yml is non-causal)")
        if encoder_in_nchannels != 1:
            raise NotImplementedError("encoder_in_nchannels must be 1")
        self.cfg = dict(
            enc_dim=encoder_out_nchannels, win=encoder_kernel_size, chunk=masknet_chunksize, num_blocks=masknet_numlayers,
            num_spk=masknet_numspks, intra_layers=intra_numlayers, inter_layers=inter_numlayers, intra_heads=intra_nhead,
            inter_heads=inter_nhead, intra_dffn=intra_dffn, inter_dffn=inter_dffn, intra_pe=int(bool(intra_use_positional)),
            inter_pe=int(bool(inter_use_positional)), intra_norm_before=int(bool(intra_norm_before)),
            inter_norm_before=int(bool(inter_norm_before)),
        )
        self.encoder = _Encoder(encoder_kernel_size, encoder_out_nchannels, encoder_in_nchannels)
        intra_model = _TransformerBlock(intra_numlayers, encoder_out_nchannels, intra_nhead, intra_dffn, intra_use_positional)
        inter_model = _TransformerBlock(inter_numlayers, encoder_out_nchannels, inter_nhead, inter_dffn, inter_use_positional)
        self.masknet = _DualPathModel(encoder_out_nchannels, encoder_out_nchannels, intra_model, inter_model, masknet_numlayers,
                                      masknet_chunksize, masknet_numspks)
        self.decoder = nn.ConvTranspose1d(encoder_out_nchannels, encoder_in_nchannels, encoder_kernel_size,
                                          stride=encoder_kernel_size // 2, bias=False)
        self.num_spks = masknet_numspks
        self.model_name = "Sepformer"
        for module in [self.encoder, self.masknet, self.decoder]:
            _reset_layer_recursively(module)
        # 'fp32' (bf16x3 tensor-core products, fp32 parity) or 'bf16'
        self.precision = os.environ.get("DUALPATH_PRECISION", "fp32")
        self.dropout = 0.1  # TransformerBlock default (sepformer.py:507); see the module docstring
        self._handle = None
        self._flat = None
        self._views: List[torch.Tensor] = []
        self._view_off: List[int] = []
        self._pack = None
        self._pack_sig = None
        self._ws = None
        self.last_launches = 0

    # ------------------------------------------------------------------ reference API
    def forward(self, mix):
        was_one_d = False
        x = mix
        if x.ndim == 1:
            was_one_d = True
            x = x.unsqueeze(0)
        if x.ndim == 3:
            x = x.squeeze(1)
        if x.ndim != 2:
            raise ValueError(f"expected [T], [B, T] or [B, 1, T], got {tuple(mix.shape)}")
        _lib.require_cuda(x.contiguous(), "Sepformer input")
        xin = x.contiguous().float()
        self._sync_flat(xin.device)
        params = [e for e in self._views if not isinstance(e, tuple)]
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            est = _SepformerFunction.apply(self, xin, *params)
        else:
            est = self._engine_forward(xin)
        if est.dtype != mix.dtype and mix.dtype.is_floating_point:
            est = est.to(mix.dtype)
        return est.squeeze(0) if was_one_d else est

    def get_model_args(self):
        return {"n_src": 2}  # sepformer.py:1018-1020

    # ------------------------------------------------------------------ flat parameters / engine
    def _table(self):
        """Tensors (parameters and pe buffers) in the order of the offset table of include/dualpath_b200.h."""
        mk = self.masknet
        table = [self.encoder.conv1d.weight, mk.norm.weight, mk.norm.bias, mk.conv1d.weight, mk.prelu.weight, mk.conv2d.weight,
                 mk.conv2d.bias, mk.output[0].weight, mk.output[0].bias, mk.output_gate[0].weight, mk.output_gate[0].bias,
                 mk.end_conv1x1.weight, self.decoder.weight]
        for blk in mk.dual_mdl:
            for tb, gln in ((blk.intra_mdl, blk.intra_norm), (blk.inter_mdl, blk.inter_norm)):
                table.append(("pe", tb.pos_enc) if hasattr(tb, "pos_enc") else None)
                for ly in tb.mdl.layers:
                    att, ffn = ly.self_att.att, ly.pos_ffn.ffn
                    table += [att.in_proj_weight, att.in_proj_bias, att.out_proj.weight, att.out_proj.bias, ffn[0].weight, ffn[0].bias,
                              ffn[3].weight, ffn[3].bias, ly.norm1.weight, ly.norm1.bias, ly.norm2.weight, ly.norm2.bias]
                table += [tb.mdl.norm.weight, tb.mdl.norm.bias, gln.gamma, gln.beta]
        return table

    @staticmethod
    def _tensor_of(entry):
        return entry[1].pe if isinstance(entry, tuple) else entry

    def _flat_is_valid(self, device) -> bool:
        if self._flat is None or self._flat.device != device:
            return False
        base = self._flat.data_ptr()
        return all(self._tensor_of(e).data_ptr() == base + 4 * o and self._tensor_of(e).dtype == torch.float32
                   for e, o in zip(self._views, self._view_off))

    def _sync_flat(self, device):
        if self._flat_is_valid(device):
            return
        table = self._table()
        for e in table:
            if e is not None and self._tensor_of(e).device != device:
                raise RuntimeError(f"Sepformer parameters are on {self._tensor_of(e).device} but the input is on {device}; move the model "
                                   "with .to(device) (the dual-path kernels are CUDA-only, there is no CPU path)")
        views, offs, total = [], [], 0
        for e in table:
            if e is None:
                continue
            views.append(e)
            offs.append(total)
            total += (self._tensor_of(e).numel() + 7) // 8 * 8
        flat = torch.zeros(total, device=device, dtype=torch.float32)
        with torch.no_grad():
            for e, o in zip(views, offs):
                t = self._tensor_of(e)
                flat[o : o + t.numel()].copy_(t.detach().reshape(-1).to(device=device, dtype=torch.float32))
                view = flat[o : o + t.numel()].view(t.shape)
                if isinstance(e, tuple):
                    e[1]._buffers["pe"] = view
                else:
                    e.data = view
        self._flat, self._views, self._view_off = flat, views, offs
        it = iter(offs)
        offsets = [(-1 if e is None else next(it)) for e in table]
        if self._handle is not None:
            lib().dp_sepformer_destroy(self._handle)
            self._handle = None
        cfg = _lib.SepformerConfig(**self.cfg)
        arr = (C.c_int64 * len(offsets))(*offsets)
        h = C.c_void_p()
        check(lib().dp_sepformer_create(C.byref(cfg), arr, len(offsets), total, C.byref(h)), "dp_sepformer_create")
        self._handle = h
        self._pack = torch.empty(lib().dp_sepformer_pack_bytes(h), device=device, dtype=torch.uint8)
        self._pack_sig = None

    def __del__(self):
        try:
            if self._handle is not None:
                lib().dp_sepformer_destroy(self._handle)
        except Exception:
            pass

    def mark_params_dirty(self):
        self._pack_sig = None

    def _prec(self) -> int:
        p = str(self.precision).lower()
        if p in ("fp32", "float32"):
            return _lib.PREC_FP32
        if p in ("bf16", "bfloat16"):
            return _lib.PREC_BF16
        raise ValueError(f"precision must be 'fp32' or 'bf16', got {self.precision!r}")

    def _ensure_pack(self):
        sig = tuple(self._tensor_of(e)._version for e in self._views)
        if sig != self._pack_sig:
            check(lib().dp_sepformer_pack(self._handle, ptr(self._flat), ptr(self._pack), stream_ptr()), "dp_sepformer_pack")
            self._pack_sig = sig

    # est[b] of a B > 1 call holds rows of other utterances (the reference's reshape quirk, sepformer.py:1004): callers that pair est[b]
    # with targets[b] (metrics.evaluate) must run one utterance per call
    batch_rows_scrambled = True
    # fused training interface (DualPathTrainer): the flat buffer also holds the pe buffers (their gradient stays zero)
    flat_has_buffers = True

    @property
    def pack_launches(self) -> int:
        return 2

    def _draw_dropout(self):
        """(p, seed): the masks are a pure function of (seed, layer, site, element); one seed per forward from torch's generator."""
        p = float(self.dropout) if self.training else 0.0
        seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item()) if p > 0.0 else 0
        return p, seed

    def _train_forward(self, mixture, ws=None):
        drop = self._draw_dropout()
        check(lib().dp_sepformer_set_dropout(self._handle, drop[0], drop[1]), "dp_sepformer_set_dropout")
        est, ws = self._engine_forward_train(mixture, ws)
        return est, (ws, drop)

    def _train_backward(self, d_est, gflat, ctx, B, T):
        ws, drop = ctx
        check(lib().dp_sepformer_set_dropout(self._handle, drop[0], drop[1]), "dp_sepformer_set_dropout")
        check(lib().dp_sepformer_backward(self._handle, ptr(self._flat), ptr(self._pack), ptr(d_est), ptr(gflat), ptr(ws), B, T, self._prec(),
                                          stream_ptr()), "dp_sepformer_backward")
        self.last_launches = lib().dp_sepformer_last_launches(self._handle)

    def _engine_forward_train(self, mixture, ws=None):
        B, T = mixture.shape
        self._ensure_pack()
        nbytes = lib().dp_sepformer_train_workspace_bytes(self._handle, B, T)
        if nbytes < 0:
            check(1, "dp_sepformer_train_workspace_bytes")
        if ws is None or ws.numel() < nbytes or ws.device != mixture.device:
            ws = torch.empty(nbytes, device=mixture.device, dtype=torch.uint8)
        est = torch.empty(B, self.num_spks, T, device=mixture.device, dtype=torch.float32)
        check(lib().dp_sepformer_forward_train(self._handle, ptr(self._flat), ptr(self._pack), ptr(mixture), ptr(est), ptr(ws), B, T,
                                               self._prec(), stream_ptr()), "dp_sepformer_forward_train")
        self.last_launches = lib().dp_sepformer_last_launches(self._handle)
        return est, ws

    def _engine_forward(self, mixture, est=None):
        B, T = mixture.shape
        self._ensure_pack()
        nbytes = lib().dp_sepformer_workspace_bytes(self._handle, B, T)
        if nbytes < 0:
            check(1, "dp_sepformer_workspace_bytes")
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != mixture.device:
            self._ws = torch.empty(nbytes, device=mixture.device, dtype=torch.uint8)
        if est is None:
            est = torch.empty(B, self.num_spks, T, device=mixture.device, dtype=torch.float32)
        check(lib().dp_sepformer_forward(self._handle, ptr(self._flat), ptr(self._pack), ptr(mixture), ptr(est), ptr(self._ws), B, T,
                                         self._prec(), stream_ptr()), "dp_sepformer_forward")
        self.last_launches = lib().dp_sepformer_last_launches(self._handle)
        return est
