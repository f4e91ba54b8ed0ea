"""Drop-in ``TasNet`` (look2hear/models/gc3_network.py:7-188) whose forward/backward run in the CUDA engine.

The constructor signature, ``forward(mixture) -> est_sources`` contract, ``model_name`` attribute and every
``state_dict`` key/shape are those of the reference, so ``audio_train.py`` / ``audio_test.py`` style lookups
(``getattr(models, "TasNet")(sample_rate=..., **audionet_config)``, ``from_pretrain``) keep working and existing
``best_model.pth`` files load.  The sub-modules below are *parameter containers only* (built in the reference's
construction order so the default initialisation is reproduced bit for bit under the same seed); none of their
``forward`` methods is ever called.  All parameters live as views of ONE flat fp32 buffer, which is what the engine,
the fused Adam step and the single NCCL gradient all-reduce operate on.

Supported on this path: ``module="DPRNN"`` and ``module="DPTNet"``, ``group_size=1``, ``enc_dim=bn_dim=64``,
``hidden_dim=128``, ``win=16`` (every DPRNN / DPTNet config of the reference), ``unfold`` True or False; and the GroupComm
variant ``group_size in (8, 16, 32)`` (both modules, ``unfold`` True or False) with per-group widths (bn_dim/G, hidden_dim/G) = (4, 8) or (8, 16)
(``unit_tests.py:69-86`` of the reference), inference only, on its own fp32 engine (csrc/groupcomm.cu).  Anything else raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List

import torch
import torch.nn as nn

from .. import _lib
from .._lib import check, lib, ptr, stream_ptr
from .base_model import BaseModel


class _ProjRNN(nn.Module):
    """Parameter container with the keys of ``ProjRNN`` (gc3_basics.py:7-17): ``rnn.*``, ``proj.*``."""

    def __init__(self, input_size, hidden_size, bidirectional=True):
        super().__init__()
        self.rnn = nn.LSTM(input_size, hidden_size, 1, dropout=0, batch_first=True, bidirectional=bidirectional)
        self.proj = nn.Linear(hidden_size * (int(bidirectional) + 1), input_size)


class _TAC(nn.Module):
    """Parameter container with the keys of ``TAC`` (gc3_basics.py:28-36)."""

    def __init__(self, input_size, hidden_size):
        super().__init__()
        self.TAC_input = nn.Sequential(nn.Linear(input_size, hidden_size), nn.PReLU())
        self.TAC_mean = nn.Sequential(nn.Linear(hidden_size, hidden_size), nn.PReLU())
        self.TAC_output = nn.Sequential(nn.Linear(hidden_size * 2, input_size), nn.PReLU())
        self.TAC_norm = nn.GroupNorm(1, input_size)


class _GCRNN(nn.Module):
    """Parameter container with the keys of ``GC_RNN`` (groupcomm.py:10-24): ``TAC.i.*``, ``rnn.i.*``, ``LN.i.*``."""

    def __init__(self, input_size, hidden_size, num_group, num_layers):
        super().__init__()
        self.TAC = nn.ModuleList([])
        self.rnn = nn.ModuleList([])
        self.LN = nn.ModuleList([])
        for _ in range(num_layers):
            self.TAC.append(_TAC(input_size // num_group, hidden_size * 3 // num_group))
            self.rnn.append(_ProjRNN(input_size // num_group, hidden_size // num_group))
            self.LN.append(nn.GroupNorm(1, input_size // num_group))


class _DPRNN(nn.Module):
    """Parameter container with the keys of ``DPRNN`` (dprnn.py:7-51); ``num_group > 1`` adds ``TAC.i.*`` and narrows everything to
    ``input_size // num_group`` (shared by the groups)."""

    def __init__(self, input_size, hidden_size, output_size, num_layers, unfold, num_group=1):
        super().__init__()
        if num_group > 1:
            self.TAC = nn.ModuleList([])
            tac_in, tac_hid = input_size // num_group, hidden_size * 3 // num_group
            input_size, hidden_size, output_size = input_size // num_group, hidden_size // num_group, output_size // num_group
        self.row_rnn = nn.ModuleList([])
        self.col_rnn = nn.ModuleList([])
        self.row_norm = nn.ModuleList([])
        self.col_norm = nn.ModuleList([])
        if unfold:  # one shared instance of everything, reused by all layers (dprnn.py:26-34)
            row_rnn = _ProjRNN(input_size, hidden_size)
            col_rnn = _ProjRNN(input_size, hidden_size)
            row_norm = nn.GroupNorm(1, input_size, eps=1e-8)
            col_norm = nn.GroupNorm(1, input_size, eps=1e-8)
            self.concat_block = nn.Sequential(nn.Conv2d(input_size, input_size, 1, 1, groups=input_size), nn.PReLU())
        for _ in range(num_layers):
            if num_group > 1:
                self.TAC.append(_TAC(tac_in, tac_hid))
            self.row_rnn.append(row_rnn if unfold else _ProjRNN(input_size, hidden_size))
            self.col_rnn.append(col_rnn if unfold else _ProjRNN(input_size, hidden_size))
            self.row_norm.append(row_norm if unfold else nn.GroupNorm(1, input_size, eps=1e-8))
            self.col_norm.append(col_norm if unfold else nn.GroupNorm(1, input_size, eps=1e-8))
        self.output = nn.Conv2d(input_size, output_size, 1)


class _TransformerEncoderLayer(nn.Module):
    """Parameter container with the keys of DPTNet's ``TransformerEncoderLayer`` (dptnet.py:45-56): attention, a BiLSTM
    as the "feed-forward", Linear(4*d -> d), two LayerNorms.  Built in the reference's order (same default init)."""

    def __init__(self, d_model, nhead):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=0)
        self.linear1 = nn.LSTM(d_model, d_model * 2, 1, bidirectional=True)
        self.linear2 = nn.Linear(d_model * 2 * 2, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)


class _SingleTransformer(nn.Module):
    """Key prefix ``transformer.*`` of ``SingleTransformer`` (dptnet.py:85-89)."""

    def __init__(self, input_size):
        super().__init__()
        self.transformer = _TransformerEncoderLayer(input_size, 4)


class _DPTNet(nn.Module):
    """Parameter container with the keys of ``DPTNet`` (dptnet.py:99-131), ``num_group == 1``."""

    def __init__(self, input_size, hidden_size, output_size, num_layers, unfold, num_group=1):
        super().__init__()
        if num_group > 1:   # dptnet.py:111-131: TAC per layer, everything else at width input_size // num_group
            self.TAC = nn.ModuleList([])
            tac_in, tac_hid = input_size // num_group, hidden_size * 3 // num_group
            input_size, output_size = input_size // num_group, output_size // num_group
        self.row_xfmr = nn.ModuleList([])
        self.col_xfmr = nn.ModuleList([])
        if unfold:
            row_xfmr = _SingleTransformer(input_size)
            col_xfmr = _SingleTransformer(input_size)
            self.concat_block = nn.Sequential(nn.Conv2d(input_size, input_size, 1, 1, groups=input_size), nn.PReLU())
        for _ in range(num_layers):
            if num_group > 1:
                self.TAC.append(_TAC(tac_in, tac_hid))
            self.row_xfmr.append(row_xfmr if unfold else _SingleTransformer(input_size))
            self.col_xfmr.append(col_xfmr if unfold else _SingleTransformer(input_size))
        self.output = nn.Conv2d(input_size, output_size, 1)


class _DPWrapper(nn.Module):
    """Key prefix of ``DP_Wrapper`` (groupcomm.py:49-98): ``seq_model.*``."""

    def __init__(self, input_dim, hidden_dim, output_dim, layer, unfold, module="DPRNN", num_group=1):
        super().__init__()
        cls = _DPRNN if module == "DPRNN" else _DPTNet
        self.seq_model = cls(input_dim, hidden_dim, output_dim, layer, unfold, num_group)


class _TasNetFunction(torch.autograd.Function):
    """Autograd node for the whole network: forward and backward are single engine calls."""

    @staticmethod
    def forward(ctx, model, mixture, *params):
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("TasNet: the engine has no gradient with respect to the input waveform (the reference's training "
                                      "step never asks for it); detach the mixture")
        train = any(ctx.needs_input_grad[2:])
        est, ws = model._engine_forward(mixture, train)
        ctx.model, ctx.ws, ctx.dims = model, ws, mixture.shape
        return est

    @staticmethod
    def backward(ctx, d_est):
        model = ctx.model
        if ctx.ws is None:
            raise RuntimeError("TasNet backward: forward ran without gradient tracking")
        gflat = torch.zeros_like(model._flat)
        model._engine_backward(d_est.contiguous().float(), gflat, ctx.ws, *ctx.dims)
        ctx.ws = None
        grads = [gflat[o : o + p.numel()].view(p.shape) for p, o in zip(model._uniq, model._uniq_off)]
        return (None, None, *grads)


class _GcTasNetFunction(torch.autograd.Function):
    """Autograd node of the GroupComm engine (group_size > 1, DPRNN or DPTNet stack): training forward + backward of csrc/groupcomm.cu."""

    @staticmethod
    def forward(ctx, model, mixture, *params):
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("TasNet: the engine has no gradient with respect to the input waveform; detach the mixture")
        est, saved = model._gc_train_forward(mixture)
        ctx.model, ctx.saved, ctx.dims = model, saved, mixture.shape
        return est

    @staticmethod
    def backward(ctx, d_est):
        model = ctx.model
        gflat = torch.zeros_like(model._flat)
        model._gc_train_backward(d_est.contiguous().float(), gflat, ctx.saved, *ctx.dims)
        ctx.saved = None
        grads = [gflat[o : o + p.numel()].view(p.shape) for p, o in zip(model._uniq, model._uniq_off)]
        return (None, None, *grads)


class TasNet(BaseModel):
    def __init__(
        self,
        enc_dim=64,
        bn_dim=64,
        hidden_dim=128,
        win=16,
        layer=6,
        num_spk=2,
        module="DPRNN",
        context_size=24,
        group_size=1,
        block_size=100,
        sample_rate=16000,
        unfold=False,
    ):
        super().__init__(sample_rate=sample_rate)
        assert module in ["DPRNN", "DPTNet", "TCN", "SudoRMRF", "GC_TCN", "GC_SudoRMRF"]  # gc3_network.py:25-32
        if module not in ("DPRNN", "DPTNet"):
            raise NotImplementedError(f"module={module!r}: this build accelerates the dual-path modules 'DPRNN' and 'DPTNet' "
                                      "(see DESIGN.md scope table)")
        self.num_spk = num_spk
        self.enc_dim = enc_dim
        self.bn_dim = bn_dim
        self.hidden_dim = hidden_dim
        self.context_size = context_size
        self.group_size = group_size
        self.win = win
        self.stride = self.win // 2
        self.model_name = module
        self.unfold = unfold
        self.layer = layer
        self.block_size = block_size
        # 'fp32' (bf16x3 tensor-core products, fp32 parity) or 'bf16'
        self.precision = os.environ.get("DUALPATH_PRECISION", "fp32")

        # ---- parameters, created in the reference's order (gc3_network.py:48-106) ----
        self.encoder = nn.Conv1d(1, self.enc_dim, self.win, bias=False, stride=self.stride)
        torch.nn.init.xavier_uniform_(self.encoder.weight)
        self.bottleneck = nn.Sequential(
            nn.GroupNorm(1, self.enc_dim, eps=torch.finfo(torch.float32).eps),
            nn.Conv1d(self.enc_dim, self.bn_dim, 1, bias=False),
        )
        if self.group_size > 1:  # context encoder / decoder (gc3_network.py:59-61)
            self.context_enc = _GCRNN(self.bn_dim, self.hidden_dim, self.group_size, 2)
            self.context_dec = _GCRNN(self.bn_dim, self.hidden_dim, self.group_size, 2)
        self.seq_model = _DPWrapper(self.bn_dim, self.hidden_dim, self.bn_dim, layer, unfold, module, self.group_size)
        self.mask = nn.Sequential(
            nn.Conv1d(self.bn_dim // self.group_size, self.enc_dim * self.num_spk // self.group_size, 1), nn.ReLU(inplace=True)
        )
        self.decoder = nn.ConvTranspose1d(self.enc_dim, 1, self.win, bias=False, stride=self.stride)
        torch.nn.init.xavier_uniform_(self.decoder.weight)

        self._handle = None
        self._flat = None
        self._uniq: List[nn.Parameter] = []
        self._uniq_off: List[int] = []
        self._pack = None
        self._pack_sig = None
        self.last_launches = 0

    # ------------------------------------------------------------------ reference API
    def pad_input(self, input):
        """Shape normalisation of gc3_network.py:108-131 (the zero padding itself happens inside the engine)."""
        was_one_d = False
        if input.ndim == 1:
            was_one_d = True
            input = input.unsqueeze(0)
        if input.ndim == 3:
            input = input.squeeze(1)
        nsample = input.shape[1]
        rest = self.win - (self.stride + nsample % self.win) % self.win
        return input, rest, was_one_d

    def forward(self, input):
        x, _, was_one_d = self.pad_input(input)
        if x.ndim != 2:
            raise ValueError(f"expected [T], [B, T] or [B, 1, T], got {tuple(input.shape)}")
        _lib.require_cuda(x.contiguous(), "TasNet input")
        xin = x.contiguous().float()
        self._sync_flat(xin.device)
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self._uniq)
        if self.cuda_graph and not needs_grad and not torch.cuda.is_current_stream_capturing():
            est = self._graph_forward(xin)
        elif self.group_size > 1:
            est = _GcTasNetFunction.apply(self, xin, *self._uniq) if needs_grad else self._gc_forward(xin)
        else:
            est = _TasNetFunction.apply(self, xin, *self._uniq)
        if est.dtype != input.dtype and input.dtype.is_floating_point:
            est = est.to(input.dtype)
        return est.squeeze(0) if was_one_d else est

    # ---- inference forwards as CUDA graphs -------------------------------------------------------------------------------------------
    # One forward is 70-odd dependent launches (DESIGN.md sections 5 / 5b); at small batch the GroupComm engine is launch-bound.  The second
    # forward with the same (batch, samples, precision, weights) is captured (torch.cuda.CUDAGraph: the engine makes no allocation and no
    # synchronisation, its workspace comes from the graph's private pool) and later calls replay it.  Set ``model.cuda_graph = False`` to opt out.
    cuda_graph = True
    _GRAPH_SLOTS = 8

    def _eager_inference(self, xin):
        if self.group_size > 1:
            return self._gc_forward(xin)
        return self._engine_forward(xin, False)[0]

    def _graph_forward(self, xin):
        if self.group_size == 1:
            self._ensure_pack()
        key = (tuple(xin.shape), str(self.precision), self._flat.data_ptr(), tuple(p._version for p in self._uniq), xin.device.index)
        graphs = self.__dict__.setdefault("_graphs", {})
        seen = self.__dict__.setdefault("_graph_seen", {})
        ent = graphs.get(key)
        if ent is None:
            if key not in seen:            # first sight of this shape: run it eagerly (variable-length evaluation never pays for a capture)
                if len(seen) >= 64:
                    seen.clear()
                seen[key] = True
                return self._eager_inference(xin)
            if len(graphs) >= self._GRAPH_SLOTS:
                graphs.clear()
            static_in = xin.clone()
            side = torch.cuda.Stream(device=xin.device)
            side.wait_stream(torch.cuda.current_stream(xin.device))
            with torch.cuda.stream(side):      # warm-up outside the capture (function attributes, lazily loaded kernels)
                self._eager_inference(static_in)
            torch.cuda.current_stream(xin.device).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                static_out = self._eager_inference(static_in)
            ent = graphs[key] = (g, static_in, static_out)
        g, static_in, static_out = ent
        static_in.copy_(xin)
        g.replay()
        return static_out.clone()

    def get_model_args(self):
        return {"n_src": 2}  # gc3_network.py:186-188

    # ------------------------------------------------------------------ flat parameters / engine
    @staticmethod
    def _tac_params(t):
        return [t.TAC_input[0].weight, t.TAC_input[0].bias, t.TAC_input[1].weight, t.TAC_mean[0].weight, t.TAC_mean[0].bias,
                t.TAC_mean[1].weight, t.TAC_output[0].weight, t.TAC_output[0].bias, t.TAC_output[1].weight, t.TAC_norm.weight,
                t.TAC_norm.bias]

    @staticmethod
    def _rnn_params(rnn, norm):
        r = rnn.rnn
        return [r.weight_ih_l0, r.weight_hh_l0, r.bias_ih_l0, r.bias_hh_l0, r.weight_ih_l0_reverse, r.weight_hh_l0_reverse,
                r.bias_ih_l0_reverse, r.bias_hh_l0_reverse, rnn.proj.weight, rnn.proj.bias, norm.weight, norm.bias]

    def _gc_param_table(self):
        """Order of the dp_gctasnet parameter table (include/dualpath_b200.h)."""
        sm = self.seq_model.seq_model
        cat = sm.concat_block if self.unfold else None
        table = [self.encoder.weight, self.bottleneck[0].weight, self.bottleneck[0].bias, self.bottleneck[1].weight, sm.output.weight,
                 sm.output.bias, self.mask[0].weight, self.mask[0].bias, self.decoder.weight,
                 cat[0].weight if cat is not None else None, cat[0].bias if cat is not None else None,
                 cat[1].weight if cat is not None else None]
        for gc in (self.context_enc, self.context_dec):
            for i in range(2):
                table += self._tac_params(gc.TAC[i]) + self._rnn_params(gc.rnn[i], gc.LN[i])
        for i in range(self.layer):
            table += self._tac_params(sm.TAC[i])
            if self.model_name == "DPTNet":
                for x in (sm.row_xfmr[i].transformer, sm.col_xfmr[i].transformer):
                    r = x.linear1
                    table += [r.weight_ih_l0, r.weight_hh_l0, r.bias_ih_l0, r.bias_hh_l0, r.weight_ih_l0_reverse, r.weight_hh_l0_reverse,
                              r.bias_ih_l0_reverse, r.bias_hh_l0_reverse, x.linear2.weight, x.linear2.bias, x.norm2.weight, x.norm2.bias,
                              x.self_attn.in_proj_weight, x.self_attn.in_proj_bias, x.self_attn.out_proj.weight, x.self_attn.out_proj.bias,
                              x.norm1.weight, x.norm1.bias]
            else:
                table += self._rnn_params(sm.row_rnn[i], sm.row_norm[i]) + self._rnn_params(sm.col_rnn[i], sm.col_norm[i])
        return table

    def _gc_forward(self, xin):
        """GroupComm engine, inference forward (nothing saved)."""
        B, T = xin.shape
        nbytes = lib().dp_gctasnet_workspace_bytes(self._handle, B, T)
        if nbytes < 0:
            check(1, "dp_gctasnet_workspace_bytes")
        ws = torch.empty(nbytes, device=xin.device, dtype=torch.uint8)
        est = torch.empty(B, self.num_spk, T, device=xin.device, dtype=torch.float32)
        check(lib().dp_gctasnet_forward(self._handle, ptr(self._flat), ptr(xin), ptr(est), ptr(ws), B, T, stream_ptr()), "dp_gctasnet_forward")
        self.last_launches = lib().dp_gctasnet_last_launches(self._handle)
        return est

    def _destroy_handle(self):
        if self._handle is not None:
            (lib().dp_gctasnet_destroy if self.group_size > 1 else lib().dp_tasnet_destroy)(self._handle)
            self._handle = None

    def _param_table(self):
        if self.group_size > 1:
            return self._gc_param_table()
        sm = self.seq_model.seq_model
        cat = sm.concat_block if self.unfold else None
        table = [
            self.encoder.weight,
            self.bottleneck[0].weight,
            self.bottleneck[0].bias,
            self.bottleneck[1].weight,
            sm.output.weight,
            sm.output.bias,
            self.mask[0].weight,
            self.mask[0].bias,
            self.decoder.weight,
            cat[0].weight if cat is not None else None,
            cat[0].bias if cat is not None else None,
            cat[1].weight if cat is not None else None,
        ]
        for i in range(self.layer):
            if self.model_name == "DPTNet":  # order of DP_TASNET_PATH_PARAMS_DPTNET in include/dualpath_b200.h
                for x in (sm.row_xfmr[i].transformer, sm.col_xfmr[i].transformer):
                    r = x.linear1
                    table += [
                        r.weight_ih_l0, r.weight_hh_l0, r.bias_ih_l0, r.bias_hh_l0,
                        r.weight_ih_l0_reverse, r.weight_hh_l0_reverse, r.bias_ih_l0_reverse, r.bias_hh_l0_reverse,
                        x.linear2.weight, x.linear2.bias, x.norm2.weight, x.norm2.bias,
                        x.self_attn.in_proj_weight, x.self_attn.in_proj_bias, x.self_attn.out_proj.weight, x.self_attn.out_proj.bias,
                        x.norm1.weight, x.norm1.bias,
                    ]
                continue
            for rnn, norm in ((sm.row_rnn[i], sm.row_norm[i]), (sm.col_rnn[i], sm.col_norm[i])):
                r = rnn.rnn
                table += [
                    r.weight_ih_l0, r.weight_hh_l0, r.bias_ih_l0, r.bias_hh_l0,
                    r.weight_ih_l0_reverse, r.weight_hh_l0_reverse, r.bias_ih_l0_reverse, r.bias_hh_l0_reverse,
                    rnn.proj.weight, rnn.proj.bias, norm.weight, norm.bias,
                ]
        return table

    def _flat_is_valid(self, device) -> bool:
        if self._flat is None or self._flat.device != device:
            return False
        base = self._flat.data_ptr()
        return all(p.data_ptr() == base + 4 * o and p.dtype == torch.float32 for p, o in zip(self._uniq, self._uniq_off))

    def _sync_flat(self, device):
        """(Re)build the flat parameter buffer and the engine handle when parameters moved (``.to()``, ``.cuda()``)."""
        if self._flat_is_valid(device):
            return
        table = self._param_table()
        for p in table:
            if p is not None and p.device != device:
                raise RuntimeError(
                    f"TasNet parameters are on {p.device} but the input is on {device}; move the model with .to(device) "
                    "(the dual-path kernels are CUDA-only, there is no CPU path)"
                )
        uniq, offs, seen, total = [], [], {}, 0
        for p in table:
            if p is None or id(p) in seen:
                continue
            seen[id(p)] = total
            uniq.append(p)
            offs.append(total)
            total += (p.numel() + 7) // 8 * 8  # 32-byte aligned: the bf16 copies are read as 16-byte vectors
        flat = torch.zeros(total, device=device, dtype=torch.float32)
        with torch.no_grad():
            for p, o in zip(uniq, offs):
                flat[o : o + p.numel()].copy_(p.data.reshape(-1).to(device=device, dtype=torch.float32))
                p.data = flat[o : o + p.numel()].view(p.shape)
        self._flat, self._uniq, self._uniq_off = flat, uniq, offs
        offsets = [(-1 if p is None else seen[id(p)]) for p in table]
        self._destroy_handle()
        if self.group_size > 1:
            cfg = _lib.GcTasnetConfig(self.enc_dim, self.bn_dim, self.hidden_dim, self.win, self.layer, self.num_spk, self.context_size,
                                      self.group_size, self.block_size, int(self.unfold),
                                      _lib.MODULE_DPTNET if self.model_name == "DPTNet" else _lib.MODULE_DPRNN)
            arr = (C.c_int64 * len(offsets))(*offsets)
            h = C.c_void_p()
            check(lib().dp_gctasnet_create(C.byref(cfg), arr, len(offsets), total, C.byref(h)), "dp_gctasnet_create")
            self._handle = h
            return
        cfg = _lib.TasnetConfig(self.enc_dim, self.bn_dim, self.hidden_dim, self.win, self.layer, self.num_spk, self.block_size,
                                int(self.unfold), _lib.MODULE_DPTNET if self.model_name == "DPTNet" else _lib.MODULE_DPRNN)
        arr = (C.c_int64 * len(offsets))(*offsets)
        h = C.c_void_p()
        check(lib().dp_tasnet_create(C.byref(cfg), arr, len(offsets), total, C.byref(h)), "dp_tasnet_create")
        self._handle = h
        self._pack = torch.empty(lib().dp_tasnet_pack_bytes(h), device=device, dtype=torch.uint8)
        self._pack_sig = None

    def __del__(self):
        try:
            self._destroy_handle()
        except Exception:
            pass

    def mark_params_dirty(self):
        """Call after writing the flat buffer directly (the fused optimizer step does)."""
        self._pack_sig = None

    def _prec(self) -> int:
        p = str(self.precision).lower()
        if p in ("fp32", "float32"):
            return _lib.PREC_FP32
        if p in ("bf16", "bfloat16"):
            return _lib.PREC_BF16
        raise ValueError(f"precision must be 'fp32' or 'bf16', got {self.precision!r}")

    def _ensure_pack(self):
        self._require_training_engine()
        sig = tuple(p._version for p in self._uniq)
        if sig != self._pack_sig:
            check(lib().dp_tasnet_pack(self._handle, ptr(self._flat), ptr(self._pack), stream_ptr()), "dp_tasnet_pack")
            self._pack_sig = sig

    def _require_training_engine(self):
        if self.group_size > 1:  # the handle is a dp_gctasnet: never hand it to the dp_tasnet_* entry points
            raise NotImplementedError("TasNet(group_size > 1) runs on the GroupComm engine (dp_gctasnet_*): its handle must not reach the "
                                      "dp_tasnet_* entry points; training goes through _train_forward / _train_backward (autograd, DualPathTrainer, fit)")

    def _engine_forward(self, mixture, train: bool, est=None, ws=None):
        self._require_training_engine()
        B, T = mixture.shape
        self._ensure_pack()
        nbytes = lib().dp_tasnet_workspace_bytes(self._handle, B, T, int(train))
        if nbytes < 0:
            check(1, "dp_tasnet_workspace_bytes")
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(nbytes, device=mixture.device, dtype=torch.uint8)
        if est is None:
            est = torch.empty(B, self.num_spk, T, device=mixture.device, dtype=torch.float32)
        check(
            lib().dp_tasnet_forward(self._handle, ptr(self._flat), ptr(self._pack), ptr(mixture), ptr(est), ptr(ws), B, T, int(train),
                                    self._prec(), stream_ptr()),
            "dp_tasnet_forward",
        )
        self.last_launches = lib().dp_tasnet_last_launches(self._handle)
        return est, (ws if train else None)

    # fused training interface (DualPathTrainer): forward that saves, backward into a flat gradient buffer
    @property
    def pack_launches(self) -> int:
        return 0 if self.group_size > 1 else 1 + 2 * self.layer

    # the fused step draws nothing on the host (no dropout): DualPathTrainer(cuda_graph=True) may replay it as a CUDA graph
    graph_safe_training = True

    def _train_forward(self, mixture, ws=None):
        if self.group_size > 1:
            return self._gc_train_forward(mixture, ws)
        est, ws = self._engine_forward(mixture, True, ws=ws)
        return est, (ws,)

    def _train_backward(self, d_est, gflat, ctx, B, T):
        if self.group_size > 1:
            return self._gc_train_backward(d_est, gflat, ctx, B, T)
        self._engine_backward(d_est, gflat, ctx[0], B, T)

    # GroupComm engine (group_size > 1): training forward that keeps every stage's tensors, backward into the flat gradient buffer
    def _gc_train_forward(self, mixture, ws=None):
        B, T = mixture.shape
        nbytes = lib().dp_gctasnet_train_workspace_bytes(self._handle, B, T)
        if nbytes < 0:
            check(1, "dp_gctasnet_train_workspace_bytes")
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(nbytes, device=mixture.device, dtype=torch.uint8)
        est = torch.empty(B, self.num_spk, T, device=mixture.device, dtype=torch.float32)
        check(lib().dp_gctasnet_forward_train(self._handle, ptr(self._flat), ptr(mixture), ptr(est), ptr(ws), B, T, stream_ptr()),
              "dp_gctasnet_forward_train")
        self.last_launches = lib().dp_gctasnet_last_launches(self._handle)
        return est, (ws, mixture)

    def _gc_train_backward(self, d_est, gflat, ctx, B, T):
        ws, mixture = ctx
        check(lib().dp_gctasnet_backward(self._handle, ptr(self._flat), ptr(gflat), ptr(mixture), ptr(d_est), ptr(ws), B, T, stream_ptr()),
              "dp_gctasnet_backward")
        self.last_launches = lib().dp_gctasnet_last_launches(self._handle)

    def _engine_backward(self, d_est, gflat, ws, B, T):
        self._require_training_engine()
        check(
            lib().dp_tasnet_backward(self._handle, ptr(self._flat), ptr(self._pack), ptr(d_est), ptr(gflat), ptr(ws), B, T, self._prec(),
                                     stream_ptr()),
            "dp_tasnet_backward",
        )
        self.last_launches = lib().dp_tasnet_last_launches(self._handle)
