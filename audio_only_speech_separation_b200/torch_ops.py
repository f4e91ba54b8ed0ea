"""``torch.library`` custom operators over the C-ABI (namespace ``dualpath::``).

The operators are registered for the CUDA dispatch key only: a CPU tensor reaches no kernel and raises (there is no CPU path).
Each has a fake (meta) implementation for shape inference / ``torch.compile`` tracing and a registered backward, itself an
operator of the same namespace, so graphs stay inside the extension.

    torch.ops.dualpath.segment(x[B,N,L], K)                    -> y[B,N,K,S]      split_feature   gc3_basics.py:79-91
    torch.ops.dualpath.overlap_add(y[B,N,K,S], rest)           -> x[B,N,L]        merge_feature   gc3_basics.py:94-109
    torch.ops.dualpath.attention(qkv[B,S,K,3E], heads, layout) -> (o, lse)        MHA core        dptnet.py:48, sepformer.py:124-133
    torch.ops.dualpath.add_layernorm(a, b, gamma, beta, eps)   -> (out, z)        LayerNorm(a+b)  dptnet.py:57-63
    torch.ops.dualpath.pit_sdr_loss(ests, targets, mode, thr)  -> (loss, perm)    PIT loss        pit_wrapper.py:30-67 + matrix.py:13-57

Segmentation and overlap-add are each other's adjoint (every frame lands in exactly two chunks), which is what the engines use for
their backward passes as well.
"""
from __future__ import annotations

from typing import Tuple

import torch
from torch.library import custom_op

from . import _lib, ops
from ._lib import check, lib, ptr, stream_ptr
from .losses.matrix import pit_sdr_forward

__all__ = ["segment", "overlap_add", "attention", "attention_backward", "add_layernorm", "layernorm_backward", "pit_sdr_loss",
           "pit_sdr_loss_backward", "OP_NAMES"]

OP_NAMES = ("segment", "overlap_add", "attention", "attention_backward", "add_layernorm", "layernorm_backward", "pit_sdr_loss",
            "pit_sdr_loss_backward")


# ------------------------------------------------------------------------------------------ segmentation / overlap-add
@custom_op("dualpath::segment", mutates_args=(), device_types="cuda")
def segment(x: torch.Tensor, block_size: int) -> torch.Tensor:
    return ops.split_feature(x.contiguous(), block_size)[0]


@segment.register_fake
def _(x, block_size):
    B, N, L = x.shape
    _, S = _lib.seg_geometry(int(L), int(block_size))
    return x.new_empty(B, N, block_size, S)


@custom_op("dualpath::overlap_add", mutates_args=(), device_types="cuda")
def overlap_add(y: torch.Tensor, rest: int) -> torch.Tensor:
    return ops.merge_feature(y.contiguous(), rest)


@overlap_add.register_fake
def _(y, rest):
    B, N, K, S = y.shape
    return y.new_empty(B, N, (S // 2) * K - K // 2 - rest)   # gc3_basics.py:94-109


def _segment_setup(ctx, inputs, output):
    x, block_size = inputs
    ctx.rest = _lib.seg_geometry(int(x.shape[-1]), int(block_size))[0]


def _segment_backward(ctx, grad):
    return torch.ops.dualpath.overlap_add(grad.contiguous(), ctx.rest), None


segment.register_autograd(_segment_backward, setup_context=_segment_setup)


def _ola_setup(ctx, inputs, output):
    y, rest = inputs
    ctx.block_size = int(y.shape[2])


def _ola_backward(ctx, grad):
    return torch.ops.dualpath.segment(grad.contiguous(), ctx.block_size), None


overlap_add.register_autograd(_ola_backward, setup_context=_ola_setup)


# ------------------------------------------------------------------------------------------ attention core
@custom_op("dualpath::attention", mutates_args=(), device_types="cuda")
def attention(qkv: torch.Tensor, heads: int, layout: str) -> Tuple[torch.Tensor, torch.Tensor]:
    o, lse = ops.attention(qkv.contiguous(), heads, layout, save=True)
    return o, lse


@attention.register_fake
def _(qkv, heads, layout):
    B, S, K, E3 = qkv.shape
    return qkv.new_empty(B, S, K, E3 // 3), qkv.new_empty(B * S * K, heads)


@custom_op("dualpath::attention_backward", mutates_args=(), device_types="cuda")
def attention_backward(qkv: torch.Tensor, o: torch.Tensor, lse: torch.Tensor, d_o: torch.Tensor, heads: int, layout: str) -> torch.Tensor:
    return ops.attention_backward(qkv.contiguous(), o, lse, d_o.contiguous(), heads, layout)


@attention_backward.register_fake
def _(qkv, o, lse, d_o, heads, layout):
    return torch.empty_like(qkv)


def _attn_setup(ctx, inputs, output):
    qkv, heads, layout = inputs
    o, lse = output
    ctx.save_for_backward(qkv, o, lse)
    ctx.heads, ctx.layout = heads, layout


def _attn_backward(ctx, d_o, _d_lse):
    qkv, o, lse = ctx.saved_tensors
    return torch.ops.dualpath.attention_backward(qkv, o, lse, d_o.contiguous(), ctx.heads, ctx.layout), None, None


attention.register_autograd(_attn_backward, setup_context=_attn_setup)


# ------------------------------------------------------------------------------------------ residual add + LayerNorm
@custom_op("dualpath::add_layernorm", mutates_args=(), device_types="cuda")
def add_layernorm(a: torch.Tensor, b: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float) -> Tuple[torch.Tensor, torch.Tensor]:
    out, z = ops.add_layernorm(a.contiguous(), b.contiguous(), gamma, beta, eps, save_z=True)
    return out, z


@add_layernorm.register_fake
def _(a, b, gamma, beta, eps):
    return torch.empty_like(a), torch.empty_like(a)


@custom_op("dualpath::layernorm_backward", mutates_args=(), device_types="cuda")
def layernorm_backward(dy: torch.Tensor, z: torch.Tensor, gamma: torch.Tensor, eps: float) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    return ops.layernorm_backward(dy.contiguous(), z, gamma, eps)


@layernorm_backward.register_fake
def _(dy, z, gamma, eps):
    return torch.empty_like(dy), torch.empty_like(gamma), torch.empty_like(gamma)


def _ln_setup(ctx, inputs, output):
    a, b, gamma, beta, eps = inputs
    ctx.save_for_backward(output[1], gamma)
    ctx.eps = eps


def _ln_backward(ctx, d_out, _d_z):
    z, gamma = ctx.saved_tensors
    dz, dgamma, dbeta = torch.ops.dualpath.layernorm_backward(d_out.contiguous(), z, gamma, ctx.eps)
    return dz, dz, dgamma, dbeta, None


add_layernorm.register_autograd(_ln_backward, setup_context=_ln_setup)


# ------------------------------------------------------------------------------------------ fused PIT SNR / SI-SDR / SD-SDR loss
@custom_op("dualpath::pit_sdr_loss", mutates_args=(), device_types="cuda")
def pit_sdr_loss(ests: torch.Tensor, targets: torch.Tensor, sdr_type: str, threshold_byloss: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    loss, _pw, perm, ws = pit_sdr_forward(ests.contiguous(), targets.contiguous(), sdr_type, threshold_byloss)
    return loss.reshape(()), perm, ws


@pit_sdr_loss.register_fake
def _(ests, targets, sdr_type, threshold_byloss):
    B = ests.shape[0]
    ws = ests.new_empty(int(lib().dp_pit_loss_workspace_bytes(int(B))), dtype=torch.uint8)
    return ests.new_empty(()), ests.new_empty(B, dtype=torch.int32), ws


@custom_op("dualpath::pit_sdr_loss_backward", mutates_args=(), device_types="cuda")
def pit_sdr_loss_backward(ests: torch.Tensor, targets: torch.Tensor, ws: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    B, _, T = ests.shape
    # the kernel reads raw pointers: saved tensors may be the caller's non-contiguous / non-fp32 views
    ests, targets = ests.contiguous().float(), targets.contiguous().float()
    d = torch.empty_like(ests)
    check(lib().dp_pit_loss_backward(ptr(ests), ptr(targets), B, T, ptr(ws), 1.0, ptr(d), stream_ptr()), "dp_pit_loss_backward")
    return d * g   # the upstream gradient is a device scalar: folded in without a host sync


@pit_sdr_loss_backward.register_fake
def _(ests, targets, ws, g):
    return torch.empty_like(ests)


def _pit_setup(ctx, inputs, output):
    ests, targets, _, _ = inputs
    ctx.save_for_backward(ests, targets, output[2])


def _pit_backward(ctx, g, _d_perm, _d_ws):
    ests, targets, ws = ctx.saved_tensors
    if ctx.needs_input_grad[1]:
        raise NotImplementedError("dualpath::pit_sdr_loss has no gradient with respect to the targets (the reference never asks for it)")
    return torch.ops.dualpath.pit_sdr_loss_backward(ests, targets, ws, g), None, None, None


pit_sdr_loss.register_autograd(_pit_backward, setup_context=_pit_setup)
