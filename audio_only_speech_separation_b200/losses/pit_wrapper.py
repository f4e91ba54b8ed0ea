"""``PITLossWrapper`` with the contract of look2hear/losses/pit_wrapper.py:15-67 on the fused CUDA loss."""
import torch
from torch import nn

from .._lib import check, lib, ptr, stream_ptr
from .matrix import PairwiseNegSDR, pit_sdr_forward


class _PitSdrFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ests, targets, sdr_type, threshold):
        loss, pw, perm, ws = pit_sdr_forward(ests, targets, sdr_type, threshold)
        ctx.save_for_backward(ests, targets)
        ctx.ws = ws
        ctx.mark_non_differentiable(perm)
        return loss.reshape(()), perm

    @staticmethod
    def backward(ctx, g, _):
        ests, targets = ctx.saved_tensors
        B, _, T = ests.shape
        d = torch.empty_like(ests)
        # the upstream gradient is a device scalar: fold it in without a host sync
        check(lib().dp_pit_loss_backward(ptr(ests), ptr(targets), B, T, ptr(ctx.ws), 1.0, ptr(d), stream_ptr()), "dp_pit_loss_backward")
        return d * g, None, None, None


class PITLossWrapper(nn.Module):
    def __init__(self, loss_func, pit_from="pw_mtx", perm_reduce=None, threshold_byloss=True):
        super().__init__()
        self.loss_func = loss_func
        self.pit_from = pit_from
        self.perm_reduce = perm_reduce
        self.threshold_byloss = threshold_byloss
        if self.pit_from not in ["pw_mtx", "pw_pt", "perm_avg"]:
            raise ValueError(
                "Unsupported loss function type {} for now. Expected" "one of [`pw_mtx`, `pw_pt`, `perm_avg`]".format(self.pit_from)
            )

    def forward(self, ests, targets, return_ests=False, reduce_kwargs=None, **kwargs):
        if self.pit_from != "pw_mtx" or not isinstance(self.loss_func, PairwiseNegSDR) or self.perm_reduce is not None:
            raise NotImplementedError(
                "the fused path covers pit_from='pw_mtx' with a PairwiseNegSDR loss (what every config of the reference uses)"
            )
        ests_c, targets_c = ests.contiguous(), targets.contiguous()
        mean_loss, perm = _PitSdrFunction.apply(ests_c, targets_c, self.loss_func.sdr_type, bool(self.threshold_byloss))
        if not return_ests:
            return mean_loss
        return mean_loss, self.reordered_sources(ests_c, perm)

    @staticmethod
    def reordered_sources(source, perm):
        """pit_wrapper.py:90-94 with ``perm[b]`` = 0 (identity) or 1 (swapped) as produced by the fused kernel."""
        B, _, T = source.shape
        out = torch.empty_like(source)
        check(lib().dp_pit_reorder(ptr(source), ptr(perm), ptr(out), B, T, stream_ptr()), "dp_pit_reorder")
        return out
