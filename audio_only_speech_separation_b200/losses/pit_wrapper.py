"""``PITLossWrapper`` with the contract of look2hear/losses/pit_wrapper.py:15-67 on the fused CUDA loss."""
import torch
from torch import nn

from .._lib import check, lib, ptr, stream_ptr
from .matrix import MultiSrcNegSDR, PairwiseNegSDR, SingleSrcNegSDR, pit_sdr_forward, pitn_sdr_forward


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


class _PitNSdrFunction(torch.autograd.Function):
    """General n_src / pit_from path (csrc/loss_n.cu): pair matrix, best permutation (itertools order, first minimum), threshold, mean."""

    @staticmethod
    def forward(ctx, ests, targets, sdr_type, threshold):
        loss, pw, perm, ws = pitn_sdr_forward(ests, targets, sdr_type, threshold)
        ctx.save_for_backward(ests, targets)
        ctx.ws = ws
        ctx.mark_non_differentiable(perm)
        return loss.reshape(()), perm

    @staticmethod
    def backward(ctx, g, _):
        ests, targets = ctx.saved_tensors
        B, N, T = ests.shape
        d = torch.empty_like(ests)
        check(lib().dp_pitn_loss_backward(ptr(ests), ptr(targets), B, N, T, ptr(ctx.ws), 1.0, ptr(d), stream_ptr()), "dp_pitn_loss_backward")
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
        # pit_wrapper.py:30-67.  Every (pit_from, loss class) pair of the reference reduces to the same pair matrix [b, est, tgt]:
        #   pw_mtx + PairwiseNegSDR (matrix.py:22-57); pw_pt + SingleSrcNegSDR per pair (pit_wrapper.py:69-77, matrix.py:75-106);
        #   perm_avg + MultiSrcNegSDR = mean over targets of the permuted pairs, without threshold_byloss (pit_wrapper.py:37-45,79-88)
        expected = {"pw_mtx": PairwiseNegSDR, "pw_pt": SingleSrcNegSDR, "perm_avg": MultiSrcNegSDR}[self.pit_from]
        if not isinstance(self.loss_func, expected) or self.perm_reduce is not None or kwargs or reduce_kwargs:
            raise NotImplementedError(
                f"the fused path covers pit_from='{self.pit_from}' with a {expected.__name__} loss, perm_reduce=None and no extra kwargs"
            )
        if targets.size() != ests.size() or targets.ndim != 3:
            raise TypeError(f"Inputs must be of shape [batch, n_src, time], got {targets.size()} and {ests.size()} instead")
        ests_c, targets_c = ests.contiguous(), targets.contiguous()
        threshold = bool(self.threshold_byloss) and self.pit_from != "perm_avg"
        if self.pit_from == "pw_mtx" and targets.shape[1] == 2:   # the path of every config of the reference
            mean_loss, perm = _PitSdrFunction.apply(ests_c, targets_c, self.loss_func.sdr_type, threshold)
            if not return_ests:
                return mean_loss
            return mean_loss, self.reordered_sources(ests_c, perm)
        mean_loss, perm = _PitNSdrFunction.apply(ests_c, targets_c, self.loss_func.sdr_type, threshold)
        if not return_ests:
            return mean_loss
        return mean_loss, self.reordered_sources(ests_c, perm)

    @staticmethod
    def reordered_sources(source, perm):
        """pit_wrapper.py:90-94.  ``perm[b]`` = 0 (identity) / 1 (swapped) from the n_src = 2 kernel, or ``perm[b, i]`` = estimate index for
        target i (the reference's ``batch_indices``)."""
        B, N, T = source.shape
        out = torch.empty_like(source)
        if perm.ndim == 1:
            check(lib().dp_pit_reorder(ptr(source), ptr(perm), ptr(out), B, T, stream_ptr()), "dp_pit_reorder")
        else:
            check(lib().dp_pitn_reorder(ptr(source), ptr(perm.to(torch.int32).contiguous()), ptr(out), B, N, T, stream_ptr()), "dp_pitn_reorder")
        return out
