"""``PairwiseNegSDR`` with the contract of look2hear/losses/matrix.py:13-57, computed by the fused CUDA kernels."""
import torch
from torch.nn.modules.loss import _Loss

from .._lib import check, lib, ptr, require_cuda, stream_ptr

SDR_TYPES = {"snr": 0, "sisdr": 1, "sdsdr": 2}


MAX_SRC = 4  # the general kernels (csrc/loss_n.cu) are instantiated for n_src = 1 .. 4


def _check_inputs(ests, targets):
    if targets.size() != ests.size() or targets.ndim != 3:
        raise TypeError(f"Inputs must be of shape [batch, n_src, time], got {targets.size()} and {ests.size()} instead")
    if not 1 <= targets.shape[1] <= MAX_SRC:
        raise NotImplementedError(f"the fused PIT/SDR kernels are built for n_src <= {MAX_SRC}")
    require_cuda(ests, "ests")
    require_cuda(targets, "targets")
    if ests.dtype != torch.float32 or targets.dtype != torch.float32:
        raise TypeError("ests/targets must be float32")


def pitn_sdr_forward(ests, targets, sdr_type: str, threshold_byloss: bool):
    """General n_src (1 .. 4): returns ``(loss[1], pw[B,N,N], perm[B,N] int32 = estimate index per target, ws)``."""
    _check_inputs(ests, targets)
    B, N, T = ests.shape
    dev = ests.device
    ws = torch.empty(lib().dp_pitn_loss_workspace_bytes(B, N), device=dev, dtype=torch.uint8)
    pw = torch.empty(B, N, N, device=dev, dtype=torch.float32)
    loss = torch.empty(1, device=dev, dtype=torch.float32)
    perm = torch.empty(B, N, device=dev, dtype=torch.int32)
    check(
        lib().dp_pitn_loss_forward(ptr(ests), ptr(targets), B, N, T, SDR_TYPES[sdr_type], int(bool(threshold_byloss)), ptr(ws), ptr(pw),
                                   ptr(loss), ptr(perm), stream_ptr()),
        "dp_pitn_loss_forward",
    )
    return loss, pw, perm, ws


def pit_sdr_forward(ests, targets, sdr_type: str, threshold_byloss: bool):
    """One fused pass for n_src = 2: returns ``(loss[1], pw[B,2,2], perm[B] int32, ws)``."""
    _check_inputs(ests, targets)
    if ests.shape[1] != 2:
        raise NotImplementedError("pit_sdr_forward is the n_src == 2 path; use pitn_sdr_forward")
    B, _, T = ests.shape
    dev = ests.device
    ws = torch.empty(lib().dp_pit_loss_workspace_bytes(B), device=dev, dtype=torch.uint8)
    pw = torch.empty(B, 2, 2, device=dev, dtype=torch.float32)
    loss = torch.empty(1, device=dev, dtype=torch.float32)
    perm = torch.empty(B, device=dev, dtype=torch.int32)
    check(
        lib().dp_pit_loss_forward(ptr(ests), ptr(targets), B, T, SDR_TYPES[sdr_type], int(bool(threshold_byloss)), ptr(ws), ptr(pw),
                                  ptr(loss), ptr(perm), stream_ptr()),
        "dp_pit_loss_forward",
    )
    return loss, pw, perm, ws


class PairwiseNegSDR(_Loss):
    """Pairwise negative SNR / SI-SDR / SD-SDR ``[batch, n_est, n_tgt]`` (matrix.py:13-57).

    Only the configuration the reference instantiates (``zero_mean=True, take_log=True, EPS=1e-8``) is built.
    The pair matrix itself is returned without an autograd graph; gradients flow through
    :class:`~.pit_wrapper.PITLossWrapper`, which fuses the permutation search and the backward.
    """

    def __init__(self, sdr_type, zero_mean=True, take_log=True, EPS=1e-8):
        super().__init__()
        assert sdr_type in ["snr", "sisdr", "sdsdr"]
        if not zero_mean or not take_log or EPS != 1e-8:
            raise NotImplementedError("only zero_mean=True, take_log=True, EPS=1e-8 (the reference's singletons) are built")
        self.sdr_type = sdr_type
        self.zero_mean = zero_mean
        self.take_log = take_log
        self.EPS = EPS

    def forward(self, ests, targets):
        _check_inputs(ests, targets)
        if ests.shape[1] == 2:
            return pit_sdr_forward(ests.contiguous(), targets.contiguous(), self.sdr_type, False)[1]
        return pitn_sdr_forward(ests.contiguous(), targets.contiguous(), self.sdr_type, False)[1]


class SingleSrcNegSDR(_Loss):
    """Negative SNR / SI-SDR / SD-SDR of one (estimate, target) pair per batch row, ``[batch, time] -> [batch]`` (matrix.py:60-106):
    the n_src = 1 case of the pair matrix.  Like :class:`PairwiseNegSDR` the value carries no autograd graph; used through
    ``PITLossWrapper(pit_from="pw_pt")`` it is differentiated by the fused backward."""

    def __init__(self, sdr_type, zero_mean=True, take_log=True, reduction="none", EPS=1e-8):
        assert reduction != "sum", NotImplementedError
        super().__init__(reduction=reduction)
        assert sdr_type in ["snr", "sisdr", "sdsdr"]
        if not zero_mean or not take_log or EPS != 1e-8:
            raise NotImplementedError("only zero_mean=True, take_log=True, EPS=1e-8 (the reference's singletons) are built")
        self.sdr_type = sdr_type
        self.zero_mean = zero_mean
        self.take_log = take_log
        self.EPS = 1e-8

    def forward(self, ests, targets):
        if targets.size() != ests.size() or targets.ndim != 2:
            raise TypeError(f"Inputs must be of shape [batch, time], got {targets.size()} and {ests.size()} instead")
        pw = pitn_sdr_forward(ests.contiguous().unsqueeze(1), targets.contiguous().unsqueeze(1), self.sdr_type, False)[1]
        losses = pw.reshape(-1)
        return losses.mean() if self.reduction == "mean" else losses


class MultiSrcNegSDR(_Loss):
    """Mean over sources of the negative SDR of estimate i against target i, ``[batch, n_src, time] -> [batch]`` (matrix.py:109-152):
    the mean of the pair matrix's diagonal.  No autograd graph (see :class:`PairwiseNegSDR`)."""

    def __init__(self, sdr_type, zero_mean=True, take_log=True, EPS=1e-8):
        super().__init__()
        assert sdr_type in ["snr", "sisdr", "sdsdr"]
        if not zero_mean or not take_log or EPS != 1e-8:
            raise NotImplementedError("only zero_mean=True, take_log=True, EPS=1e-8 (the reference's singletons) are built")
        self.sdr_type = sdr_type
        self.zero_mean = zero_mean
        self.take_log = take_log
        self.EPS = 1e-8

    def forward(self, ests, targets):
        _check_inputs(ests, targets)
        pw = pitn_sdr_forward(ests.contiguous(), targets.contiguous(), self.sdr_type, False)[1]
        return torch.diagonal(pw, dim1=1, dim2=2).mean(dim=-1)


pairwise_neg_sisdr = PairwiseNegSDR("sisdr")
pairwise_neg_sdsdr = PairwiseNegSDR("sdsdr")
pairwise_neg_snr = PairwiseNegSDR("snr")
singlesrc_neg_sisdr = SingleSrcNegSDR("sisdr")
singlesrc_neg_sdsdr = SingleSrcNegSDR("sdsdr")
singlesrc_neg_snr = SingleSrcNegSDR("snr")
multisrc_neg_sisdr = MultiSrcNegSDR("sisdr")
multisrc_neg_sdsdr = MultiSrcNegSDR("sdsdr")
multisrc_neg_snr = MultiSrcNegSDR("snr")
