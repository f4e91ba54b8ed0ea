"""``PairwiseNegSDR`` with the contract of look2hear/losses/matrix.py:13-57, computed by the fused CUDA kernels."""
import torch
from torch.nn.modules.loss import _Loss

from .._lib import check, lib, ptr, require_cuda, stream_ptr

SDR_TYPES = {"snr": 0, "sisdr": 1, "sdsdr": 2}


def _check_inputs(ests, targets):
    if targets.size() != ests.size() or targets.ndim != 3:
        raise TypeError(f"Inputs must be of shape [batch, n_src, time], got {targets.size()} and {ests.size()} instead")
    if targets.shape[1] != 2:
        raise NotImplementedError("the fused PIT/SDR kernels are specialised for n_src == 2 (every config of the reference)")
    require_cuda(ests, "ests")
    require_cuda(targets, "targets")
    if ests.dtype != torch.float32 or targets.dtype != torch.float32:
        raise TypeError("ests/targets must be float32")


def pit_sdr_forward(ests, targets, sdr_type: str, threshold_byloss: bool):
    """One fused pass: returns ``(loss[1], pw[B,2,2], perm[B] int32, ws)``."""
    _check_inputs(ests, targets)
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
        return pit_sdr_forward(ests.contiguous(), targets.contiguous(), self.sdr_type, False)[1]


pairwise_neg_sisdr = PairwiseNegSDR("sisdr")
pairwise_neg_sdsdr = PairwiseNegSDR("sdsdr")
pairwise_neg_snr = PairwiseNegSDR("snr")
