"""Loss lookup surface of look2hear/losses/__init__.py for the hot path."""
from .matrix import (
    MultiSrcNegSDR,
    PairwiseNegSDR,
    SingleSrcNegSDR,
    multisrc_neg_sdsdr,
    multisrc_neg_sisdr,
    multisrc_neg_snr,
    pairwise_neg_sdsdr,
    pairwise_neg_sisdr,
    pairwise_neg_snr,
    singlesrc_neg_sdsdr,
    singlesrc_neg_sisdr,
    singlesrc_neg_snr,
)
from .pit_wrapper import PITLossWrapper

__all__ = [
    "PITLossWrapper",
    "PairwiseNegSDR",
    "SingleSrcNegSDR",
    "MultiSrcNegSDR",
    "pairwise_neg_sisdr",
    "pairwise_neg_sdsdr",
    "pairwise_neg_snr",
    "singlesrc_neg_sisdr",
    "singlesrc_neg_sdsdr",
    "singlesrc_neg_snr",
    "multisrc_neg_sisdr",
    "multisrc_neg_sdsdr",
    "multisrc_neg_snr",
]
