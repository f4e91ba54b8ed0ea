"""Loss lookup surface of look2hear/losses/__init__.py for the hot path."""
from .matrix import PairwiseNegSDR, pairwise_neg_sdsdr, pairwise_neg_sisdr, pairwise_neg_snr
from .pit_wrapper import PITLossWrapper

__all__ = ["PITLossWrapper", "PairwiseNegSDR", "pairwise_neg_sisdr", "pairwise_neg_sdsdr", "pairwise_neg_snr"]
