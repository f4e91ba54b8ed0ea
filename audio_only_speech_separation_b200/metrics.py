"""Evaluation loop and metrics of the reference on the fused loss kernels.

``MetricsTracker`` keeps the interface of look2hear/metrics/wrapper.py:18-80 (``tracker(mix, clean, estimate, key)``, ``update()``,
``final()``, the ``metrics.csv`` columns).  SI-SNR and SI-SNRi come from the fused pairwise SI-SDR + PIT kernel (the same
``PITLossWrapper(PairwiseNegSDR("sisdr"))`` the reference builds, wrapper.py:28,33-36) and stay on the device: nothing is synchronised
per utterance (the reference calls ``.item()`` six times per utterance); rows are materialised in ``update()`` / ``final()``.

The ``sdr`` / ``sdr_i`` columns of the reference come from ``fast_bss_eval.sdr_pit_loss`` (BSS-eval SDR with a 512-tap distortion
filter, wrapper.py:38-41), a third-party package that is neither vendored in the reference nor installed in this image.  Its published
algorithm is restated as CUDA kernels (csrc/bss_sdr.cu: unit-norm rows, 512-lag correlations, Levinson solve of the Toeplitz systems in
fp64, coherence -> dB, best permutation) with the reference's argument order kept: ``sdr_pit_loss(clean, estimate)`` passes the clean
sources as the *estimates* and the network output as the *references* of the projection (and ``(mix, clean)`` for the baseline).
Pinned against a numpy / scipy fp64 restatement and known-answer properties (tests/test_metrics_sdr.py), not against the package.

``evaluate`` is ``audio_test.py:72-81`` with utterances of equal length batched together (every utterance is independent through the
network -- DESIGN.md section 6 -- so a batch of equal-length mixtures gives the per-utterance results of the reference's one-by-one loop).
"""
from __future__ import annotations

import csv
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from ._lib import check, lib, ptr, require_cuda, stream_ptr
from .losses.matrix import pit_sdr_forward

CSV_COLUMNS = ["snt_id", "sdr", "sdr_i", "si-snr", "si-snr_i"]


def pit_si_snr(estimates: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    """Per-utterance permutation-invariant SI-SNR in dB (``[B]``, device tensor): ``-min_perm mean_i pairwise_neg_sisdr``."""
    _, pw, _, _ = pit_sdr_forward(estimates.contiguous(), targets.contiguous(), "sisdr", False)
    # pw[b, est, tgt]; n_src = 2: identity vs swapped assignment (pit_wrapper.py:96-131)
    ident = 0.5 * (pw[:, 0, 0] + pw[:, 1, 1])
    swap = 0.5 * (pw[:, 1, 0] + pw[:, 0, 1])
    return -torch.minimum(ident, swap)


def bss_sdr_pit(est: torch.Tensor, ref: torch.Tensor, filter_length: int = 512, return_matrix: bool = False):
    """``-fast_bss_eval.sdr_pit_loss(est, ref).mean()`` per batch row: ``est, ref [B, n_src, T]`` (fp32, CUDA) -> ``[B]`` mean SDR in dB under
    the best permutation (device tensor, no host synchronisation); ``return_matrix`` adds the ``[B, n_ref, n_est]`` SDR matrix."""
    if est.shape != ref.shape or est.ndim != 3:
        raise TypeError(f"expected est/ref [B,n_src,T]; got {tuple(est.shape)}, {tuple(ref.shape)}")
    require_cuda(est, "est")
    require_cuda(ref, "ref")
    est, ref = est.float().contiguous(), ref.float().contiguous()
    B, n, T = est.shape
    ws = torch.empty(lib().dp_bss_sdr_workspace_bytes(B, n, filter_length), device=est.device, dtype=torch.uint8)
    out = torch.empty(B, device=est.device, dtype=torch.float32)
    mat = torch.empty(B, n, n, device=est.device, dtype=torch.float32) if return_matrix else None
    check(lib().dp_bss_sdr_pit(ptr(est), ptr(ref), B, n, T, filter_length, ptr(ws), ptr(out), ptr(mat), stream_ptr()), "dp_bss_sdr_pit")
    return (out, mat) if return_matrix else out


class MetricsTracker:
    def __init__(self, save_file: str = ""):
        self.all_sdrs: List[float] = []
        self.all_sdrs_i: List[float] = []
        self.all_sisnrs: List[float] = []
        self.all_sisnrs_i: List[float] = []
        self.results_csv = open(save_file, "w") if save_file else None
        self.writer = csv.DictWriter(self.results_csv, fieldnames=CSV_COLUMNS) if self.results_csv else None
        if self.writer:
            self.writer.writeheader()
        self._pending: List[Tuple] = []   # (keys, si_snr[B], si_snr_i[B], sdr[B], sdr_i[B]) on the device

    # -- reference interface: one utterance -------------------------------------------------------------------------------
    def __call__(self, mix: torch.Tensor, clean: torch.Tensor, estimate: torch.Tensor, key: str):
        """``mix [T]``, ``clean [n_src, T]``, ``estimate [n_src, T]`` (wrapper.py:31)."""
        self.add_batch(mix.unsqueeze(0), clean.unsqueeze(0), estimate.unsqueeze(0), [key])

    # -- batched entry point ------------------------------------------------------------------------------------------------
    def add_batch(self, mix: torch.Tensor, clean: torch.Tensor, estimate: torch.Tensor, keys: Sequence[str]):
        """``mix [B,T]``, ``clean [B,n_src,T]``, ``estimate [B,n_src,T]``; no host synchronisation."""
        if clean.shape != estimate.shape or clean.ndim != 3 or mix.shape != (clean.shape[0], clean.shape[2]):
            raise TypeError(f"expected mix [B,T], clean/estimate [B,n_src,T]; got {tuple(mix.shape)}, {tuple(clean.shape)}, {tuple(estimate.shape)}")
        si = pit_si_snr(estimate, clean)
        mixes = mix.unsqueeze(1).expand(-1, clean.shape[1], -1).contiguous()   # the mixture as every estimate (wrapper.py:34)
        base = pit_si_snr(mixes, clean)
        # wrapper.py:38-40, argument order as written there: sdr_pit_loss(est = clean, ref = estimate), baseline sdr_pit_loss(est = mix, ref = clean)
        sdr = bss_sdr_pit(clean, estimate)
        sdr_base = bss_sdr_pit(mixes, clean)
        self._pending.append((list(keys), si, si - base, sdr, sdr - sdr_base))

    def _flush(self):
        for keys, si, si_i, sdr, sdr_i in self._pending:
            cols = [t.detach().cpu().tolist() for t in (si, si_i, sdr, sdr_i)]
            for k, a, b, c, d in zip(keys, *cols):
                row = {"snt_id": k, "sdr": c, "sdr_i": d, "si-snr": a, "si-snr_i": b}
                if self.writer:
                    self.writer.writerow(row)
                self.all_sdrs.append(c)
                self.all_sdrs_i.append(d)
                self.all_sisnrs.append(a)
                self.all_sisnrs_i.append(b)
        self._pending = []

    def update(self):
        self._flush()
        return {"sdr_i": float(np.array(self.all_sdrs_i).mean()) if self.all_sdrs_i else float("nan"),
                "si-snr_i": float(np.array(self.all_sisnrs_i).mean()) if self.all_sisnrs_i else float("nan")}

    def final(self):
        self._flush()
        for name, fn in (("avg", np.mean), ("std", np.std)):
            row = {"snt_id": name, "sdr": float(fn(np.array(self.all_sdrs))), "sdr_i": float(fn(np.array(self.all_sdrs_i))),
                   "si-snr": float(fn(np.array(self.all_sisnrs))), "si-snr_i": float(fn(np.array(self.all_sisnrs_i)))}
            if self.writer:
                self.writer.writerow(row)
        if self.results_csv:
            self.results_csv.close()
            self.results_csv = None
        return {"si-snr": float(np.mean(self.all_sisnrs)), "si-snr_i": float(np.mean(self.all_sisnrs_i))}


@torch.no_grad()
def evaluate(model, test_set: Iterable, metrics: Optional[MetricsTracker] = None, batch_size: int = 16, device="cuda"):
    """``audio_test.py:72-81``: forward every ``(mix [T], sources [n_src, T], key)`` of ``test_set`` and track the metrics.

    Utterances are grouped by exact length and run ``batch_size`` at a time (equal-length utterances share a batch without changing
    any utterance's result; different lengths are never padded together, which would change the normalisation statistics).
    ``Sepformer`` is the exception: the engine reproduces the reference's ``(spk, batch) -> (batch, spk)`` reshape quirk
    (sepformer.py:1004), so for B > 1 row ``b`` of its output mixes utterances; like the reference's ``audio_test.py``
    (one utterance per call) it is evaluated one utterance at a time.  Returns the tracker.
    """
    metrics = metrics if metrics is not None else MetricsTracker()
    buckets = {}
    if getattr(model, "batch_rows_scrambled", False):
        batch_size = 1

    def run(items):
        mix = torch.stack([m for m, _, _ in items]).to(device, non_blocking=True)
        src = torch.stack([s for _, s, _ in items]).to(device, non_blocking=True)
        est = model(mix)
        metrics.add_batch(mix, src, est, [k for _, _, k in items])

    for mix, sources, key in test_set:
        items = buckets.setdefault(int(mix.shape[-1]), [])
        items.append((mix, sources, key))
        if len(items) == batch_size:
            run(items)
            items.clear()
    for items in buckets.values():
        if items:
            run(items)
    return metrics
