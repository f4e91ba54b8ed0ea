"""Golden vectors for the evaluation metrics, from the REAL reference (build container only).

``look2hear.metrics.wrapper`` itself cannot be imported here (it needs ``fast_bss_eval``); its SI-SNR / SI-SNRi arithmetic
(wrapper.py:28,33-36) is replayed with the reference's own ``look2hear.losses`` classes, exactly as ``MetricsTracker.__call__`` does.
Writes ``tests/golden/metrics.npz``.   python -B tests/golden/make_golden_metrics.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")

from look2hear.losses import PITLossWrapper, PairwiseNegSDR  # noqa: E402

pit_sisnr = PITLossWrapper(PairwiseNegSDR("sisdr", zero_mean=True), pit_from="pw_mtx")   # wrapper.py:28
g = torch.Generator().manual_seed(77)
out = {}
for i, T in enumerate([4000, 8001, 16000, 1234]):
    clean = torch.randn(2, T, generator=g) * 0.1
    mix = clean.sum(0)
    noise = torch.randn(2, T, generator=g) * (0.01 * (i + 1))
    estimate = clean.flip(0) + noise if i % 2 else clean + noise          # odd cases: the estimates come out swapped
    sisnr = pit_sisnr(estimate.unsqueeze(0), clean.unsqueeze(0))          # wrapper.py:33
    mix2 = torch.stack([mix] * clean.shape[0], dim=0)                      # wrapper.py:34
    base = pit_sisnr(mix2.unsqueeze(0), clean.unsqueeze(0))               # wrapper.py:35
    out[f"mix{i}"], out[f"clean{i}"], out[f"est{i}"] = mix.numpy(), clean.numpy(), estimate.numpy()
    out[f"si_snr{i}"] = np.float64(-sisnr.item())                          # the "si-snr" column (wrapper.py:46)
    out[f"si_snr_i{i}"] = np.float64(-(sisnr - base).item())               # the "si-snr_i" column (wrapper.py:47)
np.savez_compressed(os.path.join(HERE, "metrics.npz"), **out)
print({k: float(v) for k, v in out.items() if k.startswith("si_")})
