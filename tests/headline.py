"""Parity checks at the BASELINE.json headline shapes against the committed reference goldens (tests/golden/headline_*.npz).

Shared by the ``-m gpu`` tests and by ``bench.py`` (which reports the parity of the very run it times).  Nothing here touches
``oracle/`` or ``/root/reference``: the goldens were written by ``tests/golden/make_golden_headline.py`` in the build container,
the input is regenerated from its seed and pinned by the manifest's checksums.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SEED_X = 4321


def manifest():
    return json.load(open(os.path.join(GOLDEN, "headline_manifest.json")))


def headline_input(T):
    """(mixture [1, T], sources [1, 2, T]); same definition as make_golden_headline.headline_input."""
    g = torch.Generator().manual_seed(SEED_X)
    s = torch.randn(1, 2, T, generator=g) * 0.1
    return s.sum(1).contiguous(), s.contiguous()


def load_case(name):
    """(mixture [1,T], sources [1,2,T], reference output [1,2,T], manifest entry); asserts that the regenerated input is the pinned one."""
    meta = manifest()["cases"][name]
    x, s = headline_input(meta["T"])
    assert abs(float(x.double().sum()) - meta["x_sum"]) < 1e-6 * max(1.0, abs(meta["x_sum"])), "regenerated input differs from the golden's"
    assert abs(float(x.double().abs().sum()) - meta["x_abs_sum"]) < 1e-6 * meta["x_abs_sum"]
    y = torch.from_numpy(np.load(os.path.join(GOLDEN, f"{name}.npz"))["y"])
    return x, s, y, meta


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _sisdr_db(est, ref):
    """SI-SDR of ``est`` against ``ref`` per row, dB (zero-mean, scale-invariant)."""
    est = est.double() - est.double().mean(-1, keepdim=True)
    ref = ref.double() - ref.double().mean(-1, keepdim=True)
    proj = (est * ref).sum(-1, keepdim=True) / (ref.pow(2).sum(-1, keepdim=True) + 1e-30) * ref
    return 10 * torch.log10(proj.pow(2).sum(-1) / ((est - proj).pow(2).sum(-1) + 1e-30))


def pit_sisnr_db(est, src):
    """PIT SI-SNR (2 sources) of ``est [1,2,T]`` against ``src [1,2,T]`` in dB: what ``PITLossWrapper(pairwise_neg_sisdr)`` negates."""
    e, s = est.cpu()[0], src.cpu()[0]
    a = (_sisdr_db(e[0], s[0]) + _sisdr_db(e[1], s[1])) / 2
    b = (_sisdr_db(e[0], s[1]) + _sisdr_db(e[1], s[0])) / 2
    return float(torch.maximum(a, b))


def bf16_gate(y_new, y_ref, src):
    """SURVEY 8d gate (iii): |PIT-SI-SNR(new) - PIT-SI-SNR(fp32 reference)| in dB, plus SI-SDR(new || reference) as the stricter proxy."""
    return {"delta_pit_sisnr_db": abs(pit_sisnr_db(y_new, src) - pit_sisnr_db(y_ref, src)),
            "sisdr_vs_reference_db": float(_sisdr_db(y_new.cpu().reshape(-1), y_ref.cpu().reshape(-1)))}


def embed_batch(x, B, seed=77):
    """Batch of ``B`` utterances whose row ``B // 2`` is the golden utterance ``x [1,T]`` (the others are seeded noise mixtures)."""
    g = torch.Generator().manual_seed(seed)
    xb = torch.randn(B, x.shape[1], generator=g) * 0.1
    row = B // 2
    xb[row] = x[0]
    return xb.contiguous(), row
