"""CPU restatement (numpy / scipy, fp64) of the SDR metric of the reference's evaluation loop -- TEST INFRASTRUCTURE, never imported by
the product.

The reference computes ``sdr = -fast_bss_eval.sdr_pit_loss(clean, estimate).mean()`` (look2hear/metrics/wrapper.py:38-41).
``fast_bss_eval`` (pip dependency of the reference, not vendored, absent from this image and from /opt/wheelhouse) implements the BSS-eval
v4 SDR with an L = 512 tap distortion filter (R. Scheibler, "SDR -- Medium Rare with Fast Computations", ICASSP 2022).  Its published
algorithm, with the defaults the reference uses (filter_length = 512, zero_mean = False, no diagonal loading, direct solve):

    est, ref <- rows scaled to unit L2 norm
    acf_i[l]    = irfft(|rfft(ref_i, n_fft)|^2)[l]              = sum_t ref_i[t] ref_i[t + l]         (n_fft >= T + L: linear correlation)
    xcorr_ij[l] = irfft(conj(rfft(ref_i)) rfft(est_j))[l]       = sum_t ref_i[t] est_j[t + l]
    h_ij        = solve(Toeplitz(acf_i), xcorr_ij)              (least-squares projection of est_j on the L shifts of ref_i)
    coh_ij      = <xcorr_ij, h_ij>;   SDR_ij = 10 log10(coh_ij / (1 - coh_ij))
    sdr_pit_loss = -SDR under the permutation of the estimates that maximises the summed SDR (linear_sum_assignment)

PARITY UNPINNED against the package itself (it cannot be installed here: no network); pinned instead by known-answer properties of the
definition (tests/test_metrics_sdr.py): an estimate that is an FIR-filtered reference (<= 512 taps) plus orthogonal noise has
SDR = 10 log10(|filtered|^2 / |noise_perp|^2), invariance to scaling, permutation solving.
"""
import itertools

import numpy as np
from scipy.linalg import solve_toeplitz


def _normalize(x, eps=1e-12):
    return x / np.maximum(np.linalg.norm(x, axis=-1, keepdims=True), eps)


def sdr_matrix(est, ref, filter_length=512):
    """``est, ref [n, T]`` -> SDR in dB ``[n_ref, n_est]`` of every estimate against every reference."""
    est = _normalize(np.asarray(est, dtype=np.float64))
    ref = _normalize(np.asarray(ref, dtype=np.float64))
    n, T = ref.shape
    L = filter_length
    n_fft = 2 ** int(np.ceil(np.log2(T + L)))
    R = np.fft.rfft(ref, n=n_fft, axis=-1)
    E = np.fft.rfft(est, n=n_fft, axis=-1)
    acf = np.fft.irfft(R.real**2 + R.imag**2, n=n_fft, axis=-1)[:, :L]
    out = np.empty((n, est.shape[0]))
    for i in range(n):
        for j in range(est.shape[0]):
            xcorr = np.fft.irfft(np.conj(R[i]) * E[j], n=n_fft)[:L]
            h = solve_toeplitz(acf[i], xcorr)
            coh = float(np.clip(xcorr @ h, np.finfo(np.float32).eps, 1.0 - np.finfo(np.float32).eps))
            out[i, j] = 10.0 * np.log10(coh / (1.0 - coh))
    return out


def sdr_pit_mean(est, ref, filter_length=512):
    """``-fast_bss_eval.sdr_pit_loss(est, ref).mean()``: mean SDR under the best assignment of estimates to references."""
    s = sdr_matrix(est, ref, filter_length)
    n = s.shape[0]
    return max(sum(s[i, p[i]] for i in range(n)) for p in itertools.permutations(range(n))) / n
