"""SDR / SDRi of the evaluation loop (look2hear/metrics/wrapper.py:38-41 = fast_bss_eval.sdr_pit_loss, restated: csrc/bss_sdr.cu).
CPU: the fp64 numpy / scipy restatement (oracle/bss_sdr_oracle.py) against known answers of the definition.  GPU: the CUDA path against it."""
import numpy as np
import pytest
import torch

from oracle import bss_sdr_oracle as BO


def _signals(seed, n=2, T=6000):
    rng = np.random.default_rng(seed)
    # coloured (speech-like spectrum) references: white noise through a short AR filter
    ref = rng.standard_normal((n, T + 200))
    for k in range(1, ref.shape[1]):
        ref[:, k] += 0.9 * ref[:, k - 1]
    return ref[:, 200:] * 0.1, rng


def test_oracle_known_answers():
    ref, rng = _signals(0)
    n, T = ref.shape
    # an estimate that is a causal FIR-filtered reference (40 taps) plus white noise: the filtered part lies in the span of the 512 shifts,
    # so SDR ~= 10 log10(|filtered|^2 / |noise|^2) (the noise is almost orthogonal to that 512-dimensional subspace: T >> 512)
    h = rng.standard_normal(40) * np.exp(-np.arange(40) / 8.0)
    filt = np.stack([np.convolve(ref[i], h)[:T] for i in range(n)])
    noise = rng.standard_normal((n, T)) * 0.05
    est = filt + noise
    s = BO.sdr_matrix(est, ref)
    for i in range(n):
        expect = 10 * np.log10((filt[i] ** 2).sum() / (noise[i] ** 2).sum())
        assert abs(s[i, i] - expect) < 0.6, (s[i, i], expect)          # the projection also absorbs ~512/T of the noise energy
        assert s[i, i] - s[i, 1 - i] > 10                                # the wrong pairing is far worse
    # invariance to the scale of either argument; permutation solving
    assert np.allclose(BO.sdr_matrix(3.0 * est, 0.2 * ref), s, atol=1e-8)
    assert abs(BO.sdr_pit_mean(est[::-1], ref) - BO.sdr_pit_mean(est, ref)) < 1e-9
    assert abs(BO.sdr_pit_mean(est, ref) - 0.5 * (s[0, 0] + s[1, 1])) < 1e-12
    # a pure delay of up to 511 samples costs nothing, a delay beyond the filter does (references that end in silence, so that the delayed
    # copies are not truncated by the end of the signal)
    quiet = ref.copy()
    quiet[:, -800:] = 0.0
    d300 = np.concatenate([np.zeros((n, 300)), quiet[:, :-300]], axis=1)
    d700 = np.concatenate([np.zeros((n, 700)), quiet[:, :-700]], axis=1)
    assert BO.sdr_matrix(d300, quiet)[0, 0] > 40 and BO.sdr_matrix(d700, quiet)[0, 0] < 15


@pytest.mark.gpu
@pytest.mark.parametrize("n,T,seed", [(2, 6000, 1), (2, 32000, 2), (3, 5000, 3), (1, 4000, 4)])
def test_cuda_sdr_matches_oracle(n, T, seed):
    from audio_only_speech_separation_b200.metrics import bss_sdr_pit

    B = 3
    ests, refs = [], []
    for b in range(B):
        ref, rng = _signals(seed * 10 + b, n, T)
        perm = rng.permutation(n)
        est = ref[perm] * rng.uniform(0.5, 2.0) + rng.standard_normal((n, T)) * 0.1 * rng.uniform(0.02, 1.0)
        est = np.stack([np.convolve(e, [1.0, 0.4, -0.2])[:T] for e in est])
        ests.append(est)
        refs.append(ref)
    est_t = torch.tensor(np.stack(ests), dtype=torch.float32).cuda()
    ref_t = torch.tensor(np.stack(refs), dtype=torch.float32).cuda()
    out, mat = bss_sdr_pit(est_t, ref_t, return_matrix=True)
    for b in range(B):
        want_m = BO.sdr_matrix(est_t[b].cpu().numpy(), ref_t[b].cpu().numpy())
        assert np.abs(mat[b].cpu().numpy() - want_m).max() < 2e-3, (b, mat[b].cpu().numpy(), want_m)     # dB
        assert abs(out[b].item() - BO.sdr_pit_mean(est_t[b].cpu().numpy(), ref_t[b].cpu().numpy())) < 2e-3


@pytest.mark.gpu
def test_tracker_sdr_columns_follow_the_reference_call(tmp_path):
    """wrapper.py:38-40: sdr = -sdr_pit_loss(clean, estimate).mean(), baseline -sdr_pit_loss(mix, clean).mean(), sdr_i = sdr - baseline."""
    from audio_only_speech_separation_b200.metrics import MetricsTracker

    ref, rng = _signals(7, 2, 8000)
    est = ref[::-1] + rng.standard_normal(ref.shape) * 0.02
    mix = ref.sum(0)
    t = MetricsTracker(save_file=str(tmp_path / "m.csv"))
    t(torch.tensor(mix, dtype=torch.float32).cuda(), torch.tensor(ref, dtype=torch.float32).cuda(), torch.tensor(est.copy(), dtype=torch.float32).cuda(), "u")
    t.final()
    f32 = lambda a: np.asarray(a, dtype=np.float32)
    sdr = BO.sdr_pit_mean(f32(ref), f32(est))
    base = BO.sdr_pit_mean(f32(np.stack([mix, mix])), f32(ref))
    assert abs(t.all_sdrs[0] - sdr) < 2e-3 and abs(t.all_sdrs_i[0] - (sdr - base)) < 4e-3
