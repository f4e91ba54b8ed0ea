"""Evaluation metrics / eval loop: interface on CPU, values against the reference's golden vectors on the GPU."""
import csv
import math

import numpy as np
import pytest
import torch

from conftest import load_npz


def test_tracker_interface_and_csv(tmp_path):
    from audio_only_speech_separation_b200.metrics import CSV_COLUMNS, MetricsTracker

    assert CSV_COLUMNS == ["snt_id", "sdr", "sdr_i", "si-snr", "si-snr_i"]          # look2hear/metrics/wrapper.py:24
    t = MetricsTracker(save_file=str(tmp_path / "metrics.csv"))
    with pytest.raises(TypeError):
        t.add_batch(torch.zeros(2, 10), torch.zeros(2, 2, 11), torch.zeros(2, 2, 11), ["a", "b"])
    # rows are buffered on the device; inject two by hand to exercise update() / final() without a GPU
    t._pending.append((["u1", "u2"], torch.tensor([10.0, 12.0]), torch.tensor([9.0, 11.0]), torch.tensor([11.0, 13.0]), torch.tensor([8.0, 10.0])))
    assert t.update() == {"sdr_i": 9.0, "si-snr_i": 10.0}
    out = t.final()
    assert out["si-snr"] == 11.0 and out["si-snr_i"] == 10.0
    rows = list(csv.DictReader(open(tmp_path / "metrics.csv")))
    assert [r["snt_id"] for r in rows] == ["u1", "u2", "avg", "std"]
    assert float(rows[2]["si-snr"]) == 11.0 and float(rows[3]["si-snr"]) == 1.0 and float(rows[0]["sdr"]) == 11.0
    assert float(rows[2]["sdr"]) == 12.0 and float(rows[2]["sdr_i"]) == 9.0


def test_evaluate_buckets_by_exact_length():
    from audio_only_speech_separation_b200.metrics import evaluate

    calls = []

    class Stub:
        def add_batch(self, mix, clean, est, keys):
            calls.append((tuple(mix.shape), list(keys)))

    def model(mix):
        return torch.stack([mix, mix], 1)

    data = [(torch.zeros(T), torch.zeros(2, T), f"k{i}") for i, T in enumerate([100, 100, 50, 100, 50, 100, 70])]
    evaluate(model, data, Stub(), batch_size=3, device="cpu")
    assert calls[0] == ((3, 100), ["k0", "k1", "k3"])            # a full batch of equal-length utterances runs as soon as it is complete
    assert sorted(c[0] for c in calls[1:]) == [(1, 70), (1, 100), (2, 50)]
    assert sorted(k for _, ks in calls for k in ks) == [f"k{i}" for i in range(7)]


@pytest.mark.gpu
def test_si_snr_matches_reference_golden(tmp_path):
    from audio_only_speech_separation_b200.metrics import MetricsTracker

    z = load_npz("metrics.npz")
    t = MetricsTracker(save_file=str(tmp_path / "m.csv"))
    for i in range(4):
        t(torch.from_numpy(z[f"mix{i}"]).cuda(), torch.from_numpy(z[f"clean{i}"]).cuda(), torch.from_numpy(z[f"est{i}"]).cuda(), f"utt{i}")
    t.final()
    for i in range(4):
        assert abs(t.all_sisnrs[i] - float(z[f"si_snr{i}"])) < 1e-3, i          # dB
        assert abs(t.all_sisnrs_i[i] - float(z[f"si_snr_i{i}"])) < 1e-3, i


@pytest.mark.gpu
def test_batched_evaluation_equals_one_by_one():
    from audio_only_speech_separation_b200.metrics import MetricsTracker, evaluate
    from audio_only_speech_separation_b200.models import TasNet

    torch.manual_seed(0)
    model = TasNet(sample_rate=8000, layer=2).cuda().eval()
    g = torch.Generator().manual_seed(5)
    data = []
    for i, T in enumerate([4000, 4000, 3000, 4000, 3000, 4000]):
        src = torch.randn(2, T, generator=g) * 0.1
        data.append((src.sum(0), src, f"utt{i}"))
    one = evaluate(model, data, MetricsTracker(), batch_size=1)
    one.final()
    many = evaluate(model, data, MetricsTracker(), batch_size=4)
    many.final()
    # batching changes the order rows are produced in, not their values
    a = dict(zip(["utt0", "utt1", "utt2", "utt3", "utt4", "utt5"], one.all_sisnrs_i))
    got = sorted(many.all_sisnrs_i)
    assert np.allclose(sorted(a.values()), got, atol=1e-4)
