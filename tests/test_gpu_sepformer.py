"""GPU parity tests of the SepFormer path (look2hear/models/sepformer.py) against the committed reference outputs and the oracle."""
import pytest
import torch

from conftest import load_npz, record, rel_l2
from oracle import dualpath_oracle as O
from oracle import sepformer_oracle as SO

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4      # rel-L2, BASELINE.json north_star
BF16_TOL_DB = 0.05   # |delta PIT SI-SNR| in dB


def _model(manifest, case, precision="fp32"):
    from audio_only_speech_separation_b200.models import Sepformer

    c = manifest["cases"][case]
    torch.manual_seed(c["seed"])
    m = Sepformer(sample_rate=c["sample_rate"], **c["audionet_config"])
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.cuda().eval()
    m.precision = precision
    return m, sd, c["audionet_config"]


@pytest.mark.parametrize("case", ["sepformer_small_b2_t3000", "sepformer_small_b1_t1001_1d", "sepformer_smallpost_b1_t2001",
                                  "sepformer_base_b1_t16000"])
def test_forward_matches_reference_golden(manifest, case):
    m, _, _ = _model(manifest, case)
    z = load_npz(f"model_{case}.npz")
    with torch.no_grad():
        y = m(torch.from_numpy(z["x"]).cuda())
    ref = torch.from_numpy(z["y"])
    assert tuple(y.shape) == tuple(ref.shape)
    err = rel_l2(y, ref)
    record("sepformer_fwd_fp32", case=case, rel_l2=err, launches=m.last_launches)
    assert err < FP32_TOL


def test_forward_lengths_and_state_dict_reload(manifest):
    m, sd, cfg = _model(manifest, "sepformer_small_b2_t3000")
    g = torch.Generator().manual_seed(9)
    for shape in ((1, 16), (3, 23), (2, 1, 777), (1, 4096)):
        x = torch.randn(*shape, generator=g) * 0.1
        with torch.no_grad():
            y = m(x.cuda()).cpu()
            ref = SO.sepformer_forward(sd, x, **cfg)
        assert tuple(y.shape) == tuple(ref.shape)
        assert rel_l2(y, ref) < FP32_TOL, shape
    from audio_only_speech_separation_b200.models import Sepformer

    torch.manual_seed(321)
    other = Sepformer(sample_rate=8000, **cfg)
    m.load_state_dict(other.state_dict())
    x = torch.randn(1, 2000, generator=g) * 0.1
    with torch.no_grad():
        y = m(x.cuda()).cpu()
        ref = SO.sepformer_forward({k: v.detach() for k, v in other.state_dict().items()}, x, **cfg)
    assert rel_l2(y, ref) < FP32_TOL
    with pytest.raises(Exception):
        m(torch.randn(1, 15).cuda())  # shorter than the encoder kernel: the reference's conv1d fails too


def test_forward_bf16_within_si_snr_budget(manifest):
    """bf16 mode (single tensor-core product) on structured input: PIT SI-SNR within 0.05 dB of the fp32 reference output."""
    m, sd, cfg = _model(manifest, "sepformer_base_b1_t16000", precision="bf16")
    g = torch.Generator().manual_seed(21)
    src = torch.randn(1, 2, 16000, generator=g) * 0.1
    mix = src.sum(1)
    with torch.no_grad():
        ref = SO.sepformer_forward(sd, mix, **cfg)
        y = m(mix.cuda()).cpu()
    si_ref = -O.pit_loss(ref, src, "sisdr", False).item()
    si_new = -O.pit_loss(y, src, "sisdr", False).item()
    proxy = -O.pairwise_neg_sdr(y, ref, "sisdr").diagonal(dim1=1, dim2=2).mean().item()
    record("sepformer_fwd_bf16", si_ref=si_ref, si_new=si_new, si_sdr_vs_ref=proxy, rel_l2=rel_l2(y, ref))
    assert abs(si_new - si_ref) <= BF16_TOL_DB
    assert proxy > 30.0


def test_training_mode_is_refused_loudly(manifest):
    m, _, _ = _model(manifest, "sepformer_small_b2_t3000")
    m.train()
    with pytest.raises(NotImplementedError):
        m(torch.randn(1, 2000).cuda())
