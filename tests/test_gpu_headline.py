"""Parity at the BASELINE.json headline shapes against goldens written by the REAL reference:
unfolded DPRNN with the golden utterance embedded in a batch of 32 (C3), DPTNet at T = 32000 alone and embedded in a batch of 16
(C4, fp32 gate and bf16 0.05 dB gate), SepFormer base at T = 128000 and T = 256000 (C5: 130 / 258-position inter-chunk sequences).
Gates: fp32 mode rel-L2 <= 1e-4; bf16 mode |dPIT-SI-SNR| <= 0.05 dB against the fp32 reference output."""
import os
import sys

import pytest
import torch
import yaml

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import headline as HL  # noqa: E402
from conftest import record  # noqa: E402

pytestmark = pytest.mark.gpu

CFG = {
    "dprnn_lrs2_unfolded": dict(enc_dim=64, bn_dim=64, hidden_dim=128, win=16, layer=6, num_spk=2, module="DPRNN", group_size=1, block_size=100,
                                unfold=True),
    "dptnet_wsj0": dict(enc_dim=64, bn_dim=64, hidden_dim=128, win=16, layer=6, num_spk=2, module="DPTNet", group_size=1, block_size=100, unfold=False),
}


def _tasnet(cfgname, precision):
    from audio_only_speech_separation_b200.models import TasNet

    torch.manual_seed(0)
    m = TasNet(sample_rate=8000, **CFG[cfgname]).cuda().eval()
    m.precision = precision
    return m


def _sepformer(precision):
    from audio_only_speech_separation_b200.models import Sepformer

    torch.manual_seed(0)
    m = Sepformer(sample_rate=8000).cuda().eval()   # defaults == configs/sepformer_base.yml
    m.precision = precision
    return m


# (case, config, batch sizes, bf16 gate in dB).  BASELINE.json names bf16 for DPTNet and SepFormer only; for the fp32 unfolded-DPRNN config
# the bf16 figure is recorded and bounded loosely: with random weights its PIT-SI-SNR sits at -33 dB, where a 0.7 % output change that is
# correlated with the mixture (SI-SDR(new || reference) = 43.5 dB, better than the reference's own autocast run, SURVEY 7 hard part 8)
# moves the projection on the target by 0.11 dB.
@pytest.mark.parametrize("case,cfgname,batches,bf16_db", [("headline_dprnn_unfold_t32000", "dprnn_lrs2_unfolded", (1, 32), 0.25),
                                                          ("headline_dptnet_t32000", "dptnet_wsj0", (1, 16), 0.05)])
def test_tasnet_headline_shapes(case, cfgname, batches, bf16_db):
    x, s, y_ref, meta = HL.load_case(case)
    for B in batches:
        xb, row = HL.embed_batch(x, B)
        with torch.no_grad():
            y32 = _tasnet(cfgname, "fp32")(xb.cuda())[row : row + 1].cpu()
            y16 = _tasnet(cfgname, "bf16")(xb.cuda())[row : row + 1].cpu()
        r = HL.rel_l2(y32, y_ref)
        gate = HL.bf16_gate(y16, y_ref, s)
        record(f"{case}_B{B}", rel_l2_fp32=r, **gate)
        assert r <= 1e-4, (case, B, r)
        assert gate["delta_pit_sisnr_db"] <= bf16_db and gate["sisdr_vs_reference_db"] >= 38.0, (case, B, gate)


@pytest.mark.parametrize("case", ["headline_sepformer_t128000", "headline_sepformer_t256000"])
def test_sepformer_headline_shapes(case):
    x, s, y_ref, meta = HL.load_case(case)
    with torch.no_grad():
        y32 = _sepformer("fp32")(x.cuda()).cpu()
        y16 = _sepformer("bf16")(x.cuda()).cpu()
    r = HL.rel_l2(y32, y_ref)
    gate = HL.bf16_gate(y16, y_ref, s)
    record(case, rel_l2_fp32=r, **gate)
    assert r <= 1e-4, (case, r)
    assert gate["delta_pit_sisnr_db"] <= 0.05, (case, gate)


def test_headline_configs_match_reference_yaml():
    """The ctor kwargs used above are the reference's YAML files (staged copy on the GPU box, /root/reference in the build container)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for base in ("/root/reference/configs", os.path.join(root, "baseline", "_ref", "configs")):
        if os.path.isdir(base):
            for name, kw in CFG.items():
                ac = yaml.safe_load(open(os.path.join(base, name + ".yml")))["audionet"]["audionet_config"]
                assert {k: ac[k] for k in kw if k in ac} == {k: kw[k] for k in kw if k in ac}, name
            return
    pytest.skip("no copy of the reference configs here")
