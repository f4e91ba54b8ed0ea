"""CPU tests of the GroupComm path: the oracle against the reference's golden outputs, and the drop-in model's parameter containers
(state-dict keys, shapes and default initialisation of the reference)."""
import json
import os

import pytest
import torch

from conftest import GOLDEN, load_npz, rel_l2
from oracle import groupcomm_oracle as GO

GC_MANIFEST = json.load(open(os.path.join(GOLDEN, "groupcomm_manifest.json")))


def _model(case):
    from audio_only_speech_separation_b200.models import TasNet

    c = GC_MANIFEST["cases"][case]
    torch.manual_seed(c["seed"])
    return TasNet(**c["kwargs"]), c


@pytest.mark.parametrize("case", ["g16_b1_t300", "g16_b1_t3999_1d", "g16_unfold_b2_t4001", "dpt_g16_unfold_b1_t2000", "dpt_g8_l2_b1_t3000",
                                  "g8_l2_b2_t4000"])
def test_oracle_matches_reference_golden(case):
    m, c = _model(case)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    z = load_npz(f"groupcomm_{case}.npz")
    taps = {}
    with torch.no_grad():
        y = GO.tasnet_gc_forward(sd, torch.from_numpy(z["x"]), group_size=c["kwargs"]["group_size"], layer=c["kwargs"].get("layer", 6),
                                 unfold=c["kwargs"].get("unfold", False), module=c["kwargs"]["module"], lstm_impl="loop", taps=taps)
    assert rel_l2(y, torch.from_numpy(z["y"])) < 5e-6
    assert rel_l2(taps["squeeze_mean"], torch.from_numpy(z["squeeze_mean"])) < 5e-6
    assert rel_l2(taps["feature_map"], torch.from_numpy(z["feature_map"])) < 5e-6


@pytest.mark.parametrize("case", ["g16_b2_t8001", "g16_unfold_b2_t4001", "dpt_g16_b2_t4001", "dpt_g16_unfold_b1_t2000", "g8_l2_b2_t4000"])
def test_state_dict_is_the_reference_one(case):
    m, c = _model(case)
    sd, ref = m.state_dict(), GC_MANIFEST["state_dicts"][case]
    assert list(sd.keys()) == list(ref.keys())
    for k, v in sd.items():
        assert list(v.shape) == ref[k][2:], k
        assert abs(float(v.double().sum()) - ref[k][0]) <= 1e-9 * max(1.0, abs(ref[k][0])), k
        assert abs(float(v.double().abs().sum()) - ref[k][1]) <= 1e-9 * max(1.0, ref[k][1]), k
    assert sum(p.numel() for p in m.parameters()) == c["n_params"]
    from audio_only_speech_separation_b200 import _lib

    per_layer = 47 if c["kwargs"]["module"] == "DPTNet" else 35
    assert len(m._gc_param_table()) == 12 + 4 * 23 + per_layer * c["kwargs"].get("layer", 6)


@pytest.mark.parametrize("name", ["g16", "dpt_g16"])
def test_oracle_autograd_matches_reference_gradients(name):
    """Pins the oracle's backward for this path (the check the CUDA training backward will be held to): PIT-SNR loss and every
    parameter gradient of the reference's own ``loss.backward()``."""
    from oracle import dualpath_oracle as O

    m, c = _model(f"grads_{name}")
    z = load_npz(f"groupcomm_grads_{name}.npz")
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    kw = c["kwargs"]
    est = GO.tasnet_gc_forward(leaf, torch.from_numpy(z["x"]), group_size=kw["group_size"], layer=kw["layer"], module=kw["module"])
    loss = O.pit_loss(est, torch.from_numpy(z["tgt"]), "snr", False)
    loss.backward()
    assert abs(loss.item() - float(z["loss"])) < 1e-5
    for k, v in leaf.items():
        ref = torch.from_numpy(z["grad::" + k])
        assert rel_l2(v.grad, ref) < 1e-3, k


def test_unsupported_variants_raise():
    from audio_only_speech_separation_b200.models import TasNet

    gc = TasNet(module="DPRNN", group_size=16)
    with pytest.raises(NotImplementedError):   # the fused training step must never reach the dual-path engine with a GroupComm handle
        gc._engine_forward(torch.zeros(1, 800), True)
    with pytest.raises(NotImplementedError):
        gc._engine_backward(torch.zeros(1, 2, 800), None, None, 1, 800)
    from audio_only_speech_separation_b200.losses import PITLossWrapper, pairwise_neg_snr
    from audio_only_speech_separation_b200.trainer import DualPathTrainer

    DualPathTrainer(gc, PITLossWrapper(pairwise_neg_snr, pit_from="pw_mtx", threshold_byloss=False))   # GroupComm models train (csrc/groupcomm.cu)
    DualPathTrainer(TasNet(module="DPTNet", group_size=16), PITLossWrapper(pairwise_neg_snr, pit_from="pw_mtx", threshold_byloss=False))

    with pytest.raises(RuntimeError):   # CUDA-only: CPU tensors are refused, there is no CPU path
        TasNet(module="DPRNN", group_size=16).eval()(torch.zeros(1, 800))
