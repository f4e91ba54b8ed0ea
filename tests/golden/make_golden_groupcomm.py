"""Golden vectors for the GroupComm path (``TasNet(..., group_size > 1)``) from the REAL reference (build container only:
``python -B tests/golden/make_golden_groupcomm.py``).

Runs the unmodified ``look2hear.models.TasNet`` with the configuration of the reference's own ``unit_tests.py:79-80``
(``module="DPRNN", enc_dim=64, bn_dim=64, group_size=16``, and ``unfold=True`` of ``:85-86``) and a ``group_size=8`` variant, asserts that
``oracle/groupcomm_oracle.py`` reproduces it (rel-L2 <= 5e-6, explicit LSTM time loop) and stores inputs, outputs and intermediate
taps of the oracle in ``tests/golden/groupcomm_*.npz`` plus ``groupcomm_manifest.json`` (per-key checksums of the default init).
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

from look2hear.models import TasNet  # noqa: E402

from oracle import groupcomm_oracle as GO  # noqa: E402

torch.set_num_threads(8)
CASES = [
    # name, constructor kwargs, B, T, input kind
    ("g16_b2_t8001", dict(module="DPRNN", enc_dim=64, bn_dim=64, group_size=16), 2, 8001, "2d"),
    ("g16_b1_t3999_1d", dict(module="DPRNN", enc_dim=64, bn_dim=64, group_size=16), 1, 3999, "1d"),
    ("g16_b1_t300", dict(module="DPRNN", enc_dim=64, bn_dim=64, group_size=16), 1, 300, "2d"),
    ("g16_unfold_b2_t4001", dict(module="DPRNN", enc_dim=64, bn_dim=64, group_size=16, unfold=True), 2, 4001, "2d"),
    ("dpt_g16_b2_t4001", dict(module="DPTNet", enc_dim=64, bn_dim=64, group_size=16), 2, 4001, "2d"),
    ("dpt_g16_unfold_b1_t2000", dict(module="DPTNet", enc_dim=64, bn_dim=64, group_size=16, unfold=True), 1, 2000, "2d"),
    ("dpt_g8_l2_b1_t3000", dict(module="DPTNet", enc_dim=64, bn_dim=64, group_size=8, layer=2), 1, 3000, "2d"),
    ("g8_l2_b2_t4000", dict(module="DPRNN", enc_dim=64, bn_dim=64, group_size=8, layer=2, sample_rate=8000), 2, 4000, "3d"),
]


def checksum(sd):
    return {k: [float(v.double().sum()), float(v.double().abs().sum())] + list(v.shape) for k, v in sd.items()}


manifest = {"torch": torch.__version__, "cases": {}, "state_dicts": {}}
for name, kw, B, T, kind in CASES:
    torch.manual_seed(0)
    m = TasNet(**kw).eval()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(4321)
    x = torch.randn(B, T, generator=g) * 0.1
    xin = {"2d": x, "1d": x[0], "3d": x.unsqueeze(1)}[kind]
    taps = {}
    with torch.no_grad():
        y = m(xin)
        yo = GO.tasnet_gc_forward(sd, xin, group_size=kw["group_size"], layer=kw.get("layer", 6), unfold=kw.get("unfold", False),
                                  module=kw["module"], lstm_impl="loop", taps=taps)
    r = ((yo - y).norm() / y.norm()).item()
    assert r < 5e-6, (name, r)
    np.savez_compressed(os.path.join(HERE, f"groupcomm_{name}.npz"), x=xin.numpy(), y=y.numpy(),
                        squeeze_mean=taps["squeeze_mean"].numpy(), feature_map=taps["feature_map"].numpy())
    key = json.dumps(kw, sort_keys=True)
    manifest["cases"][name] = {"kwargs": kw, "seed": 0, "oracle_rel_l2": r, "y_abs_sum": float(y.double().abs().sum()),
                               "n_params": sum(p.numel() for p in m.parameters())}
    manifest["state_dicts"][name] = checksum(sd)
    print(name, "oracle rel-L2", r, "params", manifest["cases"][name]["n_params"])
# ---------------------------------------------------------------- gradients (pins the oracle's autograd for the training backward of
# this path, which the CUDA engine does not have yet): reference loss.backward() on the unit-test configuration, all gradients stored
from look2hear.losses import PITLossWrapper, pairwise_neg_snr  # noqa: E402
from oracle import dualpath_oracle as O  # noqa: E402

for name, kw in [("g16", dict(module="DPRNN", enc_dim=64, bn_dim=64, group_size=16, layer=2)),
                 ("dpt_g16", dict(module="DPTNet", enc_dim=64, bn_dim=64, group_size=16, layer=2))]:
    torch.manual_seed(0)
    m = TasNet(**kw).train()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(99)
    x = torch.randn(2, 3000, generator=g) * 0.1
    tgt = torch.randn(2, 2, 3000, generator=g) * 0.1
    loss = PITLossWrapper(pairwise_neg_snr, pit_from="pw_mtx", threshold_byloss=False)(m(x), tgt)
    m.zero_grad()
    loss.backward()
    ref_grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    lo = O.pit_loss(GO.tasnet_gc_forward(leaf, x, group_size=16, layer=2, module=kw["module"], lstm_impl="loop"), tgt, "snr", False)
    lo.backward()
    worst = max(((leaf[k].grad - ref_grads[k]).norm() / ref_grads[k].norm().clamp_min(1e-12)).item() for k in ref_grads)
    assert abs(lo.item() - loss.item()) < 1e-5 and worst < 2e-3, (name, lo.item(), loss.item(), worst)
    gnpz = {"x": x.numpy(), "tgt": tgt.numpy(), "loss": np.array(loss.item())}
    for k, v in ref_grads.items():
        gnpz["grad::" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, f"groupcomm_grads_{name}.npz"), **gnpz)
    manifest["cases"][f"grads_{name}"] = {"kwargs": kw, "seed": 0, "oracle_worst_rel": worst, "loss": loss.item()}
    print("grads", name, "loss", loss.item(), "worst oracle-vs-reference rel", worst)

with open(os.path.join(HERE, "groupcomm_manifest.json"), "w") as f:
    json.dump(manifest, f, indent=1)
