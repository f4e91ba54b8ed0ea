"""Generate the committed SepFormer golden vectors from the REAL reference (build container only).

    python -B tests/golden/make_golden_sepformer.py

Runs the unmodified ``look2hear.models.Sepformer`` imported from ``/root/reference``, asserts that
``oracle/sepformer_oracle.py`` reproduces it (rel-L2 <= 2e-6) and that the drop-in model's default initialisation is
bit-identical under the same seed, then stores inputs / outputs in ``tests/golden/model_sepformer_*.npz`` and adds the
cases (config, seed, per-key float64 checksums of the weights) to ``manifest.json``.
"""
import json
import os
import sys

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from look2hear.models import Sepformer  # noqa: E402

from audio_only_speech_separation_b200.models import Sepformer as Ours  # noqa: E402
from oracle import sepformer_oracle as SO  # noqa: E402

torch.set_num_threads(8)
manifest = json.load(open(os.path.join(HERE, "manifest.json")))

base = yaml.safe_load(open(f"{REF}/configs/sepformer_base.yml"))["audionet"]["audionet_config"]
small = dict(encoder_out_nchannels=64, masknet_chunksize=50, masknet_numlayers=2, intra_numlayers=2, inter_numlayers=2, intra_nhead=4,
             inter_nhead=4, intra_dffn=128, inter_dffn=128)
CASES = [
    # name, config name, config, B, T, input kind
    ("sepformer_small_b2_t3000", "sepformer_small", small, 2, 3000, "2d"),          # B = 2 pins the (spk, batch) row scramble
    ("sepformer_small_b1_t1001_1d", "sepformer_small", small, 1, 1001, "1d"),
    ("sepformer_smallpost_b1_t2001", "sepformer_smallpost",
     dict(small, intra_norm_before=False, inter_norm_before=False, inter_use_positional=False), 1, 2001, "3d"),
    ("sepformer_base_b1_t16000", "sepformer_base", base, 1, 16000, "2d"),           # configs/sepformer_base.yml, 2 s @ 8 kHz
]


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def checksum(sd):
    return {k: [float(v.double().sum()), float(v.double().abs().sum())] + list(v.shape) for k, v in sd.items()}


done = set()
for name, cfgname, cfg, B, T, kind in CASES:
    torch.manual_seed(0)
    m = Sepformer(sample_rate=8000, **cfg).eval()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(B, T, generator=g) * 0.1
    xin = {"2d": x, "1d": x[0], "3d": x.unsqueeze(1)}[kind]
    with torch.no_grad():
        y = m(xin)
        yo = SO.sepformer_forward(sd, xin, **cfg)
    r = rel(yo, y)
    assert tuple(yo.shape) == tuple(y.shape) and r < 2e-6, (name, r)
    np.savez_compressed(os.path.join(HERE, f"model_{name}.npz"), x=xin.numpy(), y=y.numpy())
    manifest["cases"][name] = {"config": cfgname, "audionet_config": cfg, "sample_rate": 8000, "seed": 0, "oracle_rel_l2": r,
                               "y_abs_sum": float(y.double().abs().sum())}
    if cfgname not in done:
        done.add(cfgname)
        torch.manual_seed(0)
        ours = Ours(sample_rate=8000, **cfg)
        osd = ours.state_dict()
        assert list(osd.keys()) == list(sd.keys()) and all(torch.equal(osd[k], sd[k]) for k in sd), cfgname
        manifest["state_dicts"][cfgname] = checksum(sd)
        manifest["n_params"][cfgname] = sum(p.numel() for p in m.parameters())
    print(name, "oracle rel-L2", r)

json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)
print("manifest updated")

# ---------------------------------------------------------------- gradients (training semantics with dropout disabled)
from look2hear.losses import PITLossWrapper, pairwise_neg_snr  # noqa: E402
from oracle import dualpath_oracle as O  # noqa: E402

torch.manual_seed(0)
m = Sepformer(sample_rate=8000, **small).train()
for mod in m.modules():  # the reference hard-codes dropout 0.1 (sepformer.py:507); parity is only defined without it
    if isinstance(mod, torch.nn.Dropout):
        mod.p = 0.0
    if isinstance(mod, torch.nn.MultiheadAttention):
        mod.dropout = 0.0
sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
g = torch.Generator().manual_seed(99)
x = torch.randn(1, 2500, generator=g) * 0.1
tgt = torch.randn(1, 2, 2500, generator=g) * 0.1
loss = PITLossWrapper(pairwise_neg_snr, pit_from="pw_mtx", threshold_byloss=False)(m(x), tgt)
m.zero_grad()
loss.backward()
ref_grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
lo = O.pit_loss(SO.sepformer_forward(leaf, x, **small), tgt, "snr", False)
lo.backward()
worst = max(rel(leaf[k].grad, ref_grads[k]) for k in ref_grads)
assert abs(lo.item() - loss.item()) < 1e-5 and worst < 1e-3, (lo.item(), loss.item(), worst)
keep = ["encoder.conv1d.weight", "decoder.weight", "masknet.prelu.weight", "masknet.conv2d.bias", "masknet.norm.weight",
        "masknet.dual_mdl.0.intra_mdl.mdl.layers.0.self_att.att.in_proj_weight", "masknet.dual_mdl.1.inter_mdl.mdl.layers.1.pos_ffn.ffn.3.weight",
        "masknet.dual_mdl.0.inter_norm.gamma", "masknet.dual_mdl.1.intra_mdl.mdl.norm.bias", "masknet.output_gate.0.weight"]
gnpz = {"x": x.numpy(), "tgt": tgt.numpy(), "loss": np.array(loss.item())}
for k in keep:
    gnpz["grad::" + k] = ref_grads[k].numpy()
np.savez_compressed(os.path.join(HERE, "grads_sepformer_small.npz"), **gnpz)
manifest["grad_norms_sepformer_small"] = {k: float(v.double().norm()) for k, v in ref_grads.items()}
manifest["grad_oracle_worst_rel_sepformer"] = worst
json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)
print("sepformer grads ok, worst oracle-vs-reference rel", worst)
