"""Generate the committed golden vectors from the REAL reference.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):  ``python -B tests/golden/make_golden.py``.

For every case the script (1) runs the unmodified reference imported from
``/root/reference`` (``look2hear.models.TasNet``, ``look2hear.losses``), (2)
asserts that ``oracle/dualpath_oracle.py`` reproduces it (bit-exact for the
index ops, rel-L2 <= 2e-6 for floating point), and (3) stores inputs / outputs
in ``tests/golden/*.npz`` plus ``manifest.json``.  Model weights are not stored:
they are the reference's default init under ``torch.manual_seed(seed)``; the
manifest pins them with per-key float64 checksums so the tests can prove the
drop-in model reproduces the same init on the GPU box.
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

from look2hear.losses import PITLossWrapper, pairwise_neg_sdsdr, pairwise_neg_sisdr, pairwise_neg_snr  # noqa: E402
from look2hear.models import TasNet  # noqa: E402
from look2hear.models.utils.gc3_basics import merge_feature, split_feature  # noqa: E402

from oracle import dualpath_oracle as O  # noqa: E402

torch.set_num_threads(8)
manifest = {"torch": torch.__version__, "cases": {}}


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


# ---------------------------------------------------------------- seg / ola
seg = {}
for i, (L, K) in enumerate(
    [(4002, 100), (3999, 250), (1999, 100), (100, 100), (50, 100), (4000, 100), (4050, 100), (7, 24), (1, 100), (401, 24)]
):
    g = torch.Generator().manual_seed(100 + i)
    B, N = (2, 3) if L < 3000 else (1, 2)
    x = torch.randn(B, N, L, generator=g)
    blk, rest = split_feature(x, K)
    y = torch.randn(blk.shape, generator=g)
    mrg = merge_feature(y, rest)
    ob, orest = O.split_feature(x, K)
    assert orest == rest and torch.equal(ob, blk)
    assert torch.equal(O.merge_feature(y, rest), mrg)
    assert torch.equal(merge_feature(blk, rest), 2 * x)
    seg[f"x{i}"], seg[f"blk{i}"], seg[f"y{i}"], seg[f"mrg{i}"] = x.numpy(), blk.numpy(), y.numpy(), mrg.numpy()
    seg[f"meta{i}"] = np.array([L, K, rest, blk.shape[3]])
seg["n"] = np.array(10)
np.savez_compressed(os.path.join(HERE, "seg_ola.npz"), **seg)
print("seg_ola ok")

# ---------------------------------------------------------------- models
MODEL_CASES = [
    # name, config, B, T, input kind
    ("dprnn_wsj0_b2_t8001", "dprnn_wsj0", 2, 8001, "2d"),
    ("dprnn_wsj0_b1_t32000", "dprnn_wsj0", 1, 32000, "2d"),
    ("dprnn_wsj0_1d_t4000", "dprnn_wsj0", 1, 4000, "1d"),
    ("dprnn_wsj0_3d_t1234", "dprnn_wsj0", 2, 1234, "3d"),
    ("dprnn_unfold_b2_t8000", "dprnn_lrs2_unfolded", 2, 8000, "2d"),
    ("dptnet_wsj0_b1_t8000", "dptnet_wsj0", 1, 8000, "2d"),
]


def checksum(sd):
    return {k: [float(v.double().sum()), float(v.double().abs().sum())] + list(v.shape) for k, v in sd.items()}


models = {}
for name, cfgname, B, T, kind in MODEL_CASES:
    cfg = yaml.safe_load(open(f"{REF}/configs/{cfgname}.yml"))
    ac = cfg["audionet"]["audionet_config"]
    torch.manual_seed(0)
    m = TasNet(sample_rate=cfg["datamodule"]["data_config"]["sample_rate"], **ac).eval()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(B, T, generator=g) * 0.1
    xin = {"2d": x, "1d": x[0], "3d": x.unsqueeze(1)}[kind]
    with torch.no_grad():
        y = m(xin)
        yo = O.tasnet_forward(sd, xin, module=ac["module"], unfold=ac["unfold"], lstm_impl="loop")
    r = rel(yo, y)
    assert r < 2e-6, (name, r)
    np.savez_compressed(os.path.join(HERE, f"model_{name}.npz"), x=xin.numpy(), y=y.numpy())
    manifest["cases"][name] = {
        "config": cfgname,
        "audionet_config": ac,
        "sample_rate": cfg["datamodule"]["data_config"]["sample_rate"],
        "seed": 0,
        "oracle_rel_l2": r,
        "y_abs_sum": float(y.double().abs().sum()),
    }
    if cfgname not in models:
        models[cfgname] = (m, sd, ac)
        manifest.setdefault("state_dicts", {})[cfgname] = checksum(sd)
        manifest.setdefault("n_params", {})[cfgname] = sum(p.numel() for p in m.parameters())
    print(name, "oracle rel-L2", r)

# ---------------------------------------------------------------- losses
g = torch.Generator().manual_seed(7)
e = torch.randn(6, 2, 3000, generator=g)
t = torch.randn(6, 2, 3000, generator=g)
e[1] = t[1].flip(0) + 0.01 * torch.randn(2, 3000, generator=g)  # swapped, ~40 dB
e[2] = t[2] + 1e-3 * torch.randn(2, 3000, generator=g)  # ~60 dB (threshold_byloss drops it)
e[3] = 0.5 * t[3] + 3.0  # scaled + DC offset
t[4, 1] = 0.0  # silent target
loss_npz = {"ests": e.numpy(), "targets": t.numpy()}
for sname, fn in [("snr", pairwise_neg_snr), ("sisdr", pairwise_neg_sisdr), ("sdsdr", pairwise_neg_sdsdr)]:
    pw = fn(e, t)
    assert torch.allclose(pw, O.pairwise_neg_sdr(e, t, sname), rtol=1e-6, atol=1e-6)
    loss_npz[f"pw_{sname}"] = pw.numpy()
    for thr in (True, False):
        lr, rr = PITLossWrapper(fn, pit_from="pw_mtx", threshold_byloss=thr)(e, t, return_ests=True)
        lo, ro, perm = O.pit_loss(e, t, sname, thr, True)
        assert torch.allclose(lr, lo, rtol=1e-6, atol=1e-6) and torch.equal(rr, ro)
        loss_npz[f"loss_{sname}_{int(thr)}"] = lr.numpy()
        loss_npz[f"perm_{sname}"] = perm.numpy()
np.savez_compressed(os.path.join(HERE, "loss.npz"), **loss_npz)
print("loss ok")

# ---------------------------------------------------------------- gradients (training-step semantics)
m, sd, ac = models["dprnn_wsj0"]
m.train()
g = torch.Generator().manual_seed(99)
x = torch.randn(2, 4000, generator=g) * 0.1
tgt = torch.randn(2, 2, 4000, generator=g) * 0.1
loss = PITLossWrapper(pairwise_neg_snr, pit_from="pw_mtx", threshold_byloss=False)(m(x), tgt)
m.zero_grad()
loss.backward()
ref_grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
# oracle autograd on the same weights
leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
lo = O.pit_loss(O.tasnet_forward(leaf, x, lstm_impl="loop"), tgt, "snr", False)
lo.backward()
worst = max(rel(leaf[k].grad, ref_grads[k]) for k in ref_grads)
assert abs(lo.item() - loss.item()) < 1e-5 and worst < 1e-3, (lo.item(), loss.item(), worst)
keep = ["encoder.weight", "decoder.weight", "mask.0.bias", "bottleneck.0.weight", "seq_model.seq_model.row_rnn.0.rnn.bias_ih_l0",
        "seq_model.seq_model.col_rnn.5.proj.weight", "seq_model.seq_model.row_norm.3.weight", "seq_model.seq_model.output.bias"]
gnpz = {"x": x.numpy(), "tgt": tgt.numpy(), "loss": np.array(loss.item())}
for k in keep:
    gnpz["grad::" + k] = ref_grads[k].numpy()
np.savez_compressed(os.path.join(HERE, "grads_dprnn_wsj0.npz"), **gnpz)
manifest["grad_norms_dprnn_wsj0"] = {k: float(v.double().norm()) for k, v in ref_grads.items()}
manifest["grad_oracle_worst_rel"] = worst
print("grads ok, worst oracle-vs-reference rel", worst)

json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)
print("wrote manifest")
