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
