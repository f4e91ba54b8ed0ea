"""Debug helper: per-parameter gradient errors of the CUDA path vs the oracle's autograd (run on the GPU box)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dualpath_oracle as O  # noqa: E402
from audio_only_speech_separation_b200.losses import PITLossWrapper, pairwise_neg_snr  # noqa: E402
from audio_only_speech_separation_b200.models import TasNet  # noqa: E402

module = sys.argv[1] if len(sys.argv) > 1 else "DPTNet"
layer = int(sys.argv[2]) if len(sys.argv) > 2 else 6
torch.manual_seed(0)
m = TasNet(sample_rate=8000, module=module, layer=layer)
sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
m = m.cuda().train()
g = torch.Generator().manual_seed(99)
x = torch.randn(2, 4000, generator=g) * 0.1
tgt = torch.randn(2, 2, 4000, generator=g) * 0.1
leaf = {k: v.clone().double().requires_grad_(True) for k, v in sd.items()}
ref = O.pit_loss(O.tasnet_forward(leaf, x.double(), module=module, layer=layer, lstm_impl="loop"), tgt.double(), "snr", False)
ref.backward()
loss = PITLossWrapper(pairwise_neg_snr, pit_from="pw_mtx", threshold_byloss=False)(m(x.cuda()), tgt.cuda())
loss.backward()
rows = []
for k, p in m.named_parameters():
    gr = leaf[k].grad
    rows.append((float((p.grad.cpu().double() - gr).norm() / gr.norm()), k, float(gr.norm())))
rows.sort(reverse=True)
print("loss", loss.item(), ref.item())
for r in rows[:40]:
    print("%.3e  %-70s |g|=%.3e" % r)
