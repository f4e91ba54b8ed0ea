"""Time one training step of the drop-in models on the GPU (CUDA events, synthetic data, random-init weights).

    python tests/tools/time_train.py <dprnn|dprnn_unfold|dptnet|dptnet_unfold|sepformer> <B> <T> <fp32|bf16> [sample_rate]

TasNet models: ``DualPathTrainer.step`` (forward + PIT loss + backward + clip + Adam fused).  Sepformer: the autograd path the
reference's Lightning module uses (``loss.backward()`` through the single engine node, ``clip_grad_norm_`` + ``torch.optim.Adam``),
with the reference's dropout 0.1 (``DROPOUT=0`` in the environment switches the sites off).
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from audio_only_speech_separation_b200.losses import PITLossWrapper, pairwise_neg_sisdr, pairwise_neg_snr  # noqa: E402
from audio_only_speech_separation_b200.models import Sepformer, TasNet  # noqa: E402
from audio_only_speech_separation_b200.trainer import DualPathTrainer  # noqa: E402

name, B, T, prec = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
sr = int(sys.argv[5]) if len(sys.argv) > 5 else 8000
torch.manual_seed(0)
g = torch.Generator().manual_seed(1234)
src = torch.randn(B, 2, T, generator=g) * 0.1
mix, tgt = src.sum(1).cuda(), src.cuda()
if name == "sepformer":
    m = Sepformer(sample_rate=sr).cuda().train()
    m.dropout = float(os.environ.get("DROPOUT", "0.1"))   # the reference's default; DROPOUT=0 switches the four sites off
    m.precision = prec
    loss_fn = PITLossWrapper(pairwise_neg_sisdr, pit_from="pw_mtx", threshold_byloss=True)
    opt = torch.optim.Adam(m.parameters(), lr=1.5e-4)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = loss_fn(m(mix), tgt)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 5.0)
        opt.step()
        if hasattr(m, "mark_params_dirty"):
            m.mark_params_dirty()
        return loss
else:
    m = TasNet(sample_rate=sr, module="DPTNet" if name.startswith("dptnet") else "DPRNN", unfold=name.endswith("unfold")).cuda().train()
    m.precision = prec
    tr = DualPathTrainer(m, PITLossWrapper(pairwise_neg_snr, pit_from="pw_mtx", threshold_byloss=False))

    def step():
        return tr.step(mix, tgt)

for _ in range(3):
    loss = step()
torch.cuda.synchronize()
n = 8
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(json.dumps({"model": name, "B": B, "T": T, "precision": prec, "train_ms": ms, "samples_per_s": B / (ms * 1e-3),
                  "audio_s_per_s": B * T / sr / (ms * 1e-3), "loss": float(loss)}), flush=True)
