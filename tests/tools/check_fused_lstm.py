"""Debug helper: model forward / gradients with the fused tcgen05 LSTM kernel vs the unfused path (GPU box)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from audio_only_speech_separation_b200._lib import check, lib  # noqa: E402
from audio_only_speech_separation_b200.losses import PITLossWrapper, pairwise_neg_snr  # noqa: E402
from audio_only_speech_separation_b200.models import TasNet  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


for prec in ("fp32", "bf16"):
    for B, T, layer in ((1, 1600, 1), (2, 4000, 2), (3, 8000, 6)):
        out = {}
        for fused in (0, 1):
            check(lib().dp_set_fused_lstm(2 * fused))
            torch.manual_seed(0)
            m = TasNet(sample_rate=8000, layer=layer).cuda()
            m.precision = prec
            g = torch.Generator().manual_seed(1)
            x = (torch.randn(B, T, generator=g) * 0.1).cuda()
            tgt = (torch.randn(B, 2, T, generator=g) * 0.1).cuda()
            with torch.no_grad():
                m.eval()
                y_inf = m(x).clone()
            m.train()
            y = m(x)
            PITLossWrapper(pairwise_neg_snr, pit_from="pw_mtx", threshold_byloss=False)(y, tgt).backward()
            torch.cuda.synchronize()
            out[fused] = (y_inf, y.detach().clone(), torch.cat([p.grad.reshape(-1) for p in m.parameters()]))
        print(prec, B, T, layer, "infer", rel(out[1][0], out[0][0]), "train-fwd", rel(out[1][1], out[0][1]), "grads", rel(out[1][2], out[0][2]),
              flush=True)
check(lib().dp_set_fused_lstm(1))
