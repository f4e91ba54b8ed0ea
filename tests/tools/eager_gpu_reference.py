"""Calibration only: the oracle port of the reference (plain torch ops: cuDNN LSTM, cuBLAS, ATen) run eagerly on the B200.

This is what the reference's own code path would dispatch to on this GPU; TF32 is disabled so the
arithmetic matches the fp32 reference.  Prints one JSON line per case.  Test infrastructure, never imported by the product.
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import dualpath_oracle as O  # noqa: E402
from audio_only_speech_separation_b200.models import TasNet  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda")


def run(B, T, train, unfold=False, iters=5):
    torch.manual_seed(0)
    m = TasNet(sample_rate=8000, unfold=unfold)
    sd = {k: v.detach().clone().to(dev).requires_grad_(train) for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(1)
    src = (torch.randn(B, 2, T, generator=g) * 0.1).to(dev)
    mix = src.sum(1)

    def step():
        if train:
            for p in sd.values():
                p.grad = None
            loss = O.pit_loss(O.tasnet_forward(sd, mix, unfold=unfold, lstm_impl="aten"), src, "snr", False)
            loss.backward()
        else:
            with torch.no_grad():
                O.tasnet_forward(sd, mix, unfold=unfold, lstm_impl="aten")

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / iters
    print(json.dumps({"case": f"eager torch on GPU, DPRNN B={B} T={T} train={train} unfold={unfold}", "ms": dt * 1e3,
                      "samples_per_s": B / dt, "audio_s_per_s": B * T / 8000 / dt}), flush=True)


if __name__ == "__main__":
    run(1, 32000, False)
    run(16, 32000, False)
    run(16, 32000, True)
    run(32, 32000, False, unfold=True)
