"""Time the forward (inference) of the drop-in models on the GPU: audio-seconds per second, CUDA events.

    python tests/tools/time_forward.py <dprnn|dprnn_unfold|dptnet|sepformer> <B> <T> <fp32|bf16> [sample_rate]
"""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from audio_only_speech_separation_b200.models import Sepformer, TasNet  # noqa: E402

name, B, T, prec = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
sr = int(sys.argv[5]) if len(sys.argv) > 5 else 8000
torch.manual_seed(0)
if name == "sepformer":
    m = Sepformer(sample_rate=sr)
else:
    m = TasNet(sample_rate=sr, module="DPTNet" if name == "dptnet" else "DPRNN", unfold=name == "dprnn_unfold")
m = m.cuda().eval()
m.precision = prec
if "LSTM_CLUSTER" in os.environ:   # 0 off / 1 automatic / 2 always: the four-CTA-cluster recurrence of small inference passes
    from audio_only_speech_separation_b200 import _lib
    _lib.check(_lib.lib().dp_set_lstm_cluster(int(os.environ["LSTM_CLUSTER"])))
x = torch.randn(B, T, device="cuda") * 0.1
with torch.no_grad():
    for _ in range(3):
        m(x)
    torch.cuda.synchronize()
    n = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        m(x)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(json.dumps({"model": name, "B": B, "T": T, "precision": prec, "ms": ms, "audio_s_per_s": B * T / sr / (ms * 1e-3),
                  "launches": m.last_launches}), flush=True)
