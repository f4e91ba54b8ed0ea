"""Evaluation-loop throughput: N synthetic test utterances through model + metrics, one by one (the reference's
audio_test.py loop shape) and batched by equal length.   python tests/tools/time_eval.py [N] [seconds] [batch]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from audio_only_speech_separation_b200.metrics import MetricsTracker, evaluate  # noqa: E402
from audio_only_speech_separation_b200.models import TasNet  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sec = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 16
T = int(8000 * sec)
torch.manual_seed(0)
model = TasNet(sample_rate=8000).cuda().eval()
g = torch.Generator().manual_seed(1)
data = []
for i in range(N):
    src = (torch.randn(2, T, generator=g) * 0.1).pin_memory()
    data.append((src.sum(0).pin_memory(), src, f"utt{i}"))
for bs in (1, batch):
    evaluate(model, data[: 2 * bs], MetricsTracker(), batch_size=bs)   # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m = evaluate(model, data, MetricsTracker(), batch_size=bs)
    res = m.final()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"utterances": N, "seconds_each": sec, "batch": bs, "wall_s": dt, "audio_s_per_s": N * sec / dt, "si-snr_i": res["si-snr_i"]}), flush=True)
