"""Experiment: the fused DPRNN training step captured in a CUDA graph against eager launches (the captured step bakes the Adam step number:
timing only).  Usage: python tests/tools/time_train_graph.py [B]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from audio_only_speech_separation_b200.losses import PITLossWrapper, pairwise_neg_snr  # noqa: E402
from audio_only_speech_separation_b200.models import TasNet  # noqa: E402
from audio_only_speech_separation_b200.trainer import DualPathTrainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.manual_seed(0)
m = TasNet(sample_rate=8000).cuda().train()
tr = DualPathTrainer(m, PITLossWrapper(pairwise_neg_snr, pit_from="pw_mtx", threshold_byloss=False))
g = torch.Generator().manual_seed(1)
src = (torch.randn(B, 2, 32000, generator=g) * 0.1).cuda()
mix = src.sum(1).contiguous()
for _ in range(3):
    tr.step(mix, src)
torch.cuda.synchronize()


def timeit(fn, n=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


eager = timeit(lambda: tr.step(mix, src))
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    tr.step(mix, src)
torch.cuda.current_stream().wait_stream(side)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    loss = tr.step(mix, src)
torch.cuda.synchronize()
for _ in range(3):
    graph.replay()
torch.cuda.synchronize()
replay = timeit(graph.replay)
print(json.dumps({"B": B, "eager_ms": eager, "graph_ms": replay, "loss": float(loss)}))
