"""Profiling driver: W warm-up training steps, then ONE step inside cudaProfilerStart/Stop.

    python profiles/run_step.py [--batch 16] [--precision fp32] [--mode train|fwd]
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... python profiles/run_step.py
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from audio_only_speech_separation_b200.losses import PITLossWrapper, pairwise_neg_snr
from audio_only_speech_separation_b200.models import TasNet
from audio_only_speech_separation_b200.trainer import DualPathTrainer

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--precision", default="fp32")
ap.add_argument("--mode", default="train")
ap.add_argument("--warm", type=int, default=2)
a = ap.parse_args()
torch.manual_seed(0)
model = TasNet(sample_rate=8000, **bench.CFG_DPRNN).cuda()
model.precision = a.precision
mix, tgt = bench.synthetic(a.batch, 1234)
mix, tgt = mix.cuda(), tgt.cuda()
if a.mode == "train":
    tr = DualPathTrainer(model, PITLossWrapper(pairwise_neg_snr, pit_from="pw_mtx", threshold_byloss=False))
    step = lambda: tr.step(mix, tgt)
else:
    model.eval()
    def step():
        with torch.no_grad():
            return model(mix)
for _ in range(a.warm):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one", a.mode, "step, batch", a.batch, a.precision)
