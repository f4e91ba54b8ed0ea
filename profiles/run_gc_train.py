"""One fused training step of the GroupComm TasNet (G = 16, B = 16 x 4 s) inside a profiler range.
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python profiles/run_gc_train.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_only_speech_separation_b200.losses import PITLossWrapper, pairwise_neg_snr  # noqa: E402
from audio_only_speech_separation_b200.models import TasNet  # noqa: E402
from audio_only_speech_separation_b200.trainer import DualPathTrainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.manual_seed(0)
m = TasNet(module=os.environ.get("MODULE", "DPRNN"), enc_dim=64, bn_dim=64, group_size=16).cuda().train()
tr = DualPathTrainer(m, PITLossWrapper(pairwise_neg_snr, pit_from="pw_mtx", threshold_byloss=False))
g = torch.Generator().manual_seed(1)
src = (torch.randn(B, 2, 32000, generator=g) * 0.1).cuda()
mix = src.sum(1).contiguous()
for _ in range(2):
    tr.step(mix, src)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step(mix, src)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one GroupComm training step, batch", B)
