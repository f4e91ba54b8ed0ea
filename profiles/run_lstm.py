"""Profiling driver: the persistent BiLSTM forward + BPTT kernels alone at the bench shape (B=16: 1312 x 100 intra)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from audio_only_speech_separation_b200 import ops

B, S, K = int(os.environ.get("B", 16)), 82, 100
layout = os.environ.get("LAYOUT", "intra")
prec = os.environ.get("PREC", "fp32")
torch.manual_seed(0)
from audio_only_speech_separation_b200 import _lib
_lib.check(_lib.lib().dp_set_lstm_pipeline(int(os.environ.get("PIPE", 1))))
lstm = torch.nn.LSTM(64, 128, 1, batch_first=True, bidirectional=True).cuda()
pack = ops.LstmPack(lstm)
x = torch.randn(B, S, K, 64, device="cuda")
dH = torch.randn(B, S, K, 256, device="cuda")
for it in range(3):
    if it == 2:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    H, G, Cst = ops.bilstm_forward(pack, x, layout, save=True, precision=prec)
    dx, db = ops.bilstm_backward(pack, G, Cst, dH, (B, S, K), layout, precision=prec)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(H.abs().mean()), float(dx.abs().mean()))
