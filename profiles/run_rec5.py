"""One forward (training mode) and one BPTT launch of the tcgen05 recurrence kernels (csrc/lstm_rec5.cu) at the bench shape
(B = 16: 1 312 intra-chunk sequences of 100 steps per direction), fp32-parity mode -- the command profiled by ncu.
Usage: python profiles/run_rec5.py [B] [tc5 mode: 2 = tcgen05 kernels, 0 = mma.sync kernels]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_only_speech_separation_b200 import _lib, ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 2
S, K = 82, 100
P = B * S * K
dev = torch.device("cuda")
torch.manual_seed(0)
lstm = torch.nn.LSTM(64, 128, 1, batch_first=True, bidirectional=True).cuda()
pack = ops.LstmPack(lstm)
G0 = torch.randn(P, 1024, device=dev) * 0.5
dH = torch.randn(P, 256, device=dev) * 0.1
L = _lib.lib()
_lib.check(L.dp_set_lstm_tcgen05(mode))
nseq, ln, qdiv, s_hi, s_lo, s_t = B * S, K, 1 << 30, 0, K, 1
H = torch.empty(P, 256, device=dev)
C = torch.empty(P, 256, device=dev)
dbias = torch.zeros(1024, device=dev)
for it in range(2):
    G = G0.clone()
    _lib.check(L.dp_lstm_recurrence_f32(_lib.ptr(pack.buf), _lib.ptr(G), _lib.ptr(H), _lib.ptr(C), nseq, ln, qdiv, s_hi, s_lo, s_t, 1, _lib.PREC_FP32,
                                        _lib.stream_ptr()))
    _lib.check(L.dp_bilstm_backward_f32(_lib.ptr(pack.buf), _lib.ptr(G), _lib.ptr(C), _lib.ptr(dH), None, 0, _lib.ptr(dbias), P, nseq, ln, qdiv, s_hi,
                                        s_lo, s_t, _lib.PREC_FP32, _lib.stream_ptr()))
torch.cuda.synchronize()
print("ok", float(H.abs().mean()), float(G.abs().mean()))
