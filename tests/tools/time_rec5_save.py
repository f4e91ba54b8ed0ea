"""Forward recurrence with and without the training-mode saves (activated gates, c_t), tcgen05 kernels (dp_set_lstm_tcgen05 2) against the
mma.sync kernels (0), fp32-parity and bf16 mode: how much of the step is the global-store path, and which kernel serves inference.
Usage: python tests/tools/time_rec5_save.py [B]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from audio_only_speech_separation_b200 import _lib, ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S, K = 82, 100
P = B * S * K
dev = torch.device("cuda")
torch.manual_seed(0)
lstm = torch.nn.LSTM(64, 128, 1, batch_first=True, bidirectional=True).cuda()
pack = ops.LstmPack(lstm)
G0 = torch.randn(P, 1024, device=dev) * 0.5
L = _lib.lib()
try:
    for prec, pname in ((_lib.PREC_FP32, "fp32"), (_lib.PREC_BF16, "bf16")):
        for mode in (2, 0):
            _lib.check(L.dp_set_lstm_tcgen05(mode))
            for save in (1, 0):
                G = torch.empty_like(G0)
                H = torch.empty(P, 256, device=dev)
                C = torch.empty(P, 256, device=dev)
                ts = []
                for it in range(5):
                    G.copy_(G0)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    _lib.check(L.dp_lstm_recurrence_f32(_lib.ptr(pack.buf), _lib.ptr(G), _lib.ptr(H), _lib.ptr(C), B * S, K, 1 << 30, 0, K, 1, save,
                                                        prec, _lib.stream_ptr()))
                    e1.record()
                    torch.cuda.synchronize()
                    if it >= 2:
                        ts.append(e0.elapsed_time(e1) * 1e3)
                print(json.dumps({"prec": pname, "kernels": "tcgen05" if mode else "mma.sync", "save": save, "fwd_us": round(sum(ts) / len(ts), 1)}),
                      flush=True)
finally:
    _lib.check(L.dp_set_lstm_tcgen05(1))
