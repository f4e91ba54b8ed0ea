"""Forward recurrence of small inference passes: four-CTA-cluster kernel (dp_set_lstm_cluster 2) against the 16-warp kernel (0), per pass.
Usage: python tests/tools/time_cluster_rec.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from audio_only_speech_separation_b200 import _lib, ops  # noqa: E402

S, K = 82, 100
dev = torch.device("cuda")
torch.manual_seed(0)
lstm = torch.nn.LSTM(64, 128, 1, batch_first=True, bidirectional=True).cuda()
pack = ops.LstmPack(lstm)
L = _lib.lib()
try:
    for B in (1, 2, 3, 4):
        P = B * S * K
        G = torch.randn(P, 1024, device=dev) * 0.5
        H = torch.empty(P, 256, device=dev)
        for layout in ("intra", "inter"):
            nseq, ln, qdiv, s_hi, s_lo, s_t = (B * S, K, 1 << 30, 0, K, 1) if layout == "intra" else (B * K, S, K, S * K, 1, K)
            for prec, pname in ((_lib.PREC_FP32, "fp32"), (_lib.PREC_BF16, "bf16")):
                res = {}
                for mode in (0, 2):
                    _lib.check(L.dp_set_lstm_cluster(mode))
                    _lib.check(L.dp_set_lstm_tcgen05(0))
                    ts = []
                    for it in range(6):
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        _lib.check(L.dp_lstm_recurrence_f32(_lib.ptr(pack.buf), _lib.ptr(G), _lib.ptr(H), None, nseq, ln, qdiv, s_hi, s_lo, s_t, 0, prec,
                                                            _lib.stream_ptr()))
                        e1.record()
                        torch.cuda.synchronize()
                        if it >= 2:
                            ts.append(e0.elapsed_time(e1) * 1e3)
                    res["cluster_us" if mode else "warp16_us"] = round(sum(ts) / len(ts), 1)
                print(json.dumps({"B": B, "layout": layout, "nseq": nseq, "prec": pname, **res}), flush=True)
finally:
    _lib.check(L.dp_set_lstm_cluster(1))
    _lib.check(L.dp_set_lstm_tcgen05(1))
