"""Time the tcgen05 recurrence kernels (csrc/lstm_rec5.cu) against the mma.sync kernels at the bench shape and report how far their
outputs are apart (different summation order: not bit-identical).  Usage: python tests/tools/time_rec5.py [B]"""
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
dH = torch.randn(P, 256, device=dev) * 0.1
L = _lib.lib()


def run(tc5, layout, prec):
    _lib.check(L.dp_set_lstm_tcgen05(tc5))
    nseq, ln, qdiv, s_hi, s_lo, s_t = (B * S, K, 1 << 30, 0, K, 1) if layout == "intra" else (B * K, S, K, S * K, 1, K)
    G = torch.empty_like(G0)
    H = torch.empty(P, 256, device=dev)
    C = torch.empty(P, 256, device=dev)
    Gs = None
    dbias = torch.zeros(1024, device=dev)
    tf, tb = [], []
    for it in range(4):
        G.copy_(G0)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        _lib.check(L.dp_lstm_recurrence_f32(_lib.ptr(pack.buf), _lib.ptr(G), _lib.ptr(H), _lib.ptr(C), nseq, ln, qdiv, s_hi, s_lo, s_t, 1, prec,
                                            _lib.stream_ptr()))
        e1.record()
        Gs = G.clone() if it == 3 else Gs
        e1b = torch.cuda.Event(enable_timing=True)
        e1b.record()
        dbias.zero_()
        _lib.check(L.dp_bilstm_backward_f32(_lib.ptr(pack.buf), _lib.ptr(G), _lib.ptr(C), _lib.ptr(dH), None, 0, _lib.ptr(dbias), P, nseq, ln,
                                            qdiv, s_hi, s_lo, s_t, prec, _lib.stream_ptr()))
        e2.record()
        torch.cuda.synchronize()
        if it >= 1:
            tf.append(e0.elapsed_time(e1))
            tb.append(e1b.elapsed_time(e2))
    return sum(tf) / len(tf), sum(tb) / len(tb), H.clone(), Gs, C.clone(), G.clone(), dbias.clone()


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


try:
    for prec, pname in ((_lib.PREC_FP32, "fp32"), (_lib.PREC_BF16, "bf16")):
        for layout in ("intra", "inter"):
            ref = run(0, layout, prec)
            new = run(2, layout, prec)
            print(json.dumps({"prec": pname, "layout": layout, "B": B, "mma_sync_fwd_us": round(ref[0] * 1e3, 1), "mma_sync_bwd_us": round(ref[1] * 1e3, 1),
                              "tcgen05_fwd_us": round(new[0] * 1e3, 1), "tcgen05_bwd_us": round(new[1] * 1e3, 1),
                              "rel_l2(H,gates,c,dG,dbias)": [rel(new[i], ref[i]) for i in (2, 3, 4, 5, 6)]}), flush=True)
finally:
    _lib.check(L.dp_set_lstm_tcgen05(1))
