"""Time the forward / backward recurrence kernels per pipelining mode at the bench shape and check that every mode
reproduces mode 0 bit for bit (same arithmetic, same order).  Usage: python tests/tools/time_recurrence.py [B]"""
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


def run(mode, layout, prec, bwd):
    _lib.check(L.dp_set_lstm_pipeline(mode))
    nseq, ln, qdiv, s_hi, s_lo, s_t = (B * S, K, 1 << 30, 0, K, 1) if layout == "intra" else (B * K, S, K, S * K, 1, K)
    G = torch.empty_like(G0)
    H = torch.empty(P, 256, device=dev)
    C = torch.empty(P, 256, device=dev)
    dbias = torch.zeros(1024, device=dev)
    tf, tb = [], []
    for it in range(4):
        G.copy_(G0)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        _lib.check(L.dp_lstm_recurrence_f32(_lib.ptr(pack.buf), _lib.ptr(G), _lib.ptr(H), _lib.ptr(C), nseq, ln, qdiv, s_hi, s_lo, s_t, 1, prec,
                                            _lib.stream_ptr()))
        e1.record()
        if bwd:
            dbias.zero_()
            _lib.check(L.dp_bilstm_backward_f32(_lib.ptr(pack.buf), _lib.ptr(G), _lib.ptr(C), _lib.ptr(dH), None, 0, _lib.ptr(dbias), P, nseq, ln,
                                                qdiv, s_hi, s_lo, s_t, prec, _lib.stream_ptr()))
        e2.record()
        torch.cuda.synchronize()
        if it >= 1:
            tf.append(e0.elapsed_time(e1))
            tb.append(e1.elapsed_time(e2))
    return sum(tf) / len(tf), sum(tb) / len(tb), H, G, C, dbias


for prec, pname in ((_lib.PREC_FP32, "fp32"), (_lib.PREC_BF16, "bf16")):
    for layout in ("intra", "inter"):
        ref = None
        for mode in (0, 2, 3):
            f, b, H, G, C, db = run(mode, layout, prec, True)
            if ref is None:
                ref = (H, G, C, db)
                same = None
            else:
                same = [bool(torch.equal(a, r)) for a, r in zip((H, G, C), ref[:3])] + [float((db - ref[3]).abs().max() / ref[3].abs().max())]
                same.append(float((G - ref[1]).norm() / ref[1].norm()))   # dG rel-L2 (the K-split BPTT kernel sums in another order)
            print(json.dumps({"prec": pname, "layout": layout, "B": B, "mode": mode, "fwd_us": round(f * 1e3, 1), "bwd_us": round(b * 1e3, 1),
                              "same_as_mode0(H,dG,C,dbias_relerr,dG_rel_l2)": same}), flush=True)
_lib.check(L.dp_set_lstm_pipeline(1))
