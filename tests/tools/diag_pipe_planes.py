"""Plane outputs (h, h_prev) of the forward recurrence kernels against the fp32 H of the plain kernel, one- and two-wave sizes."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from audio_only_speech_separation_b200 import _lib, ops
S, K = 82, 100
dev = torch.device("cuda"); torch.manual_seed(0)
lstm = torch.nn.LSTM(64, 128, 1, batch_first=True, bidirectional=True).cuda()
pack = ops.LstmPack(lstm)
L = _lib.lib()
for B in (16, 24):
    P = B * S * K
    G0 = torch.randn(P, 1024, device=dev) * 0.5
    for layout in ("intra", "inter"):
        nseq, ln, qdiv, s_hi, s_lo, s_t = (B * S, K, 1 << 30, 0, K, 1) if layout == "intra" else (B * K, S, K, S * K, 1, K)
        _lib.check(L.dp_set_lstm_pipeline(0))
        G = G0.clone(); Href = torch.empty(P, 256, device=dev)
        _lib.check(L.dp_lstm_recurrence_f32(_lib.ptr(pack.buf), _lib.ptr(G), _lib.ptr(Href), None, nseq, ln, qdiv, s_hi, s_lo, s_t, 0, 0, _lib.stream_ptr()))
        for mode in (0, 2, 2, 3):
            for save in (0, 1):
                _lib.check(L.dp_set_lstm_pipeline(mode))
                G = G0.clone()
                C = torch.empty(P, 256, device=dev)
                hh, hl, ph, plo = (torch.full((P, 256), float("nan"), device=dev, dtype=torch.bfloat16) for _ in range(4))
                _lib.check(L.dp_lstm_recurrence_planes_f32(_lib.ptr(pack.buf), _lib.ptr(G), None, _lib.ptr(C) if save else None, _lib.ptr(hh), _lib.ptr(hl),
                                                          _lib.ptr(ph) if save else None, _lib.ptr(plo) if save else None, nseq, ln, qdiv, s_hi, s_lo, s_t,
                                                          save, 0, _lib.stream_ptr()))
                torch.cuda.synchronize()
                h = hh.float() + hl.float()
                bad = (h - Href).abs() > 1e-5
                rows = bad.any(1).nonzero().flatten()
                print(json.dumps({"B": B, "layout": layout, "mode": mode, "save": save, "nan": int(torch.isnan(h).sum()), "maxdiff": float((h - Href).abs().max()),
                                  "bad_rows": int(rows.numel()), "first_bad": rows[:6].tolist()}), flush=True)
_lib.check(L.dp_set_lstm_pipeline(1))
