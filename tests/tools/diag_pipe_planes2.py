import json, os, sys, torch
sys.path.insert(0, os.getcwd())
from audio_only_speech_separation_b200 import _lib, ops
S, K = 82, 100
dev = torch.device("cuda"); torch.manual_seed(0)
lstm = torch.nn.LSTM(64, 128, 1, batch_first=True, bidirectional=True).cuda()
pack = ops.LstmPack(lstm)
L = _lib.lib()
B = 24; P = B * S * K
G0 = torch.randn(P, 1024, device=dev) * 0.5
nseq, ln, qdiv, s_hi, s_lo, s_t = (B * S, K, 1 << 30, 0, K, 1)
_lib.check(L.dp_set_lstm_pipeline(0))
G = G0.clone(); Href = torch.empty(P, 256, device=dev)
_lib.check(L.dp_lstm_recurrence_f32(_lib.ptr(pack.buf), _lib.ptr(G), _lib.ptr(Href), None, nseq, ln, qdiv, s_hi, s_lo, s_t, 0, 0, _lib.stream_ptr()))
_lib.check(L.dp_set_lstm_pipeline(2))
for rep in range(3):
    G = G0.clone(); C = torch.empty(P, 256, device=dev)
    hh, hl, ph, plo = (torch.full((P, 256), float("nan"), device=dev, dtype=torch.bfloat16) for _ in range(4))
    _lib.check(L.dp_lstm_recurrence_planes_f32(_lib.ptr(pack.buf), _lib.ptr(G), None, _lib.ptr(C), _lib.ptr(hh), _lib.ptr(hl), _lib.ptr(ph), _lib.ptr(plo),
                                              nseq, ln, qdiv, s_hi, s_lo, s_t, 1, 0, _lib.stream_ptr()))
    torch.cuda.synchronize()
    h = hh.float() + hl.float()
    bad = ((h - Href).abs() > 1e-5)
    rows = bad.any(1).nonzero().flatten().tolist()
    for r in rows[:6]:
        cols = bad[r].nonzero().flatten()
        seq, t = divmod(r, K)
        half = "fwd" if cols.max() < 128 else ("bwd" if cols.min() >= 128 else "both")
        other = r + 1 if half == "bwd" else r - 1   # the step processed just before the last one
        stale = float((h[r, cols] - Href[other, cols]).abs().max()) if 0 <= other < P else None
        print(json.dumps({"rep": rep, "row": r, "seq": seq, "slot": seq % 24, "t": t, "half": half, "ncols": int(cols.numel()), "col_min": int(cols.min()), "col_max": int(cols.max()),
                          "diff_vs_prev_step_value": stale, "units_mod16": sorted(set((cols % 128 % 16).tolist()))[:16]}), flush=True)
    # cell state / gates consistency
    print("rep", rep, "bad rows", len(rows))
