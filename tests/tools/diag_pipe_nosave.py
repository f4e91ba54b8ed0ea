"""Forward recurrence, inference mode (save = 0): pipelined / 16-warp kernels against the plain one at one- and two-wave sizes, twice each."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from audio_only_speech_separation_b200 import _lib, ops
S, K = 82, 100
dev = torch.device("cuda"); torch.manual_seed(0)
lstm = torch.nn.LSTM(64, 128, 1, batch_first=True, bidirectional=True).cuda()
pack = ops.LstmPack(lstm)
L = _lib.lib()
for B in (16, 24, 40):
    P = B * S * K
    G0 = torch.randn(P, 1024, device=dev) * 0.5
    for layout in ("intra", "inter"):
        nseq, ln, qdiv, s_hi, s_lo, s_t = (B * S, K, 1 << 30, 0, K, 1) if layout == "intra" else (B * K, S, K, S * K, 1, K)
        outs = {}
        for mode in (0, 2, 2, 3):
            _lib.check(L.dp_set_lstm_pipeline(mode))
            G = G0.clone(); H = torch.full((P, 256), float("nan"), device=dev)
            _lib.check(L.dp_lstm_recurrence_f32(_lib.ptr(pack.buf), _lib.ptr(G), _lib.ptr(H), None, nseq, ln, qdiv, s_hi, s_lo, s_t, 0, 0, _lib.stream_ptr()))
            torch.cuda.synchronize()
            ref = outs.setdefault(0, H) if mode == 0 else outs[0]
            print(json.dumps({"B": B, "layout": layout, "mode": mode, "nan": int(torch.isnan(H).sum()), "maxdiff_vs_plain": float((H - ref).abs().max()),
                              "G_untouched": bool(torch.equal(G, G0))}), flush=True)
_lib.check(L.dp_set_lstm_pipeline(1))
