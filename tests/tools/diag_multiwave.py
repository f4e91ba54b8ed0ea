import os, sys, torch
sys.path.insert(0, os.getcwd())
from audio_only_speech_separation_b200 import _lib
from audio_only_speech_separation_b200.models import TasNet
torch.manual_seed(0)
m = TasNet(sample_rate=8000).cuda().eval()
g = torch.Generator().manual_seed(99)
x = (torch.randn(48, 32000, generator=g) * 0.1).cuda()
def rel(a, b): return float((a - b).norm() / b.norm())
with torch.no_grad():
    singles = {i: m(x[i:i+1]) for i in (0, 7, 23, 39)}
    for mode in (1, 0, 3):
        _lib.check(_lib.lib().dp_set_lstm_pipeline(mode))
        for B in (16, 24, 32, 40, 48):
            y1 = m(x[:B]); y2 = m(x[:B])
            errs = {i: rel(y1[i:i+1], s) for i, s in singles.items() if i < B}
            print("mode", mode, "B", B, "rerun", rel(y1, y2), "vs single", {k: f"{v:.1e}" for k, v in errs.items()}, flush=True)
_lib.check(_lib.lib().dp_set_lstm_pipeline(1))
