"""Time one training / inference forward of the model per backend variant (fused tcgen05 LSTM on/off)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from audio_only_speech_separation_b200._lib import check, lib
from audio_only_speech_separation_b200.models import TasNet

for B in (1, 16):
    for prec in ("fp32", "bf16"):
        for fused in (0, 1):
            check(lib().dp_set_fused_lstm(2 * fused))
            torch.manual_seed(0)
            m = TasNet(sample_rate=8000).cuda().eval()
            m.precision = prec
            x = torch.randn(B, 32000, device="cuda") * 0.1
            with torch.no_grad():
                for _ in range(3):
                    m(x)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    m(x)
                e1.record()
                torch.cuda.synchronize()
            print(json.dumps({"B": B, "prec": prec, "fused": fused, "fwd_ms": e0.elapsed_time(e1) / 10}), flush=True)
check(lib().dp_set_fused_lstm(1))
