import os, sys, torch
sys.path.insert(0, os.getcwd())
from audio_only_speech_separation_b200 import ops
torch.manual_seed(0)
for (M, N, K) in ((2000, 768, 256), (32500, 768, 256), (32500, 1024, 256), (131200, 192, 64)):
    a = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") * 0.1
    ah, al = ops.split_rows(a); wh, wl = ops.split_rows(w)
    for prec in ("bf16", "fp32"):
        out, planes = ops.linear_planes(ah, al if prec == "fp32" else None, wh, wl if prec == "fp32" else None, planes_out=True, precision=prec)
        torch.cuda.synchronize()
        ref = (ah.float() @ wh.float().t()) if prec == "bf16" else a @ w.t()
        got = planes[0].float() + (planes[1].float() if prec == "fp32" else 0)
        print(M, N, K, prec, float((got - ref).norm() / ref.norm()), float((out - ref).norm() / ref.norm()), flush=True)
