"""Time the weight-gradient GEMM (gemm_tma_tn_kernel) at the bench shape -- dW_ih | dW_hh of one LSTM direction: A = dG[:, dir] [P, 512],
B = [X [P, 64] | h_prev [P, 128]] -- with and without the cluster multicast of the B tiles.  Usage: python tests/tools/time_wgrad.py [B]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from audio_only_speech_separation_b200 import _lib, ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
P = B * 82 * 100
g = torch.Generator().manual_seed(0)
dG = torch.randn(P, 1024, generator=g).cuda() * 0.1
X = torch.randn(P, 64, generator=g).cuda()
Hp = torch.randn(P, 256, generator=g).cuda()
dGh, dGl = ops.split_rows(dG)
Xh, Xl = ops.split_rows(X)
Hh, Hl = ops.split_rows(Hp)
L = _lib.lib()
res = {}
for mc in (0, 1):
    L.dp_set_wgrad_multicast(mc)
    o0, o1 = torch.zeros(512, 64).cuda(), torch.zeros(512, 128).cuda()
    ts = []
    for it in range(6):
        o0.zero_(); o1.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.linear_wgrad_planes((dGh[:, :512], dGl[:, :512]), (Xh, Xl), o0, (Hh[:, :128], Hl[:, :128]), o1)
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(e0.elapsed_time(e1) * 1e3)
    ref0 = dG[:, :512].double().t() @ X.double()
    ref1 = dG[:, :512].double().t() @ Hp[:, :128].double()
    res["multicast" if mc else "plain"] = {"us": round(sum(ts) / len(ts), 1), "rel_l2": [float((o0.double() - ref0).norm() / ref0.norm()),
                                                                                           float((o1.double() - ref1).norm() / ref1.norm())]}
L.dp_set_wgrad_multicast(0)
print(json.dumps(res))
