"""Micro-benchmark: TMA-fed tcgen05 GEMM on planes vs the mma.sync GEMM on fp32 operands (CUDA events)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from audio_only_speech_separation_b200 import ops  # noqa: E402


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for M, N, K in [(131200, 1024, 64), (131200, 64, 256), (131200, 64, 1024), (131200, 256, 64), (32500, 768, 256), (32500, 1024, 256),
                (32500, 256, 1024), (32500, 256, 256)]:
    a = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    out = torch.empty(M, N, device="cuda")
    ah, al = ops.split_rows(a)
    wh, wl = ops.split_rows(w)
    whs, wls = ops.split_bf16(w)
    for prec in ("fp32", "bf16"):
        t_new = timeit(lambda: ops.linear_planes(ah, al, wh, wl, out=out, precision=prec))
        t_old = timeit(lambda: ops.linear(a, w, out=out, precision=prec))
        flops = 2.0 * M * N * K
        byt = M * K * 4 + M * N * 4
        print(json.dumps({"M": M, "N": N, "K": K, "prec": prec, "tma_us": t_new, "mma_sync_us": t_old, "tma_tflops": flops / t_new / 1e6,
                          "tma_gbs": byt / t_new / 1e3}), flush=True)
