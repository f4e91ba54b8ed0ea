"""Attention forward / backward kernels at the engines' shapes: tcgen05 (planes in) vs warp-level tensor cores (fp32 in) vs CUDA cores."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from audio_only_speech_separation_b200 import _lib, ops
from audio_only_speech_separation_b200._lib import check, lib, ptr, stream_ptr

def t(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

for name, (B, S, K, E, H) in {"sepformer 16 s": (1, 130, 250, 256, 8), "dptnet B=16": (16, 82, 100, 64, 4), "sepformer 16 kHz": (1, 258, 250, 256, 8)}.items():
    qkv = torch.randn(B, S, K, 3 * E, device="cuda") * 0.5
    d_o = torch.randn(B, S, K, E, device="cuda")
    for layout in ("intra", "inter"):
        L = K if layout == "intra" else S
        for prec in ("bf16", "fp32"):
            r = {"shape": name, "layout": layout, "L": L, "prec": prec}
            if L <= 256:
                hi, lo = ops.split_rows(qkv.reshape(-1, 3 * E).contiguous())
                o32 = torch.empty(B, S, K, E, device="cuda"); oh = torch.empty(B * S * K, E, device="cuda", dtype=torch.bfloat16); ol = torch.empty_like(oh)
                lse0 = torch.empty(B * S * K, H, device="cuda")
                pr = _lib.PREC_FP32 if prec == "fp32" else _lib.PREC_BF16
                r["fwd_tcgen05_us"] = round(t(lambda: check(lib().dp_attention_forward_planes_f32(ptr(hi), ptr(lo), ptr(o32), ptr(oh), ptr(ol), ptr(lse0), E, H,
                                                                                                  int(layout == "inter"), B, S, K, pr, stream_ptr()))), 1)
            r["fwd_mma_us"] = round(t(lambda: ops.attention_tensor_cores(qkv, H, layout, precision=prec, save=True)), 1)
            o, lse = ops.attention_tensor_cores(qkv, H, layout, precision=prec, save=True)
            r["bwd_mma_us"] = round(t(lambda: ops.attention_backward(qkv, o, lse, d_o, H, layout, tensor_cores=True, precision=prec)), 1)
            if prec == "fp32":
                r["fwd_cuda_core_us"] = round(t(lambda: ops.attention(qkv, H, layout, save=True), 3), 1)
            print(json.dumps(r), flush=True)
