"""Golden vectors for the general PIT loss (n_src = 1 .. 4, pit_from = pw_mtx / pw_pt / perm_avg) from the REFERENCE itself:
imports /root/reference/look2hear/losses (build container only) and writes tests/golden/pit_general.npz.
Run: python tests/golden/make_golden_pit.py"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/look2hear/losses"


def _load(name):
    spec = importlib.util.spec_from_file_location(f"ref_losses.{name}", os.path.join(REF, f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


matrix = _load("matrix")
pitw = _load("pit_wrapper")

CASES = [  # name, n_src, B, T, pit_from, sdr_type, threshold_byloss, kind of estimates
    ("n3_pwmtx_sisdr", 3, 4, 1500, "pw_mtx", "sisdr", True, "mixed"),
    ("n3_pwmtx_snr_nothr", 3, 3, 1200, "pw_mtx", "snr", False, "mixed"),
    ("n3_pwmtx_sdsdr", 3, 3, 1200, "pw_mtx", "sdsdr", True, "mixed"),
    ("n3_pwpt_sisdr", 3, 3, 1500, "pw_pt", "sisdr", True, "mixed"),
    ("n2_pwpt_snr", 2, 4, 1500, "pw_pt", "snr", True, "mixed"),
    ("n3_permavg_sisdr", 3, 3, 1500, "perm_avg", "sisdr", True, "mixed"),
    ("n2_permavg_snr", 2, 4, 1000, "perm_avg", "snr", True, "mixed"),
    ("n4_pwmtx_sisdr", 4, 3, 900, "pw_mtx", "sisdr", False, "mixed"),
    ("n3_pwmtx_snr_threshold_hits", 3, 4, 1000, "pw_mtx", "snr", True, "accurate"),
    ("n1_pwmtx_sisdr", 1, 3, 800, "pw_mtx", "sisdr", True, "mixed"),
]
LOSS = {"pw_mtx": matrix.PairwiseNegSDR, "pw_pt": matrix.SingleSrcNegSDR, "perm_avg": matrix.MultiSrcNegSDR}

out = {}
for ci, (name, N, B, T, pit_from, sdr, thr, kind) in enumerate(CASES):
    g = torch.Generator().manual_seed(1000 + ci)
    tgt = torch.randn(B, N, T, generator=g) * 0.3 + 0.05 * torch.randn(B, N, 1, generator=g)
    perm = torch.stack([torch.randperm(N, generator=g) for _ in range(B)])
    noise = 0.4 if kind == "mixed" else 1e-3
    est = torch.stack([tgt[b, perm[b]] for b in range(B)]) + noise * torch.randn(B, N, T, generator=g)
    if kind == "accurate":   # some utterances below -30 dB (threshold_byloss drops them), one above
        est[0] = tgt[0, perm[0]] + 0.5 * torch.randn(N, T, generator=g)
    est = est.clone().requires_grad_(True)
    wrapper = pitw.PITLossWrapper(LOSS[pit_from](sdr), pit_from=pit_from, threshold_byloss=thr)
    loss, reordered = wrapper(est, tgt, return_ests=True)
    loss.backward()
    out[f"{name}::est"] = est.detach().numpy()
    out[f"{name}::tgt"] = tgt.numpy()
    out[f"{name}::loss"] = np.float32(loss.item())
    out[f"{name}::grad"] = est.grad.numpy()
    out[f"{name}::reordered"] = reordered.detach().numpy()
    with torch.no_grad():
        pw = matrix.PairwiseNegSDR(sdr)(est, tgt)
        out[f"{name}::pw"] = pw.numpy()
        out[f"{name}::multisrc"] = matrix.MultiSrcNegSDR(sdr)(est, tgt).numpy()
        out[f"{name}::singlesrc0"] = matrix.SingleSrcNegSDR(sdr)(est[:, 0], tgt[:, 0]).numpy()
    print(name, float(loss))
out["cases"] = np.array([f"{c[0]}|{c[1]}|{c[4]}|{c[5]}|{int(c[6])}" for c in CASES])
np.savez_compressed(os.path.join(HERE, "pit_general.npz"), **out)
print("wrote", os.path.join(HERE, "pit_general.npz"), os.path.getsize(os.path.join(HERE, "pit_general.npz")) // 1024, "KiB")
