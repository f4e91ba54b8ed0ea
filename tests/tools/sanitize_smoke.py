"""Small forward + backward of every engine for compute-sanitizer (memcheck): tiny shapes, all kernel families.

    compute-sanitizer --tool memcheck python tests/tools/sanitize_smoke.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from audio_only_speech_separation_b200.losses import PITLossWrapper, pairwise_neg_sisdr, pairwise_neg_snr  # noqa: E402
from audio_only_speech_separation_b200.models import Sepformer, TasNet  # noqa: E402
from audio_only_speech_separation_b200.trainer import DualPathTrainer  # noqa: E402

torch.manual_seed(0)
g = torch.Generator().manual_seed(0)
src = torch.randn(2, 2, 2403, generator=g) * 0.1
mix, tgt = src.sum(1).cuda(), src.cuda()
for module, unfold in (("DPRNN", False), ("DPRNN", True), ("DPTNet", False), ("DPTNet", True)):
    for prec in ("fp32", "bf16"):
        m = TasNet(sample_rate=8000, layer=1, module=module, unfold=unfold).cuda().train()
        m.precision = prec
        tr = DualPathTrainer(m, PITLossWrapper(pairwise_neg_snr, pit_from="pw_mtx", threshold_byloss=False))
        loss = tr.step(mix, tgt)
        m.eval()
        with torch.no_grad():
            y = m(mix[:1])
        print(module, unfold, prec, float(loss), tuple(y.shape), flush=True)
cfg = dict(encoder_out_nchannels=128, intra_dffn=256, inter_dffn=256, intra_nhead=4, inter_nhead=4, intra_numlayers=1, inter_numlayers=1,
           masknet_chunksize=50, masknet_numlayers=1)
for prec in ("fp32", "bf16"):
    m = Sepformer(sample_rate=8000, **cfg).cuda().train()
    m.precision = prec
    loss = PITLossWrapper(pairwise_neg_sisdr, pit_from="pw_mtx", threshold_byloss=True)(m(mix), tgt)
    loss.backward()
    m.eval()
    with torch.no_grad():
        y = m(mix[:1])
    print("Sepformer", prec, float(loss), tuple(y.shape), flush=True)
torch.cuda.synchronize()
print("sanitize smoke done")
