"""Time the GroupComm TasNet forward on the GPU (CUDA events; synthetic data, random-init weights) and, with ``cpu`` as the last
argument, the CPU oracle on the same input (test infrastructure: the oracle is only the timed baseline here).

    python tests/tools/time_groupcomm.py <B> <T> [group_size] [cpu]        (MODULE=DPTNet / UNFOLD=1 in the environment select the variant)
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from audio_only_speech_separation_b200.models import TasNet  # noqa: E402

B, T = int(sys.argv[1]), int(sys.argv[2])
G = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3] != "cpu" else 16
torch.manual_seed(0)
MODULE, UNFOLD = os.environ.get("MODULE", "DPRNN"), os.environ.get("UNFOLD", "0") == "1"
m = TasNet(module=MODULE, enc_dim=64, bn_dim=64, group_size=G, unfold=UNFOLD)
sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
m = m.cuda().eval()
x = torch.randn(B, T, generator=torch.Generator().manual_seed(1)) * 0.1
xd = x.cuda()
with torch.no_grad():
    for _ in range(3):
        y = m(xd)
    torch.cuda.synchronize()
    n = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        y = m(xd)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
out = {"module": MODULE, "unfold": UNFOLD, "B": B, "T": T, "group_size": G, "forward_ms": ms, "launches": m.last_launches, "audio_s_per_s_8k": B * T / 8000 / (ms * 1e-3)}
with torch.no_grad():   # the same forward without the CUDA graph (eager engine call)
    m.cuda_graph = False
    for _ in range(3):
        m(xd)
    e0.record()
    for _ in range(n):
        m(xd)
    e1.record()
    torch.cuda.synchronize()
    out["forward_ms_no_graph"] = e0.elapsed_time(e1) / n
    m.cuda_graph = True
if MODULE == "DPRNN" and os.environ.get("TRAIN", "1") == "1":   # fused training step (forward + PIT-SNR + backward + clip + Adam)
    from audio_only_speech_separation_b200.losses import PITLossWrapper, pairwise_neg_snr
    from audio_only_speech_separation_b200.trainer import DualPathTrainer

    m.train()
    tr = DualPathTrainer(m, PITLossWrapper(pairwise_neg_snr, pit_from="pw_mtx", threshold_byloss=False))
    tgt = (torch.randn(B, 2, T, generator=torch.Generator().manual_seed(2)) * 0.1).cuda()
    for _ in range(2):
        tr.step(xd, tgt)
    e0.record()
    for _ in range(5):
        tr.step(xd, tgt)
    e1.record()
    torch.cuda.synchronize()
    out["train_step_ms"] = e0.elapsed_time(e1) / 5
    out["train_launches"] = tr.launches_per_step
    m.eval()
if sys.argv[-1] == "cpu":
    from oracle import groupcomm_oracle as GO

    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        GO.tasnet_gc_forward(sd, x[:1], group_size=G, module=MODULE, unfold=UNFOLD)
        t0 = time.perf_counter()
        yo = GO.tasnet_gc_forward(sd, x, group_size=G, module=MODULE, unfold=UNFOLD)
        out["cpu_oracle_ms"] = (time.perf_counter() - t0) * 1e3
    out["cpu_threads"] = torch.get_num_threads()
    out["rel_l2_vs_oracle"] = ((y.cpu() - yo).norm() / yo.norm()).item()
print(json.dumps(out), flush=True)
