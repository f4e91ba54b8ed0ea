"""Debug helper: per-parameter gradient errors of the SepFormer CUDA path vs the oracle's fp64 autograd (GPU box)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import dualpath_oracle as O
from oracle import sepformer_oracle as SO
from audio_only_speech_separation_b200.losses import PITLossWrapper, pairwise_neg_snr
from audio_only_speech_separation_b200.models import Sepformer

small = dict(encoder_out_nchannels=64, masknet_chunksize=50, masknet_numlayers=2, intra_numlayers=2, inter_numlayers=2, intra_nhead=4,
             inter_nhead=4, intra_dffn=128, inter_dffn=128)
torch.manual_seed(0)
m = Sepformer(sample_rate=8000, **small)
sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
m = m.cuda().train()
m.dropout = 0.0
g = torch.Generator().manual_seed(99)
x = torch.randn(1, 2500, generator=g) * 0.1
tgt = torch.randn(1, 2, 2500, generator=g) * 0.1
leaf = {k: v.clone().double().requires_grad_(not k.endswith("pos_enc.pe")) for k, v in sd.items()}
ref = O.pit_loss(SO.sepformer_forward(leaf, x.double(), **small), tgt.double(), "snr", False)
ref.backward()
# sensitivity of the reference algorithm itself: 1e-5 relative noise on its matmuls is not available here, so compare fp32 vs fp64
leaf32 = {k: v.clone().requires_grad_(not k.endswith("pos_enc.pe")) for k, v in sd.items()}
O.pit_loss(SO.sepformer_forward(leaf32, x, **small), tgt, "snr", False).backward()
loss = PITLossWrapper(pairwise_neg_snr, pit_from="pw_mtx", threshold_byloss=False)(m(x.cuda()), tgt.cuda())
loss.backward()
rows = []
for k, p in m.named_parameters():
    gr = leaf[k].grad
    rows.append((float((p.grad.cpu().double() - gr).norm() / gr.norm()), float((leaf32[k].grad.double() - gr).norm() / gr.norm()), k, float(gr.norm())))
rows.sort(reverse=True)
print("loss", loss.item(), ref.item())
for r in rows[:45]:
    print("%.3e (oracle fp32: %.1e)  %-75s |g|=%.3e" % r)
