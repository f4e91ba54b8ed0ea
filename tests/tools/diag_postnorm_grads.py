import os, sys, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import dropout_ref as DR
from audio_only_speech_separation_b200 import _lib
from audio_only_speech_separation_b200.losses import PITLossWrapper, pairwise_neg_snr
from audio_only_speech_separation_b200.models import Sepformer
from oracle import dualpath_oracle as O, sepformer_oracle as SO
def rel(a, b): return float((a.cpu().double() - b.double()).norm() / (b.double().norm() + 1e-30))
for drop in (0.0, 0.1):
    cfg = dict(encoder_out_nchannels=128, intra_dffn=256, inter_dffn=256, intra_nhead=4, inter_nhead=4, intra_numlayers=2, inter_numlayers=1,
               masknet_chunksize=50, masknet_numlayers=1, intra_norm_before=False, inter_norm_before=False)
    torch.manual_seed(5)
    m = Sepformer(sample_rate=8000, **cfg); sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.cuda().train(); m.dropout = drop
    g = torch.Generator().manual_seed(21)
    B, T, K = 2, 2400, 50
    x = torch.randn(B, T, generator=g) * 0.1; tgt = torch.randn(B, 2, T, generator=g) * 0.1
    torch.manual_seed(77); seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
    _, S = _lib.seg_geometry((T - 16) // 8 + 1, K)
    cb = DR.oracle_dropout(0.1, seed, B, S, K, 4, 4, 2, 1) if drop > 0 else None
    leaf = {k: v.clone().requires_grad_(not k.endswith("pos_enc.pe")) for k, v in sd.items()}
    ref_loss = O.pit_loss(SO.sepformer_forward(leaf, x, dropout=cb, **cfg), tgt, "snr", False); ref_loss.backward()
    torch.manual_seed(77)
    loss = PITLossWrapper(pairwise_neg_snr, pit_from="pw_mtx", threshold_byloss=False)(m(x.cuda()), tgt.cuda()); loss.backward()
    print("drop", drop, "loss", loss.item(), ref_loss.item())
    errs = sorted(((rel(p.grad, leaf[k].grad), k) for k, p in m.named_parameters()), reverse=True)
    for e, k in errs[:14]: print(f"  {e:.2e} {k}")
    print("  median", errs[len(errs)//2][0])
