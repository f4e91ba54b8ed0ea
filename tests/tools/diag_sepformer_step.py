import os, sys, time, torch
sys.path.insert(0, os.getcwd())
from audio_only_speech_separation_b200.losses import PITLossWrapper, pairwise_neg_sisdr
from audio_only_speech_separation_b200.models import Sepformer
torch.manual_seed(0)
g = torch.Generator().manual_seed(1234)
src = torch.randn(1, 2, 128000, generator=g) * 0.1
mix, tgt = src.sum(1).cuda(), src.cuda()
m = Sepformer(sample_rate=8000).cuda().train(); m.dropout = 0.0; m.precision = "bf16"
loss_fn = PITLossWrapper(pairwise_neg_sisdr, pit_from="pw_mtx", threshold_byloss=True)
opt = torch.optim.Adam(m.parameters(), lr=1.5e-4)
def phase(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); r = fn(); e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return r, (t1 - t0) * 1e3, e0.elapsed_time(e1)
for it in range(4):
    opt.zero_grad(set_to_none=True)
    est, cf, gf = phase(lambda: m(mix))
    loss, cl, gl = phase(lambda: loss_fn(est, tgt))
    _, cb, gb = phase(lambda: loss.backward())
    _, cc, gc = phase(lambda: torch.nn.utils.clip_grad_norm_(m.parameters(), 5.0))
    _, co, go = phase(lambda: opt.step())
    m.mark_params_dirty()
    if it >= 2:
        print(f"fwd cpu {cf:.1f} gpu {gf:.1f} | loss cpu {cl:.1f} gpu {gl:.1f} | bwd cpu {cb:.1f} gpu {gb:.1f} | clip cpu {cc:.1f} gpu {gc:.1f} | adam cpu {co:.1f} gpu {go:.1f}", flush=True)
