"""PITLossWrapper beyond the reference configs' setting: n_src = 1 .. 4 and pit_from = pw_mtx / pw_pt / perm_avg (csrc/loss_n.cu) against
golden outputs of the reference's own PITLossWrapper / PairwiseNegSDR / SingleSrcNegSDR / MultiSrcNegSDR
(look2hear/losses/pit_wrapper.py:30-131, matrix.py:13-152; tests/golden/make_golden_pit.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pit_general.npz"))
CASES = [str(c).split("|") for c in Z["cases"]]


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm())


@pytest.mark.parametrize("name,n_src,pit_from,sdr,thr", CASES, ids=[c[0] for c in CASES])
def test_pit_general_matches_reference(name, n_src, pit_from, sdr, thr):
    from audio_only_speech_separation_b200 import losses as L

    cls = {"pw_mtx": L.PairwiseNegSDR, "pw_pt": L.SingleSrcNegSDR, "perm_avg": L.MultiSrcNegSDR}[pit_from]
    est = torch.from_numpy(Z[f"{name}::est"]).cuda().requires_grad_(True)
    tgt = torch.from_numpy(Z[f"{name}::tgt"]).cuda()
    assert est.shape[1] == int(n_src)
    wrapper = L.PITLossWrapper(cls(sdr), pit_from=pit_from, threshold_byloss=bool(int(thr)))
    loss, reordered = wrapper(est, tgt, return_ests=True)
    loss.backward()
    ref_loss = float(Z[f"{name}::loss"])
    assert abs(loss.item() - ref_loss) < 1e-4 * max(1.0, abs(ref_loss))
    assert torch.equal(reordered.cpu(), torch.from_numpy(Z[f"{name}::reordered"]))   # a pure gather: bit-exact
    assert rel_l2(est.grad, Z[f"{name}::grad"]) < 2e-4
    # the loss classes on their own (no autograd graph, like the n_src = 2 PairwiseNegSDR)
    with torch.no_grad():
        pw = L.PairwiseNegSDR(sdr)(est.detach(), tgt)
        assert float((pw.cpu() - torch.from_numpy(Z[f"{name}::pw"])).abs().max()) < 2e-4
        assert float((L.MultiSrcNegSDR(sdr)(est.detach(), tgt).cpu() - torch.from_numpy(Z[f"{name}::multisrc"])).abs().max()) < 2e-4
        s0 = L.SingleSrcNegSDR(sdr)(est.detach()[:, 0].contiguous(), tgt[:, 0].contiguous())
        assert float((s0.cpu() - torch.from_numpy(Z[f"{name}::singlesrc0"])).abs().max()) < 2e-4


def test_pit_n2_general_path_equals_specialised_kernels():
    """n_src = 2 through the general kernels (pw_pt) and through the specialised ones (pw_mtx): same loss, gradient and order."""
    from audio_only_speech_separation_b200 import losses as L

    g = torch.Generator().manual_seed(5)
    tgt = (torch.randn(6, 2, 4000, generator=g) * 0.2).cuda()
    est0 = (tgt.flip(1) + 0.05 * torch.randn(6, 2, 4000, generator=g).cuda())
    outs = []
    for cls, pf in ((L.PairwiseNegSDR, "pw_mtx"), (L.SingleSrcNegSDR, "pw_pt")):
        est = est0.clone().requires_grad_(True)
        loss, re = L.PITLossWrapper(cls("sisdr"), pit_from=pf, threshold_byloss=True)(est, tgt, return_ests=True)
        loss.backward()
        outs.append((loss.item(), est.grad.clone(), re))
    assert abs(outs[0][0] - outs[1][0]) < 1e-5 * max(1.0, abs(outs[0][0]))
    assert rel_l2(outs[1][1], outs[0][1].cpu()) < 1e-5 and torch.equal(outs[0][2], outs[1][2])


def test_pit_wrapper_rejects_what_is_not_built():
    from audio_only_speech_separation_b200 import losses as L

    x = torch.randn(2, 2, 100).cuda()
    with pytest.raises(NotImplementedError):
        L.PITLossWrapper(L.pairwise_neg_snr, pit_from="pw_pt")(x, x)          # loss class does not match pit_from
    with pytest.raises(NotImplementedError):
        L.PITLossWrapper(L.pairwise_neg_snr, perm_reduce=lambda a: a.sum(-1))(x, x)
    with pytest.raises(NotImplementedError):
        L.pairwise_neg_snr(torch.randn(2, 5, 100).cuda(), torch.randn(2, 5, 100).cuda())   # n_src > 4
