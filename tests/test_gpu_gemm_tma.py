"""GPU parity of the TMA-fed tcgen05 GEMMs on pre-split bf16 planes against fp64 torch."""
import pytest
import torch

from conftest import record, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (1000, 64, 256), (4100, 1024, 64), (777, 256, 1024), (3000, 768, 256), (260, 192, 64),
                                   (2, 128, 128)])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_linear_planes_matches_torch(M, N, K, precision):
    from audio_only_speech_separation_b200 import ops

    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    ref = a.double() @ w.double().t() + b.double()
    ah, al = ops.split_rows(a.cuda())
    wh, wl = ops.split_rows(w.cuda())
    c, planes = ops.linear_planes(ah, al, wh, wl, b.cuda(), planes_out=True, precision=precision)
    err = rel_l2(c, ref)
    record("linear_planes", M=M, N=N, K=K, precision=precision, rel_l2=err)
    assert err < (2e-5 if precision == "fp32" else 1e-2)
    ch, cl = planes
    assert rel_l2(ch.float() + cl.float(), c) < 1e-5  # the plane output is the split of the fp32 output
    # epilogue variants: ReLU, accumulate
    c2, _ = ops.linear_planes(ah, al, wh, wl, b.cuda(), act=1, precision=precision)
    assert torch.equal(c2, torch.relu(c))
    base = torch.randn(M, N, generator=g).cuda()
    c3, _ = ops.linear_planes(ah, al, wh, wl, None, out=base.clone(), accumulate=True, precision=precision)
    assert rel_l2(c3, base.double().cpu() + (ref - b.double())) < (2e-5 if precision == "fp32" else 1e-2)


def test_linear_planes_agrees_with_mma_sync_backend():
    from audio_only_speech_separation_b200 import ops

    g = torch.Generator().manual_seed(0)
    a = torch.randn(5000, 256, generator=g).cuda()
    w = (torch.randn(64, 256, generator=g) / 16).cuda()
    ref = ops.linear(a, w)
    ah, al = ops.split_rows(a)
    wh, wl = ops.split_rows(w)
    c, _ = ops.linear_planes(ah, al, wh, wl)
    assert rel_l2(c, ref) < 1e-5


@pytest.fixture(params=[0, 1], ids=["unicast", "multicast"])
def wgrad_multicast(request):
    """Both variants of the weight-gradient GEMM: plain, and 4-CTA clusters multicasting the shared operand tiles (Mo % 512 == 0 only)."""
    from audio_only_speech_separation_b200 import _lib

    _lib.lib().dp_set_wgrad_multicast(request.param)
    yield request.param
    _lib.lib().dp_set_wgrad_multicast(0)


@pytest.mark.parametrize("P,Mo,nb0,nb1", [(1000, 128, 64, 0), (5000, 512, 64, 128), (131, 256, 64, 0), (20000, 1024, 64, 0), (9000, 512, 128, 64)])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_wgrad_planes_matches_torch(wgrad_multicast, P, Mo, nb0, nb1, precision):
    from audio_only_speech_separation_b200 import ops

    g = torch.Generator().manual_seed(P + Mo)
    wide = torch.randn(P, Mo + 64, generator=g)       # A is a column slice of a wider tensor (like one direction of dG)
    b0 = torch.randn(P, nb0, generator=g)
    b1 = torch.randn(P, nb1 + 32, generator=g) if nb1 else None
    a = wide[:, 64:]
    ref0 = a.double().t() @ b0.double()
    wh, wl = ops.split_rows(wide.cuda())
    b0p = ops.split_rows(b0.cuda())
    out0 = torch.ones(Mo, nb0).cuda()
    if nb1:
        b1h, b1l = ops.split_rows(b1.cuda())
        out1 = torch.zeros(nb1, Mo).cuda()        # transposed store
        ops.linear_wgrad_planes((wh[:, 64:], wl[:, 64:]), b0p, out0, (b1h[:, 32:], b1l[:, 32:]), out1, tr1=True, precision=precision)
        ref1 = a.double().t() @ b1[:, 32:].double()
        e1 = rel_l2(out1.t(), ref1)
    else:
        ops.linear_wgrad_planes((wh[:, 64:], wl[:, 64:]), b0p, out0, precision=precision)
        e1 = 0.0
    e0 = rel_l2(out0 - 1.0, ref0)
    record("wgrad_planes", P=P, Mo=Mo, nb0=nb0, nb1=nb1, precision=precision, e0=e0, e1=e1)
    tol = 3e-5 if precision == "fp32" else 1e-2
    assert e0 < tol and e1 < tol
