"""tcgen05 / tensor-memory recurrence kernels (csrc/lstm_rec5.cu, forced with dp_set_lstm_tcgen05(2)) against the oracle restatement of
nn.LSTM (look2hear/models/utils/gc3_basics.py:16,22) and against the mma.sync kernels."""
import pytest
import torch

from oracle import dualpath_oracle as O

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm())


@pytest.fixture(scope="module")
def ops():
    from audio_only_speech_separation_b200 import ops as _ops

    return _ops


@pytest.fixture
def tc5():
    from audio_only_speech_separation_b200 import _lib

    _lib.check(_lib.lib().dp_set_lstm_tcgen05(2))
    yield _lib
    _lib.check(_lib.lib().dp_set_lstm_tcgen05(1))


def _lstm_and_pack(ops, seed=0):
    torch.manual_seed(seed)
    lstm = torch.nn.LSTM(64, 128, 1, batch_first=True, bidirectional=True)
    sd = {"rnn." + k: v.detach() for k, v in lstm.state_dict().items()}
    return lstm, sd, ops.LstmPack(lstm.cuda())


def _oracle_bilstm(x, sd, layout, impl="aten"):
    B, S, K, N = x.shape
    if layout == "intra":
        return O.bilstm(x.reshape(B * S, K, N), sd, "rnn.", impl).reshape(B, S, K, 256)
    xi = x.permute(0, 2, 1, 3).reshape(B * K, S, N)
    return O.bilstm(xi, sd, "rnn.", impl).reshape(B, K, S, 256).permute(0, 2, 1, 3)


@pytest.mark.parametrize("layout", ["intra", "inter"])
@pytest.mark.parametrize("B,S,K", [(1, 5, 7), (3, 21, 9), (2, 6, 10), (16, 82, 6), (1, 82, 100), (5, 33, 12)])   # ragged tiles and full tiles
def test_tc5_forward_backward_parity(ops, tc5, layout, B, S, K):
    lstm, sd, pack = _lstm_and_pack(ops, seed=3)
    g = torch.Generator().manual_seed(B * 100 + S + K)
    x = torch.randn(B, S, K, 64, generator=g)
    dH = torch.randn(B, S, K, 256, generator=g)
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    ref = _oracle_bilstm(xr, leaf, layout, impl="aten")
    ref.backward(dH)
    for prec, tol in (("fp32", 2e-5), ("bf16", 3e-2)):
        H, _, _ = ops.bilstm_forward(pack, x.cuda(), layout, precision=prec)
        Hs, G, Cst = ops.bilstm_forward(pack, x.cuda(), layout, save=True, precision=prec)
        assert torch.equal(H, Hs)
        assert rel_l2(H, ref.detach()) < tol, prec
        dx, dbias = ops.bilstm_backward(pack, G, Cst, dH.cuda(), (B, S, K), layout, precision=prec)
        assert rel_l2(dx, xr.grad) < (5e-5 if prec == "fp32" else 5e-2), prec
        if prec == "fp32":
            assert rel_l2(dbias, G.double().sum(0)) < 1e-5  # bias gradient fused into the BPTT kernel
            perm = torch.tensor([(r % 4) * 128 + r // 4 for r in range(512)])
            db = dbias.double().cpu()
            for d, name in enumerate(["rnn.bias_ih_l0", "rnn.bias_ih_l0_reverse"]):
                got = torch.empty(512, dtype=torch.float64)
                got[perm] = db[d * 512 : (d + 1) * 512]
                assert rel_l2(got, leaf[name].grad) < 5e-5


@pytest.mark.parametrize("layout", ["intra", "inter"])
def test_tc5_plane_outputs_two_waves(ops, tc5, layout):
    """Operand planes (h = hi + lo at every step, h_prev = previous step's h, zeros at a sequence's first step), saved gates / cell states
    and the BPTT outputs (dG planes, bias gradient) at B = 40 (3 280 / 4 000 sequences per direction: more than one wave of 32-sequence
    tiles), against the mma.sync kernels; twice, to catch run-to-run differences."""
    _lib = tc5
    lstm, sd, pack = _lstm_and_pack(ops, seed=4)
    B, S, K = 40, 82, 100
    P = B * S * K
    g = torch.Generator().manual_seed(11)
    G0 = (torch.randn(P, 1024, generator=g) * 0.5).cuda()
    dH = (torch.randn(P, 256, generator=g) * 0.1).cuda()
    nseq, ln, qdiv, s_hi, s_lo, s_t = (B * S, K, 1 << 30, 0, K, 1) if layout == "intra" else (B * K, S, K, S * K, 1, K)
    L = _lib.lib()

    def run(save):
        Gw, C = G0.clone(), torch.empty(P, 256, device="cuda")
        hh, hl, ph, plo = (torch.full((P, 256), float("nan"), device="cuda", dtype=torch.bfloat16) for _ in range(4))
        _lib.check(L.dp_lstm_recurrence_planes_f32(_lib.ptr(pack.buf), _lib.ptr(Gw), None, _lib.ptr(C) if save else None, _lib.ptr(hh), _lib.ptr(hl),
                                                  _lib.ptr(ph) if save else None, _lib.ptr(plo) if save else None, nseq, ln, qdiv, s_hi, s_lo,
                                                  s_t, save, 0, _lib.stream_ptr()))
        out = {"h": hh.float() + hl.float()}
        if save:
            out.update(hp=ph.float() + plo.float(), gates=Gw.clone(), c=C.clone())
            dbias = torch.zeros(1024, device="cuda")
            _lib.check(L.dp_bilstm_backward_f32(_lib.ptr(pack.buf), _lib.ptr(Gw), _lib.ptr(C), _lib.ptr(dH), None, 0, _lib.ptr(dbias), P, nseq, ln,
                                                qdiv, s_hi, s_lo, s_t, 0, _lib.stream_ptr()))
            out.update(dG=Gw, dbias=dbias)
        return out

    _lib.check(L.dp_set_lstm_tcgen05(0))
    ref_inf, ref = run(0), run(1)
    _lib.check(L.dp_set_lstm_tcgen05(2))
    first = None
    for rep in range(2):
        inf, trn = run(0), run(1)
        assert float((inf["h"] - ref_inf["h"]).abs().max()) < 2e-5
        for k in ("h", "hp", "gates", "c"):
            assert float((trn[k] - ref[k]).abs().max()) < 3e-5, (k, rep)
        assert rel_l2(trn["dG"], ref["dG"]) < 2e-5 and rel_l2(trn["dbias"], ref["dbias"]) < 2e-5
        if first is None:
            first = trn
        else:
            for k in ("h", "hp", "gates", "c", "dG"):
                assert torch.equal(trn[k], first[k]), k   # deterministic (the bias gradient is an atomic sum)


def test_tc5_model_forward_and_gradients(tc5):
    """The whole DPRNN engine with the tcgen05 recurrence forced: forward against the oracle, parameter gradients against oracle autograd."""
    from audio_only_speech_separation_b200.models import TasNet

    torch.manual_seed(0)
    model = TasNet(sample_rate=8000, layer=2)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.cuda()
    model.train()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 4000, generator=g) * 0.1
    w = torch.randn(2, 2, 4000, generator=g)
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = O.tasnet_forward(leaf, x, layer=2)
    (ref * w).sum().backward()
    y = model(x.cuda())
    assert rel_l2(y, ref) < 1e-4
    (y * w.cuda()).sum().backward()
    num = den = 0.0
    for k, p in model.named_parameters():
        if leaf[k].grad is None:
            continue
        num += float((p.grad.double().cpu() - leaf[k].grad.double()).pow(2).sum())
        den += float(leaf[k].grad.double().pow(2).sum())
    assert (num / den) ** 0.5 < 1e-4
