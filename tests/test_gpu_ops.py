"""GPU parity tests of the individual kernels, through the C-ABI, against the CPU oracle and the golden vectors."""
import numpy as np
import pytest
import torch

from conftest import load_npz, record, rel_l2
from oracle import dualpath_oracle as O

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4  # rel-L2 gate of BASELINE.json north_star for fp32 mode


@pytest.fixture(scope="module")
def ops():
    from audio_only_speech_separation_b200 import ops as _ops

    return _ops


# ------------------------------------------------------------------------------- (a) segmentation / overlap-add
def test_segment_overlap_add_golden_bit_exact(ops):
    z = load_npz("seg_ola.npz")
    for i in range(int(z["n"])):
        L, K, rest, S = [int(v) for v in z[f"meta{i}"]]
        x, blk, y, mrg = (torch.from_numpy(z[f"{k}{i}"]).cuda() for k in ("x", "blk", "y", "mrg"))
        out, r = ops.split_feature(x, K)
        assert r == rest and tuple(out.shape) == tuple(blk.shape)
        assert torch.equal(out, blk), (L, K)
        assert torch.equal(ops.merge_feature(y, rest), mrg), (L, K)


@pytest.mark.parametrize("B,N,L,K", [(3, 5, 4002, 100), (2, 4, 15999, 250), (1, 3, 31999, 250), (2, 64, 1999, 100), (1, 1, 1, 2),
                                     (2, 2, 63, 64), (1, 2, 4000, 100), (1, 2, 5000, 2000)])
def test_segment_overlap_add_vs_oracle(ops, B, N, L, K):
    g = torch.Generator().manual_seed(L + K)
    x = torch.randn(B, N, L, generator=g)
    ref, rest = O.split_feature(x, K)
    out, r = ops.split_feature(x.cuda(), K)
    assert r == rest and torch.equal(out.cpu(), ref)
    y = torch.randn(ref.shape, generator=g)
    assert torch.equal(ops.merge_feature(y.cuda(), rest).cpu(), O.merge_feature(y, rest))
    # channels-last variants used inside the engine
    f = x.permute(0, 2, 1).contiguous().cuda()
    xcl = ops.segment_channels_last(torch.cat([f] * 4, dim=2), K)  # C multiple of 4
    assert torch.equal(xcl[..., :N].permute(0, 3, 2, 1).cpu(), ref)
    ycl = torch.cat([y.permute(0, 3, 2, 1)] * 4, dim=3).contiguous().cuda()
    fcl = ops.overlap_add_channels_last(ycl, L)
    assert torch.equal(fcl[..., :N].permute(0, 2, 1).cpu(), O.merge_feature(y, rest))


def test_segment_full_size_round_trip(ops):
    """BASELINE size (B=16 utterances of 4 s): merge(split(x)) == 2x exactly, and padding positions are zero."""
    x = torch.randn(16, 64, 4002, device="cuda")
    blk, rest = ops.split_feature(x, 100)
    assert tuple(blk.shape) == (16, 64, 100, 82) and rest == 48
    assert torch.equal(ops.merge_feature(blk, rest), 2 * x)
    assert float(blk[:, :, :50, 0].abs().max()) == 0.0  # leading half chunk is padding
    assert float(blk.double().sum()) == pytest.approx(2 * float(x.double().sum()), rel=1e-9)


def test_segment_rejects_bad_arguments(ops):
    from audio_only_speech_separation_b200._lib import DualPathError

    with pytest.raises(DualPathError):
        ops.split_feature(torch.randn(1, 2, 100, device="cuda"), 25)
    with pytest.raises(ValueError):
        ops.split_feature(torch.randn(2, 100, device="cuda"), 100)


# ------------------------------------------------------------------------------- (b) GEMMs
@pytest.mark.parametrize("M,N,K,w_kn,relu", [(1000, 64, 64, False, False), (777, 1024, 64, False, False), (513, 64, 256, False, True),
                                            (300, 16, 64, True, False), (4002, 128, 64, False, True), (260, 256, 64, True, False),
                                            (129, 64, 1024, True, False)])
def test_linear_fp32_parity(ops, M, N, K, w_kn, relu):
    g = torch.Generator().manual_seed(M)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K**0.5
    b = torch.randn(N, generator=g)
    ref = a.double() @ w.double().t() + b.double()
    if relu:
        ref = ref.clamp_min(0)
    wd = (w.t().contiguous() if w_kn else w).cuda()
    out = ops.linear(a.cuda(), wd, b.cuda(), w_kn=w_kn, relu=relu)
    err = rel_l2(out, ref)
    record("linear_fp32", M=M, N=N, K=K, w_kn=w_kn, rel_l2=err)
    assert err < 2e-5
    out16 = ops.linear(a.cuda(), wd, b.cuda(), w_kn=w_kn, relu=relu, precision="bf16")
    assert rel_l2(out16, ref) < 2e-2


def test_linear_accumulate_stats_and_frames(ops):
    g = torch.Generator().manual_seed(3)
    # overlapping frames (encoder Conv1d as a GEMM): rows at stride 8, K = 16
    x = torch.randn(8 * 501, generator=g)
    w = torch.randn(64, 16, generator=g)
    frames = x.unfold(0, 16, 8)
    out = ops.linear(x.cuda(), w.cuda(), None, lda=8, rows=frames.shape[0])
    assert rel_l2(out, frames.double() @ w.double().t()) < 2e-5
    # accumulate + per-group statistics
    a = torch.randn(600, 64, generator=g)
    w2 = torch.randn(64, 64, generator=g)
    base = torch.randn(600, 64, generator=g)
    stats = torch.zeros(3, 2, dtype=torch.float64, device="cuda")
    out = ops.linear(a.cuda(), w2.cuda(), None, out=base.clone().cuda(), accumulate=True, stats=stats, rows_per_group=200)
    ref = base.double() + a.double() @ w2.double().t()
    assert rel_l2(out, ref) < 2e-5
    rs = torch.stack([ref.reshape(3, -1).sum(1), (ref.reshape(3, -1) ** 2).sum(1)], dim=1)
    assert torch.allclose(stats.cpu(), rs, rtol=1e-4, atol=1e-3)


@pytest.fixture(params=[1, 0], ids=["tcgen05", "mma_sync"])
def backend(request):
    from audio_only_speech_separation_b200._lib import check, lib

    check(lib().dp_set_gemm_backend(request.param))
    yield request.param
    check(lib().dp_set_gemm_backend(2))  # the library default


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (1000, 256, 64), (8200, 1024, 64), (3333, 64, 1024), (700, 128, 256)])
def test_linear_backends_agree(ops, backend, M, N, K):
    """Same GEMM through the tcgen05/TMEM kernel and the mma.sync kernel, both against fp64."""
    g = torch.Generator().manual_seed(M + N)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K**0.5
    b = torch.randn(N, generator=g)
    base = torch.randn(M, N, generator=g)
    ref = a.double() @ w.double().t() + b.double()
    out = ops.linear(a.cuda(), w.cuda(), b.cuda())
    err = rel_l2(out, ref)
    record("linear_backend", backend=backend, M=M, N=N, K=K, rel_l2=err)
    assert err < 2e-5
    out2 = ops.linear(a.cuda(), w.cuda(), None, out=base.clone().cuda(), accumulate=True)
    assert rel_l2(out2, base.double() + a.double() @ w.double().t()) < 2e-5
    assert rel_l2(ops.linear(a.cuda(), w.cuda(), b.cuda(), precision="bf16"), ref) < 2e-2


@pytest.mark.parametrize("P,Mo,No", [(5000, 64, 256), (3000, 1024, 64), (1111, 512, 128), (4002, 64, 16), (2500, 256, 64), (40, 512, 128)])
def test_linear_wgrad(ops, backend, P, Mo, No):
    g = torch.Generator().manual_seed(P)
    a = torch.randn(P, Mo, generator=g)
    b = torch.randn(P, No, generator=g)
    out = torch.zeros(Mo, No, device="cuda")
    ops.linear_wgrad(a.cuda(), b.cuda(), out)
    err = rel_l2(out, a.double().t() @ b.double())
    record("linear_wgrad", backend=backend, P=P, Mo=Mo, No=No, rel_l2=err)
    assert err < 2e-5


# ------------------------------------------------------------------------------- (c) persistent BiLSTM
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
@pytest.mark.parametrize("B,S,K", [(2, 6, 10), (1, 82, 100), (16, 82, 20), (3, 14, 100)])
def test_bilstm_forward_parity(ops, layout, B, S, K):
    lstm, sd, pack = _lstm_and_pack(ops)
    g = torch.Generator().manual_seed(B * 1000 + S)
    x = torch.randn(B, S, K, 64, generator=g)
    with torch.no_grad():
        ref = _oracle_bilstm(x, sd, layout)
    H, _, _ = ops.bilstm_forward(pack, x.cuda(), layout)
    err = rel_l2(H, ref)
    record("bilstm_fwd_fp32", layout=layout, B=B, S=S, K=K, rel_l2=err)
    assert err < 2e-5
    H16, _, _ = ops.bilstm_forward(pack, x.cuda(), layout, precision="bf16")
    err16 = rel_l2(H16, ref)
    record("bilstm_fwd_bf16", layout=layout, B=B, S=S, K=K, rel_l2=err16)
    assert err16 < 3e-2


@pytest.mark.parametrize("layout", ["intra", "inter"])
@pytest.mark.parametrize("mode", [0, 2, 3])
@pytest.mark.parametrize("B,S,K", [(1, 5, 7), (3, 21, 9), (16, 82, 6)])   # ragged tiles (5, 63, 27 sequences) and full-size sequence counts
def test_bilstm_forward_kernel_variants(ops, layout, mode, B, S, K):
    """dp_set_lstm_pipeline: plain 8-warp, software-pipelined (16 + 8 sequence groups) and 16-warp forward kernels against the
    oracle, with and without saving the activated gates / cell states (the backward then consumes what each variant saved)."""
    from audio_only_speech_separation_b200 import _lib

    lstm, sd, pack = _lstm_and_pack(ops, seed=3)
    g = torch.Generator().manual_seed(B * 100 + S + K)
    x = torch.randn(B, S, K, 64, generator=g)
    dH = torch.randn(B, S, K, 256, generator=g)
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    ref = _oracle_bilstm(xr, leaf, layout, impl="aten")
    ref.backward(dH)
    _lib.check(_lib.lib().dp_set_lstm_pipeline(mode))
    _lib.check(_lib.lib().dp_set_lstm_cluster(0))   # the inference call below must run the selected variant, not the cluster kernel
    try:
        for prec, tol in (("fp32", 2e-5), ("bf16", 3e-2)):
            H, _, _ = ops.bilstm_forward(pack, x.cuda(), layout, precision=prec)
            Hs, G, Cst = ops.bilstm_forward(pack, x.cuda(), layout, save=True, precision=prec)
            assert torch.equal(H, Hs)
            err = rel_l2(H, ref.detach())
            record("bilstm_fwd_variant", mode=mode, prec=prec, layout=layout, B=B, S=S, K=K, rel_l2=err)
            assert err < tol
            if prec == "fp32":
                dx, _ = ops.bilstm_backward(pack, G, Cst, dH.cuda(), (B, S, K), layout)
                assert rel_l2(dx, xr.grad) < 5e-5
    finally:
        _lib.check(_lib.lib().dp_set_lstm_pipeline(1))
        _lib.check(_lib.lib().dp_set_lstm_cluster(1))


@pytest.mark.parametrize("layout", ["intra", "inter"])
@pytest.mark.parametrize("B,S,K", [(1, 5, 7), (2, 1, 3), (1, 82, 100), (2, 82, 100), (3, 21, 9), (5, 82, 12)])
def test_bilstm_forward_cluster_kernel(ops, layout, B, S, K):
    """dp_set_lstm_cluster: the four-CTA-cluster forward recurrence (gate rows split over the cluster, h exchanged through distributed
    shared memory) against the oracle and bit for bit against the 16-warp kernel: ragged tiles (5 / 7 sequences), one- and three-step
    sequences, the B = 1 and B = 2
    pass sizes of the bench geometry (8- and 16-sequence tiles), several waves of clusters (410 sequences, forced), fp32 and bf16 mode,
    fp32 H and the hi / lo operand planes."""
    from audio_only_speech_separation_b200 import _lib

    lstm, sd, pack = _lstm_and_pack(ops, seed=5)
    g = torch.Generator().manual_seed(B * 100 + S + K)
    x = torch.randn(B, S, K, 64, generator=g)
    with torch.no_grad():
        ref = _oracle_bilstm(x, sd, layout)
    L = _lib.lib()
    try:
        for prec, tol in (("fp32", 2e-5), ("bf16", 3e-2)):
            _lib.check(L.dp_set_lstm_cluster(0))
            _lib.check(L.dp_set_lstm_pipeline(3))
            _lib.check(L.dp_set_lstm_tcgen05(0))
            H0, _, _ = ops.bilstm_forward(pack, x.cuda(), layout, precision=prec)
            _lib.check(L.dp_set_lstm_cluster(2))
            H1, _, _ = ops.bilstm_forward(pack, x.cuda(), layout, precision=prec)
            for _ in range(3):   # run-to-run: the exchange protocol has no data race that timing could expose
                H2, _, _ = ops.bilstm_forward(pack, x.cuda(), layout, precision=prec)
                assert torch.equal(H1, H2)
            err = rel_l2(H1, ref)
            record("bilstm_fwd_cluster", prec=prec, layout=layout, B=B, S=S, K=K, rel_l2=err, bit_equal=bool(torch.equal(H0, H1)))
            assert err < tol
            assert torch.equal(H0, H1)
        # operand planes (what the engines consume in inference)
        P = B * S * K
        nseq, ln, qdiv, s_hi, s_lo, s_t = (B * S, K, 1 << 30, 0, K, 1) if layout == "intra" else (B * K, S, K, S * K, 1, K)
        G0 = (torch.randn(P, 1024, generator=g) * 0.5).cuda()
        outs = []
        for mode in (0, 2):
            _lib.check(L.dp_set_lstm_cluster(mode))
            hh, hl = (torch.full((P, 256), float("nan"), device="cuda", dtype=torch.bfloat16) for _ in range(2))
            Gw = G0.clone()
            _lib.check(L.dp_lstm_recurrence_planes_f32(_lib.ptr(pack.buf), _lib.ptr(Gw), None, None, _lib.ptr(hh), _lib.ptr(hl), None, None, nseq, ln,
                                                      qdiv, s_hi, s_lo, s_t, 0, 0, _lib.stream_ptr()))
            assert torch.equal(Gw, G0)   # inference leaves the pre-activations alone
            outs.append((hh, hl))
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    finally:
        _lib.check(L.dp_set_lstm_cluster(1))
        _lib.check(L.dp_set_lstm_pipeline(1))
        _lib.check(L.dp_set_lstm_tcgen05(1))


@pytest.mark.parametrize("layout", ["intra", "inter"])
@pytest.mark.parametrize("B,S,K", [(2, 6, 10), (16, 82, 12)])
def test_bilstm_backward_parity(ops, layout, B, S, K):
    lstm, sd, pack = _lstm_and_pack(ops, seed=1)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, S, K, 64, generator=g)
    dH = torch.randn(B, S, K, 256, generator=g)
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    ref = _oracle_bilstm(xr, leaf, layout, impl="loop")
    ref.backward(dH)
    H, G, Cst = ops.bilstm_forward(pack, x.cuda(), layout, save=True)
    assert rel_l2(H, ref.detach()) < 2e-5
    dx, dbias = ops.bilstm_backward(pack, G, Cst, dH.cuda(), (B, S, K), layout)
    err = rel_l2(dx, xr.grad)
    record("bilstm_bwd_dx", layout=layout, B=B, S=S, K=K, rel_l2=err)
    assert err < 5e-5
    # weight gradient of the input projection from the d(pre-activation) buffer (packed row order u*4 + gate)
    dW = torch.zeros(1024, 64, device="cuda")
    ops.linear_wgrad(G, x.reshape(-1, 64).cuda(), dW)
    perm = torch.tensor([(r % 4) * 128 + r // 4 for r in range(512)])
    for d, name in enumerate(["rnn.weight_ih_l0", "rnn.weight_ih_l0_reverse"]):
        got = torch.empty(512, 64)
        got[perm] = dW[d * 512 : (d + 1) * 512].cpu()
        e = rel_l2(got, leaf[name].grad)
        record("bilstm_bwd_dwih", layout=layout, dir=d, rel_l2=e)
        assert e < 5e-5
    assert rel_l2(dbias, G.double().sum(0)) < 1e-5  # bias gradient fused into the BPTT kernel
    db = dbias.double().cpu()
    for d, name in enumerate(["rnn.bias_ih_l0", "rnn.bias_ih_l0_reverse"]):
        got = torch.empty(512, dtype=torch.float64)
        got[perm] = db[d * 512 : (d + 1) * 512]
        assert rel_l2(got, leaf[name].grad) < 5e-5


# ------------------------------------------------------------------------------- (d) GroupNorm + residual
@pytest.mark.parametrize("unfold", [False, True])
def test_groupnorm_residual(ops, unfold):
    g = torch.Generator().manual_seed(11)
    B, P, C = 3, 700, 64
    y = torch.randn(B * P, C, generator=g) * 2 + 0.3
    res = torch.randn(B * P, C, generator=g)
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    stats = torch.stack([y.double().reshape(B, -1).sum(1), (y.double().reshape(B, -1) ** 2).sum(1)], dim=1).cuda()
    ref = res + O.group_norm1(y.reshape(B, P, C).permute(0, 2, 1), gamma, beta, 1e-8).permute(0, 2, 1).reshape(B * P, C)
    concat = None
    if unfold:
        cw, cb, sl = torch.randn(C, generator=g), torch.randn(C, generator=g), torch.tensor([0.25])
        ref = O.prelu(ref * cw + cb, sl)
        concat = (cw.cuda(), cb.cuda(), sl.cuda())
    out = ops.groupnorm_residual(y.cuda(), res.cuda(), gamma.cuda(), beta.cuda(), stats, P, 1e-8, concat=concat)
    assert rel_l2(out, ref) < 1e-5


# ------------------------------------------------------------------------------- (e) fused PIT loss
def _losses():
    from audio_only_speech_separation_b200 import losses

    return losses


def test_pit_loss_golden():
    L = _losses()
    z = load_npz("loss.npz")
    e, t = torch.from_numpy(z["ests"]).cuda(), torch.from_numpy(z["targets"]).cuda()
    fns = {"snr": L.pairwise_neg_snr, "sisdr": L.pairwise_neg_sisdr, "sdsdr": L.pairwise_neg_sdsdr}
    for s, fn in fns.items():
        pw = fn(e, t).cpu()
        ref = torch.from_numpy(z[f"pw_{s}"])
        # entries beyond +-90 dB are rounding noise even in the reference (its own fp64 evaluation moves them by
        # 0.05 dB: utterance 3 of tests/golden/make_golden.py): compare those loosely
        sane = ref.abs() < 90
        assert torch.allclose(pw[sane], ref[sane], rtol=1e-4, atol=2e-3), (s, (pw - ref).abs().max())
        assert torch.allclose(pw[~sane], ref[~sane], rtol=0, atol=0.5), (s, (pw - ref).abs().max())
        for thr in (0, 1):
            loss, reordered = L.PITLossWrapper(fn, pit_from="pw_mtx", threshold_byloss=bool(thr))(e, t, return_ests=True)
            assert abs(loss.item() - float(z[f"loss_{s}_{thr}"])) < 1e-3 * max(1.0, abs(float(z[f"loss_{s}_{thr}"])))
            perm = torch.from_numpy(z[f"perm_{s}"])
            expect = torch.stack([ee[p] for ee, p in zip(e.cpu(), perm)])
            assert torch.equal(reordered.cpu(), expect)


@pytest.mark.parametrize("sdr", ["snr", "sisdr", "sdsdr"])
@pytest.mark.parametrize("thr", [False, True])
def test_pit_loss_gradient(sdr, thr):
    L = _losses()
    g = torch.Generator().manual_seed(5)
    B, T = 5, 3001
    t = torch.randn(B, 2, T, generator=g)
    e = torch.randn(B, 2, T, generator=g) * 0.5 + 0.1
    e[1] = t[1].flip(0) + 0.05 * torch.randn(2, T, generator=g)
    e[2] = t[2] + 1e-3 * torch.randn(2, T, generator=g)
    er = e.clone().requires_grad_(True)
    ref = O.pit_loss(er, t, sdr, thr)
    ref.backward()
    ec = e.clone().cuda().requires_grad_(True)
    fn = {"snr": L.pairwise_neg_snr, "sisdr": L.pairwise_neg_sisdr, "sdsdr": L.pairwise_neg_sdsdr}[sdr]
    loss = L.PITLossWrapper(fn, pit_from="pw_mtx", threshold_byloss=thr)(ec, t.cuda())
    (loss * 3.0).backward()
    assert abs(loss.item() - ref.item()) < 1e-4 * max(1.0, abs(ref.item()))
    err = rel_l2(ec.grad / 3.0, er.grad)
    record("pit_loss_grad", sdr=sdr, thr=thr, rel_l2=err)
    assert err < 1e-4


def test_pit_loss_full_size_properties():
    """B=16 x 4 s: permutation/scale invariances that do not need the oracle."""
    L = _losses()
    g = torch.Generator().manual_seed(9)
    t = torch.randn(16, 2, 32000, generator=g).cuda()
    e = (t + 0.3 * torch.randn(16, 2, 32000, generator=g).cuda())
    w = L.PITLossWrapper(L.pairwise_neg_sisdr, pit_from="pw_mtx", threshold_byloss=False)
    a = w(e, t).item()
    assert abs(w(e.flip(1).contiguous(), t).item() - a) < 1e-4  # permutation invariant
    assert abs(w(e * 3.7, t).item() - a) < 1e-3                  # SI-SDR is scale invariant
    assert abs(w(e + 0.5, t).item() - a) < 1e-3                  # zero-mean
    snr = 10 * np.log10(1 / 0.09)
    assert abs(-a - snr) < 0.2


# ------------------------------------------------------------------------------- optimizer
def test_adam_clip_matches_oracle():
    from audio_only_speech_separation_b200._lib import check, lib, ptr, stream_ptr

    g = torch.Generator().manual_seed(2)
    n = 100_003
    p0 = torch.randn(n, generator=g)
    p_ref, m_ref, v_ref = [p0.clone()], [torch.zeros(n)], [torch.zeros(n)]
    p, m, v = p0.clone().cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    norm2 = torch.zeros(1, dtype=torch.float64, device="cuda")
    for step in range(1, 4):
        grad = torch.randn(n, generator=g) * (0.001 if step == 2 else 0.1)  # step 2 is below the clip threshold
        total = O.adam_clip_step(p_ref, [grad.clone()], m_ref, v_ref, step)
        check(lib().dp_adam_clip_step(ptr(p), ptr(grad.cuda()), ptr(m), ptr(v), n, ptr(norm2), 1.0, 5.0, 1e-3, 0.9, 0.999, 1e-8, step, 0.0,
                                      stream_ptr()))
        assert abs(float(norm2.sqrt()) - total) < 1e-4 * total
        assert torch.allclose(p.cpu(), p_ref[0], rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------- torch.library operators (dualpath::)
def test_torch_library_ops_forward_and_autograd():
    """torch.ops.dualpath.*: values equal the plain operator functions; registered backwards against torch autograd references."""
    from audio_only_speech_separation_b200 import torch_ops  # noqa: F401  (registers the operators)

    g = torch.Generator().manual_seed(3)
    # segmentation / overlap-add: bit-exact forward, and each is the other's adjoint: <seg(x), y> == <x, ola(y)>
    x = torch.randn(2, 8, 1003, generator=g).cuda().requires_grad_(True)
    y = torch.ops.dualpath.segment(x, 50)
    ref, rest = O.split_feature(x.detach().cpu(), 50)
    assert torch.equal(y.detach().cpu(), ref)
    w = torch.randn(*y.shape, generator=g).cuda()
    (y * w).sum().backward()
    assert torch.equal(x.grad.cpu(), O.merge_feature(w.cpu(), rest))
    wl = w.clone().requires_grad_(True)
    z = torch.ops.dualpath.overlap_add(wl, rest)
    v = torch.randn(*z.shape, generator=g).cuda()
    (z * v).sum().backward()
    assert torch.equal(wl.grad.cpu(), O.split_feature(v.cpu(), 50)[0])
    # attention core against torch SDPA (fp32 math), forward and gradient
    B, S, K, E, H = 1, 5, 37, 64, 4
    qkv = (torch.randn(B, S, K, 3 * E, generator=g) * 0.5).cuda().requires_grad_(True)
    o, _ = torch.ops.dualpath.attention(qkv, H, "intra")
    go = torch.randn(B, S, K, E, generator=g).cuda()
    (o * go).sum().backward()
    q2 = qkv.detach().double().cpu().requires_grad_(True)
    q, k, vv = (t.reshape(B * S, K, H, E // H).transpose(1, 2) for t in q2.split(E, dim=-1))
    ro = torch.nn.functional.scaled_dot_product_attention(q, k, vv).transpose(1, 2).reshape(B, S, K, E)
    (ro * go.double().cpu()).sum().backward()
    assert rel_l2(o.detach(), ro.detach()) < 1e-5 and rel_l2(qkv.grad, q2.grad) < 1e-5
    # residual add + LayerNorm
    a = torch.randn(300, 64, generator=g).cuda().requires_grad_(True)
    b = torch.randn(300, 64, generator=g).cuda().requires_grad_(True)
    gam = (1 + 0.1 * torch.randn(64, generator=g)).cuda().requires_grad_(True)
    bet = (0.1 * torch.randn(64, generator=g)).cuda().requires_grad_(True)
    out, _ = torch.ops.dualpath.add_layernorm(a, b, gam, bet, 1e-5)
    gout = torch.randn(300, 64, generator=g).cuda()
    (out * gout).sum().backward()
    a2, b2, g2, be2 = (t.detach().double().cpu().requires_grad_(True) for t in (a, b, gam, bet))
    r = torch.nn.functional.layer_norm(a2 + b2, (64,), g2, be2, 1e-5)
    (r * gout.double().cpu()).sum().backward()
    assert rel_l2(out.detach(), r.detach()) < 1e-5
    for got, want in ((a.grad, a2.grad), (b.grad, b2.grad), (gam.grad, g2.grad), (bet.grad, be2.grad)):
        assert rel_l2(got, want) < 1e-5
    # fused PIT loss against the oracle, with gradient
    est = (torch.randn(3, 2, 2000, generator=g) * 0.1).cuda().requires_grad_(True)
    tgt = (torch.randn(3, 2, 2000, generator=g) * 0.1).cuda()
    loss, perm, _ = torch.ops.dualpath.pit_sdr_loss(est, tgt, "sisdr", False)
    loss.backward()
    e2 = est.detach().cpu().requires_grad_(True)
    rl = O.pit_loss(e2, tgt.cpu(), "sisdr", False)
    rl.backward()
    assert abs(loss.item() - rl.item()) < 1e-4 * max(1.0, abs(rl.item())) and rel_l2(est.grad, e2.grad) < 1e-4


@pytest.mark.parametrize("layout", ["intra", "inter"])
@pytest.mark.parametrize("mode", [0, 2, 3])
def test_recurrence_plane_outputs_two_waves(ops, layout, mode):
    """The operand planes the engines consume (h = hi + lo of every step, h_prev = the previous step's h, zeros at a sequence's first
    step) from every forward kernel variant, at B = 24 (1 968 / 2 400 sequences: more than one wave of full 24-sequence tiles --
    the size at which a missing barrier before the last cell update of the pipelined kernel showed), inference and training mode."""
    from audio_only_speech_separation_b200 import _lib

    lstm, sd, pack = _lstm_and_pack(ops, seed=4)
    B, S, K = 24, 82, 100
    P = B * S * K
    g = torch.Generator().manual_seed(11)
    G0 = (torch.randn(P, 1024, generator=g) * 0.5).cuda()
    nseq, ln, qdiv, s_hi, s_lo, s_t = (B * S, K, 1 << 30, 0, K, 1) if layout == "intra" else (B * K, S, K, S * K, 1, K)
    L = _lib.lib()
    try:
        _lib.check(L.dp_set_lstm_pipeline(0))
        Gr, Href = G0.clone(), torch.empty(P, 256, device="cuda")
        _lib.check(L.dp_lstm_recurrence_f32(_lib.ptr(pack.buf), _lib.ptr(Gr), _lib.ptr(Href), None, nseq, ln, qdiv, s_hi, s_lo, s_t, 0, 0, _lib.stream_ptr()))
        # h_prev reference: h shifted by one step along the sequence, per direction (forward: t - 1, backward: t + 1)
        Hs = Href.reshape(B, S, K, 256)
        ax = 2 if layout == "intra" else 1
        prev = torch.zeros_like(Hs)
        sl = [slice(None)] * 4
        a, b = list(sl), list(sl)
        a[ax], b[ax] = slice(1, None), slice(0, -1)
        prev[tuple(a)][..., :128] = Hs[tuple(b)][..., :128]
        prev[tuple(b)][..., 128:] = Hs[tuple(a)][..., 128:]
        _lib.check(L.dp_set_lstm_pipeline(mode))
        for rep in range(2):
            for save in (0, 1):
                Gw, C = G0.clone(), torch.empty(P, 256, device="cuda")
                hh, hl, ph, plo = (torch.full((P, 256), float("nan"), device="cuda", dtype=torch.bfloat16) for _ in range(4))
                _lib.check(L.dp_lstm_recurrence_planes_f32(_lib.ptr(pack.buf), _lib.ptr(Gw), None, _lib.ptr(C) if save else None, _lib.ptr(hh), _lib.ptr(hl),
                                                          _lib.ptr(ph) if save else None, _lib.ptr(plo) if save else None, nseq, ln, qdiv, s_hi, s_lo,
                                                          s_t, save, 0, _lib.stream_ptr()))
                assert float(((hh.float() + hl.float()) - Href).abs().max()) < 1e-5, (mode, save, rep)
                if save:
                    assert float(((ph.float() + plo.float()).reshape(B, S, K, 256) - prev).abs().max()) < 1e-5, (mode, save, rep)
    finally:
        _lib.check(L.dp_set_lstm_pipeline(1))
