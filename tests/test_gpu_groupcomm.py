"""GPU parity tests of the GroupComm path: drop-in ``TasNet(module="DPRNN" | "DPTNet", group_size > 1)`` forward (csrc/groupcomm.cu) against the
reference's golden outputs and the CPU oracle (oracle/groupcomm_oracle.py)."""
import json
import os

import pytest
import torch

from conftest import GOLDEN, load_npz, record, rel_l2
from oracle import groupcomm_oracle as GO

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4   # rel-L2, BASELINE.json north_star
GC_MANIFEST = json.load(open(os.path.join(GOLDEN, "groupcomm_manifest.json")))


def _model(case):
    from audio_only_speech_separation_b200.models import TasNet

    c = GC_MANIFEST["cases"][case]
    torch.manual_seed(c["seed"])
    m = TasNet(**c["kwargs"])
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    return m.cuda().eval(), sd, c


@pytest.mark.parametrize("case", [c for c in GC_MANIFEST["cases"] if not c.startswith("grads_")])
def test_forward_matches_reference_golden(case):
    m, _, _ = _model(case)
    z = load_npz(f"groupcomm_{case}.npz")
    with torch.no_grad():
        y = m(torch.from_numpy(z["x"]).cuda())
    ref = torch.from_numpy(z["y"])
    assert tuple(y.shape) == tuple(ref.shape)
    err = rel_l2(y, ref)
    record("groupcomm_fwd_fp32", case=case, rel_l2=err, launches=m.last_launches)
    assert err < FP32_TOL
    assert m.last_launches > 0


@pytest.mark.parametrize("case,B,T", [("g16_b2_t8001", 3, 2777), ("g16_unfold_b2_t4001", 2, 2500), ("dpt_g16_b2_t4001", 2, 2777),
                                      ("dpt_g8_l2_b1_t3000", 2, 30000), ("g8_l2_b2_t4000", 2, 1601), ("g16_b2_t8001", 1, 97)])
def test_forward_matches_oracle_on_other_shapes(case, B, T):
    """Ragged lengths (short last context block, a single DPRNN chunk pair) and batch independence."""
    m, sd, c = _model(case)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, T, generator=g) * 0.1
    with torch.no_grad():
        y = m(x.cuda())
        ref = GO.tasnet_gc_forward(sd, x, group_size=c["kwargs"]["group_size"], layer=c["kwargs"].get("layer", 6), unfold=c["kwargs"].get("unfold", False), module=c["kwargs"]["module"])
        y0 = m(x[:1].cuda())
    err = rel_l2(y, ref)
    record("groupcomm_fwd_oracle", case=case, B=B, T=T, rel_l2=err)
    assert err < FP32_TOL
    assert rel_l2(y[:1], y0) < 1e-6   # every utterance is independent through the path


def test_other_context_and_block_sizes_match_oracle():
    """context_size 40 takes the unfused context stage (LSTM / projection / norm as separate kernels), 16 the fused one; block_size 50."""
    from audio_only_speech_separation_b200.models import TasNet

    for ctx, G in ((40, 16), (16, 8)):
        torch.manual_seed(3)
        m = TasNet(module="DPRNN", enc_dim=64, bn_dim=64, group_size=G, context_size=ctx, block_size=50, layer=2)
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        m = m.cuda().eval()
        x = torch.randn(2, 5003, generator=torch.Generator().manual_seed(5)) * 0.1
        with torch.no_grad():
            y = m(x.cuda())
            ref = GO.tasnet_gc_forward(sd, x, group_size=G, layer=2, context_size=ctx, block_size=50)
        err = rel_l2(y, ref)
        record("groupcomm_fwd_oracle_ctx", context_size=ctx, group_size=G, rel_l2=err, launches=m.last_launches)
        assert err < FP32_TOL


def test_full_size_batch_properties():
    """4 s at 8 kHz, B = 16: finite, scale-equivariant in the way the path is (the bottleneck GroupNorm makes the masks invariant to
    the input gain up to its eps = 1.2e-7 against the frame variance, so est(a * x) = a * est(x) for unit-scale inputs), and
    deterministic up to the order of the statistics atomics."""
    m, _, _ = _model("g16_b2_t8001")
    g = torch.Generator().manual_seed(11)
    x = torch.randn(16, 32000, generator=g).cuda()
    with torch.no_grad():
        y1, y2, y3 = m(x), m(x), m(2.0 * x)
    assert torch.isfinite(y1).all()
    assert rel_l2(y2, y1) < 1e-6
    assert rel_l2(y3, 2.0 * y1) < 1e-4


def test_training_and_unsupported_configurations_fail_loudly():
    from audio_only_speech_separation_b200 import _lib
    from audio_only_speech_separation_b200.models import TasNet

    m, _, _ = _model("g16_b1_t300")
    x = torch.randn(1, 800).cuda()
    m.train()
    assert m(x).requires_grad                # the training engine serves it (gradient tests below)
    with torch.no_grad():
        assert m(x).shape == (1, 2, 800)   # no graph requested: the inference engine serves it
    with pytest.raises(_lib.DualPathError):
        TasNet(module="DPRNN", group_size=4).cuda().eval()(x)   # per-group widths (16, 32): not built
    with pytest.raises(RuntimeError):
        TasNet(module="DPRNN", group_size=16).eval()(x)          # parameters on the CPU: no CPU path


def test_evaluation_loop_with_groupcomm_model():
    """The evaluation loop (audio_test.py:72-81) over a GroupComm model: batched by length = one by one."""
    import numpy as np

    from audio_only_speech_separation_b200.metrics import MetricsTracker, evaluate

    m, _, _ = _model("g16_b1_t300")
    g = torch.Generator().manual_seed(5)
    data = []
    for i, T in enumerate([4000, 4000, 3000, 4000, 3000]):
        src = torch.randn(2, T, generator=g) * 0.1
        data.append((src.sum(0), src, f"utt{i}"))
    one = evaluate(m, data, MetricsTracker(), batch_size=1)
    one.final()
    many = evaluate(m, data, MetricsTracker(), batch_size=4)
    many.final()
    assert np.isfinite(one.all_sisnrs_i).all()
    assert np.allclose(sorted(one.all_sisnrs_i), sorted(many.all_sisnrs_i), atol=1e-4)


def test_staged_and_prefetching_recurrence_agree():
    """The narrow recurrence with its inputs staged in shared memory (default) against the global-prefetch path kept for very long sequences."""
    from audio_only_speech_separation_b200._lib import lib

    for case in ("g16_b2_t8001", "dpt_g8_l2_b1_t3000"):
        m, _, _ = _model(case)
        x = (torch.randn(2, 6000, generator=torch.Generator().manual_seed(2)) * 0.1).cuda()
        with torch.no_grad():
            y1 = m(x)
            prev = lib().dp_gctasnet_set_lstm_staging(0)
            try:
                y0 = m(x)
            finally:
                lib().dp_gctasnet_set_lstm_staging(prev)
        assert prev == 1
        assert rel_l2(y1, y0) < 1e-6


# ------------------------------------------------------------------------------------------------ training backward (grouped DPRNN)
def _gc_grads(kwargs, seed, x, tgt):
    from audio_only_speech_separation_b200.losses import PITLossWrapper, pairwise_neg_snr
    from audio_only_speech_separation_b200.models import TasNet

    torch.manual_seed(seed)
    m = TasNet(**kwargs)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.cuda().train()
    loss = PITLossWrapper(pairwise_neg_snr, pit_from="pw_mtx", threshold_byloss=False)(m(x.cuda()), tgt.cuda())
    loss.backward()
    return m, sd, loss


@pytest.mark.parametrize("name", ["g16", "dpt_g16"])
def test_groupcomm_gradients_match_reference_golden(name):
    """Every parameter gradient of the reference's own loss.backward() (tests/golden/groupcomm_grads_{g16,dpt_g16}.npz: grouped DPRNN and
    grouped DPTNet, G = 16, 2 layers)."""
    c = GC_MANIFEST["cases"][f"grads_{name}"]
    z = load_npz(f"groupcomm_grads_{name}.npz")
    m, sd, loss = _gc_grads(c["kwargs"], c["seed"], torch.from_numpy(z["x"]), torch.from_numpy(z["tgt"]))
    assert abs(loss.item() - float(z["loss"])) < 1e-4
    num = den = 0.0
    worst = (0.0, None)
    for k, p in m.named_parameters():
        ref = torch.from_numpy(z["grad::" + k])
        assert p.grad is not None, k
        e = rel_l2(p.grad, ref)
        num += float((p.grad.cpu().double() - ref.double()).pow(2).sum())
        den += float(ref.double().pow(2).sum())
        worst = max(worst, (e, k))
    record("groupcomm_grads", case=name, total_rel_l2=(num / den) ** 0.5, worst=worst[0], worst_key=worst[1])
    assert (num / den) ** 0.5 < 1e-4, worst
    assert worst[0] < 5e-3, worst


@pytest.mark.parametrize("kw", [dict(module="DPRNN", enc_dim=64, bn_dim=64, group_size=16, layer=2, unfold=True),
                                dict(module="DPRNN", enc_dim=64, bn_dim=64, hidden_dim=128, group_size=8, layer=1, context_size=16, block_size=20),
                                dict(module="DPRNN", enc_dim=64, bn_dim=128, hidden_dim=256, group_size=32, layer=1),
                                dict(module="DPTNet", enc_dim=64, bn_dim=64, group_size=16, layer=2, unfold=True),
                                dict(module="DPTNet", enc_dim=64, bn_dim=64, hidden_dim=128, group_size=8, layer=1, block_size=30)],
                         ids=["g16_unfold", "g8_ctx16", "g32", "dpt_g16_unfold", "dpt_g8"])
def test_groupcomm_gradients_match_oracle_autograd(kw):
    """Other configurations (unfold with the shared concat_block, per-group widths (8, 16), another context / block size, G = 32) against
    autograd through the oracle (itself pinned to the reference's gradients, tests/test_groupcomm_oracle.py)."""
    from oracle import dualpath_oracle as O

    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 2600, generator=g) * 0.1
    tgt = torch.randn(2, 2, 2600, generator=g) * 0.1
    m, sd, loss = _gc_grads(kw, 1, x, tgt)
    by_storage, leaf = {}, {}
    for k, v in m.state_dict().items():   # unfold: aliased entries share one leaf so that autograd sums their gradients
        if v.data_ptr() not in by_storage:
            by_storage[v.data_ptr()] = sd[k].clone().requires_grad_(True)
        leaf[k] = by_storage[v.data_ptr()]
    est = GO.tasnet_gc_forward(leaf, x, enc_dim=kw["enc_dim"], bn_dim=kw["bn_dim"], group_size=kw["group_size"], layer=kw["layer"], module=kw["module"],
                               unfold=kw.get("unfold", False), context_size=kw.get("context_size", 24), block_size=kw.get("block_size", 100))
    ref_loss = O.pit_loss(est, tgt, "snr", False)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) < 1e-4 * max(1.0, abs(ref_loss.item()))
    num = den = 0.0
    for k, p in m.named_parameters():
        gr = leaf[k].grad
        num += float((p.grad.cpu().double() - gr.double()).pow(2).sum())
        den += float(gr.double().pow(2).sum())
    assert (num / den) ** 0.5 < 1e-4


def test_groupcomm_fused_training_steps_lower_the_loss():
    from audio_only_speech_separation_b200.losses import PITLossWrapper, pairwise_neg_snr
    from audio_only_speech_separation_b200.models import TasNet
    from audio_only_speech_separation_b200.trainer import DualPathTrainer

    torch.manual_seed(0)
    m = TasNet(module="DPRNN", enc_dim=64, bn_dim=64, group_size=16, layer=2).cuda().train()
    tr = DualPathTrainer(m, PITLossWrapper(pairwise_neg_snr, pit_from="pw_mtx", threshold_byloss=False))
    g = torch.Generator().manual_seed(2)
    src = (torch.randn(4, 2, 4000, generator=g) * 0.1).cuda()
    losses = [tr.step(src.sum(1).contiguous(), src).item() for _ in range(8)]
    assert all(l == l for l in losses) and losses[-1] < losses[0]
    torch.manual_seed(0)
    md = TasNet(module="DPTNet", enc_dim=64, bn_dim=64, group_size=16, layer=1).cuda().train()
    trd = DualPathTrainer(md, PITLossWrapper(pairwise_neg_snr, pit_from="pw_mtx", threshold_byloss=False))
    losses = [trd.step(src.sum(1).contiguous(), src).item() for _ in range(8)]
    assert all(l == l for l in losses) and losses[-1] < losses[0]
