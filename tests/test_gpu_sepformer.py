"""GPU parity tests of the SepFormer path (look2hear/models/sepformer.py) against the committed reference outputs and the oracle."""
import pytest
import torch

from conftest import load_npz, record, rel_l2
from oracle import dualpath_oracle as O
from oracle import sepformer_oracle as SO

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4      # rel-L2, BASELINE.json north_star
BF16_TOL_DB = 0.05   # |delta PIT SI-SNR| in dB


def _model(manifest, case, precision="fp32"):
    from audio_only_speech_separation_b200.models import Sepformer

    c = manifest["cases"][case]
    torch.manual_seed(c["seed"])
    m = Sepformer(sample_rate=c["sample_rate"], **c["audionet_config"])
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.cuda().eval()
    m.precision = precision
    return m, sd, c["audionet_config"]


@pytest.mark.parametrize("case", ["sepformer_small_b2_t3000", "sepformer_small_b1_t1001_1d", "sepformer_smallpost_b1_t2001",
                                  "sepformer_base_b1_t16000"])
def test_forward_matches_reference_golden(manifest, case):
    m, _, _ = _model(manifest, case)
    z = load_npz(f"model_{case}.npz")
    with torch.no_grad():
        y = m(torch.from_numpy(z["x"]).cuda())
    ref = torch.from_numpy(z["y"])
    assert tuple(y.shape) == tuple(ref.shape)
    err = rel_l2(y, ref)
    record("sepformer_fwd_fp32", case=case, rel_l2=err, launches=m.last_launches)
    assert err < FP32_TOL


def test_forward_lengths_and_state_dict_reload(manifest):
    m, sd, cfg = _model(manifest, "sepformer_small_b2_t3000")
    g = torch.Generator().manual_seed(9)
    for shape in ((1, 16), (3, 23), (2, 1, 777), (1, 4096)):
        x = torch.randn(*shape, generator=g) * 0.1
        with torch.no_grad():
            y = m(x.cuda()).cpu()
            ref = SO.sepformer_forward(sd, x, **cfg)
        assert tuple(y.shape) == tuple(ref.shape)
        assert rel_l2(y, ref) < FP32_TOL, shape
    from audio_only_speech_separation_b200.models import Sepformer

    torch.manual_seed(321)
    other = Sepformer(sample_rate=8000, **cfg)
    m.load_state_dict(other.state_dict())
    x = torch.randn(1, 2000, generator=g) * 0.1
    with torch.no_grad():
        y = m(x.cuda()).cpu()
        ref = SO.sepformer_forward({k: v.detach() for k, v in other.state_dict().items()}, x, **cfg)
    assert rel_l2(y, ref) < FP32_TOL
    with pytest.raises(Exception):
        m(torch.randn(1, 15).cuda())  # shorter than the encoder kernel: the reference's conv1d fails too


def test_forward_bf16_within_si_snr_budget(manifest):
    """bf16 mode (single tensor-core product) on structured input: PIT SI-SNR within 0.05 dB of the fp32 reference output."""
    m, sd, cfg = _model(manifest, "sepformer_base_b1_t16000", precision="bf16")
    g = torch.Generator().manual_seed(21)
    src = torch.randn(1, 2, 16000, generator=g) * 0.1
    mix = src.sum(1)
    with torch.no_grad():
        ref = SO.sepformer_forward(sd, mix, **cfg)
        y = m(mix.cuda()).cpu()
    si_ref = -O.pit_loss(ref, src, "sisdr", False).item()
    si_new = -O.pit_loss(y, src, "sisdr", False).item()
    proxy = -O.pairwise_neg_sdr(y, ref, "sisdr").diagonal(dim1=1, dim2=2).mean().item()
    record("sepformer_fwd_bf16", si_ref=si_ref, si_new=si_new, si_sdr_vs_ref=proxy, rel_l2=rel_l2(y, ref))
    assert abs(si_new - si_ref) <= BF16_TOL_DB
    assert proxy > 30.0




def test_training_needs_dropout_opt_in_and_gradients_match_oracle(manifest):
    """Forward + backward engine calls behind one autograd node (dropout sites off): loss and every parameter gradient
    against autograd through the oracle; a few of them also against the real reference's gradients (golden)."""
    from audio_only_speech_separation_b200.losses import PITLossWrapper, pairwise_neg_snr

    m, sd, cfg = _model(manifest, "sepformer_small_b2_t3000")
    z = load_npz("grads_sepformer_small.npz")
    x, tgt = torch.from_numpy(z["x"]), torch.from_numpy(z["tgt"])
    from audio_only_speech_separation_b200 import _lib

    m.train()
    assert m.dropout == 0.1               # the reference's default (sepformer.py:507)
    with pytest.raises(_lib.DualPathError):
        m(x.cuda())                       # enc_dim = 64 layers run on the mma.sync engine, which has no dropout: fails loudly
    m.dropout = 0.0
    leaf = {k: v.clone().requires_grad_(not k.endswith("pos_enc.pe")) for k, v in sd.items()}
    ref_loss = O.pit_loss(SO.sepformer_forward(leaf, x, **cfg), tgt, "snr", False)
    ref_loss.backward()
    loss = PITLossWrapper(pairwise_neg_snr, pit_from="pw_mtx", threshold_byloss=False)(m(x.cuda()), tgt.cuda())
    loss.backward()
    assert abs(loss.item() - ref_loss.item()) < 1e-4 * max(1.0, abs(ref_loss.item()))
    worst, worst_key, num, den, errs = 0.0, None, 0.0, 0.0, []
    named = dict(m.named_parameters())
    for k, p in named.items():
        assert p.grad is not None, k
        gr = leaf[k].grad
        e = rel_l2(p.grad, gr)
        errs.append(e)
        num += float((p.grad.cpu().double() - gr.double()).pow(2).sum())
        den += float(gr.double().pow(2).sum())
        if e > worst:
            worst, worst_key = e, k
    total = (num / den) ** 0.5
    errs.sort()
    median = errs[len(errs) // 2]
    record("sepformer_grads", worst_rel_l2=worst, worst_key=worst_key, total_rel_l2=total, median_rel_l2=median, loss=loss.item())
    # ReLU-flip conditioning: on this input 7 of the 716 800 FFN pre-activations of the reference lie within
    # 1e-5 of zero (4 of them in block 0 / intra / layer 1); our forward agrees with the reference to ~1e-5, so exactly those
    # units can take the other side of the ReLU and move the gradients of their layer (and of everything upstream) by ~1e-2,
    # while all other layers agree to ~3e-5.  Hence a strict bound on the typical key and loose bounds on the total / worst.
    assert median < 1e-4
    assert total < 5e-3
    assert worst < 5e-2, worst_key
    for key in z.files:
        if key.startswith("grad::"):
            assert rel_l2(named[key[6:]].grad, torch.from_numpy(z[key])) < 5e-2, key


def test_training_batch2_and_optimizer_step(manifest):
    m, sd, cfg = _model(manifest, "sepformer_small_b2_t3000")
    from audio_only_speech_separation_b200.losses import PITLossWrapper, pairwise_neg_snr

    m.train()
    m.dropout = 0.0
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    lossf = PITLossWrapper(pairwise_neg_snr, pit_from="pw_mtx", threshold_byloss=False)
    g = torch.Generator().manual_seed(8)
    x = (torch.randn(2, 2000, generator=g) * 0.1).cuda()
    tgt = (torch.randn(2, 2, 2000, generator=g) * 0.1).cuda()
    # gradient of a B = 2 batch against the oracle (covers the (spk, batch) row scramble in the backward)
    leaf = {k: v.clone().requires_grad_(not k.endswith("pos_enc.pe")) for k, v in sd.items()}
    O.pit_loss(SO.sepformer_forward(leaf, x.cpu(), **cfg), tgt.cpu(), "snr", False).backward()
    loss = lossf(m(x), tgt)
    loss.backward()
    errs = sorted(rel_l2(p.grad, leaf[k].grad) for k, p in m.named_parameters())
    num = sum(float((p.grad.cpu().double() - leaf[k].grad.double()).pow(2).sum()) for k, p in m.named_parameters())
    den = sum(float(leaf[k].grad.double().pow(2).sum()) for k, _ in m.named_parameters())
    assert errs[len(errs) // 2] < 1e-4 and (num / den) ** 0.5 < 5e-3  # same ReLU-flip conditioning as above
    losses = []
    for _ in range(4):
        opt.zero_grad()
        loss = lossf(m(x), tgt)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 5.0)
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0]
    assert m._flat_is_valid(x.device)


def test_training_tma_path_gradients_match_oracle():
    """enc_dim = 128 / d_ffn = 256 layers run the training forward and backward on the TMA-fed tcgen05 GEMMs (operand planes,
    transposed weight planes, tcgen05 attention forward): loss and gradients against autograd through the oracle and against the
    mma.sync engine (dp_set_gemm_backend(0)) on the same weights."""
    from audio_only_speech_separation_b200 import _lib
    from audio_only_speech_separation_b200.losses import PITLossWrapper, pairwise_neg_snr
    from audio_only_speech_separation_b200.models import Sepformer

    cfg = dict(encoder_out_nchannels=128, intra_dffn=256, inter_dffn=256, intra_nhead=4, inter_nhead=4, intra_numlayers=2, inter_numlayers=1,
               masknet_chunksize=50, masknet_numlayers=1)
    torch.manual_seed(5)
    m = Sepformer(sample_rate=8000, **cfg)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.cuda().train()
    m.dropout = 0.0
    g = torch.Generator().manual_seed(21)
    x = torch.randn(2, 2400, generator=g) * 0.1
    tgt = torch.randn(2, 2, 2400, generator=g) * 0.1
    leaf = {k: v.clone().requires_grad_(not k.endswith("pos_enc.pe")) for k, v in sd.items()}
    ref_loss = O.pit_loss(SO.sepformer_forward(leaf, x, **cfg), tgt, "snr", False)
    ref_loss.backward()
    lossf = PITLossWrapper(pairwise_neg_snr, pit_from="pw_mtx", threshold_byloss=False)

    def run(backend):
        _lib.check(_lib.lib().dp_set_gemm_backend(backend))
        try:
            for p in m.parameters():
                p.grad = None
            loss = lossf(m(x.cuda()), tgt.cuda())
            loss.backward()
            return loss.item(), {k: p.grad.detach().clone() for k, p in m.named_parameters()}
        finally:
            _lib.check(_lib.lib().dp_set_gemm_backend(2))

    loss_t, gt = run(2)
    loss_m, gm = run(0)
    assert abs(loss_t - ref_loss.item()) < 1e-4 * max(1.0, abs(ref_loss.item()))
    assert abs(loss_t - loss_m) < 1e-4 * max(1.0, abs(loss_m))

    def summary(grads):
        errs = sorted(rel_l2(grads[k], leaf[k].grad) for k in grads)
        num = sum(float((grads[k].cpu().double() - leaf[k].grad.double()).pow(2).sum()) for k in grads)
        den = sum(float(leaf[k].grad.double().pow(2).sum()) for k in grads)
        return errs[len(errs) // 2], (num / den) ** 0.5, errs[-1]

    med_t, tot_t, worst_t = summary(gt)
    med_m, tot_m, worst_m = summary(gm)
    record("sepformer_grads_tma", median=med_t, total=tot_t, worst=worst_t, median_mma=med_m, total_mma=tot_m, worst_mma=worst_m)
    assert med_t < 1e-4 and tot_t < 5e-3 and worst_t < 5e-2    # same ReLU-flip conditioning as the tests above
    # bf16 mode: runs, finite, close to the fp32 loss
    m.precision = "bf16"
    loss_b, gb = run(2)
    m.precision = "fp32"
    assert abs(loss_b - loss_t) < 5e-2 * max(1.0, abs(loss_t))
    assert all(torch.isfinite(v).all() for v in gb.values())


@pytest.mark.parametrize("chunk,T,pre", [(50, 2400, True), (8, 8800, True), (50, 2400, False), (8, 4400, False)])
def test_training_with_dropout_matches_oracle_given_the_same_masks(chunk, T, pre):   # (8, 8800): 276-position inter sequences; pre = False: post-norm layers
    """The reference's four dropout sites per layer (attention probabilities, attention output, FFN hidden, FFN output; p = 0.1) on the
    TMA engine.  The masks are a counter-based function of (seed, layer, site, element); tests/dropout_ref.py replays it in numpy, so the
    oracle runs with the SAME masks and the loss and every gradient can be compared exactly (the backward regenerates the masks)."""
    import dropout_ref as DR
    from audio_only_speech_separation_b200 import _lib
    from audio_only_speech_separation_b200.losses import PITLossWrapper, pairwise_neg_snr
    from audio_only_speech_separation_b200.models import Sepformer

    cfg = dict(encoder_out_nchannels=128, intra_dffn=256, inter_dffn=256, intra_nhead=4, inter_nhead=4, intra_numlayers=2, inter_numlayers=1,
               masknet_chunksize=chunk, masknet_numlayers=2, intra_norm_before=pre, inter_norm_before=pre)
    torch.manual_seed(5)
    m = Sepformer(sample_rate=8000, **cfg)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.cuda().train()
    assert m.dropout == 0.1
    g = torch.Generator().manual_seed(21)
    B, K = 2, chunk
    x = torch.randn(B, T, generator=g) * 0.1
    tgt = torch.randn(B, 2, T, generator=g) * 0.1
    torch.manual_seed(77)
    seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())      # the draw Sepformer.forward makes
    _, S = _lib.seg_geometry((T - 16) // 8 + 1, K)
    cb = DR.oracle_dropout(0.1, seed, B, S, K, 4, 4, 2, 1)
    leaf = {k: v.clone().requires_grad_(not k.endswith("pos_enc.pe")) for k, v in sd.items()}
    ref_loss = O.pit_loss(SO.sepformer_forward(leaf, x, dropout=cb, **cfg), tgt, "snr", False)
    ref_loss.backward()
    nodrop = O.pit_loss(SO.sepformer_forward(sd, x, **cfg), tgt, "snr", False).item()
    lossf = PITLossWrapper(pairwise_neg_snr, pit_from="pw_mtx", threshold_byloss=False)
    torch.manual_seed(77)
    loss = lossf(m(x.cuda()), tgt.cuda())
    loss.backward()
    assert abs(ref_loss.item() - nodrop) > 1e-5 * abs(nodrop)                 # the masks do change the loss ...
    assert abs(loss.item() - ref_loss.item()) < 1e-4 * max(1.0, abs(ref_loss.item()))   # ... and the engine applies the same ones
    errs = sorted(rel_l2(p.grad, leaf[k].grad) for k, p in m.named_parameters())
    num = sum(float((p.grad.cpu().double() - leaf[k].grad.double()).pow(2).sum()) for k, p in m.named_parameters())
    den = sum(float(leaf[k].grad.double().pow(2).sum()) for k, _ in m.named_parameters())
    record("sepformer_grads_dropout", median=errs[len(errs) // 2], total=(num / den) ** 0.5, worst=errs[-1], loss=loss.item(), loss_nodrop=nodrop)
    if pre:
        assert errs[len(errs) // 2] < 1e-4 and (num / den) ** 0.5 < 5e-3 and errs[-1] < 5e-2   # same ReLU-flip conditioning as above
    else:
        # post-norm layers (LayerNorm AFTER each residual sum) pass forward rounding differences on to every gradient far more strongly:
        # with the loss equal to 1e-7 the gradients of ALL keys, the tail's included, move together by 2e-4 ... 2e-3 depending on the
        # draw (tests/tools/diag_postnorm_grads.py: 1.4e-5 with one mask draw, 1.9e-4 without dropout); a wrong or missing term in the
        # backward would show as O(0.1 ... 1) on the keys it touches
        assert errs[len(errs) // 2] < 5e-3 and (num / den) ** 0.5 < 2e-2 and errs[-1] < 1e-1
    # a different seed gives different masks; eval mode ignores dropout
    torch.manual_seed(78)
    assert abs(lossf(m(x.cuda()), tgt.cuda()).item() - loss.item()) > 1e-4 * abs(loss.item())
    m.eval()
    with torch.no_grad():
        assert abs(lossf(m(x.cuda()), tgt.cuda()).item() - nodrop) < 1e-4 * max(1.0, abs(nodrop))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_base_config_long_input_batch_properties(precision):
    """configs/sepformer_base.yml at 8 s, B = 2: determinism, forward independence of the two utterances (up to the reference's
    (spk, batch) row scramble), gradients without dropout as the mean of the per-utterance gradients."""
    from audio_only_speech_separation_b200.losses import PITLossWrapper, pairwise_neg_snr
    from audio_only_speech_separation_b200.models import Sepformer

    torch.manual_seed(2)
    m = Sepformer(sample_rate=8000).cuda().eval()
    m.precision = precision
    g = torch.Generator().manual_seed(9)
    x = (torch.randn(2, 64000, generator=g) * 0.1).cuda()
    with torch.no_grad():
        y, y2 = m(x), m(x)
        singles = torch.cat([m(x[i : i + 1]) for i in range(2)])          # [2, 2, T]: rows (b, spk)
    assert bool(torch.isfinite(y).all()) and rel_l2(y, y2) == 0.0
    # batch-2 rows are the reference's reshape of (spk, b)-ordered rows: row r of the flattened output holds (spk, b) = divmod(r, 2)
    scr = torch.stack([singles[b, s] for s in range(2) for b in range(2)]).reshape(2, 2, -1)
    assert rel_l2(y, scr) < (3e-5 if precision == "fp32" else 2e-3)
    m.train()
    m.dropout = 0.0
    tgt = (torch.randn(2, 2, 64000, generator=g) * 0.1).cuda()

    def grad_of(fn):
        for p in m.parameters():
            p.grad = None
        fn().backward()
        return torch.cat([p.grad.flatten() for p in m.parameters()]).clone()

    g1 = grad_of(lambda: m(x).pow(2).mean())
    g2 = grad_of(lambda: 0.5 * (m(x[0:1]).pow(2).mean() + m(x[1:2]).pow(2).mean()))
    assert rel_l2(g1, g2) < (1e-3 if precision == "fp32" else 3e-2)


@pytest.mark.parametrize("mode", [1, 2])
def test_engine_attention_forward_kernels_agree_with_reference(manifest, mode):
    """dp_set_attention_forward: the tcgen05 kernel (1) and the warp-level tensor-core kernel (2) inside the engines, both against the
    reference golden (fp32 mode) and within the bf16 budget; the default picks per shape / precision from measurements."""
    from audio_only_speech_separation_b200 import _lib

    case = "sepformer_base_b1_t16000"
    m, _, _ = _model(manifest, case)
    z = load_npz(f"model_{case}.npz")
    _lib.check(_lib.lib().dp_set_attention_forward(mode))
    try:
        with torch.no_grad():
            y = m(torch.from_numpy(z["x"]).cuda())
            m.precision = "bf16"
            y16 = m(torch.from_numpy(z["x"]).cuda())
            m.precision = "fp32"
    finally:
        _lib.check(_lib.lib().dp_set_attention_forward(0))
    ref = torch.from_numpy(z["y"])
    record("sepformer_fwd_attention_mode", mode=mode, fp32=rel_l2(y, ref), bf16=rel_l2(y16, ref))
    assert rel_l2(y, ref) < FP32_TOL and rel_l2(y16, ref) < 3e-2
