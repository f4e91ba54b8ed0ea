"""CPU: the oracle (oracle/) against the golden vectors generated from the real reference (tests/golden/make_golden.py)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT, load_npz, rel_l2
from oracle import dualpath_oracle as O


def test_seg_ola_oracle_bit_exact():
    z = load_npz("seg_ola.npz")
    for i in range(int(z["n"])):
        L, K, rest, S = [int(v) for v in z[f"meta{i}"]]
        x, blk, y, mrg = (torch.from_numpy(z[f"{k}{i}"]) for k in ("x", "blk", "y", "mrg"))
        ob, orest = O.split_feature(x, K)
        assert orest == rest and ob.shape[3] == S
        assert torch.equal(ob, blk)
        assert torch.equal(O.merge_feature(y, rest), mrg)
        assert torch.equal(O.merge_feature(ob, rest), 2 * x)  # encode -> decode round trip
        assert O.seg_rest(L, K) == rest and O.num_chunks(L, K) == S


def test_seg_ola_c_oracle_bit_exact():
    so = os.path.join(ROOT, "oracle", "_build", "liboracle_seg.so")
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    lib = ctypes.CDLL(so)
    fp = ctypes.POINTER(ctypes.c_float)
    z = load_npz("seg_ola.npz")
    for i in range(int(z["n"])):
        L, K, rest, S = [int(v) for v in z[f"meta{i}"]]
        assert lib.oracle_seg_rest(L, K) == rest and lib.oracle_num_chunks(L, K) == S
        x, blk, y, mrg = (np.ascontiguousarray(z[f"{k}{i}"]) for k in ("x", "blk", "y", "mrg"))
        rows = x.shape[0] * x.shape[1]
        out = np.empty_like(blk)
        lib.oracle_segment_f32(x.ctypes.data_as(fp), out.ctypes.data_as(fp), rows, L, K)
        assert np.array_equal(out, blk)
        out2 = np.empty_like(mrg)
        lib.oracle_overlap_add_f32(y.ctypes.data_as(fp), out2.ctypes.data_as(fp), rows, K, S, rest)
        assert np.array_equal(out2, mrg)


def _init_state_dict(manifest, case):
    """Reference default init under the manifest's seed, via the drop-in model's parameter containers."""
    from audio_only_speech_separation_b200.models import TasNet

    c = manifest["cases"][case]
    torch.manual_seed(c["seed"])
    m = TasNet(sample_rate=c["sample_rate"], **c["audionet_config"])
    return {k: v.detach() for k, v in m.state_dict().items()}, c


@pytest.mark.parametrize("case", ["dprnn_wsj0_b2_t8001", "dprnn_wsj0_1d_t4000", "dprnn_wsj0_3d_t1234", "dprnn_unfold_b2_t8000",
                                  "dprnn_wsj0_b1_t32000"])
@pytest.mark.parametrize("impl", ["aten", "loop"])
def test_model_oracle_vs_golden(manifest, case, impl):
    if impl == "loop" and case == "dprnn_wsj0_b1_t32000":
        pytest.skip("covered by aten at full size")
    sd, c = _init_state_dict(manifest, case)
    z = load_npz(f"model_{case}.npz")
    ac = c["audionet_config"]
    with torch.no_grad():
        y = O.tasnet_forward(sd, torch.from_numpy(z["x"]), module=ac["module"], unfold=ac["unfold"], lstm_impl=impl)
    ref = torch.from_numpy(z["y"])
    assert y.shape == ref.shape
    assert rel_l2(y, ref) < 5e-6


def test_loss_oracle_vs_golden():
    z = load_npz("loss.npz")
    e, t = torch.from_numpy(z["ests"]), torch.from_numpy(z["targets"])
    for s in ("snr", "sisdr", "sdsdr"):
        assert torch.allclose(O.pairwise_neg_sdr(e, t, s), torch.from_numpy(z[f"pw_{s}"]), rtol=1e-5, atol=1e-5)
        for thr in (0, 1):
            loss, _, perm = O.pit_loss(e, t, s, bool(thr), True)
            assert abs(loss.item() - float(z[f"loss_{s}_{thr}"])) < 1e-5
            assert np.array_equal(perm.numpy(), z[f"perm_{s}"])


def test_grad_oracle_vs_golden(manifest):
    sd, _ = _init_state_dict(manifest, "dprnn_wsj0_b2_t8001")
    z = load_npz("grads_dprnn_wsj0.npz")
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    loss = O.pit_loss(O.tasnet_forward(leaf, torch.from_numpy(z["x"]), lstm_impl="aten"), torch.from_numpy(z["tgt"]), "snr", False)
    loss.backward()
    assert abs(loss.item() - float(z["loss"])) < 1e-5
    for key in z.files:
        if key.startswith("grad::"):
            assert rel_l2(leaf[key[6:]].grad, torch.from_numpy(z[key])) < 1e-4, key
    for k, n in manifest["grad_norms_dprnn_wsj0"].items():
        assert abs(leaf[k].grad.double().norm().item() - n) <= 1e-4 * max(n, 1e-6) + 1e-9, k


def test_adam_clip_oracle_matches_torch():
    torch.manual_seed(0)
    ps = [torch.randn(7, 5), torch.randn(11)]
    gs = [torch.randn(7, 5) * 3, torch.randn(11) * 3]
    ref = [p.clone().requires_grad_(True) for p in ps]
    opt = torch.optim.Adam(ref, lr=1e-3)
    mine = [p.clone() for p in ps]
    m = [torch.zeros_like(p) for p in ps]
    v = [torch.zeros_like(p) for p in ps]
    for step in range(1, 4):
        for r, g in zip(ref, gs):
            r.grad = g.clone() * step
        torch.nn.utils.clip_grad_norm_(ref, 5.0)
        opt.step()
        O.adam_clip_step(mine, [g * step for g in gs], m, v, step)
    for a, b in zip(mine, ref):
        assert torch.allclose(a, b.detach(), rtol=1e-6, atol=1e-7)


# ------------------------------------------------------------------------------------------------ SepFormer
SEPFORMER_CASES = ["sepformer_small_b2_t3000", "sepformer_small_b1_t1001_1d", "sepformer_smallpost_b1_t2001", "sepformer_base_b1_t16000"]


@pytest.mark.parametrize("case", SEPFORMER_CASES)
def test_sepformer_oracle_vs_golden(manifest, case):
    """oracle/sepformer_oracle.py against outputs of the real look2hear Sepformer; the drop-in model's default
    initialisation must reproduce the reference's weights (checksums pinned in the manifest)."""
    from audio_only_speech_separation_b200.models import Sepformer
    from oracle import sepformer_oracle as SO

    c = manifest["cases"][case]
    torch.manual_seed(c["seed"])
    m = Sepformer(sample_rate=c["sample_rate"], **c["audionet_config"])
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    ref = manifest["state_dicts"][c["config"]]
    assert set(sd) == set(ref)
    for k, (s, a, *shape) in ref.items():
        assert list(sd[k].shape) == shape, k
        assert abs(float(sd[k].double().sum()) - s) < 1e-9 and abs(float(sd[k].double().abs().sum()) - a) < 1e-9, k
    assert sum(p.numel() for p in m.parameters()) == manifest["n_params"][c["config"]]
    z = load_npz(f"model_{case}.npz")
    with torch.no_grad():
        y = SO.sepformer_forward(sd, torch.from_numpy(z["x"]), **c["audionet_config"])
    assert tuple(y.shape) == z["y"].shape
    assert rel_l2(y, torch.from_numpy(z["y"])) < 5e-6


def test_sepformer_batch_scramble_quirk(manifest):
    """sepformer.py:1004: decoder rows are ordered (spk, b) but reshaped as (b, spk) -- kept for parity."""
    from audio_only_speech_separation_b200.models import Sepformer
    from oracle import sepformer_oracle as SO

    c = manifest["cases"]["sepformer_small_b2_t3000"]
    torch.manual_seed(0)
    sd = {k: v.detach() for k, v in Sepformer(sample_rate=8000, **c["audionet_config"]).state_dict().items()}
    x = torch.from_numpy(load_npz("model_sepformer_small_b2_t3000.npz")["x"])
    with torch.no_grad():
        both = SO.sepformer_forward(sd, x, **c["audionet_config"])
        one = [SO.sepformer_forward(sd, x[i : i + 1], **c["audionet_config"])[0] for i in range(2)]  # [spk, T] each
    # row r = spk*B + b of the decoder lands at est[r // spks, r % spks]
    assert rel_l2(both[0, 0], one[0][0]) < 1e-5 and rel_l2(both[0, 1], one[1][0]) < 1e-5
    assert rel_l2(both[1, 0], one[0][1]) < 1e-5 and rel_l2(both[1, 1], one[1][1]) < 1e-5


def test_sepformer_oracle_gradients_vs_reference_golden(manifest):
    """Autograd through oracle/sepformer_oracle.py against gradients of the real reference (dropout disabled there)."""
    from audio_only_speech_separation_b200.models import Sepformer
    from oracle import sepformer_oracle as SO

    c = manifest["cases"]["sepformer_small_b2_t3000"]
    torch.manual_seed(0)
    sd = {k: v.detach() for k, v in Sepformer(sample_rate=8000, **c["audionet_config"]).state_dict().items()}
    z = load_npz("grads_sepformer_small.npz")
    x, tgt = torch.from_numpy(z["x"]), torch.from_numpy(z["tgt"])
    leaf = {k: v.clone().requires_grad_(v.dtype.is_floating_point and not k.endswith("pos_enc.pe")) for k, v in sd.items()}
    loss = O.pit_loss(SO.sepformer_forward(leaf, x, **c["audionet_config"]), tgt, "snr", False)
    loss.backward()
    assert abs(loss.item() - float(z["loss"])) < 1e-5
    for key in z.files:
        if key.startswith("grad::"):
            assert rel_l2(leaf[key[6:]].grad, torch.from_numpy(z[key])) < 1e-4, key
