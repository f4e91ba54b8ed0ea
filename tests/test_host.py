"""CPU: host logic, C-ABI surface, drop-in contract (no compute call needs a GPU)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
from audio_only_speech_separation_b200 import _lib
from oracle import dualpath_oracle as O


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "dualpath_b200.h")).read()
    declared = set(re.findall(r"\b(dp_[a-z0-9_]+)\s*\(", header))
    declared -= {"dp_tasnet_config"}
    assert len(declared) >= 25
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    handle = ctypes.CDLL(_lib.LIB_PATH)  # fails loudly if the library was not built
    for name in declared:
        assert hasattr(handle, name), name
    assert _lib.lib().dp_version() >= 100


@pytest.mark.parametrize("L,K", [(4002, 100), (3999, 250), (15999, 250), (31999, 250), (1999, 100), (100, 100), (50, 100),
                                 (4000, 100), (4050, 100), (7, 24), (1, 2)])
def test_geometry_matches_oracle(L, K):
    rest, S = _lib.seg_geometry(L, K)
    assert rest == O.seg_rest(L, K) and S == O.num_chunks(L, K)


@pytest.mark.parametrize("T", [1, 7, 8, 15, 16, 17, 1234, 8001, 32000])
def test_wave_geometry_matches_oracle(T):
    rest, frames = _lib.wave_geometry(T, 16)
    assert rest == O.wave_rest(T, 16) and frames == O.num_frames(T, 16)


def test_geometry_rejects_bad_arguments():
    with pytest.raises(_lib.DualPathError):
        _lib.seg_geometry(100, 25)  # odd chunk size
    with pytest.raises(_lib.DualPathError):
        _lib.seg_geometry(0, 100)


def test_state_dict_matches_reference_inventory(manifest):
    from audio_only_speech_separation_b200.models import TasNet

    for cfgname, case in (("dprnn_wsj0", "dprnn_wsj0_b2_t8001"), ("dprnn_lrs2_unfolded", "dprnn_unfold_b2_t8000")):
        c = manifest["cases"][case]
        torch.manual_seed(c["seed"])
        m = TasNet(sample_rate=c["sample_rate"], **c["audionet_config"])
        sd, ref = m.state_dict(), manifest["state_dicts"][cfgname]
        assert set(sd.keys()) == set(ref.keys())  # (the manifest json is key-sorted)
        for k, v in sd.items():
            assert list(v.shape) == ref[k][2:], k
            assert abs(float(v.double().sum()) - ref[k][0]) < 1e-9 and abs(float(v.double().abs().sum()) - ref[k][1]) < 1e-9, k
        assert sum(p.numel() for p in m.parameters()) == manifest["n_params"][cfgname]
        assert m.model_name == "DPRNN" and m.get_model_args() == {"n_src": 2} and m.sample_rate() == c["sample_rate"]


def test_unfold_aliases_parameters():
    from audio_only_speech_separation_b200.models import TasNet

    m = TasNet(unfold=True)
    sm = m.seq_model.seq_model
    assert all(sm.row_rnn[i] is sm.row_rnn[0] for i in range(6)) and all(sm.col_norm[i] is sm.col_norm[0] for i in range(6))
    table = m._param_table()
    assert len(table) == 12 + 12 * 12 and table[9] is sm.concat_block[0].weight
    assert table[12] is table[12 + 24]  # layer 0 and layer 1 row weight_ih are the same tensor


def test_registry_and_serialize(tmp_path):
    from audio_only_speech_separation_b200 import models

    assert models.get("tasnet") is models.TasNet and models.get("TasNet") is models.TasNet
    with pytest.raises(ValueError):
        models.get("nope")
    with pytest.raises(ValueError):
        models.register_model(models.TasNet)
    m = models.TasNet(sample_rate=8000)
    conf = m.serialize()
    assert set(conf) == {"model_name", "state_dict", "model_args", "infos"} and conf["model_name"] == "TasNet"
    path = tmp_path / "best_model.pth"
    torch.save(conf, path)
    m2 = models.TasNet.from_pretrain(str(path), sample_rate=8000)
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)


def test_unsupported_configs_fail_loudly():
    from audio_only_speech_separation_b200.models import TasNet

    with pytest.raises(NotImplementedError):
        TasNet(module="TCN")
    with pytest.raises(NotImplementedError):
        TasNet(module="GC_TCN", group_size=16)
    with pytest.raises(AssertionError):
        TasNet(module="bogus")


def test_no_cpu_fallback():
    from audio_only_speech_separation_b200 import ops
    from audio_only_speech_separation_b200.losses import PITLossWrapper, pairwise_neg_snr
    from audio_only_speech_separation_b200.models import TasNet

    with pytest.raises(_lib.DualPathError):
        TasNet()(torch.randn(1, 800))
    with pytest.raises(_lib.DualPathError):
        ops.split_feature(torch.randn(1, 2, 300), 100)
    with pytest.raises(_lib.DualPathError):
        PITLossWrapper(pairwise_neg_snr)(torch.randn(2, 2, 100), torch.randn(2, 2, 100))


def test_loss_argument_errors_match_reference():
    from audio_only_speech_separation_b200.losses import PITLossWrapper, PairwiseNegSDR, pairwise_neg_sisdr

    with pytest.raises(ValueError):
        PITLossWrapper(pairwise_neg_sisdr, pit_from="bogus")
    with pytest.raises(TypeError):
        pairwise_neg_sisdr(torch.randn(2, 2, 100), torch.randn(2, 2, 99))
    with pytest.raises(TypeError):
        pairwise_neg_sisdr(torch.randn(2, 100), torch.randn(2, 100))
    with pytest.raises(AssertionError):
        PairwiseNegSDR("nope")
    assert PITLossWrapper(pairwise_neg_sisdr).threshold_byloss is True  # reference default


def test_shard_range_partitions_batch():
    from audio_only_speech_separation_b200.parallel import shard_range

    for n in (1, 7, 16, 33):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_torch_library_operators_registered():
    """dualpath:: custom operators: schemas exist, fake (meta) shapes follow the reference, CUDA dispatch key only."""
    from torch._subclasses.fake_tensor import FakeTensorMode

    from audio_only_speech_separation_b200 import torch_ops

    for name in torch_ops.OP_NAMES:
        assert hasattr(torch.ops.dualpath, name), name
    assert "Tensor x, SymInt block_size" in str(torch.ops.dualpath.segment.default._schema)
    with FakeTensorMode():
        x = torch.empty(2, 64, 4002, device="cuda")
        y = torch.ops.dualpath.segment(x, 100)
        assert tuple(y.shape) == (2, 64, 100, O.num_chunks(4002, 100))
        assert tuple(torch.ops.dualpath.overlap_add(y, O.seg_rest(4002, 100)).shape) == (2, 64, 4002)
        o, lse = torch.ops.dualpath.attention(torch.empty(1, 82, 100, 192, device="cuda"), 4, "intra")
        assert tuple(o.shape) == (1, 82, 100, 64) and tuple(lse.shape) == (8200, 4)
        loss, perm, _ = torch.ops.dualpath.pit_sdr_loss(torch.empty(4, 2, 800, device="cuda"), torch.empty(4, 2, 800, device="cuda"), "snr", False)
        assert loss.ndim == 0 and tuple(perm.shape) == (4,)
    with pytest.raises(NotImplementedError):   # no CPU kernel is registered
        torch.ops.dualpath.segment(torch.zeros(1, 4, 100), 10)
