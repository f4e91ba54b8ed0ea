"""Training shell: scheduler / early-stop semantics on CPU, a short end-to-end run on the GPU."""
import os

import pytest
import torch

CONFIG = {   # configs/dprnn_wsj0.yml of the reference, model shrunk to 1 layer for the test
    "audionet": {"audionet_name": "TasNet", "audionet_config": dict(enc_dim=64, bn_dim=64, hidden_dim=128, win=16, layer=1, num_spk=2,
                                                                     module="DPRNN", group_size=1, block_size=100, unfold=False)},
    "loss": {"train": {"loss_func": "PITLossWrapper", "sdr_type": "pairwise_neg_snr", "config": {"pit_from": "pw_mtx", "threshold_byloss": False}},
             "val": {"loss_func": "PITLossWrapper", "sdr_type": "pairwise_neg_sisdr", "config": {"pit_from": "pw_mtx", "threshold_byloss": False}}},
    "training": {"epochs": 500, "early_stop": {"monitor": "val_loss/dataloader_idx_0", "mode": "min", "patience": 30, "verbose": True}},
    "optimizer": {"optim_name": "adam", "lr": 0.001, "weight_decay": 0},
    "scheduler": {"sche_name": "ReduceLROnPlateau", "sche_config": {"patience": 15, "factor": 0.5}},
    "datamodule": {"data_name": "LRS2DataModule", "data_config": {"sample_rate": 8000, "segment": 4.0, "batch_size": 2}},
}


def test_plateau_scheduler_matches_torch():
    from audio_only_speech_separation_b200.fit import PlateauScheduler

    class T:
        lr = 1e-3

    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=1e-3)
    ref = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, patience=2, factor=0.5)
    mine = PlateauScheduler(T(), patience=2, factor=0.5)
    g = torch.Generator().manual_seed(0)
    metrics = [1.0, 0.9, 0.95, 0.91, 0.92, 0.93, 0.5, 0.50001, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1] + torch.rand(40, generator=g).tolist()
    for m in metrics:
        ref.step(m)
        assert mine.step(m) == pytest.approx(opt.param_groups[0]["lr"], rel=1e-12), m


def test_early_stopping_and_config_lookup():
    from audio_only_speech_separation_b200.fit import EarlyStopping, build_from_config

    es = EarlyStopping(patience=3)
    assert [es.step(v) for v in [1.0, 0.9, 0.95, 0.93, 0.8, 0.81, 0.82, 0.83]] == [False, False, False, False, False, False, False, True]
    model, tr, va = build_from_config(CONFIG)
    assert model.model_name == "DPRNN" and model.sample_rate() == 8000
    assert tr.loss_func.sdr_type == "snr" and va.loss_func.sdr_type == "sisdr" and tr.threshold_byloss is False
    bad = {**CONFIG, "audionet": {"audionet_name": "NoSuchNet", "audionet_config": {}}}
    with pytest.raises(ValueError):
        build_from_config(bad)


@pytest.mark.gpu
def test_fit_runs_saves_best_checkpoint_and_lowers_the_loss(tmp_path):
    from audio_only_speech_separation_b200.fit import fit
    from audio_only_speech_separation_b200.models import BaseModel

    g = torch.Generator().manual_seed(0)
    src = torch.randn(8, 2, 4000, generator=g) * 0.1
    data = [(src[i:i + 2].sum(1), src[i:i + 2], [f"u{i}", f"u{i + 1}"]) for i in range(0, 8, 2)]
    hist = fit(CONFIG, lambda e: data[:3], lambda e: data[3:], str(tmp_path), max_epochs=6, log=lambda s: None)
    assert len(hist) == 6 and hist[-1]["train_loss"] < hist[0]["train_loss"]
    assert all(h["lr"] == 1e-3 for h in hist)
    for f in ("best_model.pth", "conf.yml", "history.json"):
        assert os.path.exists(tmp_path / f), f
    import glob

    assert len(glob.glob(str(tmp_path / "tensorboard_logs" / "events.out.tfevents.*"))) == 1   # audio_train.py:115-117
    m = BaseModel.from_pretrain(str(tmp_path / "best_model.pth"), sample_rate=8000, **CONFIG["audionet"]["audionet_config"]).cuda().eval()
    with torch.no_grad():
        assert m(data[0][0].cuda()).shape == (2, 2, 4000)


@pytest.mark.gpu
def test_end_to_end_wav_corpus_to_metrics(tmp_path):
    """configs/dprnn_wsj0.yml end to end in this image: WAV lists -> PinnedLoader -> fit() -> best_model.pth -> evaluate() -> metrics.csv."""
    from test_data import _make_corpus

    from audio_only_speech_separation_b200.data import make_loaders
    from audio_only_speech_separation_b200.fit import fit
    from audio_only_speech_separation_b200.metrics import MetricsTracker, evaluate
    from audio_only_speech_separation_b200.models import BaseModel

    corpus = str(tmp_path / "wav")
    _make_corpus(corpus, [6000, 7000, 5000, 8000, 6500, 9000])
    cfg = {**CONFIG, "datamodule": {"data_name": "LRS2DataModule", "data_config": dict(
        train_dir=corpus, valid_dir=corpus, test_dir=corpus, n_src=2, sample_rate=8000, fps=25, segment=0.5, normalize_audio=False, batch_size=2,
        num_workers=2, pin_memory=True, persistent_workers=False, audio_only=True)}}
    train, val, test = make_loaders(cfg["datamodule"]["data_config"])
    exp = str(tmp_path / "exp")
    hist = fit(cfg, lambda e: train, lambda e: val, exp, max_epochs=2, log=lambda s: None)
    assert len(hist) == 2 and all(torch.isfinite(torch.tensor([h["train_loss"], h["val_loss"]])).all() for h in hist)
    model = BaseModel.from_pretrain(os.path.join(exp, "best_model.pth"), sample_rate=8000, **cfg["audionet"]["audionet_config"]).cuda().eval()
    tracker = evaluate(model, (test[i] for i in range(len(test))), MetricsTracker(os.path.join(exp, "metrics.csv")), batch_size=4)
    res = tracker.final()
    assert len(tracker.all_sisnrs) == 6 and all(map(lambda v: v == v, tracker.all_sisnrs_i)) and "si-snr_i" in res
    assert os.path.exists(os.path.join(exp, "metrics.csv"))


@pytest.mark.gpu
def test_fit_runs_sepformer_config(tmp_path):
    """configs/sepformer_base.yml through fit(): the fused DualPathTrainer step (flat gradient buffer, clip + Adam) on the SepFormer engine,
    shrunk to one block / one layer per path; dropout is active in train() mode like in the reference."""
    from audio_only_speech_separation_b200.fit import fit
    from audio_only_speech_separation_b200.models import BaseModel

    net = dict(encoder_kernel_size=16, encoder_in_nchannels=1, encoder_out_nchannels=256, masknet_chunksize=250, masknet_numlayers=1,
               masknet_norm="gLN", masknet_numspks=2, intra_numlayers=1, inter_numlayers=1, intra_nhead=8, inter_nhead=8, intra_dffn=1024,
               inter_dffn=1024, intra_use_positional=True, inter_use_positional=True, intra_norm_before=True, inter_norm_before=True,
               intra_causal=False, inter_causal=False)
    cfg = {**CONFIG, "audionet": {"audionet_name": "Sepformer", "audionet_config": net},
           "loss": {"train": {"loss_func": "PITLossWrapper", "sdr_type": "pairwise_neg_snr", "config": {"pit_from": "pw_mtx", "threshold_byloss": True}},
                    "val": {"loss_func": "PITLossWrapper", "sdr_type": "pairwise_neg_sisdr", "config": {"pit_from": "pw_mtx", "threshold_byloss": False}}},
           "optimizer": {"optim_name": "adam", "lr": 0.00015, "weight_decay": 0}}
    g = torch.Generator().manual_seed(0)
    src = torch.randn(6, 2, 8000, generator=g) * 0.1
    data = [(src[i:i + 1].sum(1), src[i:i + 1], [f"u{i}"]) for i in range(6)]     # one utterance per batch, as the reference trains it
    hist = fit(cfg, lambda e: data[:4], lambda e: data[4:], str(tmp_path), max_epochs=4, log=lambda s: None)
    assert len(hist) == 4 and all(h["train_loss"] == h["train_loss"] and h["val_loss"] == h["val_loss"] for h in hist)
    assert hist[-1]["train_loss"] < hist[0]["train_loss"]
    m = BaseModel.from_pretrain(str(tmp_path / "best_model.pth"), sample_rate=8000, **net).cuda().eval()
    with torch.no_grad():
        assert m(data[0][0].cuda()).shape == (1, 2, 8000)
