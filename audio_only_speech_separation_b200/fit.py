"""Training shell without Lightning: what ``audio_train.py`` + ``AudioLightningModule`` do around the step.

``fit(config, train_batches, val_batches, exp_dir)`` reads the reference's ``configs/*.yml`` schema unchanged:

* model  : ``models.get(audionet.audionet_name)(sample_rate=..., **audionet.audionet_config)``            audio_train.py:41-44
* losses : ``getattr(losses, loss.{train,val}.loss_func)(getattr(losses, sdr_type), **config)``           audio_train.py:67-76
* step   : ``DualPathTrainer`` (forward + loss + backward + all-reduce + clip 5.0 + Adam fused)           audio_litmodule.py:73-88, audio_train.py:128
* epoch  : mean validation loss (per rank, then the mean over ranks), ``ReduceLROnPlateau`` on it          audio_litmodule.py:99-140, optimizers / scheduler config
* stop   : early stopping (``training.early_stop``: patience, mode min)                                  audio_train.py:95-100
* output : ``exp_dir/best_model.pth`` = ``model.serialize()`` of the best epoch, ``exp_dir/conf.yml``,    audio_train.py:52-55,141-152
           ``exp_dir/history.json`` and TensorBoard scalars under ``exp_dir/tensorboard_logs`` (``train_loss``, ``val_loss``, ``lr``,
           ``learning_rate``, ``val_pit_sisnr`` per epoch: audio_train.py:115-117, audio_litmodule.py:79-141; tb_writer.py)

``train_batches`` / ``val_batches`` are callables returning an iterable of ``(mixtures [B,T], targets [B,n_src,T], keys)`` per epoch,
the batch format of the reference's data modules (datas/lrs2datamodule.py:129-184); host tensors are copied to the device here.
"""
from __future__ import annotations

import json
import math
import os
from typing import Callable, Iterable, Optional

import torch

from . import losses as _losses
from . import models as _models
from .tb_writer import ScalarWriter
from .trainer import DualPathTrainer


class PlateauScheduler:
    """``torch.optim.lr_scheduler.ReduceLROnPlateau(mode="min", threshold=1e-4, threshold_mode="rel", cooldown=0, min_lr=0, eps=1e-8)``
    acting on ``DualPathTrainer.lr`` (the reference builds it with ``patience`` and ``factor`` from the config)."""

    def __init__(self, trainer, patience=10, factor=0.1, threshold=1e-4, min_lr=0.0, eps=1e-8):
        self.trainer, self.patience, self.factor, self.threshold, self.min_lr, self.eps = trainer, patience, factor, threshold, min_lr, eps
        self.best, self.num_bad = math.inf, 0

    def step(self, metric: float):
        if metric < self.best * (1.0 - self.threshold):
            self.best, self.num_bad = metric, 0
        else:
            self.num_bad += 1
        if self.num_bad > self.patience:
            new_lr = max(self.trainer.lr * self.factor, self.min_lr)
            if self.trainer.lr - new_lr > self.eps:
                self.trainer.lr = new_lr
            self.num_bad = 0
        return self.trainer.lr


class EarlyStopping:
    """``pytorch_lightning.callbacks.EarlyStopping(mode="min", min_delta=0)``: stop after ``patience`` epochs without improvement."""

    def __init__(self, patience=30):
        self.patience, self.best, self.wait = patience, math.inf, 0

    def step(self, metric: float) -> bool:
        if metric < self.best:
            self.best, self.wait = metric, 0
            return False
        self.wait += 1
        return self.wait >= self.patience


def build_from_config(config: dict, sample_rate: Optional[int] = None):
    """(model, train_loss, val_loss) from the reference's YAML schema."""
    net = config["audionet"]
    sr = sample_rate if sample_rate is not None else config.get("datamodule", {}).get("data_config", {}).get("sample_rate", 8000)
    model = _models.get(net["audionet_name"])(sample_rate=sr, **net["audionet_config"])
    built = {}
    for split in ("train", "val"):
        lc = config["loss"][split]
        built[split] = getattr(_losses, lc["loss_func"])(getattr(_losses, lc["sdr_type"]), **lc.get("config", {}))
    return model, built["train"], built["val"]


def fit(config: dict, train_batches: Callable[[int], Iterable], val_batches: Callable[[int], Iterable], exp_dir: str, *, device="cuda",
        distributed=False, max_epochs: Optional[int] = None, log: Callable[[str], None] = print, cuda_graph: bool = True):
    """Run the reference's training procedure; returns the per-epoch history.  cuda_graph: batches of a shape seen twice before are
    stepped by CUDA-graph replay (TasNet models; the reference trains on fixed-length segments)."""
    import yaml

    os.makedirs(exp_dir, exist_ok=True)
    with open(os.path.join(exp_dir, "conf.yml"), "w") as f:
        yaml.safe_dump(config, f)
    model, train_loss, val_loss = build_from_config(config)
    model = model.to(device)
    opt = config.get("optimizer", {})
    if str(opt.get("optim_name", "adam")).lower() != "adam":
        raise NotImplementedError("the fused step implements Adam (optimizer.optim_name: adam), what every dual-path config of the reference uses")
    trainer = DualPathTrainer(model, train_loss, lr=float(opt.get("lr", 1e-3)), weight_decay=float(opt.get("weight_decay", 0.0)), max_norm=5.0,
                              distributed=distributed, cuda_graph=cuda_graph)
    sch = config.get("scheduler", {})
    scheduler = None
    if sch.get("sche_name") == "ReduceLROnPlateau":
        scheduler = PlateauScheduler(trainer, **{k: sch.get("sche_config", {})[k] for k in ("patience", "factor") if k in sch.get("sche_config", {})})
    stopper = EarlyStopping(patience=int(config.get("training", {}).get("early_stop", {}).get("patience", 30)))
    epochs = int(config.get("training", {}).get("epochs", 1)) if max_epochs is None else max_epochs
    rank = 0
    if distributed:
        import torch.distributed as dist

        rank = dist.get_rank()
    history, best = [], math.inf
    tb = ScalarWriter(os.path.join(exp_dir, "tensorboard_logs")) if rank == 0 else None
    for epoch in range(epochs):
        model.train()
        tl, n = torch.zeros((), device=device), 0
        for mix, tgt, _ in train_batches(epoch):
            tl = tl + trainer.step(mix.to(device, non_blocking=True), tgt.to(device, non_blocking=True))   # no host sync inside the epoch
            n += 1
        model.eval()
        vl, m = torch.zeros((), device=device), 0
        with torch.no_grad():
            for mix, tgt, _ in val_batches(epoch):
                vl = vl + val_loss(model(mix.to(device, non_blocking=True)), tgt.to(device, non_blocking=True))
                m += 1
        stats = torch.stack([tl / max(n, 1), vl / max(m, 1)])
        if distributed:  # mean over ranks of the per-rank means (audio_litmodule.py:90-92,125-127)
            import torch.distributed as dist

            dist.all_reduce(stats)
            stats /= dist.get_world_size()
        train_l, val_l = (float(x) for x in stats.tolist())
        lr_used = trainer.lr
        if scheduler is not None:
            scheduler.step(val_l)
        history.append({"epoch": epoch, "train_loss": train_l, "val_loss": val_l, "val_pit_sisnr": -val_l, "lr": lr_used})
        if val_l < best and rank == 0:
            best = val_l
            torch.save(model.serialize(), os.path.join(exp_dir, "best_model.pth"))
        best = min(best, val_l)
        if rank == 0:
            log(f"epoch {epoch}: train_loss {train_l:.4f} val_loss {val_l:.4f} lr {lr_used:.2e}")
            with open(os.path.join(exp_dir, "history.json"), "w") as f:
                json.dump(history, f, indent=1)
            for tag, v in (("train_loss", train_l), ("val_loss", val_l), ("lr", lr_used), ("learning_rate", lr_used), ("val_pit_sisnr", -val_l)):
                tb.add_scalar(tag, v, epoch)
            tb.flush()
        if stopper.step(val_l):
            if rank == 0:
                log(f"early stop at epoch {epoch} (no improvement for {stopper.patience} epochs)")
            break
    if tb is not None:
        tb.close()
    return history
