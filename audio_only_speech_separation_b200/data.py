"""Data feed: the audio-only part of look2hear/datas/lrs2datamodule.py:30-184 with a pinned-memory prefetcher.

``LRS2Dataset`` reads the reference's ``mix.json`` / ``s1.json`` / ``s2.json`` lists (``[[wav_path, n_samples], ...]``), drops utterances
shorter than the segment (training), and returns ``(mixture [T], sources [n_src, T], key)`` for a random crop drawn with
``np.random.randint`` exactly like the reference (same stream under the same seed) or the whole utterance (``segment=None``: test).
WAV files are read with ``scipy.io.wavfile`` (memory-mapped, so a crop touches only its samples): PCM16 is scaled by 1/32768 as
``soundfile.read(dtype="float32")`` does, float32 files are used as they are (``soundfile`` itself is not installed in this image).

``PinnedLoader`` batches a dataset (``shuffle`` / ``drop_last`` as the reference's ``DataLoader`` calls: train shuffle=True, drop_last=True)
into pinned host tensors filled by background threads, ``prefetch`` batches ahead, so ``fit`` / ``evaluate`` overlap their non-blocking
host-to-device copies with the previous step.
"""
from __future__ import annotations

import json
import os
import queue
import threading
from typing import Iterator, List, Optional, Tuple

import numpy as np
import torch


def normalize_tensor_wav(wav_tensor, eps=1e-8, std=None):
    """lrs2datamodule.py:24-28."""
    mean = wav_tensor.mean(-1, keepdim=True)
    if std is None:
        std = wav_tensor.std(-1, keepdim=True)
    return (wav_tensor - mean) / (std + eps)


def read_wav(path: str, start: int = 0, stop: Optional[int] = None) -> np.ndarray:
    """``soundfile.read(path, start=start, stop=stop, dtype="float32")`` for mono PCM16 / PCM32 / float32 WAV files."""
    from scipy.io import wavfile

    _, data = wavfile.read(path, mmap=True)
    seg = data[start:stop]
    if seg.ndim > 1:
        seg = seg[:, 0]
    if seg.dtype == np.int16:
        return seg.astype(np.float32) * (1.0 / 32768.0)
    if seg.dtype == np.int32:
        return seg.astype(np.float32) * (1.0 / 2147483648.0)
    return np.asarray(seg, dtype=np.float32)


class LRS2Dataset:
    def __init__(self, json_dir: str = "", n_src: int = 2, sample_rate: int = 8000, fps: int = 25, segment: Optional[float] = 4.0,
                 normalize_audio: bool = False, audio_only: bool = True, log=lambda s: None):
        if json_dir is None:
            raise ValueError("JSON DIR is None!")
        if n_src != 2:
            raise NotImplementedError("the dual-path configs of the reference use n_src: 2")
        if not audio_only:
            raise NotImplementedError("audio-only data feed (audio_only: true in every dual-path config)")
        self.EPS = 1e-8
        self.json_dir, self.sample_rate, self.normalize_audio, self.n_src = json_dir, sample_rate, normalize_audio, n_src
        self.seg_len = None if segment is None else int(segment * sample_rate)
        self.test = self.seg_len is None
        with open(os.path.join(json_dir, "mix.json")) as f:
            mix_infos = json.load(f)
        sources_infos = []
        for name in ("s1", "s2"):
            with open(os.path.join(json_dir, name + ".json")) as f:
                sources_infos.append(json.load(f))
        orig_len, drop_utt, drop_len = len(mix_infos), 0, 0
        if not self.test:
            for i in range(len(mix_infos) - 1, -1, -1):   # go backward (lrs2datamodule.py:110-117)
                if mix_infos[i][1] < self.seg_len:
                    drop_utt += 1
                    drop_len += mix_infos[i][1]
                    del mix_infos[i]
                    for src_inf in sources_infos:
                        del src_inf[i]
        log(f"Drop {drop_utt} utts({drop_len / sample_rate / 3600:.2f} h) from {orig_len} (shorter than {self.seg_len} samples)")
        self.mix, self.sources, self.length = mix_infos, sources_infos, len(mix_infos)

    def __len__(self):
        return self.length

    def __getitem__(self, idx: int):
        """lrs2datamodule.py:161-193 (n_src == 2)."""
        if self.mix[idx][1] == self.seg_len or self.test:
            rand_start = 0
        else:
            rand_start = np.random.randint(0, self.mix[idx][1] - self.seg_len)
        stop = None if self.test else rand_start + self.seg_len
        x = read_wav(self.mix[idx][0], rand_start, stop)
        sources = torch.from_numpy(np.vstack([read_wav(src[idx][0], rand_start, stop) for src in self.sources]))
        mixture = torch.from_numpy(x)
        if self.normalize_audio:
            m_std = mixture.std(-1, keepdim=True)
            mixture = normalize_tensor_wav(mixture, eps=self.EPS, std=m_std)
            sources = normalize_tensor_wav(sources, eps=self.EPS, std=m_std)
        return mixture, sources, self.mix[idx][0].split("/")[-1]


class PinnedLoader:
    """Batches of ``(mixtures [B,T], sources [B,n_src,T], keys)`` in pinned memory, produced ``prefetch`` batches ahead by ``workers``
    threads (file reads and crops release the GIL in numpy / the OS).  Equal-length items only (fixed segments, or batch_size 1 for
    whole test utterances)."""

    def __init__(self, dataset, batch_size: int, shuffle: bool = False, drop_last: bool = True, workers: int = 4, prefetch: int = 4,
                 pin_memory: bool = True, seed: Optional[int] = None, rank: int = 0, world: int = 1):
        """``rank`` / ``world``: data-parallel sharding like Lightning's ``DistributedSampler`` (audio_train.py:126 trains under DDP):
        every rank draws the SAME shuffled order (the seed must be shared, it is advanced once per epoch on every rank), the order is
        padded by wrapping to a multiple of ``world`` and rank ``r`` takes items ``r, r + world, ...`` - an epoch is 1/world of the
        single-process steps and no two ranks see the same item."""
        self.dataset, self.batch_size, self.shuffle, self.drop_last = dataset, batch_size, shuffle, drop_last
        self.workers, self.prefetch = max(1, workers), max(1, prefetch)
        self.pin = pin_memory and torch.cuda.is_available()
        if not (0 <= rank < world):
            raise ValueError(f"bad rank/world {rank}/{world}")
        if world > 1 and shuffle and seed is None:
            raise ValueError("PinnedLoader(world > 1, shuffle=True) needs a seed shared by all ranks")
        self.rank, self.world = rank, world
        self.rng = np.random.default_rng(seed)

    def _n_local(self) -> int:
        return (len(self.dataset) + self.world - 1) // self.world

    def __len__(self):
        n = self._n_local()
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def _batches(self) -> List[List[int]]:
        order = self.rng.permutation(len(self.dataset)) if self.shuffle else np.arange(len(self.dataset))
        if self.world > 1:
            pad = self._n_local() * self.world - len(order)
            if pad:
                order = np.concatenate([order, order[:pad]])
            order = order[self.rank::self.world]
        out = [order[i:i + self.batch_size].tolist() for i in range(0, len(order), self.batch_size)]
        if self.drop_last and out and len(out[-1]) < self.batch_size:
            out.pop()
        return out

    def _collate(self, idxs) -> Tuple[torch.Tensor, torch.Tensor, List[str]]:
        items = [self.dataset[i] for i in idxs]
        T = items[0][0].shape[-1]
        if any(it[0].shape[-1] != T for it in items):
            raise ValueError("PinnedLoader batches equal-length items only (use a fixed segment, or batch_size=1 for whole utterances)")
        mix = torch.empty(len(items), T, dtype=torch.float32, pin_memory=self.pin)
        src = torch.empty(len(items), items[0][1].shape[0], T, dtype=torch.float32, pin_memory=self.pin)
        for j, (m, s, _) in enumerate(items):
            mix[j].copy_(m)
            src[j].copy_(s)
        return mix, src, [it[2] for it in items]

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor, List[str]]]:
        batches = self._batches()
        results: "dict[int, object]" = {}
        cond = threading.Condition()
        todo: "queue.Queue[int]" = queue.Queue()
        state = {"next_out": 0, "stop": False}

        def worker():
            while True:
                try:
                    bi = todo.get_nowait()
                except queue.Empty:
                    return
                with cond:   # stay at most `prefetch` batches ahead of the consumer
                    cond.wait_for(lambda: state["stop"] or bi < state["next_out"] + self.prefetch)
                    if state["stop"]:
                        return
                try:
                    res = self._collate(batches[bi])
                except Exception as exc:  # surfaced to the consumer
                    res = exc
                with cond:
                    results[bi] = res
                    cond.notify_all()

        for bi in range(len(batches)):
            todo.put(bi)
        threads = [threading.Thread(target=worker, daemon=True) for _ in range(min(self.workers, max(1, len(batches))))]
        for t in threads:
            t.start()
        try:
            for bi in range(len(batches)):
                with cond:
                    cond.wait_for(lambda: bi in results)
                    res = results.pop(bi)
                    state["next_out"] = bi + 1
                    cond.notify_all()
                if isinstance(res, Exception):
                    raise res
                yield res
        finally:
            with cond:
                state["stop"] = True
                cond.notify_all()


def make_loaders(data_config: dict, workers: Optional[int] = None, rank: int = 0, world: int = 1, seed: Optional[int] = None):
    """``LRS2DataModule.setup`` + ``train/val/test_dataloader`` (lrs2datamodule.py:300-370) for the ``datamodule.data_config`` block of a
    reference YAML: returns ``(train_loader, val_loader, test_set)``; the test set keeps whole utterances (``segment=None``).
    ``rank`` / ``world`` shard both loaders across data-parallel ranks (``seed`` is the shuffle seed every rank must share)."""
    c = dict(data_config)
    common = dict(n_src=c.get("n_src", 2), sample_rate=c.get("sample_rate", 8000), fps=c.get("fps", 25),
                  normalize_audio=c.get("normalize_audio", False), audio_only=c.get("audio_only", True))
    train = LRS2Dataset(c["train_dir"], segment=c.get("segment", 4.0), **common)
    val = LRS2Dataset(c["valid_dir"], segment=c.get("segment", 4.0), **common)
    test = LRS2Dataset(c["test_dir"], segment=None, **common)
    nw = c.get("num_workers", 4) if workers is None else workers
    bs = c.get("batch_size", 1)
    if world > 1 and seed is None:
        seed = 0
    kw = dict(workers=nw, pin_memory=c.get("pin_memory", True), rank=rank, world=world)
    return (PinnedLoader(train, bs, shuffle=True, drop_last=True, seed=seed, **kw), PinnedLoader(val, bs, shuffle=False, drop_last=True, **kw), test)
