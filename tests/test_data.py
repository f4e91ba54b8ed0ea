"""Data feed: LRS2Dataset semantics against the reference's code path (replayed with numpy), PinnedLoader batching."""
import json
import os

import numpy as np
import pytest
import torch
from scipy.io import wavfile


def _make_corpus(root, lengths, sr=8000, seed=0):
    rng = np.random.default_rng(seed)
    os.makedirs(root, exist_ok=True)
    lists = {"mix": [], "s1": [], "s2": []}
    waves = []
    for i, n in enumerate(lengths):
        s = (rng.standard_normal((2, n)) * 3000).astype(np.int16)
        mix = (s[0].astype(np.int32) + s[1]).clip(-32768, 32767).astype(np.int16)
        for name, w in (("mix", mix), ("s1", s[0]), ("s2", s[1])):
            path = os.path.join(root, f"{name}_{i}.wav")
            wavfile.write(path, sr, w)
            lists[name].append([path, int(n)])
        waves.append((mix, s))
    for name, lst in lists.items():
        json.dump(lst, open(os.path.join(root, name + ".json"), "w"))
    return waves


def test_dataset_crops_like_the_reference(tmp_path):
    from audio_only_speech_separation_b200.data import LRS2Dataset, normalize_tensor_wav

    lengths = [9000, 3000, 8000, 12000]           # 3000 < 1 s segment * 8000 -> dropped (lrs2datamodule.py:110-117)
    waves = _make_corpus(str(tmp_path), lengths)
    ds = LRS2Dataset(str(tmp_path), segment=1.0, sample_rate=8000)
    assert len(ds) == 3 and ds.seg_len == 8000
    kept = [0, 2, 3]
    np.random.seed(123)
    got = [ds[i] for i in range(3)]
    np.random.seed(123)                            # replay the reference's draws: randint only when the utterance is longer than the segment
    for j, i in enumerate(kept):
        n = lengths[i]
        start = 0 if n == 8000 else np.random.randint(0, n - 8000)
        mix, s = waves[i]
        m, src, key = got[j]
        assert key == f"mix_{i}.wav" and m.dtype == torch.float32 and tuple(src.shape) == (2, 8000)
        assert np.array_equal(m.numpy(), mix[start:start + 8000].astype(np.float32) / 32768.0)       # soundfile's float32 scaling of PCM16
        assert np.array_equal(src.numpy(), s[:, start:start + 8000].astype(np.float32) / 32768.0)
    # test mode: whole utterances, nothing dropped; normalisation by the mixture's std (lrs2datamodule.py:186-189)
    dt = LRS2Dataset(str(tmp_path), segment=None, normalize_audio=True)
    assert len(dt) == 4
    m, src, _ = dt[1]
    raw = torch.from_numpy(waves[1][0].astype(np.float32) / 32768.0)
    assert torch.allclose(m, normalize_tensor_wav(raw, std=raw.std(-1, keepdim=True)))
    assert m.shape[-1] == 3000 and src.shape == (2, 3000)
    with pytest.raises(NotImplementedError):
        LRS2Dataset(str(tmp_path), n_src=1)


def test_pinned_loader_batches_shuffles_and_drops_last(tmp_path):
    from audio_only_speech_separation_b200.data import LRS2Dataset, PinnedLoader, make_loaders

    _make_corpus(str(tmp_path), [8000 + 100 * i for i in range(11)])
    ds = LRS2Dataset(str(tmp_path), segment=0.5)
    ld = PinnedLoader(ds, batch_size=4, shuffle=False, drop_last=True, workers=3, prefetch=2)
    batches = list(ld)
    assert len(ld) == 2 and len(batches) == 2
    assert [k for b in batches for k in b[2]] == [f"mix_{i}.wav" for i in range(8)]          # order kept, last partial batch dropped
    assert tuple(batches[0][0].shape) == (4, 4000) and tuple(batches[0][1].shape) == (4, 2, 4000)
    keep = PinnedLoader(ds, batch_size=4, shuffle=False, drop_last=False)
    assert len(list(keep)) == 3 == len(keep)
    a = [k for b in PinnedLoader(ds, 4, shuffle=True, seed=1) for k in b[2]]
    b = [k for b in PinnedLoader(ds, 4, shuffle=True, seed=1) for k in b[2]]
    assert a == b and sorted(a) != a and len(set(a)) == 8
    # whole utterances of different lengths cannot share a batch
    with pytest.raises(ValueError):
        list(PinnedLoader(LRS2Dataset(str(tmp_path), segment=None), batch_size=2))
    tr, va, te = make_loaders(dict(train_dir=str(tmp_path), valid_dir=str(tmp_path), test_dir=str(tmp_path), n_src=2, sample_rate=8000,
                                   segment=0.5, batch_size=2, num_workers=2, pin_memory=False))
    assert len(tr) == 5 and len(va) == 5 and len(te) == 11 and te.test
