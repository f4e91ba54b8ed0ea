"""numpy replica of the counter-based dropout masks of the CUDA engines (csrc/common.cuh: hash32 / drop_site_key / drop_keep) and the
mask callback the SepFormer oracle takes.  Test infrastructure."""
import numpy as np
import torch

M32 = 0xFFFFFFFF


def hash32(x):
    x = np.asarray(x, dtype=np.uint64) & M32
    x ^= x >> 16
    x = (x * 0x7FEB352D) & M32
    x ^= x >> 15
    x = (x * 0x846CA68B) & M32
    x ^= x >> 16
    return x


def site_key(seed, layer, site):
    return int(hash32(np.uint64((seed & M32) ^ ((0x9E3779B1 * (layer * 4 + site + 1)) & M32))))


def keep(key, rows, cols, p):
    """rows [R], cols [C] (non-negative ints) -> bool [R, C]."""
    thr = int(np.float32(p) * np.float32(16777216.0))
    rows = np.asarray(rows, dtype=np.uint64)[:, None]
    cols = np.asarray(cols, dtype=np.uint64)[None, :]
    x = (rows * 0x9E3779B1 + cols * 0x85EBCA77 + key) & M32
    return (hash32(x) >> 8) >= thr


def oracle_dropout(p, seed, B, S, K, heads_intra, heads_inter, layers_intra, layers_inter):
    """Callback for ``sepformer_oracle.sepformer_forward(dropout=...)`` reproducing the engine's masks.  Layers are numbered in
    execution order (block 0 intra, block 0 inter, block 1 intra, ...); element rows are stream positions ``(b*S + s)*K + k``."""
    scale = 1.0 / (1.0 - float(np.float32(p)))

    def positions(path):
        if path == 0:   # rows (b, s), time k
            return np.arange(B * S * K).reshape(B * S, K)
        b, k, s = np.meshgrid(np.arange(B), np.arange(K), np.arange(S), indexing="ij")   # rows (b, k), time s
        return ((b * S + s) * K + k).reshape(B * K, S)

    def cb(block, path, layer, site, t):
        li = block * (layers_intra + layers_inter) + (layers_intra if path else 0) + layer
        key = site_key(seed, li, site)
        pos = positions(path)
        if site == 0:   # [Nb, h, L, L]: row = position of the query * heads + head, column = key index
            Nb, H, L, _ = t.shape
            rows = (pos[:, None, :] * H + np.arange(H)[None, :, None]).reshape(-1)
            m = keep(key, rows, np.arange(L), p).reshape(Nb, H, L, L)
        else:           # [Nb, L, C]: row = position, column = channel
            Nb, L, C = t.shape
            m = keep(key, pos.reshape(-1), np.arange(C), p).reshape(Nb, L, C)
        return t * torch.from_numpy(m.astype(np.float32)).to(t.dtype) * scale

    return cb
