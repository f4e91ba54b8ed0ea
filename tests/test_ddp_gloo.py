"""CPU: world_size-2 gloo run of the data-parallel host logic (shard by utterance, one flat-gradient all-reduce)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dualpath_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from audio_only_speech_separation_b200.parallel import allreduce_mean_, init_from_env, max_over_ranks, shard_range

    r, w, _ = init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    # a tiny "model" (the PIT loss on a linear map of the input) so per-rank gradients differ
    g = torch.Generator().manual_seed(5)
    mix = torch.randn(6, 2, 400, generator=g)
    tgt = torch.randn(6, 2, 400, generator=g)
    wgt = torch.full((2,), 0.7, requires_grad=True)
    lo, hi = shard_range(6, rank, world)
    loss = O.pit_loss(mix[lo:hi] * wgt.view(1, 2, 1), tgt[lo:hi], "snr", False)
    loss.backward()
    flat = wgt.grad.clone()
    allreduce_mean_(flat)
    slow = max_over_ranks(float(rank + 1))
    if rank == 0:
        out.put((flat.tolist(), slow))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_average_equals_global_batch():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    flat, slow = out.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # equal shards: mean of per-rank means == global-batch mean (the reference's DDP semantics)
    g = torch.Generator().manual_seed(5)
    mix = torch.randn(6, 2, 400, generator=g)
    tgt = torch.randn(6, 2, 400, generator=g)
    wgt = torch.full((2,), 0.7, requires_grad=True)
    O.pit_loss(mix * wgt.view(1, 2, 1), tgt, "snr", False).backward()
    assert torch.allclose(torch.tensor(flat), wgt.grad, rtol=1e-5, atol=1e-7)
    assert slow == 2.0
