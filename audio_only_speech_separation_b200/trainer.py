"""Fused training step with the semantics of the reference's Lightning step.

Reference: ``AudioLightningModule.training_step`` (look2hear/system/audio_litmodule.py:73-88: forward, PIT loss),
Lightning's ``backward`` + ``DDPStrategy`` gradient all-reduce (audio_train.py:120-132), ``gradient_clip_val=5.0``
(audio_train.py:128) and ``torch.optim.Adam(lr=1e-3, weight_decay=0)`` (audio_train.py:48, optimizers.py:58-75).

Everything between the host->device copy of the batch and the loss scalar is engine calls on one stream:
forward, fused PIT loss (+ its analytic backward), engine backward into ONE flat gradient buffer, one NCCL all-reduce of
that buffer when a process group is given (one process per GPU, batch sharded by utterance), and a fused
clip + Adam over the flat parameter buffer.  Like the reference's DDP, the loss mean (and ``threshold_byloss``
filtering) is per rank and the gradients are averaged over ranks.
"""
from __future__ import annotations

import torch

from ._lib import check, lib, ptr, stream_ptr
from .losses.matrix import PairwiseNegSDR, pit_sdr_forward
from .losses.pit_wrapper import PITLossWrapper


class DualPathTrainer:
    def __init__(self, model, loss: PITLossWrapper, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_norm=5.0,
                 process_group=None, distributed=False, cuda_graph=False):
        if not isinstance(loss, PITLossWrapper) or not isinstance(loss.loss_func, PairwiseNegSDR) or loss.pit_from != "pw_mtx":
            raise NotImplementedError("DualPathTrainer needs PITLossWrapper(PairwiseNegSDR(...), pit_from='pw_mtx')")
        if not (hasattr(model, "_train_forward") and hasattr(model, "_train_backward")):
            raise NotImplementedError(f"DualPathTrainer: {type(model).__name__} does not expose the fused training interface "
                                      "(_train_forward / _train_backward into a flat gradient buffer)")
        if weight_decay != 0.0 and getattr(model, "flat_has_buffers", False):
            raise NotImplementedError("weight_decay != 0 with non-parameter tensors (Sepformer's pe buffers) in the flat buffer")
        self.model, self.loss = model, loss
        self.lr, self.betas, self.eps, self.weight_decay, self.max_norm = lr, betas, eps, weight_decay, max_norm
        self.group, self.distributed = process_group, distributed
        self.step_count = 0
        self._state_for = None
        self._ws = None
        self.launches_per_step = 0
        # cuda_graph: from the third step with the same batch shape the step is replayed as two CUDA graphs (pack + forward + loss + backward:
        # ~265 launches; clip + Adam) with the NCCL all-reduce between them; learning rate and bias corrections reach the Adam kernel through
        # device memory (`lr` may change between steps; `betas`, `eps`, `weight_decay`, `max_norm` are fixed at capture: change them -> clear
        # `self._graphs`)
        self.cuda_graph = bool(cuda_graph) and bool(getattr(model, "graph_safe_training", False))
        self._graphs, self._graph_seen = {}, {}
        self._hyper_dev = None

    def _ensure_state(self, device):
        m = self.model
        m._sync_flat(device)
        if self._state_for is not m._flat:
            n = m._flat.numel()
            self.exp_avg = torch.zeros(n, device=device)
            self.exp_avg_sq = torch.zeros(n, device=device)
            self.gflat = torch.zeros(n, device=device)
            self.norm2 = torch.zeros(1, device=device, dtype=torch.float64)
            self._state_for = m._flat
            if self.distributed:
                # DDPStrategy broadcasts rank 0's parameters and buffers when it wraps the model (audio_train.py:126): replicas must not
                # depend on every rank having drawn the same initial weights
                import torch.distributed as dist

                dist.broadcast(m._flat, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
                m.mark_params_dirty()

    def world_size(self) -> int:
        if not self.distributed:
            return 1
        import torch.distributed as dist

        return dist.get_world_size(self.group)

    _GRAPH_SLOTS = 4

    def step(self, mixtures: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        """One optimisation step on a device-resident batch; returns the (device, 0-d) loss of this rank."""
        if not self.cuda_graph:
            return self._step(mixtures, targets)
        self._ensure_state(mixtures.device)
        key = (tuple(mixtures.shape), str(getattr(self.model, "precision", "")), self.model._flat.data_ptr(), mixtures.device.index)
        ent = self._graphs.get(key)
        if ent is None:
            seen = self._graph_seen.get(key, 0)
            if seen < 2:   # workspaces, lazily loaded kernels and the NCCL communicator come into being in eager steps
                if len(self._graph_seen) >= 64:
                    self._graph_seen.clear()
                self._graph_seen[key] = seen + 1
                return self._step(mixtures, targets)
            if len(self._graphs) >= self._GRAPH_SLOTS:
                self._graphs.clear()
            if self._hyper_dev is None:
                self._hyper_dev = torch.zeros(3, dtype=torch.float32, device=mixtures.device)
            static_mix, static_tgt = mixtures.clone(), targets.clone()
            torch.cuda.synchronize(mixtures.device)
            # two graphs around the gradient all-reduce, which stays an ordinary NCCL call between the replays; nothing executes while capturing
            g_grad, g_update = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_grad):
                static_loss = self._forward_backward(static_mix, static_tgt)
            with torch.cuda.graph(g_update):
                self._update(1.0 / self.world_size(), True)
            ent = self._graphs[key] = (g_grad, g_update, static_mix, static_tgt, static_loss)
        g_grad, g_update, static_mix, static_tgt, static_loss = ent
        static_mix.copy_(mixtures)
        static_tgt.copy_(targets)
        g_grad.replay()
        self._all_reduce()
        self.step_count += 1
        self._set_hyper(self.step_count)
        g_update.replay()
        self.model.mark_params_dirty()
        return static_loss.clone()

    def _set_hyper(self, step: int):
        """(lr, 1 - beta1^t, 1 - beta2^t) of the coming step into device memory, on the stream, ahead of the replay."""
        check(lib().dp_adam_set_hyper(ptr(self._hyper_dev), float(self.lr), float(self.betas[0]), float(self.betas[1]), step, stream_ptr()),
              "dp_adam_set_hyper")

    def _step(self, mixtures: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        loss = self._forward_backward(mixtures, targets)
        gscale = self._all_reduce()
        self.step_count += 1
        self._update(gscale, False)
        self.model.mark_params_dirty()
        return loss

    def _forward_backward(self, mixtures: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        """Pack (if the weights changed), forward, PIT loss, backward into the flat gradient buffer of this rank."""
        m = self.model
        self._ensure_state(mixtures.device)
        B, T = mixtures.shape
        est, ctx = m._train_forward(mixtures, self._ws)
        self._ws = ctx[0]
        launches = m.last_launches
        loss, _pw, _perm, lws = pit_sdr_forward(est, targets, self.loss.loss_func.sdr_type, self.loss.threshold_byloss)
        d_est = torch.empty_like(est)
        check(lib().dp_pit_loss_backward(ptr(est), ptr(targets), B, T, ptr(lws), 1.0, ptr(d_est), stream_ptr()), "dp_pit_loss_backward")
        launches += 4 + (1 if self.loss.loss_func.sdr_type == "sisdr" else 0)
        self.gflat.zero_()
        m._train_backward(d_est, self.gflat, ctx, B, T)
        launches += m.last_launches
        self.launches_per_step = launches + 2 + m.pack_launches  # + sumsq + adam, and the re-pack of the weights the next forward triggers
        return loss.reshape(())

    def _all_reduce(self) -> float:
        """Sum of the gradients over the ranks (one NCCL call on the flat buffer); returns the scale that turns it into the mean."""
        if not self.distributed:
            return 1.0
        import torch.distributed as dist

        dist.all_reduce(self.gflat, op=dist.ReduceOp.SUM, group=self.group)
        return 1.0 / dist.get_world_size(self.group)

    def _update(self, gscale: float, captured: bool):
        """Fused clip + Adam over the flat buffers.  captured: learning rate and bias corrections come from device memory (written before
        every replay by _set_hyper), not from the launch arguments."""
        m = self.model
        if captured:
            check(
                lib().dp_adam_clip_step_dev(ptr(m._flat), ptr(self.gflat), ptr(self.exp_avg), ptr(self.exp_avg_sq), m._flat.numel(),
                                            ptr(self.norm2), gscale, float(self.max_norm), ptr(self._hyper_dev), float(self.betas[0]),
                                            float(self.betas[1]), float(self.eps), float(self.weight_decay), stream_ptr()),
                "dp_adam_clip_step_dev",
            )
        else:
            check(
                lib().dp_adam_clip_step(ptr(m._flat), ptr(self.gflat), ptr(self.exp_avg), ptr(self.exp_avg_sq), m._flat.numel(),
                                        ptr(self.norm2), gscale, float(self.max_norm), float(self.lr), float(self.betas[0]),
                                        float(self.betas[1]), float(self.eps), self.step_count, float(self.weight_decay), stream_ptr()),
                "dp_adam_clip_step",
            )

    def grad_norm(self) -> torch.Tensor:
        """Total gradient norm of the last step (after the all-reduce average), as clip_grad_norm_ would report it."""
        return torch.sqrt(self.norm2) / self.world_size()
