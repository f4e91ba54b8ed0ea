#!/usr/bin/env python
"""Headline benchmark: DPRNN-wsj0 training step (configs/dprnn_wsj0.yml), batch 16 utterances of 4 s @ 8 kHz per GPU.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--precision fp32|bf16]

One "step" = forward + PIT neg-SNR loss + backward + (N>1: one NCCL all-reduce of the flat gradient buffer) +
clip_grad_norm_(5.0) + Adam(1e-3) on synthetic data, random-init weights of the named architecture.
* ``value``  : training samples/s over all ranks, inputs already resident in HBM, CUDA-event timed, max over ranks.
* ``e2e``    : same metric through the public trainer call with HOST (pinned) inputs: per step H2D copy of the
               mixtures and targets and a D2H read of the loss scalar inside the timed region.
* ``roofline``: the dominant kernel (persistent LSTM BPTT recurrence; the forward member reported beside it), timed alone
               with CUDA events in this process.
* ``cpu_baseline`` / ``--impl reference``: the oracle port of the reference (torch CPU, all host threads) on a bounded
  sample of the same workload.  The reference itself is pure Python and cannot travel to the GPU box.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SR, SECONDS, BATCH = 8000, 4.0, 16
CFG = dict(enc_dim=64, bn_dim=64, hidden_dim=128, win=16, layer=6, num_spk=2, module="DPRNN", group_size=1, block_size=100, unfold=False)
# dram__bytes_read.sum + dram__bytes_write.sum of one lstm_fwd_kernel launch (ncu --set full capture under profiles/), by batch
TRAFFIC_BYTES_PER_LAUNCH = {16: 1294316032}      # lstm_bwd_ks_kernel, profiles/r1_lstm_bwd_v11_full.summary.txt (dram read + write)
TRAFFIC_FWD_BYTES_PER_LAUNCH = {16: 1287374592}  # lstm_fwd_pipe_kernel, profiles/r1_lstm_fwd_v11_full.summary.txt
METRIC = "train samples/sec (DPRNN wsj0, batch 16/GPU, 4 s @ 8 kHz, fwd + PIT-SNR loss + bwd + clip + Adam)"


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled every 20 ms through NVML while the timed region runs (nvidia-smi fallback)."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.max_sm, self.reasons, self.stop_flag = index, [], 0, set(), False
        self.nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        self.sm.append(int(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        self.max_sm = max(self.max_sm, int(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)))
        r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        for name, bit in (("hw_slowdown", n.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", n.nvmlClocksThrottleReasonHwThermalSlowdown),
                          ("sw_thermal_slowdown", n.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", n.nvmlClocksThrottleReasonSwPowerCap)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        c = [v.strip() for v in out.split(",")]
        if c and c[0].isdigit():
            self.sm.append(int(c[0]))
            self.max_sm = max(self.max_sm, int(c[1]) if c[1].isdigit() else 0)
            for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
                if len(c) > 2 + i and c[2 + i].lower().startswith("active"):
                    self.reasons.add(name)

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            time.sleep(0.02 if self.nvml is not None else 0.2)

    def summary(self):
        self.stop_flag = True
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm or None, "reasons": sorted(self.reasons),
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def synthetic(batch, seed):
    g = torch.Generator().manual_seed(seed)
    T = int(SR * SECONDS)
    src = torch.randn(batch, 2, T, generator=g) * 0.1
    return src.sum(1).contiguous(), src.contiguous()


# ------------------------------------------------------------------------------------------------ reference / CPU arm
def cpu_reference_step_fn(batch):
    """Oracle port of the reference training step on the host cores (audio_litmodule.py:73-88 + audio_train.py:48,128)."""
    from audio_only_speech_separation_b200.models import TasNet
    from oracle import dualpath_oracle as O

    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    m = TasNet(sample_rate=SR, **CFG)
    params = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    keys = [k for k, _ in m.named_parameters()]
    ea = [torch.zeros_like(params[k]) for k in keys]
    es = [torch.zeros_like(params[k]) for k in keys]
    mix, tgt = synthetic(batch, 1234)
    state = {"step": 0}

    def step():
        state["step"] += 1
        for p in params.values():
            p.grad = None
        loss = O.pit_loss(O.tasnet_forward(params, mix, lstm_impl="aten"), tgt, "snr", False)
        loss.backward()
        with torch.no_grad():
            O.adam_clip_step([params[k] for k in keys], [params[k].grad for k in keys], ea, es, state["step"])
        return loss.item()

    return step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_b = 1
    step = cpu_reference_step_fn(sample_b)
    for _ in range(min(args.warmup, 2)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = sample_b * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 2), "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs/dprnn_wsj0.yml DPRNN training step, 4 s @ 8 kHz", "batch_per_step_sample": sample_b},
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{args.steps} training steps of {sample_b} utterance(s) (oracle port of the reference, torch CPU, "
                                   f"{os.cpu_count()} threads)"},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def time_recurrence(model, B, precision):
    """Average duration of the dominant kernel family (persistent BiLSTM recurrence, intra-chunk pass at the bench shape,
    training mode), each member timed alone with CUDA events on the launching stream: the forward kernel (reads the gate
    pre-activations, writes activated gates, cell states and H) and the BPTT kernel (reads gates, cell states and dH, writes
    d(pre-activations)).  Both move 84 MB per utterance; buffers (537 MB of gates at B = 16) exceed the L2."""
    from audio_only_speech_separation_b200 import _lib, ops

    dev = torch.device("cuda", torch.cuda.current_device())
    S, K = 82, 100
    P = B * S * K
    pack = ops.LstmPack(model.seq_model.seq_model.row_rnn[0].rnn)
    G0 = torch.randn(P, 1024, device=dev) * 0.5
    G = torch.empty_like(G0)
    H = torch.empty(P, 256, device=dev)
    Cst = torch.empty(P, 256, device=dev)
    dH = torch.randn(P, 256, device=dev) * 0.1
    dbias = torch.zeros(1024, device=dev)
    prec = _lib.PREC_FP32 if precision == "fp32" else _lib.PREC_BF16
    L = _lib.lib()
    tf, tb = [], []
    for it in range(6):
        G.copy_(G0)  # every timed launch starts from a cold L2
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        _lib.check(L.dp_lstm_recurrence_f32(_lib.ptr(pack.buf), _lib.ptr(G), _lib.ptr(H), _lib.ptr(Cst), B * S, K, 1 << 30, 0, K, 1, 1, prec,
                                            _lib.stream_ptr()))
        e1.record()
        _lib.check(L.dp_bilstm_backward_f32(_lib.ptr(pack.buf), _lib.ptr(G), _lib.ptr(Cst), _lib.ptr(dH), None, 0, _lib.ptr(dbias), P, B * S, K,
                                            1 << 30, 0, K, 1, prec, _lib.stream_ptr()))
        e2.record()
        torch.cuda.synchronize()
        if it >= 2:
            tf.append(e0.elapsed_time(e1) * 1e-3)
            tb.append(e1.elapsed_time(e2) * 1e-3)
    flops = 2.0 * 512 * 128 * (B * S) * K * 2   # one recurrent product per step, both directions (algorithmic: one product per MAC)
    hbm = P * (1024 + 1024 + 256 + 256) * 4.0   # 84 MB per utterance for either kernel
    return sum(tf) / len(tf), sum(tb) / len(tb), flops, hbm


def run_ours(args):
    from audio_only_speech_separation_b200.losses import PITLossWrapper, pairwise_neg_snr
    from audio_only_speech_separation_b200.models import TasNet
    from audio_only_speech_separation_b200.parallel import init_from_env
    from audio_only_speech_separation_b200.trainer import DualPathTrainer

    rank, world, local = init_from_env("nccl")
    import torch.distributed as dist

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.manual_seed(0)
    model = TasNet(sample_rate=SR, **CFG).to(dev)
    model.precision = args.precision
    model.train()
    trainer = DualPathTrainer(model, PITLossWrapper(pairwise_neg_snr, pit_from="pw_mtx", threshold_byloss=False), lr=1e-3, max_norm=5.0,
                              distributed=world > 1)
    mix_h, tgt_h = synthetic(args.batch, 1234 + rank)
    mix_h, tgt_h = mix_h.pin_memory(), tgt_h.pin_memory()
    mix_d, tgt_d = mix_h.to(dev), tgt_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        sec = e0.elapsed_time(e1) * 1e-3
        if world > 1:
            t = torch.tensor([sec], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        return sec

    def step_device():
        return trainer.step(mix_d, tgt_d)

    mix_s, tgt_s = torch.empty_like(mix_d), torch.empty_like(tgt_d)
    last = {}

    def step_e2e():
        mix_s.copy_(mix_h, non_blocking=True)
        tgt_s.copy_(tgt_h, non_blocking=True)
        last["loss"] = trainer.step(mix_s, tgt_s).item()  # D2H read of the step's result

    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sec = timed(step_device, args.steps)
    clocks = sampler.summary() if rank == 0 else None
    sec_e2e = timed(step_e2e, args.steps)
    value = args.batch * world * args.steps / sec
    e2e = args.batch * world * args.steps / sec_e2e
    # informational: the same step in the other precision mode (never the headline value)
    other = "bf16" if args.precision == "fp32" else "fp32"
    model.precision = other
    for _ in range(3):
        step_device()
    sec_other = timed(step_device, max(args.steps // 2, 2))
    other_value = args.batch * world * max(args.steps // 2, 2) / sec_other
    model.precision = args.precision

    if rank == 0:
        pk, pk_kind = peaks()
        f_sec, k_sec, k_flops, k_hbm = time_recurrence(model, args.batch, args.precision)
        tf = k_flops / k_sec / 1e12
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            stepf = cpu_reference_step_fn(1)
            stepf()
            t0 = time.perf_counter()
            n = 3
            for _ in range(n):
                stepf()
            dt = time.perf_counter() - t0
            cpu = {"value": n / dt, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": f"{n} training steps of 1 utterance (4 s @ 8 kHz) after 1 warm-up; oracle port of the reference, torch CPU"}
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (bf16x3 split tensor-core products, fp32 accumulate/state)" if args.precision == "fp32" else "bf16",
            "data": "synthetic",
            "config": {"workload": "configs/dprnn_wsj0.yml DPRNN training step, 4 s @ 8 kHz (configs[1])", "batch_per_gpu": args.batch,
                       "global_batch": args.batch * world, "samples": int(SR * SECONDS), "parallelism": f"dp{world}",
                       "l2": "per-step working set (~9.6 GB of saved activations at batch 16) is far larger than the 126 MB L2"},
            "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": int(mix_h.numel() * 4 + tgt_h.numel() * 4),
                    "d2h_bytes_per_step": 4, "ms_per_step": 1e3 * sec_e2e / args.steps, "loss": last.get("loss")},
            "gpu_launches": trainer.launches_per_step * args.steps,
            "clocks": clocks,
            "roofline": {"kernel": "lstm_bwd_ks_kernel (persistent BiLSTM BPTT recurrence, intra-chunk pass; largest single share of the step: "
                                   "the recurrence kernels fwd + bwd are ~57% of it)", "bound": "hbm",
                         "achieved": k_hbm / k_sec / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": k_hbm / k_sec / 1e9 / pk["hbm_gbs"],
                         "traffic": TRAFFIC_BYTES_PER_LAUNCH.get(args.batch),
                         "peak_source": f"{pk_kind} HBM copy bandwidth (kernel timed alone)", "ms_per_launch": 1e3 * k_sec,
                         "algorithmic_bytes_per_launch": k_hbm, "tensor_tflops": tf, "tensor_frac": tf / pk["bf16_tflops"],
                         "forward_kernel": {"kernel": "lstm_fwd_pipe_kernel (same pass, forward, training mode)", "ms_per_launch": 1e3 * f_sec,
                                            "achieved": k_hbm / f_sec / 1e9, "frac": k_hbm / f_sec / 1e9 / pk["hbm_gbs"],
                                            "traffic": TRAFFIC_FWD_BYTES_PER_LAUNCH.get(args.batch)},
                         "note": "84 MB per utterance (SURVEY 8d: bwd reads activated gates, c_t and dH and writes d(pre-activations); fwd "
                                 "reads G and writes gates + c_t + H); the kernels are bound by the per-step dependent chain (tensor-core "
                                 "issue, shared-memory operand traffic, MUFU), not by HBM: both fractions reported"},
            "cpu_baseline": cpu,
            "other_precision_mode": {"dtype": other, "value": other_value, "unit": "samples/s",
                                     "note": "informational only: the same training step with single bf16 tensor-core products and "
                                             "tanh.approx gates (forward within 0.05 dB PIT-SI-SNR of the fp32 reference); the headline "
                                             "value above is the fp32-parity mode" if other == "bf16" else "informational only"},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("DUALPATH_PRECISION", "fp32"), choices=["fp32", "bf16"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
