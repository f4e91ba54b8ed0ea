"""Profiling driver: warm-up steps, then ONE training step of any drop-in model inside cudaProfilerStart/Stop.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
        python profiles/run_train.py <dprnn|dprnn_unfold|dptnet|dptnet_unfold|sepformer> <B> <T> <fp32|bf16>
"""
import os
import runpy
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# reuse the step construction of tests/tools/time_train.py without its timing loop
src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "tools", "time_train.py")).read()
src = src[: src.index("for _ in range(3):")]
ns = {"__name__": "prof", "__file__": os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "tools", "time_train.py")}
exec(compile(src, "time_train_prefix", "exec"), ns)
step = ns["step"]
for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one training step", sys.argv[1:])
