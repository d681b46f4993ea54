"""Micro-driver for profiling the score+select kernels alone (used under ncu via gpurun).
    python profiles/probe_select.py --n 1000000 --m 20480 --engine tcgen05 --iters 5
Prints CUDA-event time per call and logits/s (not a bench number when run under ncu)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pivotcvae_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1000000)
ap.add_argument("--m", type=int, default=20480)
ap.add_argument("--engine", default="tcgen05")
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
g = torch.Generator(device="cuda").manual_seed(0)
W = torch.nn.functional.normalize(torch.randn(a.n, 8, generator=g, device="cuda"), dim=1)
Q = torch.randn(a.m, 8, generator=g, device="cuda") * 0.5
tab = ops.Table(W)
for _ in range(2):
    ops.score_select(tab, Q, "greedy", engine=a.engine)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(a.iters):
    ops.score_select(tab, Q, "greedy", engine=a.engine)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / a.iters
print("engine=%s N=%d M=%d  %.4f ms/call  %.3f T logits/s" % (a.engine, a.n, a.m, ms, a.n * a.m / ms / 1e9))
