"""CUDA-event times of the tf32 and f16 select filters at the C5 catalog size (10 M items, 32 MB column chunks)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pivotcvae_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
W = torch.nn.functional.normalize(torch.randn(10_000_000, 8, generator=g, device="cuda"), dim=1)
tab = ops.Table(W)
for M in (16384, 65536):
    Q = torch.randn(M, 8, generator=g, device="cuda") * 0.5
    for eng in ("tcgen05", "tcgen05_f16"):
        ops.score_select(tab, Q, "greedy", engine=eng)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        ops.score_select(tab, Q, "greedy", engine=eng)
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e)
        print("10M x %d %s: %.2f ms  %.2f T logits/s" % (M, eng, ms, 1e7 * M / ms / 1e9), flush=True)
