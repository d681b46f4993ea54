"""Embedding-dimension sweep of the greedy score+select: the tcgen05 filter engine and the
exact SIMT engine at D in {8, 16, 32, 64, 128}.  N x M from the command line (default 100000 items x 4096 rows;
`1000000 20480` is the C4 shape).  Prints CUDA-event time per call, logits/s and the contraction TFLOP/s."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pivotcvae_b200 import ops  # noqa: E402

N, M = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (100000, 4096)
ENGINES = sys.argv[3].split(",") if len(sys.argv) > 3 else ["tcgen05", "simt"]
g = torch.Generator(device="cuda").manual_seed(0)
print("# N=%d items x M=%d rows" % (N, M))
for D, engine in [(D, e) for D in (8, 16, 32, 64, 128) for e in ENGINES]:
    W = torch.nn.functional.normalize(torch.rand(N, D, generator=g, device="cuda") * 2 - 1, dim=1)
    Q = torch.randn(M, D, generator=g, device="cuda") * 0.5
    tab = ops.Table(W)
    for _ in range(3):
        ops.score_select(tab, Q, "greedy", engine=engine)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        ops.score_select(tab, Q, "greedy", engine=engine)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print("D=%3d %-8s %.4f ms/call  %.3f T logits/s  %.1f TFLOP/s" % (D, engine, ms, N * M / ms / 1e9, 2.0 * D * N * M / ms / 1e9),
          flush=True)
