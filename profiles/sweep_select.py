"""Shape sweep of the tcgen05 select engine (run under `ncu --metrics gpu__time_duration.sum` to get
per-kernel durations for each (N, M); prints CUDA-event times otherwise)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("PCV_LIB"):          # A/B runs: load another build of the library
    from pivotcvae_b200 import _lib
    _lib.LIB_PATH = os.environ["PCV_LIB"]
from pivotcvae_b200 import ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
shapes = [(50000, m) for m in (1024, 2560, 5120, 10240, 20480, 40960)] + [(1000000, m) for m in (1024, 4096, 20480)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in s.split("x")) for s in sys.argv[1:]]
for n, m in shapes:
    W = torch.nn.functional.normalize(torch.rand(n, 8, generator=g, device="cuda") * 2 - 1, dim=1)
    Q = torch.randn(m, 8, generator=g, device="cuda") * 0.5
    tab = ops.Table(W)
    for _ in range(3):
        ops.score_select(tab, Q, "greedy", engine="tcgen05")
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        ops.score_select(tab, Q, "greedy", engine="tcgen05")
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print("N=%d M=%d  %.4f ms/call  %.3f T logits/s" % (n, m, ms, n * m / ms / 1e9), flush=True)
