"""Phase trace (clock64) of the cluster MLP kernel, cluster 0, thread 0 of each rank; separate library
built with -DPCV_TC_TRACE under profiles/_trace/.  Events: 0 entry | 1 roles split | 2 prologue done |
3+2l layer l computed | 4+2l layer l published (cluster barrier passed) | 20 block done."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pivotcvae_b200 import build as B  # noqa: E402

out = os.path.join(ROOT, "profiles", "_trace")
lib = os.path.join(out, "libpcv_b200_trace.so")
if not os.path.exists(lib):
    os.makedirs(out, exist_ok=True)
    srcs = [os.path.join(B.CSRC, f) for f in B.SOURCES]
    subprocess.run([B._nvcc()] + B.NVCC_FLAGS + ["-DPCV_TC_TRACE", "-shared", "-o", lib] + srcs, check=True)
if "--build-only" in sys.argv:
    sys.exit(0)

import torch  # noqa: E402
from pivotcvae_b200 import _lib  # noqa: E402
_lib.LIB_PATH = lib
from pivotcvae_b200 import ops  # noqa: E402

Bn = 1024
g = torch.Generator(device="cuda").manual_seed(0)
for dims in ([27, 256, 256, 8], [80, 256, 256, 10], [35, 256, 256, 72]):
    x = torch.randn(Bn, dims[0], generator=g, device="cuda")
    layers = [(torch.randn(dims[i + 1], dims[i], generator=g, device="cuda") * 0.1,
               torch.zeros(dims[i + 1], device="cuda"), 1 if i < len(dims) - 2 else 0) for i in range(len(dims) - 1)]
    for _ in range(3):
        ops.mlp_forward([ops.Dense(x)], layers, Bn)
    torch.cuda.synchronize()
    t = (ctypes.c_longlong * 128)()
    _lib.load().pcv_debug_mlp_trace(t)
    t = list(t)
    for r in range(4):
        row = t[32 * r:32 * r + 32]
        ev = [0, 1, 2] + [3 + i for i in range(2 * (len(dims) - 1))] + [20]
        print(dims, "rank", r, " ".join("%d:%d" % (i, row[i] - row[0]) for i in ev), flush=True)
