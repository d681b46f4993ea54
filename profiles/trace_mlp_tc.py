"""Phase trace (clock64, relative to the CTA's first stamp) of mlp_tc_kernel, CTA 0: separate library built
with -DPCV_TC_TRACE under profiles/_trace/.  Role 0 = first transform warp, role 1 = MMA warp.  Event numbers: see the
MT_TRACE calls in csrc/mlp_tc.cu (block b at 26 b: 0 start | 1 buffers acquired | 2 prologue loads done | 3 prologue
published | per layer l at 4 + 6 l: accumulators ready, first chunk read, activated, buffer acquired, published, layer
done | 22 block-end sync | 23 reparam | 24 sync 2; 60 end; MMA warp: per layer first operand chunk seen, first weights
seen, first chunk issued, last chunk issued; 61 entry, 62 TMEM allocated, 63 before the final cluster barrier)."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pivotcvae_b200 import build as B  # noqa: E402

out = os.path.join(ROOT, "profiles", "_trace")
lib = os.path.join(out, "libpcv_b200_trace.so")
if not os.path.exists(lib) or "--rebuild" in sys.argv:
    os.makedirs(out, exist_ok=True)
    srcs = [os.path.join(B.CSRC, f) for f in B.SOURCES]
    subprocess.run([B._nvcc()] + B.NVCC_FLAGS + ["-DPCV_TC_TRACE", "-shared", "-o", lib] + srcs, check=True)
if "--build-only" in sys.argv:
    sys.exit(0)

import torch  # noqa: E402
from pivotcvae_b200 import _lib  # noqa: E402
_lib.LIB_PATH = lib
from pivotcvae_b200 import ops  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "profiles"))
import probe_kernels as pk  # noqa: E402

Bn = 1024
env, model, users, ctx = pk._models(100000, 100000, 5, Bn)
items = torch.randint(0, 100000, (Bn, 5), device="cuda")
pivot = torch.randint(0, 100000, (Bn,), device="cuda")
_lib.load().pcv_debug_mlp_tc_trace.argtypes = [ctypes.c_void_p]
with torch.no_grad(), ops.mlp_engine("tc"):
    r, u, _ = model._inputs(ctx, users)
    _, z, _ = model._prior_chain(r, u, model.psmMLP)
    blocks = {"prior->z->PSM chain": lambda: model._prior_chain(r, u, model.psmMLP),
              "SCM": lambda: model._scm(z, ("onehot", r), pivot, model._user_seg(u), []),
              "response MLP": lambda: env(items, users)}
    for name, fn in blocks.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t = (ctypes.c_longlong * (2 * 64))()
        _lib.load().pcv_debug_mlp_tc_trace(t)
        t = list(t)
        print("==", name)
        t0 = t[64 + 61]
        for role in range(2):
            row = t[role * 64:(role + 1) * 64]
            print(" %s: %s" % ("xform" if role == 0 else "mma  ", " ".join("%d:%d" % (i, row[i] - t0) for i in range(64) if row[i] >= t0)), flush=True)
