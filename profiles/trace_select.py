"""Phase trace of the tcgen05 filter kernel (CTA 0, clock64): builds a SEPARATE library with
-DPCV_TC_TRACE under profiles/_trace/ (never the shipped one) and prints the cycle offsets of
  0 entry | 1 setup done (barriers, TMEM alloc) | 2 query tile staged | 3 first TMA issued |
  4 first MMA issued | 5+2k / 6+2k epilogue warp 2 sees / releases tile 40+k | 13 segment handed over | 14 TMEM freed
    python profiles/trace_select.py 2048x128 50000x10240"""
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

g = torch.Generator(device="cuda").manual_seed(0)
for shp in [a for a in sys.argv[1:] if "x" in a] or ["2048x128"]:
    n, m = (int(x) for x in shp.split("x"))
    W = torch.nn.functional.normalize(torch.rand(n, 8, generator=g, device="cuda") * 2 - 1, dim=1)
    Q = torch.randn(m, 8, generator=g, device="cuda") * 0.5
    tab = ops.Table(W)
    for _ in range(3):
        ops.score_select(tab, Q, "greedy", engine=os.environ.get("TRACE_ENGINE", "tcgen05"))
    torch.cuda.synchronize()
    t = (ctypes.c_longlong * 16)()
    _lib.load().pcv_debug_tc_trace(t)
    t = list(t)
    print(shp, " ".join("%d:%d" % (i, t[i] - t[0]) for i in range(15)), flush=True)
    mm = (ctypes.c_longlong * 16)()
    _lib.load().pcv_debug_tc_mma(mm)
    mm = list(mm)
    print("   steady state (tiles 40-43, cycles from kernel entry): epilogue warp 2 [tile visible, tile released] "
          + " ".join("[%d,%d]" % (t[5 + 2 * k] - t[0], t[6 + 2 * k] - t[0]) for k in range(4))
          + " | MMA issuer [TMEM buffer free, operands ready] "
          + " ".join("[%d,%d]" % (mm[2 * k] - t[0], mm[2 * k + 1] - t[0]) for k in range(4)), flush=True)
    c = (ctypes.c_longlong * 1024)()
    _lib.load().pcv_debug_tc_cta(c)
    c = list(c)
    G = min(148, max(1, ((m + 127) // 128) * ((n + 255) // 256) // 8))
    cyc = [c[256 + i] - c[i] for i in range(G)]
    g0 = min(c[512:512 + G])
    st = [c[512 + i] - g0 for i in range(G)]
    en = [c[768 + i] - g0 for i in range(G)]
    order = sorted(range(G), key=lambda i: cyc[i])
    print("   per-CTA cycles: min %d (cta %d) median %d max %d (cta %d); start spread %d ns; last end %d ns; "
          "slowest 5: %s" % (cyc[order[0]], order[0], cyc[order[G // 2]], cyc[order[-1]], order[-1], max(st), max(en),
                             [(i, cyc[i]) for i in order[-5:]]), flush=True)
