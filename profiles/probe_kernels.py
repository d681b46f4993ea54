"""Micro-driver that runs every shipped kernel of the hot path a few times at a representative size
(used under ncu via gpurun: profiles/capture.sh).  Prints the CUDA-event time of each call, the
algorithmic bytes / logits it moves and the rate that gives (not a bench number when run under ncu).

    python profiles/probe_kernels.py [name ...]      # no names = all of them
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("PCV_LIB"):          # A/B runs: load another build of the library (profiles/_trace/)
    from pivotcvae_b200 import _lib
    _lib.LIB_PATH = os.environ["PCV_LIB"]
from pivotcvae_b200 import ops  # noqa: E402
from pivotcvae_b200.env.response_model import URM_P_MR, UserResponseModel_MLP  # noqa: E402
from pivotcvae_b200.models.pivotcvae import PIVOTCVAE_MODELS  # noqa: E402

DEV = "cuda:0"
G = torch.Generator(device=DEV).manual_seed(0)


def table(n, d=8):
    return torch.nn.functional.normalize(torch.randn(n, d, generator=G, device=DEV), dim=1)


WARM = int(os.environ.get("PROBE_WARM", "2"))     # capture.sh sets 0 / 1 under ncu: one launch per kernel
ITERS = int(os.environ.get("PROBE_ITERS", "3"))


def timed(name, fn, work, unit, iters=None):
    iters = iters or ITERS
    for _ in range(WARM):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    print("%-28s %9.4f ms/call  %10.3f %s" % (name, ms, work / ms / 1e6, unit), flush=True)


def p_select():      # score_select_tc_kernel + tc_refine_kernel at C4
    W, Q = table(1000000), torch.randn(20480, 8, generator=G, device=DEV) * 0.5
    tab = ops.Table(W)
    timed("select_c4 (1M x 20480)", lambda: ops.score_select(tab, Q, "greedy"), 1e6 * 20480 / 1e3, "T logits/s")


def p_select_f16():  # the f16 filter (packed TMEM reads, VIMNMX3.U16x2) at C4 and at the pivot-pick shape
    W = table(1000000)
    tab = ops.Table(W)
    for M in (20480, 4096):
        Q = torch.randn(M, 8, generator=G, device=DEV) * 0.5
        timed("select_c4 tf32 (1M x %d)" % M, lambda: ops.score_select(tab, Q, "greedy", engine="tcgen05"), 1e6 * M / 1e3, "T logits/s")
        timed("select_c4 f16  (1M x %d)" % M, lambda: ops.score_select(tab, Q, "greedy", engine="tcgen05_f16"), 1e6 * M / 1e3, "T logits/s")


def p_select_d128():   # the same filter at D = 128 (16 accumulating MMAs per tile): tensor-pipe evidence
    W, Q = table(1000000, 128), torch.randn(20480, 128, generator=G, device=DEV) * 0.5
    tab = ops.Table(W)
    timed("select D=128 (1M x 20480)", lambda: ops.score_select(tab, Q, "greedy", engine="tcgen05"), 2 * 128 * 1e6 * 20480 / 1e3, "TFLOP/s")


def p_topk():        # exact top-10 over 50k items for 2048 rows (the no-repeat re-selection path)
    W, Q = table(50000), torch.randn(2048, 8, generator=G, device=DEV) * 0.5
    tab = ops.Table(W)
    timed("topk 10 (50k x 2048)", lambda: ops.score_topk(tab, Q, 10), 5e4 * 2048 / 1e3, "T logits/s")


def p_select_c2():
    W, Q = table(50000), torch.randn(10240, 8, generator=G, device=DEV) * 0.5
    tab = ops.Table(W)
    timed("select_c2 (50k x 10240)", lambda: ops.score_select(tab, Q, "greedy"), 5e4 * 10240 / 1e3, "T logits/s")


def p_sampler():     # sigmoid_categorical_kernel (sampled pivot, throughput mode)
    W, Q = table(1000000), torch.randn(65536, 8, generator=G, device=DEV) * 0.5
    tab = ops.Table(W)
    timed("sigmoid_categorical 65536", lambda: ops.sigmoid_categorical(tab, Q, seed=1), 65536 / 1e3, "G rows/s")


def p_exprace():     # score_select_kernel<8, exprace> with caller-supplied noise (parity mode)
    W, Q = table(50000), torch.randn(1024, 8, generator=G, device=DEV) * 0.5
    tab = ops.Table(W)
    noise = torch.empty(1024, 50000, device=DEV).exponential_(generator=G)
    timed("exprace parity 50k x 1024", lambda: ops.score_select(tab, Q, "exprace", noise=noise), 5e4 * 1024 / 1e3, "T logits/s")


def _models(n_items, n_users, Ls, B, no_user=False):
    torch.manual_seed(0)
    D, Z = 8, 16
    u = 0 if no_user else D
    env = UserResponseModel_MLP(n_items - 1, n_users - 1, D, Ls, [Ls * D + u, 256, 256, Ls], DEV, no_user).to(DEV)
    model = PIVOTCVAE_MODELS["pivotcvae_gt_pi"](env.docEmbed, None if no_user else env.userEmbed, Ls, D, Z, Ls + 1,
                                               [Ls * D + Ls + 1 + u, 256, 256], [Z + Ls + 1 + u, 256, 256, D],
                                               [Z + Ls + 1 + D + u, 256, 256, (Ls - 1) * D], [Ls + 1 + u, 128, 128], no_user, DEV)
    users = torch.randint(0, n_users, (B,), generator=G, device=DEV)
    ctx = torch.zeros(B, Ls, device=DEV)
    ctx[:, :2] = 1
    return env, model, users, ctx


def p_mlp():         # mlp_cluster_kernel: prior->z->PSM chain, SCM, response MLP (gather + normalise prologue) at C4
    env, model, users, ctx = _models(100000, 100000, 5, 4096)

    def step():
        items, _ = model.recommend(ctx, users, return_item=True)
        env(items.view(4096, -1), users)
    # per slate: ids 8(L+1) + r 4L + rows gathered (2L+1)*32 + out 8L + 4L + z_mu 64 = 564 B
    timed("recommend+resp B=4096 (100k)", step, 4096 * 564 / 1e3, "TB/s (algorithmic)")


def p_mlp_engines():  # every MLP block of a C4 / C2 step alone, FFMA cluster engine vs the tcgen05 engine (mlp_tc_kernel)
    for B in (512, 1024, 4096, 16384):
        env, model, users, ctx = _models(100000, 100000, 5, B)
        items = torch.randint(0, 100000, (B, 5), generator=G, device=DEV)
        pivot = torch.randint(0, 100000, (B,), generator=G, device=DEV)
        with torch.no_grad():
            r, u, _ = model._inputs(ctx, users)
            _, z, _ = model._prior_chain(r, u, model.psmMLP)
            blocks = {"prior->z->PSM chain": lambda: model._prior_chain(r, u, model.psmMLP),
                      "SCM": lambda: model._scm(z, ("onehot", r), pivot, model._user_seg(u), []),
                      "response MLP": lambda: env(items, users)}
            flops = {"prior->z->PSM chain": 2 * (14 * 128 + 128 * 128 + 128 * 32 + 30 * 256 + 256 * 256 + 256 * 8),
                     "SCM": 2 * (38 * 256 + 256 * 256 + 256 * 32), "response MLP": 2 * (48 * 256 + 256 * 256 + 256 * 5)}
            for eng in ("exact", "tc"):
                with ops.mlp_engine(eng):
                    for name, fn in blocks.items():
                        # 20 launches per graph replay: the eager call is launch-bound (~60-90 us of Python per call)
                        fn()
                        torch.cuda.synchronize()
                        gr = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(gr):
                            for _ in range(20):
                                fn()
                        timed("%-20s B=%-5d %-5s x20" % (name, B, eng), gr.replay, 20 * B * flops[name] / 1e9, "TFLOP/s", iters=5)
                        del gr


def p_mlp_tc():      # mlp_tc_kernel (tcgen05 3xTF32, two-pass): the same three blocks of a C4 step, B = 4096, 20 launches per graph
    B = 4096
    env, model, users, ctx = _models(100000, 100000, 5, B)
    items = torch.randint(0, 100000, (B, 5), generator=G, device=DEV)
    pivot = torch.randint(0, 100000, (B,), generator=G, device=DEV)
    flops = 2 * (14 * 128 + 128 * 128 + 128 * 32 + 30 * 256 + 256 * 256 + 256 * 8) + 2 * (38 * 256 + 256 * 256 + 256 * 32) + \
        2 * (48 * 256 + 256 * 256 + 256 * 5)
    with torch.no_grad(), ops.mlp_engine("tc"):
        r, u, _ = model._inputs(ctx, users)
        _, z, _ = model._prior_chain(r, u, model.psmMLP)

        def step():
            model._prior_chain(r, u, model.psmMLP)
            model._scm(z, ("onehot", r), pivot, model._user_seg(u), [])
            env(items, users)
        timed("MLP blocks tc B=4096 (3 launches)", step, B * flops / 1e9, "TFLOP/s (fp32-equivalent flops)")


def p_respmlp():     # response MLP alone at a gather-bound batch
    env, model, users, ctx = _models(1000000, 1000000, 5, 65536)
    slates = torch.randint(0, 1000000, (65536, 5), generator=G, device=DEV)
    timed("response MLP B=65536", lambda: env(slates, users), 65536 * 260 / 1e3, "TB/s (algorithmic 260 B/slate)")


def p_urm():         # urm_kernel: gather-plus-dot, 284 B/slate
    B, n = 1 << 20, 1000000
    env = URM_P_MR(n - 1, n - 1, 5, 8, DEV, False, 0.2, 0.05, 0.5).to(DEV)
    slates = torch.randint(0, n, (B, 5), generator=G, device=DEV)
    users = torch.randint(0, n, (B,), generator=G, device=DEV)
    timed("urm_p_mr B=1M", lambda: env(slates, users), B * 284 / 1e3, "TB/s (algorithmic 284 B/slate)")


def p_metrics():     # slate_metrics_kernel: gather L rows + ILS + coverage bitmap
    from pivotcvae_b200 import analysis
    B, n = 1 << 20, 1000000
    W = table(n)
    slates = torch.randint(0, n, (B, 5), generator=G, device=DEV)
    timed("slate_metrics B=1M", lambda: analysis._metrics(slates, W, True, True), B * (5 * 32 + 40 + 4) / 1e3,
          "TB/s (algorithmic)")


def p_ce_tc():       # ce_tc2_kernel at C3 (tf32 logits and gradient on the tensor cores, dense mask)
    W, Q = table(100000), torch.randn(20480, 8, generator=G, device=DEV) * 0.5
    tab = ops.Table(W)
    tgt = torch.randint(0, 100000, (20480,), generator=G, device=DEV)
    timed("ce_tc C3 (100k x 20480)", lambda: ops.ce_fwd_bwd(tab, Q, tgt, engine="tf32"), 1e5 * 20480 / 1e3, "T logits/s")


def p_ce_exact():
    W, Q = table(100000), torch.randn(20480, 8, generator=G, device=DEV) * 0.5
    tab = ops.Table(W)
    tgt = torch.randint(0, 100000, (20480,), generator=G, device=DEV)
    timed("ce exact C3 (100k x 20480)", lambda: ops.ce_fwd_bwd(tab, Q, tgt), 1e5 * 20480 / 1e3, "T logits/s")


def p_ce_sparse():   # ce_sparse_kernel: Philox gap-process mask, n_neg = 1000
    W, Q = table(100000), torch.randn(20480, 8, generator=G, device=DEV) * 0.5
    tab = ops.Table(W)
    tgt = torch.randint(0, 100000, (20480,), generator=G, device=DEV)
    timed("ce_sparse C3 n_neg=1000", lambda: ops.ce_fwd_bwd(tab, Q, tgt, keep_prob=0.01, seed=3), 20480 * 1001 * 32 / 1e3,
          "TB/s (gathered rows)")


def p_cand_ce():     # cand_ce_kernel: sampled-softmax CE over 1000 candidates per row
    W, Q = table(100000), torch.randn(20480, 8, generator=G, device=DEV) * 0.5
    tab = ops.Table(W)
    cand = torch.randint(0, 100000, (20480, 1000), generator=G, device=DEV)
    pos = torch.zeros(20480, dtype=torch.int64, device=DEV)
    timed("cand_ce 20480 x 1000", lambda: ops.cand_ce_fwd_bwd(tab, Q, cand, pos), 20480 * 1000 * 40 / 1e3,
          "TB/s (ids + gathered rows)")


def p_gemm_dx():     # gemm_tn_tc_kernel, input-gradient shape: G[4096, 256] . (W^T)[256, 256]^T, act' epilogue + transposed copy
    Bsz, n = 4096, 256
    Gm = torch.randn(Bsz, n, generator=G, device=DEV)
    Wt = torch.randn(n, n, generator=G, device=DEV)
    saved = torch.randn(Bsz, n, generator=G, device=DEV)
    C, Ct = torch.empty(Bsz, n, device=DEV), torch.empty(n, Bsz, device=DEV)
    timed("gemm dX 4096x256x256", lambda: ops.gemm_tn(Gm, Wt, Bsz, n, n, C=C, Ct=Ct, dact_src=saved, dact=1),
          2.0 * Bsz * n * n / 1e3, "TFLOP/s")


def p_gemm_dw():     # weight-gradient shape: G^T[256, 4096] . (X^T)[256, 4096]^T, split-K 32
    Bsz, n = 4096, 256
    Gt = torch.randn(n, Bsz, generator=G, device=DEV)
    Xt = torch.randn(n, Bsz, generator=G, device=DEV)
    part = torch.empty(32, n, n, device=DEV)
    timed("gemm dW 256x256x4096 /32", lambda: ops.gemm_tn(Gt, Xt, n, n, Bsz, C=part, split_k=32), 2.0 * Bsz * n * n / 1e3, "TFLOP/s")


ALL = {k[2:]: v for k, v in list(globals().items()) if k.startswith("p_")}

if __name__ == "__main__":
    ops.device_ok(0)
    for name in (sys.argv[1:] or list(ALL)):
        ALL[name]()
