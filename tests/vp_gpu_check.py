"""Vocab-parallel recommend on N GPUs == single-GPU recommend, bit for bit (run under torchrun on a multi-GPU box):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/vp_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    import gpu_util
    gpu_util.dev = lambda: torch.device("cuda:%d" % torch.cuda.current_device())
    from gpu_util import N, T, build_pivot
    dev = "cuda:%d" % torch.cuda.current_device()
    ok = True
    for name in ("pivot_c1", "pivot_small"):
        fx = load_golden(name)
        for key, tag in (("pivotcvae_gt_pi", "rec_pi_k1/"), ("pivotcvae_gt_spi", "rec_spi_k2/")):
            import pivotcvae_b200.models.pivotcvae as mp
            cfg, sd = fx.cfg, fx.sub("sd/")
            m = mp.PIVOTCVAE_MODELS[key](gpu_util._Emb(sd["docEmbed.weight"]), gpu_util._Emb(sd["userEmbed.weight"]), cfg["L"],
                                         cfg["D"], cfg["Z"], cfg["L"] + 1, list(fx["cfg/enc"]), list(fx["cfg/psm"]),
                                         list(fx["cfg/scm"]), list(fx["cfg/prior"]), False, dev)
            gpu_util.load_sd(m, sd)
            m.enable_vocab_parallel()
            m.noise.push("eps", T(fx[tag + "eps"]))
            if (tag + "noise") in fx:
                m.noise.push("race", T(fx[tag + "noise"]))
            items, _ = m.recommend(T(fx[tag + "ctx"]), T(fx["in/users"]), return_item=True)
            same = np.array_equal(N(items), fx[tag + "items"])
            ok = ok and same
            if rank == 0:
                print("vocab-parallel x%d %s %s: %s" % (world, name, key, "bit-exact vs reference slates" if same else "MISMATCH"))
    # throughput mode (Philox noise): the MLP rows are sharded over the ranks too (all-gather of the queries before
    # every sharded scoring step) and a sampled pivot is drawn locally; the slates must not depend on any of it
    fx = load_golden("pivot_c1")
    cfg, sd = fx.cfg, fx.sub("sd/")
    import pivotcvae_b200.models.pivotcvae as mp
    for key in ("pivotcvae_gt_spi", "pivotcvae_gt_pi"):
        outs = []
        for vp in (False, True):
            m = mp.PIVOTCVAE_MODELS[key](gpu_util._Emb(sd["docEmbed.weight"]), gpu_util._Emb(sd["userEmbed.weight"]),
                                         cfg["L"], cfg["D"], cfg["Z"], cfg["L"] + 1, list(fx["cfg/enc"]),
                                         list(fx["cfg/psm"]), list(fx["cfg/scm"]), list(fx["cfg/prior"]), False, dev)
            gpu_util.load_sd(m, sd)
            m.noise.reseed(2024)
            if vp:
                m.enable_vocab_parallel()
                assert m._vp_row_slice(T(fx["rec_spi_k2/ctx"]).shape[0]) is not None or T(fx["rec_spi_k2/ctx"]).shape[0] % world
            items, z_mu = m.recommend(T(fx["rec_spi_k2/ctx"]), T(fx["in/users"]), return_item=True)
            outs.append((N(items), N(z_mu)))
        same = np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
        ok = ok and same
        if rank == 0:
            print("Philox noise, %s: sharded (rows + catalog) == unsharded: %s" % (key, same))
    # vocab-parallel TRAINING: CE partial records + one all-gather + merge == the unsharded get_gen_loss
    from pivotcvae_b200 import train_generative as tg
    fx = load_golden("pivot_small")
    cfg, sd = fx.cfg, fx.sub("sd/")
    rng = np.random.default_rng(5)
    B = 64
    batch = dict(slates=rng.integers(0, cfg["n_items"], (B, cfg["L"])), users=rng.integers(0, cfg["n_users"], (B,)),
                 responses=rng.integers(0, 2, (B, cfg["L"])).astype(np.float32))
    eps = torch.randn(B, cfg["Z"], generator=torch.Generator().manual_seed(3)).to(dev)
    res = []
    for vp in (False, True):
        for engine in ("exact", "tf32"):
            m = mp.PIVOTCVAE_MODELS["pivotcvae_gt_pi"](gpu_util._Emb(sd["docEmbed.weight"]), gpu_util._Emb(sd["userEmbed.weight"]),
                                                       cfg["L"], cfg["D"], cfg["Z"], cfg["L"] + 1, list(fx["cfg/enc"]),
                                                       list(fx["cfg/psm"]), list(fx["cfg/scm"]), list(fx["cfg/prior"]), False, dev)
            gpu_util.load_sd(m, sd)
            m.ce_engine = engine
            if vp:
                m.enable_vocab_parallel()
            m.noise.push("eps", eps)
            loss, rec, kld = tg.get_gen_loss(batch, m, None, 0.01, n_neg=cfg["n_items"])
            loss.backward()
            g = torch.cat([p.grad.reshape(-1) for p in m.parameters() if p.grad is not None])
            res.append((float(loss), g))
    for k, engine in enumerate(("exact", "tf32")):
        (l0, g0), (l1, g1) = res[k], res[2 + k]
        same = abs(l0 - l1) <= 1e-5 * abs(l0) and bool(torch.allclose(g0, g1, rtol=1e-3, atol=1e-6))
        ok = ok and same
        if rank == 0:
            print("vocab-parallel x%d training (%s CE): loss %.6f vs unsharded %.6f, grads max diff %.2e -> %s" % (
                world, engine, l1, l0, float((g0 - g1).abs().max()), "ok" if same else "MISMATCH"))
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(t) == 1 else 1)


if __name__ == "__main__":
    main()
