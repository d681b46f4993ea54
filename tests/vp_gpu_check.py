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
    # Philox exprace must not depend on the sharding either
    fx = load_golden("pivot_c1")
    cfg, sd = fx.cfg, fx.sub("sd/")
    import pivotcvae_b200.models.pivotcvae as mp
    outs = []
    for vp in (False, True):
        m = mp.PIVOTCVAE_MODELS["pivotcvae_gt_spi"](gpu_util._Emb(sd["docEmbed.weight"]), gpu_util._Emb(sd["userEmbed.weight"]),
                                                    cfg["L"], cfg["D"], cfg["Z"], cfg["L"] + 1, list(fx["cfg/enc"]),
                                                    list(fx["cfg/psm"]), list(fx["cfg/scm"]), list(fx["cfg/prior"]), False, dev)
        gpu_util.load_sd(m, sd)
        m.noise.reseed(2024)
        if vp:
            m.enable_vocab_parallel()
        items, _ = m.recommend(T(fx["rec_spi_k2/ctx"]), T(fx["in/users"]), return_item=True)
        outs.append(N(items))
    same = np.array_equal(outs[0], outs[1])
    ok = ok and same
    if rank == 0:
        print("Philox sampled pivot, sharded == unsharded:", same)
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(t) == 1 else 1)


if __name__ == "__main__":
    main()
