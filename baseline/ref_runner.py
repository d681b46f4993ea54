"""Runs the UNMODIFIED reference (baseline/_ref, see install_ref.py) for bench.py's baselines:

* cpu_baseline kind "reference": the reference's own classes on the box's host cores (all threads);
* gpu_library_baseline: the same classes with device="cuda:0", i.e. stock PyTorch (cuBLAS + ATen) on the B200.

Nothing of pivotcvae_b200 is on this path: models, weights loading, recommend(), the response model and
get_gen_loss are the reference's own code, imported from baseline/_ref.
"""
import contextlib
import io
import os
import sys
import time
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
_mods = None


def available():
    return os.path.exists(os.path.join(REF, "models", "pivotcvae.py"))


def load():
    """Import the reference's modules from baseline/_ref under private names, so they never shadow (or get
    shadowed by) the drop-in package.  matplotlib is stubbed (train_generative.py:9-10)."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError("baseline/_ref is missing: run `python baseline/install_ref.py` in the build container")
    for name in ["matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d"]:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.axes3d = None
            sys.modules[name] = m
    saved = {k: sys.modules.get(k) for k in ("models", "models.cvae", "models.pivotcvae", "models.listcvae", "env",
                                              "env.response_model", "train_generative", "my_utils", "settings",
                                              "data_extract", "data_loader", "analysis")}
    for k in saved:
        sys.modules.pop(k, None)
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import env.response_model as rm
            import models.listcvae as lc
            import models.pivotcvae as pc
            import train_generative as tg
        _mods = dict(rm=rm, lc=lc, pc=pc, tg=tg)
    finally:
        sys.path.remove(REF)
        for k, v in saved.items():          # leave sys.modules as we found it
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v
    return _mods


class _Emb:
    def __init__(self, w):
        self.weight = torch.from_numpy(np.ascontiguousarray(w))


def build(w, st, sd, env_sd, mode, device):
    """Reference model + response model on `device` with the given (numpy) weights."""
    mods = load()
    L, D, Z = w["L"], w["D"], w["Z"]
    uemb = None if w["no_user"] else _Emb(env_sd["userEmbed.weight"])
    with contextlib.redirect_stdout(io.StringIO()):
        if mode == "list":
            m = mods["lc"].UserListCVAEWithPrior(_Emb(env_sd["docEmbed.weight"]), uemb, L, D, Z, L + 1, st["enc"], st["dec"],
                                                 st["prior"], w["no_user"], device)
        else:
            key = "pivotcvae_gt_spi" if mode == "sampled" else "pivotcvae_gt_pi"
            m = mods["pc"].PIVOTCVAE_MODELS[key](_Emb(env_sd["docEmbed.weight"]), uemb, L, D, Z, L + 1, st["enc"], st["psm"],
                                                 st["scm"], st["prior"], w["no_user"], device)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        m.to(device)
        env = mods["rm"].UserResponseModel_MLP(w["n_items"] - 1, w["n_users"] - 1, D, L, st["resp"], device, w["no_user"])
        env.load_state_dict({k: torch.from_numpy(v) for k, v in env_sd.items()})
        env.to(device)
    return m, env


def _sync(device):
    if str(device).startswith("cuda"):
        torch.cuda.synchronize()


def time_generate(w, st, sd, env_sd, mode, device, inputs, min_seconds, warmup=1, max_steps=200):
    """slates/s of recommend(return_item=True) + resp_model(...) (train_generative.py:183-185).
    inputs(i) -> (ctx [B, L] f32, users [B] i64) CPU tensors; copies to `device` are inside the timed step,
    as the reference's own eval loop does them."""
    m, env = build(w, st, sd, env_sd, mode, device)
    times, n = [], 0
    i = 0
    with torch.no_grad():
        while True:
            ctx, users = inputs(i)
            B = ctx.shape[0]
            _sync(device)
            t0 = time.perf_counter()
            c, u = ctx.to(device), users.to(device)
            items, _ = m.recommend(c, None if w["no_user"] else u, return_item=True)
            resp = env(items.view(B, -1), u)
            resp = resp.cpu()
            _sync(device)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
                n += B
            i += 1
            if (sum(times) >= min_seconds and len(times) >= 3) or len(times) >= max_steps:
                break
    return n / sum(times), len(times), B


def time_train(w, st, sd, env_sd, device, batches, n_neg, beta, min_seconds, warmup=1, max_steps=100):
    """samples/s of get_gen_loss (mask-train) + backward + Adam.step (train_generative.py:124-134)."""
    mods = load()
    m, _ = build(w, st, sd, env_sd, "greedy", device)
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    CEL = torch.nn.CrossEntropyLoss()
    times, n, i = [], 0, 0
    while True:
        b = batches(i)
        B = b["slates"].shape[0]
        _sync(device)
        t0 = time.perf_counter()
        opt.zero_grad()
        loss, rec, kld = mods["tg"].get_gen_loss(b, m, CEL, beta, n_neg=n_neg)
        loss.backward()
        opt.step()
        loss.item()
        _sync(device)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
            n += B
        i += 1
        if (sum(times) >= min_seconds and len(times) >= 3) or len(times) >= max_steps:
            break
    return n / sum(times), len(times), B
