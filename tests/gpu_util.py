"""Helpers for the -m gpu parity tests (CUDA path through the C ABI vs the CPU oracle)."""
import numpy as np
import torch


def dev():
    return torch.device("cuda:0")


def T(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(dev())


def N(t):
    return t.detach().cpu().numpy()


def load_sd(model, sd):
    """Load a reference state_dict (numpy) into a drop-in module, bit for bit."""
    state = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}
    missing, unexpected = model.load_state_dict(state, strict=True), None
    return model


class _Emb:
    def __init__(self, w):
        self.weight = torch.from_numpy(np.ascontiguousarray(w))


def build_pivot(fx, key="pivotcvae_gt_pi"):
    from pivotcvae_b200.models.pivotcvae import PIVOTCVAE_MODELS
    cfg, sd = fx.cfg, fx.sub("sd/")
    m = PIVOTCVAE_MODELS[key](_Emb(sd["docEmbed.weight"]), None if cfg["no_user"] else _Emb(sd["userEmbed.weight"]),
                              cfg["L"], cfg["D"], cfg["Z"], cfg["L"] + 1, list(fx["cfg/enc"]), list(fx["cfg/psm"]),
                              list(fx["cfg/scm"]), list(fx["cfg/prior"]), bool(cfg["no_user"]), "cuda:0")
    return load_sd(m, sd)


def build_list(fx):
    from pivotcvae_b200.models.listcvae import UserListCVAEWithPrior
    cfg, sd = fx.cfg, fx.sub("sd/")
    L, D, Z, H, PH = cfg["L"], cfg["D"], cfg["Z"], cfg["hidden"], cfg["phidden"]
    ud = 0 if cfg["no_user"] else D
    m = UserListCVAEWithPrior(_Emb(sd["docEmbed.weight"]), None if cfg["no_user"] else _Emb(sd["userEmbed.weight"]),
                              L, D, Z, L + 1, [L * D + L + 1 + ud, H, H], [Z + L + 1 + ud, H, H, L * D],
                              [L + 1 + ud, PH, PH], bool(cfg["no_user"]), "cuda:0")
    return load_sd(m, sd)
