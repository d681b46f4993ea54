"""Generate golden fixtures by running the UNMODIFIED reference classes.

Run in the build container (the only place /root/reference exists):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports models/*.py, env/response_model.py and train_generative.py from
/root/reference (matplotlib stubbed: SURVEY F12), builds seeded models, captures
every random tensor the reference draws (normal_ for the reparameterisation,
the Exp(1) draws inside Categorical.sample, the Bernoulli mask of downsample)
and stores inputs, weights, noise and the reference's outputs as .npz files next
to this script.  The fixtures pin oracle/ (tests/test_oracle_golden.py) and the
CUDA path (tests/test_gpu_*.py).  Nothing at test time reads /root/reference.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("PCV_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

for name in ["matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d"]:
    m = types.ModuleType(name)
    m.axes3d = None
    sys.modules.setdefault(name, m)
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

import contextlib  # noqa: E402
import io  # noqa: E402

with contextlib.redirect_stdout(io.StringIO()):
    from env.response_model import URM, URM_P, URM_P_MR, UserResponseModel_MLP  # noqa: E402
    from models.listcvae import UserListCVAEWithPrior  # noqa: E402
    from models.pivotcvae import PIVOTCVAE_MODELS  # noqa: E402
    import train_generative as tg  # noqa: E402

torch.set_num_threads(1)  # the pinned path is the single-threaded CPU mm


class Capture:
    """Records the random tensors the reference draws, in call order."""

    def __init__(self):
        self.normal, self.expo, self.bern = [], [], []
        self._normal_ = torch.Tensor.normal_
        self._multinomial = torch.multinomial
        self._bernoulli = torch.bernoulli

    def __enter__(self):
        cap = self

        def normal_(t, *a, **k):
            out = cap._normal_(t, *a, **k)
            cap.normal.append(out.detach().clone())
            return out

        def multinomial(p, num_samples, replacement=False, **k):
            if num_samples != 1:
                return cap._multinomial(p, num_samples, replacement, **k)
            # torch's own fast path for one draw (ATen multinomial): argmax(p / Exp(1))
            q = torch.empty_like(p).exponential_(1)
            cap.expo.append(q.detach().clone())
            return torch.argmax(p / q, dim=-1, keepdim=True)

        def bernoulli(p, *a, **k):
            out = cap._bernoulli(p, *a, **k)
            cap.bern.append(out.detach().clone())
            return out

        torch.Tensor.normal_ = normal_
        torch.multinomial = multinomial
        torch.bernoulli = bernoulli
        return self

    def __exit__(self, *exc):
        torch.Tensor.normal_ = self._normal_
        torch.multinomial = self._multinomial
        torch.bernoulli = self._bernoulli


def check_multinomial_emulation():
    """Categorical.sample() == argmax(p_norm / Exp(1)) under the same seed."""
    from torch.distributions.categorical import Categorical
    p = torch.sigmoid(torch.randn(7, 501))
    torch.manual_seed(123)
    real = Categorical(p).sample()
    torch.manual_seed(123)
    with Capture() as cap:
        emu = Categorical(p).sample()
    assert torch.equal(real, emu), "multinomial emulation differs from torch"
    return cap.expo[0].shape == p.shape


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def sd_np(model, prefix="sd/"):
    return {prefix + k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}


def pack_bits(mask):
    mask = mask.detach().cpu().numpy() != 0
    M, N = mask.shape
    words = (N + 31) // 32
    pad = np.zeros((M, words * 32), dtype=bool)
    pad[:, :N] = mask
    b = pad.reshape(M, words, 32).astype(np.uint64)
    return (b << np.arange(32, dtype=np.uint64)).sum(-1).astype(np.uint32)


def make_env(n_items, n_users, L, D, struct, no_user, seed):
    torch.manual_seed(seed)
    return quiet(UserResponseModel_MLP, n_items - 1, n_users - 1, D, L, struct, "cpu", no_user)


def contexts(B, L, k):
    c = torch.zeros(B, L)
    c[:, :k] = 1
    return c


def pivot_fixture(path, n_items, n_users, L, D, Z, hidden, phidden, B, no_user, seed, with_grads, p_rows):
    env = make_env(n_items, n_users, L, D, [(L + (0 if no_user else 1)) * D, hidden, hidden, L], no_user, seed)
    C = L + 1
    ud = 0 if no_user else D
    enc = [L * D + C + ud, hidden, hidden]
    psm = [Z + C + ud, hidden, hidden, D]
    scm = [Z + C + D + ud, hidden, hidden, (L - 1) * D]
    pri = [C + ud, phidden, phidden]
    torch.manual_seed(seed + 1)
    models = {}
    base = None
    for key, cls in PIVOTCVAE_MODELS.items():
        m = quiet(cls, env.docEmbed, None if no_user else env.userEmbed, L, D, Z, C, enc, psm, scm, pri, no_user, "cpu")
        if base is None:
            base = m
        else:
            m.load_state_dict(base.state_dict())
        models[key] = m
    out = dict(sd_np(base))
    out["cfg"] = np.array([n_items, n_users, L, D, Z, hidden, phidden, B, int(no_user)], dtype=np.int64)
    out["cfg/enc"], out["cfg/psm"], out["cfg/scm"], out["cfg/prior"] = map(np.array, (enc, psm, scm, pri))

    g = torch.Generator().manual_seed(seed + 2)
    users = torch.randint(0, n_users, (B,), generator=g)
    slates = torch.randint(0, n_items, (B, L), generator=g)
    resp = (torch.rand(B, L, generator=g) < 0.5).float()
    out["in/users"], out["in/slates"], out["in/resp"] = users.numpy(), slates.numpy(), resp.numpy()

    # ---- recommend, every inference mode, every eval-loop context (train_generative.py:179-184)
    for infer, key in (("pi", "pivotcvae_gt_pi"), ("spi", "pivotcvae_gt_spi")):
        m = models[key]
        for k in ([1, L] if infer == "pi" else [2]):
            ctx = contexts(B, L, k)
            torch.manual_seed(seed + 10 + k)
            with Capture() as cap, torch.no_grad():
                items, zmu = m.recommend(ctx, None if no_user else users, return_item=True)
            tag = "rec_%s_k%d/" % (infer, k)
            out[tag + "ctx"] = ctx.numpy()
            out[tag + "eps"] = cap.normal[0].numpy()
            out[tag + "items"] = items.numpy()
            out[tag + "z_mu"] = zmu.numpy()
            if infer == "spi":
                out[tag + "noise"] = cap.expo[0].numpy()
            # same eps again for rx
            torch.manual_seed(seed + 10 + k)
            with Capture(), torch.no_grad():
                rx, _ = m.recommend(ctx, None if no_user else users, return_item=False)
            out[tag + "rx"] = rx.numpy()

    # ---- forward + get_gen_loss for the four training pivots
    CEL = torch.nn.CrossEntropyLoss()
    batch = {"slates": slates.numpy(), "users": users.numpy().reshape(-1, 1), "responses": resp.numpy().astype(np.float64)}
    for train, key in (("gt", "pivotcvae_gt_pi"), ("pt", "pivotcvae_pt_pi"), ("spt", "pivotcvae_spt_pi"), ("sgt", "pivotcvae_sgt_pi")):
        m = models[key]
        tag = "fwd_%s/" % train
        torch.manual_seed(seed + 20)
        with Capture() as cap, torch.no_grad():
            p, rx, z, emb, mu, lv = m.forward(slates, resp, u=None if no_user else users)
        out[tag + "eps"] = cap.normal[0].numpy()
        if cap.expo:
            out[tag + "noise"] = cap.expo[0].numpy()
        out[tag + "p_rows"] = p[:p_rows].numpy()
        out[tag + "p_argmax"] = p.argmax(1).numpy()
        out[tag + "rx"], out[tag + "z"], out[tag + "emb"] = rx.numpy(), z.numpy(), emb.numpy()
        out[tag + "z_mu"], out[tag + "z_logvar"] = mu.numpy(), lv.numpy()
        for n_neg, ntag in ((min(1000, n_items // 4), "mask"), (n_items, "full")):
            tag2 = "loss_%s_%s/" % (train, ntag)
            m.zero_grad()
            torch.manual_seed(seed + 30)
            with Capture() as cap:
                loss, rec, kld = tg.get_gen_loss(batch, m, CEL, 0.01, n_neg=n_neg)
                loss.backward()
            out[tag2 + "n_neg"] = np.array(n_neg)
            out[tag2 + "eps"] = cap.normal[0].numpy()
            if cap.expo:
                out[tag2 + "noise"] = cap.expo[0].numpy()
            out[tag2 + "bitmask"] = pack_bits(cap.bern[0])
            out[tag2 + "loss"] = np.array([loss.item(), rec.item(), kld.item()], dtype=np.float64)
            if with_grads and train == "gt":
                for name, prm in m.named_parameters():
                    if prm.grad is not None:
                        out[tag2 + "grad/" + name] = prm.grad.numpy().copy()
                    else:
                        out[tag2 + "nograd/" + name] = np.zeros(1)
    # ---- the env's own scorer on the recommended slates (train_generative.py:185)
    it = torch.from_numpy(out["rec_pi_k1/items"]).view(B, -1)
    with torch.no_grad():
        out["env/resp"] = env(it, users).numpy()
    out.update(sd_np(env, "env_sd/"))
    np.savez(path, **out)
    return out


def list_fixture(path, n_items, n_users, L, D, Z, hidden, phidden, B, no_user, seed):
    env = make_env(n_items, n_users, L, D, [(L + (0 if no_user else 1)) * D, hidden, hidden, L], no_user, seed)
    C = L + 1
    ud = 0 if no_user else D
    enc = [L * D + C + ud, hidden, hidden]
    dec = [Z + C + ud, hidden, hidden, L * D]
    pri = [C + ud, phidden, phidden]
    torch.manual_seed(seed + 1)
    m = quiet(UserListCVAEWithPrior, env.docEmbed, None if no_user else env.userEmbed, L, D, Z, C, enc, dec, pri, no_user, "cpu")
    out = dict(sd_np(m))
    out["cfg"] = np.array([n_items, n_users, L, D, Z, hidden, phidden, B, int(no_user)], dtype=np.int64)
    g = torch.Generator().manual_seed(seed + 2)
    users = torch.randint(0, n_users, (B,), generator=g)
    slates = torch.randint(0, n_items, (B, L), generator=g)
    resp = (torch.rand(B, L, generator=g) < 0.5).float()
    out["in/users"], out["in/slates"], out["in/resp"] = users.numpy(), slates.numpy(), resp.numpy()
    for k in (1, 3):
        ctx = contexts(B, L, k)
        torch.manual_seed(seed + 10 + k)
        with Capture() as cap, torch.no_grad():
            items, zmu = m.recommend(ctx, None if no_user else users, return_item=True)
        tag = "rec_k%d/" % k
        out[tag + "ctx"], out[tag + "eps"] = ctx.numpy(), cap.normal[0].numpy()
        out[tag + "items"], out[tag + "z_mu"] = items.numpy(), zmu.numpy()
        torch.manual_seed(seed + 10 + k)
        with Capture(), torch.no_grad():
            rx, _ = m.recommend(ctx, None if no_user else users, return_item=False)
        out[tag + "rx"] = rx.numpy()
    torch.manual_seed(seed + 20)
    with Capture() as cap, torch.no_grad():
        p, rx, z, emb, mu, lv = m.forward(slates, resp, u=None if no_user else users)
    out["fwd/eps"] = cap.normal[0].numpy()
    out["fwd/p_rows"], out["fwd/p_argmax"] = p[:8].numpy(), p.argmax(1).numpy()
    out["fwd/rx"], out["fwd/z"], out["fwd/emb"] = rx.numpy(), z.numpy(), emb.numpy()
    out["fwd/z_mu"], out["fwd/z_logvar"] = mu.numpy(), lv.numpy()
    CEL = torch.nn.CrossEntropyLoss()
    batch = {"slates": slates.numpy(), "users": users.numpy().reshape(-1, 1), "responses": resp.numpy().astype(np.float64)}
    for n_neg, ntag in ((n_items // 4, "mask"), (n_items, "full")):
        tag2 = "loss_%s/" % ntag
        m.zero_grad()
        torch.manual_seed(seed + 30)
        with Capture() as cap:
            loss, rec, kld = tg.get_gen_loss(batch, m, CEL, 0.01, n_neg=n_neg)
            loss.backward()
        out[tag2 + "n_neg"] = np.array(n_neg)
        out[tag2 + "eps"] = cap.normal[0].numpy()
        out[tag2 + "bitmask"] = pack_bits(cap.bern[0])
        out[tag2 + "loss"] = np.array([loss.item(), rec.item(), kld.item()], dtype=np.float64)
        for name, prm in m.named_parameters():
            if prm.grad is not None:
                out[tag2 + "grad/" + name] = prm.grad.numpy().copy()
    it = torch.from_numpy(out["rec_k1/items"]).view(B, -1)
    with torch.no_grad():
        out["env/resp"] = env(it, users).numpy()
    out.update(sd_np(env, "env_sd/"))
    np.savez(path, **out)


def env_fixture(path, seed):
    out = {}
    n_items, n_users, L, D, B = 700, 400, 5, 8, 96
    g = torch.Generator().manual_seed(seed)
    slates = torch.randint(0, n_items, (B, L), generator=g)
    users = torch.randint(0, n_users, (B,), generator=g)
    out["in/slates"], out["in/users"] = slates.numpy(), users.numpy()
    out["cfg"] = np.array([n_items, n_users, L, D, B], dtype=np.int64)
    for no_user in (False, True):
        e = make_env(n_items, n_users, L, D, [(L + (0 if no_user else 1)) * D, 64, 48, L], no_user, seed + 1)
        tag = "mlp_nouser/" if no_user else "mlp_user/"
        with torch.no_grad():
            out[tag + "out"] = e(slates, users).numpy()
        out.update(sd_np(e, tag + "sd/"))
    for name, ctor in (("urm", lambda: URM(n_items - 1, n_users - 1, L, D, "cpu", False)),
                       ("urm_p", lambda: URM_P(n_items - 1, n_users - 1, L, D, "cpu", False, 0.3, -0.1)),
                       ("urm_p_mr", lambda: URM_P_MR(n_items - 1, n_users - 1, L, D, "cpu", False, 0.3, -0.1, 0.7))):
        torch.manual_seed(seed + 2)
        e = quiet(ctor)
        with torch.no_grad():
            e.itemBias.weight.data.uniform_(-0.2, 0.2)
            e.userBias.weight.data.uniform_(-0.2, 0.2)
            out[name + "/out"] = e(slates, users).numpy()
        out.update(sd_np(e, name + "/sd/"))
        if hasattr(e, "posBias"):
            out[name + "/posBias"] = e.posBias.numpy()
            out[name + "/posDependentBias"] = e.posDependentBias.numpy()
        if hasattr(e, "mrFactor"):
            out[name + "/mrFactor"] = np.array(e.mrFactor, dtype=np.float32)
    np.savez(path, **out)


def dims_fixture(path, seed):
    """Reference scoring ops (torch.mm + torch.max, cvae.py:97-101 / pivotcvae.py:191) at other
    embedding sizes, including exact ties (duplicated rows) to pin first-index semantics."""
    import torch.nn.functional as F
    out = {}
    for D in (4, 8, 16, 32, 64, 128):
        g = torch.Generator().manual_seed(seed + D)
        N, M = 1500, 48
        W = F.normalize(torch.rand(N, D, generator=g) * 2 - 1, p=2, dim=1)
        W[700] = W[13]    # exact duplicate rows -> exact score ties
        W[1499] = W[13]
        Q = torch.randn(M, D, generator=g) * 0.5
        Q[0] = W[13] * 3  # its best item is the duplicated row: lowest index must win
        p = torch.mm(Q, W.t())
        vals, idx = torch.max(p, 1)
        idx0 = torch.mm(W, Q.t()).max(0)[1]   # pick_pivot's transposed orientation
        assert torch.equal(idx, idx0)
        out["d%d/W" % D], out["d%d/Q" % D] = W.numpy(), Q.numpy()
        out["d%d/idx" % D], out["d%d/val" % D] = idx.numpy(), vals.numpy()
        out["d%d/p_rows" % D] = p[:4].numpy()
    np.savez(path, **out)


def cand_fixture(path, seed):
    """Candidate-mode (sampled soft-max) training, the reference's default (no --mask_train):
    the reference's own Dataset sampler (data_loader.py:49-58) draws the candidates."""
    import data_loader
    n_items, n_users, L, D, Z, hidden, phidden, B, nC = 800, 200, 5, 8, 16, 64, 32, 24, 50
    env = make_env(n_items, n_users, L, D, [(L + 1) * D, hidden, hidden, L], False, seed)
    C = L + 1
    enc = [L * D + C + D, hidden, hidden]
    psm = [Z + C + D, hidden, hidden, D]
    scm = [Z + C + D + D, hidden, hidden, (L - 1) * D]
    dec = [Z + C + D, hidden, hidden, L * D]
    pri = [C + D, phidden, phidden]
    g = torch.Generator().manual_seed(seed + 2)
    users = torch.randint(0, n_users, (B,), generator=g)
    slates = torch.randint(0, n_items, (B, L), generator=g)
    slates[0, 0] = n_items - 1                       # max_iid is derived from the data
    resp = (torch.rand(B, L, generator=g) < 0.5).float()
    ds = quiet(data_loader.UserSlateResponseDataset, slates.numpy(), users.numpy(), resp.numpy(), False)
    ds.init_sampling(nC) if False else quiet(ds.init_sampling, nC)
    np.random.seed(seed)
    rows = [ds[i] for i in range(B)]
    batch = {k: np.stack([r_[k] for r_ in rows]) for k in rows[0]}
    out = {"in/users": users.numpy(), "in/slates": slates.numpy(), "in/resp": resp.numpy(),
           "in/candidates": batch["sample_candidates"], "in/targets": batch["sample_targets"],
           "cfg": np.array([n_items, n_users, L, D, Z, hidden, phidden, B, 0], dtype=np.int64),
           "cfg/enc": np.array(enc), "cfg/psm": np.array(psm), "cfg/scm": np.array(scm), "cfg/prior": np.array(pri),
           "cfg/dec": np.array(dec)}
    CEL = torch.nn.CrossEntropyLoss()
    torch.manual_seed(seed + 1)
    pm = quiet(PIVOTCVAE_MODELS["pivotcvae_gt_pi"], env.docEmbed, env.userEmbed, L, D, Z, C, enc, psm, scm, pri, False, "cpu")
    lm = quiet(UserListCVAEWithPrior, env.docEmbed, env.userEmbed, L, D, Z, C, enc, dec, pri, False, "cpu")
    for tag, m in (("pivot/", pm), ("list/", lm)):
        m.candidateFlag = True
        out.update(sd_np(m, tag + "sd/"))
        torch.manual_seed(seed + 30)
        with Capture() as cap:
            loss, rec, kld = tg.get_gen_loss(batch, m, CEL, 0.01)
            loss.backward()
        out[tag + "eps"] = cap.normal[0].numpy()
        out[tag + "loss"] = np.array([loss.item(), rec.item(), kld.item()], dtype=np.float64)
        for name, prm in m.named_parameters():
            if prm.grad is not None:
                out[tag + "grad/" + name] = prm.grad.numpy().copy()
        torch.manual_seed(seed + 30)
        with Capture(), torch.no_grad():
            p, rx, z, emb, mu, lv = m.forward(torch.from_numpy(batch["slates"]), resp, candidates=torch.from_numpy(batch["sample_candidates"]), u=users)
        out[tag + "p"] = p.numpy()
        out[tag + "rx"] = rx.numpy()
    np.savez(path, **out)


def metrics_fixture(path, seed):
    """analysis.py get_coverage / get_ILS on random slates."""
    import analysis
    g = torch.Generator().manual_seed(seed)
    n_items, B = 900, 200
    emb = torch.nn.Embedding(n_items, 8)
    with torch.no_grad():
        emb.weight.copy_(torch.randn(n_items, 8, generator=g))
    slates = torch.randint(0, n_items, (B, 5), generator=g)
    slates[3] = slates[3, 0]          # a slate of identical items: ILS = 1
    with torch.no_grad():
        ils = analysis.get_ILS(slates, emb)
    np.savez(path, **{"table": emb.weight.detach().numpy(), "slates": slates.numpy(), "ils": ils.numpy(),
                      "coverage": np.array(analysis.get_coverage(slates, n_items))})


if __name__ == "__main__":
    assert check_multinomial_emulation()
    if len(sys.argv) > 1 and sys.argv[1] == "metrics":
        metrics_fixture(os.path.join(HERE, "metrics.npz"), 99)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "cand":
        cand_fixture(os.path.join(HERE, "cand_small.npz"), 6060)
        sys.exit(0)
    # C1 shape (ML-1M): 3707 items, 6041 users, L=5, D=8, z=16, reference default structs
    pivot_fixture(os.path.join(HERE, "pivot_c1.npz"), 3707, 6041, 5, 8, 16, 256, 128, 16, False, 20211, False, 4)
    # small model, all variants, gradients
    pivot_fixture(os.path.join(HERE, "pivot_small.npz"), 500, 300, 5, 8, 16, 64, 32, 24, False, 777, True, 8)
    pivot_fixture(os.path.join(HERE, "pivot_small_nouser.npz"), 400, 10, 4, 8, 12, 48, 32, 20, True, 99, True, 8)
    list_fixture(os.path.join(HERE, "list_small.npz"), 2000, 50, 10, 8, 16, 64, 32, 24, True, 4242)
    list_fixture(os.path.join(HERE, "list_small_user.npz"), 600, 200, 5, 8, 16, 64, 32, 16, False, 4343)
    env_fixture(os.path.join(HERE, "env_small.npz"), 31337)
    dims_fixture(os.path.join(HERE, "dims.npz"), 555)
    cand_fixture(os.path.join(HERE, "cand_small.npz"), 6060)
    metrics_fixture(os.path.join(HERE, "metrics.npz"), 99)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
