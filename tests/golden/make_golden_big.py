"""Golden fixtures at the BASELINE.json config sizes, from the UNMODIFIED reference classes.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_big.py [c2] [c3] [c4]

The fixtures stay small because everything large is regenerated from seeds on both sides
(tests/golden/synth.py): tables + MLP weights = bench.make_weights(workload), the Exp(1) draws of the
sampled pivot and the Bernoulli mask of downsample() = numpy PCG64 streams that this script INJECTS into
the reference (torch.multinomial / torch.bernoulli are pointed at them for the duration of the call; the
reference's own code is untouched).  Stored: seeds, checksums of every regenerated tensor, eps, and the
reference's outputs — items, z_mu, rx, the pivot the reference picked, the loss triple, every parameter
gradient — plus the reference's top-1/top-2 logits per row, so that a test can tell a legitimate
near-tie flip (gap below the measured MLP rounding difference) from a wrong answer.

  big_c2.npz  C2: 50 000 items, L=10, nouser, B=256 — PivotCVAE greedy (2 contexts), sampled pivot, ListCVAE
  big_c3.npz  C3: 100 000 items, L=5, with user, B=256 — get_gen_loss + backward, n_neg=1000 and n_neg=N
  big_c4.npz  C4: 1 000 000 items, L=5, with user, B=64 — PivotCVAE greedy
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402  (imports the reference modules, matplotlib stubbed)
import synth  # noqa: E402

import bench  # noqa: E402


class Inject(mg.Capture):
    """Capture, but the Exp(1) draws / the Bernoulli mask come from the given arrays."""

    def __init__(self, race=None, bern=None):
        super().__init__()
        self.race, self.bern_src = race, bern

    def __enter__(self):
        cap = self
        super().__enter__()
        if self.race is not None:
            def multinomial(p, num_samples, replacement=False, **k):
                assert num_samples == 1
                q = torch.from_numpy(cap.race)
                assert q.shape == p.shape
                return torch.argmax(p / q, dim=-1, keepdim=True)
            torch.multinomial = multinomial
        if self.bern_src is not None:
            def bernoulli(p, *a, **k):
                m = torch.from_numpy(cap.bern_src).to(p.dtype)
                assert m.shape == p.shape
                return m
            torch.bernoulli = bernoulli
        return self


class Emb:
    def __init__(self, w):
        self.weight = torch.from_numpy(np.ascontiguousarray(w))


def ref_models(workload, kind, keys):
    w, sd, env_sd = synth.weights(workload, kind)
    st = bench.structs(w)
    L, D, Z = w["L"], w["D"], w["Z"]
    uemb = None if w["no_user"] else Emb(env_sd["userEmbed.weight"])
    models = {}
    for key in keys:
        if key == "list":
            m = mg.quiet(mg.UserListCVAEWithPrior, Emb(env_sd["docEmbed.weight"]), uemb, L, D, Z, L + 1, st["enc"], st["dec"],
                         st["prior"], w["no_user"], "cpu")
        else:
            m = mg.quiet(mg.PIVOTCVAE_MODELS[key], Emb(env_sd["docEmbed.weight"]), uemb, L, D, Z, L + 1, st["enc"], st["psm"],
                         st["scm"], st["prior"], w["no_user"], "cpu")
        m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        models[key] = m
    return w, sd, env_sd, models


def top2(q, table):
    """Reference scoring op (cvae.py:97-101) -> (idx int64[M], top-2 logits f32[M, 2])."""
    p = torch.mm(q, table.t())
    v, i = torch.topk(p, 2, dim=1)
    first = torch.max(p, 1)[1]
    return first.numpy(), v.numpy()


def rec_case(out, tag, m, w, users, k, seed, race_seed=None):
    """recommend(return_item=True) + everything a test needs to audit a mismatch."""
    B, L, D = users.shape[0], w["L"], w["D"]
    ctx = torch.from_numpy(synth.contexts(B, L, k))
    u = None if w["no_user"] else users
    race = synth.race_noise(race_seed, B, w["n_items"]) if race_seed is not None else None
    seen = {}
    if hasattr(m, "pick_pivot"):
        orig = m.pick_pivot

        def spy(pivot_output, true_pivot=[]):
            seen["pivot_out"] = pivot_output.detach().clone()
            e = orig(pivot_output, true_pivot)
            seen["pivot_emb"] = e.detach().clone()
            return e
        m.pick_pivot = spy
    torch.manual_seed(seed)
    with Inject(race=race) as cap, torch.no_grad():
        items, zmu = m.recommend(ctx, u, return_item=True)
    eps = cap.normal[0].numpy()
    torch.manual_seed(seed)
    with Inject(race=race), torch.no_grad():
        rx, _ = m.recommend(ctx, u, return_item=False)
    if hasattr(m, "pick_pivot"):
        del m.pick_pivot
    rx = rx.reshape(B, L, D)
    table = m.docEmbed.weight.data
    idx, t2 = top2(rx.reshape(-1, D), table)
    assert np.array_equal(idx, items.numpy())
    out[tag + "k"], out[tag + "eps"] = np.array(k), eps
    out[tag + "items"], out[tag + "z_mu"], out[tag + "rx"] = items.numpy(), zmu.numpy(), rx.numpy()
    out[tag + "slot_top2"] = t2
    if "pivot_out" in seen:
        out[tag + "pivot_out"] = seen["pivot_out"].numpy()
        # the pivot the reference picked: its embedding row is slot 0 of rx; recover the index the way
        # pick_pivot computed it (greedy) or from the injected race (sampled)
        if race is None:
            pidx, pt2 = top2(seen["pivot_out"], table)
            out[tag + "pivot_top2"] = pt2
        else:
            p = torch.sigmoid(torch.mm(seen["pivot_out"], table.t()))
            p = p / p.sum(-1, keepdim=True)
            key = p / torch.from_numpy(race)
            v, i2 = torch.topk(key, 2, dim=1)
            pidx = torch.argmax(key, dim=-1).numpy()
            out[tag + "pivot_key_top2"] = v.numpy()
            out[tag + "race_seed"] = np.array(race_seed)
            out[tag + "race_sum"] = np.array(synth.checksum(race), dtype=np.uint64)
        assert torch.equal(table[torch.from_numpy(pidx)], seen["pivot_emb"])
        out[tag + "pivot_idx"] = pidx
    return items


def make_c2(path):
    t0 = time.time()
    B = 256
    w, sd, env_sd, models = ref_models("c2", "pivot", ["pivotcvae_gt_pi", "pivotcvae_gt_spi"])
    out = {"workload": np.array("c2"), "B": np.array(B)}
    out.update(synth.weights_checksums(sd, env_sd))
    g = torch.Generator().manual_seed(5150)
    users = torch.randint(0, w["n_users"], (B,), generator=g)
    out["users"] = users.numpy()
    items = rec_case(out, "greedy_k3/", models["pivotcvae_gt_pi"], w, users, 3, 7001)
    rec_case(out, "greedy_k10/", models["pivotcvae_gt_pi"], w, users, 10, 7002)
    rec_case(out, "sampled_k2/", models["pivotcvae_gt_spi"], w, users, 2, 7003, race_seed=9001)
    # the env's scorer on the generated slates (train_generative.py:185)
    env = mg.quiet(mg.UserResponseModel_MLP, w["n_items"] - 1, w["n_users"] - 1, w["D"], w["L"], bench.structs(w)["resp"], "cpu",
                   w["no_user"])
    env.load_state_dict({k: torch.from_numpy(v) for k, v in env_sd.items()})
    with torch.no_grad():
        out["greedy_k3/resp"] = env(items.view(B, -1), users).numpy()
    # ListCVAE on the same catalog (listcvae.py:170-188)
    wl, sdl, env_sdl, lm = ref_models("c2", "list", ["list"])
    lo = {}
    lo.update(synth.weights_checksums(sdl, env_sdl))
    rec_case(lo, "list_k4/", lm["list"], wl, users, 4, 7004)
    out.update({"list/" + k: v for k, v in lo.items()})
    np.savez_compressed(path, **out)
    print("c2 done in %.1fs" % (time.time() - t0))


def make_c3(path):
    t0 = time.time()
    B = 256
    w, sd, env_sd, models = ref_models("c3", "pivot", ["pivotcvae_gt_pi"])
    m = models["pivotcvae_gt_pi"]
    out = {"workload": np.array("c3"), "B": np.array(B)}
    out.update(synth.weights_checksums(sd, env_sd))
    batch = bench.make_train_batch(w, B, 0, seed=6100)
    out["slates"], out["users"], out["responses"] = (batch["slates"].numpy(), batch["users"].numpy(), batch["responses"].numpy())
    nb = {"slates": batch["slates"].numpy(), "users": batch["users"].numpy(), "responses": batch["responses"].numpy().astype(np.float64)}
    CEL = torch.nn.CrossEntropyLoss()
    N, M = w["n_items"], B * w["L"]
    for n_neg, tag, mseed in ((1000, "nneg1000/", 8101), (N, "full/", None)):
        keep = n_neg / N
        mask = synth.bernoulli_mask(mseed, M, N, keep) if mseed is not None else np.ones((M, N), dtype=bool)
        m.zero_grad()
        torch.manual_seed(6200)
        with Inject(bern=mask) as cap:
            loss, rec, kld = mg.tg.get_gen_loss(nb, m, CEL, 0.001, n_neg=n_neg)
            loss.backward()
        out[tag + "n_neg"] = np.array(n_neg)
        out[tag + "eps"] = cap.normal[0].numpy()
        out[tag + "loss"] = np.array([loss.item(), rec.item(), kld.item()], dtype=np.float64)
        if mseed is not None:
            out[tag + "mask_seed"] = np.array(mseed)
            out[tag + "mask_sum"] = np.array(synth.checksum(synth.pack_bitmask(mask)), dtype=np.uint64)
        for name, prm in m.named_parameters():
            if prm.grad is not None:
                out[tag + "grad/" + name] = prm.grad.numpy().copy()
        print(" c3", tag, out[tag + "loss"], "%.1fs" % (time.time() - t0))
    np.savez_compressed(path, **out)
    print("c3 done in %.1fs" % (time.time() - t0))


def make_c4(path):
    t0 = time.time()
    B = 64
    w, sd, env_sd, models = ref_models("c4", "pivot", ["pivotcvae_gt_pi"])
    out = {"workload": np.array("c4"), "B": np.array(B)}
    out.update(synth.weights_checksums(sd, env_sd))
    g = torch.Generator().manual_seed(5151)
    users = torch.randint(0, w["n_users"], (B,), generator=g)
    out["users"] = users.numpy()
    rec_case(out, "greedy_k2/", models["pivotcvae_gt_pi"], w, users, 2, 7101)
    np.savez_compressed(path, **out)
    print("c4 done in %.1fs" % (time.time() - t0))


if __name__ == "__main__":
    torch.set_num_threads(1)   # the pinned path is the single-threaded CPU mm
    which = sys.argv[1:] or ["c2", "c3", "c4"]
    for name in which:
        {"c2": make_c2, "c3": make_c3, "c4": make_c4}[name](os.path.join(HERE, "big_%s.npz" % name))
    for f in sorted(os.listdir(HERE)):
        if f.startswith("big_"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
