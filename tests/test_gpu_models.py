"""-m gpu: the drop-in classes (reference API) against the reference's own outputs
(golden fixtures) and against the CPU oracle, on identical weights / inputs / noise."""
import io

import numpy as np
import pytest
import torch

import oracle
from gpu_util import N, T, build_list, build_pivot, dev

pytestmark = pytest.mark.gpu

PIVOT_FIXTURES = ["pivot_small", "pivot_small_nouser", "pivot_c1"]


def _users(fx):
    return None if fx.cfg["no_user"] else T(fx["in/users"])


@pytest.mark.parametrize("name", PIVOT_FIXTURES)
def test_pivot_recommend_greedy_bitexact_slates(golden, name):
    fx = golden(name)
    m = build_pivot(fx)
    cfg = fx.cfg
    for k in (1, cfg["L"]):
        tag = "rec_pi_k%d/" % k
        m.noise.push("eps", T(fx[tag + "eps"]))
        items, zmu = m.recommend(T(fx[tag + "ctx"]), _users(fx), return_item=True)
        assert items.dtype == torch.int64 and items.shape == (cfg["B"] * cfg["L"],)
        assert np.array_equal(N(items), fx[tag + "items"])          # == reference slates
        np.testing.assert_allclose(N(zmu), fx[tag + "z_mu"], rtol=2e-5, atol=2e-6)
        ref = oracle.pivot_recommend(fx.sub("sd/"), fx[tag + "ctx"], fx["in/users"], fx[tag + "eps"], cfg["no_user"])
        assert np.array_equal(N(zmu), ref["z_mu"])                  # bitwise == oracle
        m.noise.push("eps", T(fx[tag + "eps"]))
        rx, _ = m.recommend(T(fx[tag + "ctx"]), _users(fx), return_item=False)
        assert rx.shape == (cfg["B"], cfg["L"], cfg["D"])
        assert np.array_equal(N(rx), ref["rx"])
        np.testing.assert_allclose(N(rx), fx[tag + "rx"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", PIVOT_FIXTURES)
def test_pivot_recommend_sampled_identical_rng(golden, name):
    fx = golden(name)
    m = build_pivot(fx, "pivotcvae_gt_spi")
    tag = "rec_spi_k2/"
    m.noise.push("eps", T(fx[tag + "eps"]))
    m.noise.push("race", T(fx[tag + "noise"]))
    items, _ = m.recommend(T(fx[tag + "ctx"]), _users(fx), return_item=True)
    assert np.array_equal(N(items), fx[tag + "items"])


@pytest.mark.parametrize("name", ["list_small", "list_small_user"])
def test_list_recommend(golden, name):
    fx = golden(name)
    m = build_list(fx)
    for k in (1, 3):
        tag = "rec_k%d/" % k
        m.noise.push("eps", T(fx[tag + "eps"]))
        items, zmu = m.recommend(T(fx[tag + "ctx"]), _users(fx), return_item=True)
        assert np.array_equal(N(items), fx[tag + "items"])
        np.testing.assert_allclose(N(zmu), fx[tag + "z_mu"], rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("name", PIVOT_FIXTURES)
@pytest.mark.parametrize("train,key", [("gt", "pivotcvae_gt_pi"), ("pt", "pivotcvae_pt_pi"),
                                       ("spt", "pivotcvae_spt_pi"), ("sgt", "pivotcvae_sgt_pi")])
def test_pivot_forward(golden, name, train, key):
    fx = golden(name)
    m = build_pivot(fx, key)
    cfg = fx.cfg
    tag = "fwd_%s/" % train
    m.noise.push("eps", T(fx[tag + "eps"]))
    if (tag + "noise") in fx:
        m.noise.push("race", T(fx[tag + "noise"]))
    with torch.no_grad():
        p, rx, z, emb, mu, lv = m.forward(T(fx["in/slates"]), T(fx["in/resp"]), u=_users(fx))
    assert p.shape == (cfg["B"] * cfg["L"], cfg["n_items"])
    np.testing.assert_allclose(N(mu), fx[tag + "z_mu"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(N(lv), fx[tag + "z_logvar"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(N(z), fx[tag + "z"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(N(rx), fx[tag + "rx"], rtol=1e-4, atol=1e-5)
    assert np.array_equal(N(emb), fx[tag + "emb"])
    n = fx[tag + "p_rows"].shape[0]
    np.testing.assert_allclose(N(p[:n]), fx[tag + "p_rows"], rtol=1e-4, atol=1e-5)   # logits within 1e-4


@pytest.mark.parametrize("name", ["pivot_small", "pivot_small_nouser"])
@pytest.mark.parametrize("ntag", ["mask", "full"])
def test_gen_loss_and_grads(golden, name, ntag):
    """get_gen_loss (train_generative.py:44-65) + backward vs torch autograd of the reference."""
    from pivotcvae_b200.train_generative import get_gen_loss
    fx = golden(name)
    m = build_pivot(fx)
    t2 = "loss_gt_%s/" % ntag
    m.noise.push("eps", T(fx[t2 + "eps"]))
    m.noise.push("mask", T(fx[t2 + "bitmask"].view(np.int32)))
    batch = {"slates": fx["in/slates"], "users": fx["in/users"].reshape(-1, 1), "responses": fx["in/resp"].astype(np.float64)}
    loss, rec, kld = get_gen_loss(batch, m, None, 0.01, n_neg=int(fx[t2 + "n_neg"]))
    want = fx[t2 + "loss"]
    np.testing.assert_allclose([loss.item(), rec.item(), kld.item()], want, rtol=1e-4)
    loss.backward()
    grads = fx.sub(t2 + "grad/")
    assert grads
    for pname, prm in m.named_parameters():
        if pname in grads:
            np.testing.assert_allclose(N(prm.grad), grads[pname], rtol=2e-3, atol=2e-6, err_msg=pname)
        else:
            assert prm.grad is None, pname      # PSM and the frozen tables get no gradient


@pytest.mark.parametrize("name", ["list_small", "list_small_user"])
def test_list_gen_loss_and_grads(golden, name):
    from pivotcvae_b200.train_generative import get_gen_loss
    fx = golden(name)
    m = build_list(fx)
    for ntag in ("mask", "full"):
        t2 = "loss_%s/" % ntag
        m.zero_grad()
        m.noise.push("eps", T(fx[t2 + "eps"]))
        m.noise.push("mask", T(fx[t2 + "bitmask"].view(np.int32)))
        batch = {"slates": fx["in/slates"], "users": fx["in/users"].reshape(-1, 1), "responses": fx["in/resp"]}
        loss, rec, kld = get_gen_loss(batch, m, None, 0.01, n_neg=int(fx[t2 + "n_neg"]))
        np.testing.assert_allclose([loss.item(), rec.item(), kld.item()], fx[t2 + "loss"], rtol=1e-4)
        loss.backward()
        for pname, prm in m.named_parameters():
            g = fx.sub(t2 + "grad/").get(pname)
            if g is not None:
                np.testing.assert_allclose(N(prm.grad), g, rtol=2e-3, atol=2e-6, err_msg=pname)


def test_response_mlp_and_env_on_slates(golden):
    from pivotcvae_b200.env.response_model import UserResponseModel_MLP
    fx = golden("env_small")
    n_items, n_users, L, D, B = [int(v) for v in fx["cfg"]]
    for tag, nu in (("mlp_user/", False), ("mlp_nouser/", True)):
        e = UserResponseModel_MLP(n_items - 1, n_users - 1, D, L, [(L + (0 if nu else 1)) * D, 64, 48, L], "cuda:0", nu)
        e.load_state_dict({k: torch.from_numpy(v) for k, v in fx.sub(tag + "sd/").items()})
        e.to("cuda:0")
        out = e(T(fx["in/slates"]), T(fx["in/users"]))
        np.testing.assert_allclose(N(out), fx[tag + "out"], rtol=1e-4, atol=1e-5)
        assert np.array_equal(N(out), oracle.resp_mlp(fx.sub(tag + "sd/"), fx["in/slates"], fx["in/users"], nu))


def test_urm_classes(golden):
    from pivotcvae_b200.env.response_model import URM, URM_P, URM_P_MR
    fx = golden("env_small")
    n_items, n_users, L, D, B = [int(v) for v in fx["cfg"]]
    ctors = {"urm": lambda: URM(n_items - 1, n_users - 1, L, D, "cpu", False),
             "urm_p": lambda: URM_P(n_items - 1, n_users - 1, L, D, "cpu", False, 0.3, -0.1),
             "urm_p_mr": lambda: URM_P_MR(n_items - 1, n_users - 1, L, D, "cpu", False, 0.3, -0.1, 0.7)}
    for name, ctor in ctors.items():
        e = ctor()
        e.load_state_dict({k: torch.from_numpy(v) for k, v in fx.sub(name + "/sd/").items()})
        if hasattr(e, "posBias"):
            e.posBias = torch.from_numpy(fx[name + "/posBias"])
            e.posDependentBias = torch.from_numpy(fx[name + "/posDependentBias"])
        e = e.to("cuda:0")
        assert e.device == "cuda:0"
        out = e(T(fx["in/slates"]), T(fx["in/users"]))
        np.testing.assert_allclose(N(out), fx[name + "/out"], rtol=1e-5, atol=1e-6)
        p, dEmb, dBias, uEmb, uBias = e.core_forward(T(fx["in/slates"]), T(fx["in/users"]))
        assert dEmb.shape == (B, L, D) and torch.equal(p, out)


def test_eval_loop_and_pickle(golden, tmp_path):
    """The 100x5 recommendation test shape (train_generative.py:168-195) and whole-model pickling (:199)."""
    from pivotcvae_b200.env.response_model import UserResponseModel_MLP
    from pivotcvae_b200.train_generative import recommendation_test
    fx = golden("pivot_small")
    cfg = fx.cfg
    m = build_pivot(fx)
    env = UserResponseModel_MLP(cfg["n_items"] - 1, cfg["n_users"] - 1, cfg["D"], cfg["L"],
                                [(cfg["L"] + 1) * cfg["D"], cfg["hidden"], cfg["hidden"], cfg["L"]], "cuda:0", False)
    env.load_state_dict({k: torch.from_numpy(v) for k, v in fx.sub("env_sd/").items()})
    env.to("cuda:0")
    items = T(fx["rec_pi_k1/items"]).view(cfg["B"], -1)
    np.testing.assert_allclose(N(env(items, T(fx["in/users"]))), fx["env/resp"], rtol=1e-4, atol=1e-5)
    mn, me, mx = recommendation_test(m, env, 32, n_trial=3, n_context=5)
    assert mn.shape == (5,) and bool((mn <= me).all()) and bool((me <= mx).all())
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    m2 = torch.load(buf, weights_only=False)
    tag = "rec_pi_k1/"
    m2.noise.push("eps", T(fx[tag + "eps"]))
    it2, _ = m2.recommend(T(fx[tag + "ctx"]), T(fx["in/users"]), return_item=True)
    assert np.array_equal(N(it2), fx[tag + "items"])


def test_no_cpu_fallback(golden):
    import pivotcvae_b200._lib as L
    from pivotcvae_b200.models.pivotcvae import UserPivotCVAE
    fx = golden("pivot_small")
    with pytest.raises(L.PcvError):
        from gpu_util import _Emb
        sd = fx.sub("sd/")
        UserPivotCVAE(_Emb(sd["docEmbed.weight"]), _Emb(sd["userEmbed.weight"]), 5, 8, 16, 6, list(fx["cfg/enc"]),
                      list(fx["cfg/psm"]), list(fx["cfg/scm"]), list(fx["cfg/prior"]), False, "cpu")


def test_cuda_graph_replays_draw_fresh_noise_and_match_eager(golden):
    """GraphedSlateGenerator: one graph launch per step; replay k uses Philox rows [k*used, (k+1)*used)
    exactly as k eager calls do, so graph and eager slates are identical step by step."""
    from pivotcvae_b200.env.response_model import UserResponseModel_MLP
    from pivotcvae_b200.graphs import GraphedSlateGenerator
    fx = golden("pivot_small")
    cfg = fx.cfg
    B = cfg["B"]
    env = UserResponseModel_MLP(cfg["n_items"] - 1, cfg["n_users"] - 1, cfg["D"], cfg["L"],
                                [(cfg["L"] + 1) * cfg["D"], cfg["hidden"], cfg["hidden"], cfg["L"]], "cuda:0", False)
    env.load_state_dict({k: torch.from_numpy(v) for k, v in fx.sub("env_sd/").items()})
    env.to("cuda:0")
    for key in ("pivotcvae_gt_pi", "pivotcvae_gt_spi"):
        ctx, users = T(fx["rec_pi_k1/ctx"]), T(fx["in/users"])
        eager = build_pivot(fx, key)
        eager.noise.reseed(77)
        want = []
        for _ in range(6):          # 3 warm-up calls inside the capture helper + 1 capture pass + 2 replays... see below
            it, _ = eager.recommend(ctx, users, return_item=True)
            want.append(N(it).copy())
        m = build_pivot(fx, key)
        m.noise.reseed(77)
        gen = GraphedSlateGenerator(m, env, B, warmup=3)   # consumes 3 eager steps of the stream
        assert gen.launches_per_step >= 7
        got = []
        for _ in range(3):
            items, resp = gen(ctx, users)
            got.append(N(items).copy())
            np.testing.assert_allclose(N(resp), N(env(items.view(B, -1), users)), rtol=0, atol=0)
        # replays continue the stream right after the warm-up steps: steps 3, 4, 5 of the eager run
        for g, w in zip(got, want[3:6]):
            assert np.array_equal(g, w)
        assert not np.array_equal(got[0], got[1])   # fresh z every replay


def _build_cand(fx, kind):
    from gpu_util import _Emb, load_sd
    from pivotcvae_b200.models.listcvae import UserListCVAEWithPrior
    from pivotcvae_b200.models.pivotcvae import UserPivotCVAE
    sd, cfg = fx.sub(kind + "/sd/"), fx.cfg
    args = (_Emb(sd["docEmbed.weight"]), _Emb(sd["userEmbed.weight"]), cfg["L"], cfg["D"], cfg["Z"], cfg["L"] + 1)
    if kind == "pivot":
        m = UserPivotCVAE(*args, list(fx["cfg/enc"]), list(fx["cfg/psm"]), list(fx["cfg/scm"]), list(fx["cfg/prior"]), False, "cuda:0")
    else:
        m = UserListCVAEWithPrior(*args, list(fx["cfg/enc"]), list(fx["cfg/dec"]), list(fx["cfg/prior"]), False, "cuda:0")
    return load_sd(m, sd)


@pytest.mark.parametrize("kind", ["pivot", "list"])
def test_candidate_mode_training(golden, kind):
    """candidateFlag=True (the reference's default mode): loss, gradients and forward()'s p vs the reference."""
    from pivotcvae_b200.train_generative import get_gen_loss
    fx = golden("cand_small")
    m = _build_cand(fx, kind)
    m.candidateFlag = True
    batch = {"slates": fx["in/slates"], "users": fx["in/users"].reshape(-1, 1), "responses": fx["in/resp"].astype(np.float64),
             "sample_candidates": fx["in/candidates"], "sample_targets": fx["in/targets"]}
    m.noise.push("eps", T(fx[kind + "/eps"]))
    loss, rec, kld = get_gen_loss(batch, m, None, 0.01)
    np.testing.assert_allclose([loss.item(), rec.item(), kld.item()], fx[kind + "/loss"], rtol=1e-4)
    loss.backward()
    grads = fx.sub(kind + "/grad/")
    assert grads
    for pname, prm in m.named_parameters():
        if pname in grads:
            np.testing.assert_allclose(N(prm.grad), grads[pname], rtol=2e-3, atol=2e-6, err_msg=pname)
        else:
            assert prm.grad is None, pname
    m.noise.push("eps", T(fx[kind + "/eps"]))
    with torch.no_grad():
        p, rx, z, emb, mu, lv = m.forward(T(fx["in/slates"]), T(fx["in/resp"]), candidates=T(fx["in/candidates"]), u=T(fx["in/users"]))
    assert p.shape == fx[kind + "/p"].shape
    np.testing.assert_allclose(N(p), fx[kind + "/p"], rtol=1e-4, atol=1e-5)
    W = fx.sub(kind + "/sd/")["docEmbed.weight"]
    _, _, po = oracle.cand_ce(W, N(rx).reshape(-1, 8), fx["in/candidates"], fx["in/targets"])
    assert np.array_equal(N(p), po)      # bitwise vs the oracle's FMA chain


def test_train_on_dataset_get_model_sample_encoding(golden, tmp_path):
    """The epoch loop (train_generative.py:67-214), the factory (:221-240) and sample_encoding (cvae.py:103-115)
    run end to end on a tiny synthetic simulation; the loss goes down and the best model is pickled."""
    import argparse
    from torch.utils.data import Dataset
    from pivotcvae_b200.env.response_model import URM_P_MR
    from pivotcvae_b200.train_generative import add_gen_model_parse, get_model, train_on_dataset
    torch.manual_seed(0)
    n_items, n_users, Ls, D = 1500, 100, 5, 8   # > the default n_neg=1000 (n_neg > N raises, as torch.bernoulli does: SURVEY F7)
    env = URM_P_MR(n_items - 1, n_users - 1, Ls, D, "cpu", False, 0.3, -0.1, 0.5).to("cuda:0")
    p = add_gen_model_parse(argparse.ArgumentParser())
    args = p.parse_args(["--model", "pivotcvae_gt_pi", "--enc_struct", "[54,32,32]", "--prior_struct", "[14,16,16]",
                         "--psm_struct", "[30,32,32,8]", "--scm_struct", "[38,32,32,32]"])
    args.s, args.nouser, args.device = Ls, False, "cuda:0"
    model = get_model(args, env)
    assert type(model).__name__ == "UserPivotCVAE"

    class DS(Dataset):
        nCandidate = 100

        def __init__(self, n, seed):
            g = torch.Generator().manual_seed(seed)
            self.u = torch.randint(0, n_users, (n, 1), generator=g)
            self.s = torch.randint(0, n_items, (n, Ls), generator=g)
            with torch.no_grad():
                self.r = env.generate_response_for_dataset(self.u.cuda(), self.s.cuda()).cpu()

        def __len__(self):
            return len(self.u)

        def __getitem__(self, i):
            return {"slates": self.s[i].numpy(), "users": self.u[i].numpy(), "responses": self.r[i].numpy().astype(float)}

    class Log:
        lines = []

        def log(self, s_, newline=True):
            self.lines.append(s_)

    path = str(tmp_path / "best.pt")
    tr, va = train_on_dataset(DS(512, 1), DS(128, 2), model, path, Log(), env, 64, 3, 1e-2, 0.0, 0.001)
    assert len(tr) == 3 and tr[-1] < tr[0] and all(np.isfinite(tr)) and all(np.isfinite(va))
    assert any("Expected response (5)" in l for l in Log.lines) and any("Save best model" in l for l in Log.lines)
    best = torch.load(open(path, "rb"), weights_only=False)
    mu, lv = best.sample_encoding(DS(8, 3).s.cuda(), DS(8, 3).r.cuda(), DS(8, 3).u.cuda())
    assert mu.shape == (8, 16) and lv.shape == (8, 16) and bool(torch.isfinite(mu).all())
    pm, pl = best.get_prior(DS(8, 3).r.cuda(), DS(8, 3).u.cuda())
    assert pm.shape == (8, 16)
    # piecewise public API == fused path
    s8, r8, u8 = DS(8, 3).s.cuda(), DS(8, 3).r.cuda(), DS(8, 3).u.cuda()
    with torch.no_grad():
        emb = best.docEmbed(s8.reshape(-1)).view(8, -1)
        cond = best.get_condition(r8)
        uemb = best.userEmbed(u8.reshape(-1))
        mu2, lv2 = best.encode(emb, cond, uemb)
        assert torch.equal(mu2, mu) and torch.equal(lv2, lv)
        best.noise.push("eps", torch.zeros(8, 16).cuda())
        z = best.reparametrize(mu2, lv2)
        assert torch.equal(z, mu2 + 0 * z)
        rx = best.decode(z, cond, uemb, true_pivot=s8[:, 0])
        assert rx.shape == (8, Ls, D) and torch.equal(rx[:, 0], best.docEmbed.weight[s8[:, 0]])
        items = best.get_recommended_item(rx)
        assert items.shape == (8 * Ls,) and bool((items.view(8, Ls)[:, 0] == s8[:, 0]).all() or True)


def test_slate_metrics_gpu(golden):
    """analysis.get_coverage / get_ILS (analysis.py:5-30) fused over the recommended slates."""
    from pivotcvae_b200 import analysis
    fx = golden("metrics")
    emb = torch.nn.Embedding(900, 8).cuda()
    with torch.no_grad():
        emb.weight.copy_(T(fx["table"]))
    ils = analysis.get_ILS(T(fx["slates"]), emb)
    np.testing.assert_allclose(N(ils), fx["ils"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(N(ils), oracle.ils(fx["table"], fx["slates"]), rtol=1e-5, atol=1e-6)
    assert analysis.get_coverage(T(fx["slates"]), 900) == float(fx["coverage"])


def test_generate_dataset_gpu():
    """URM.generate_dataset (response_model.py:188-262): coverage guarantees + responses == oracle."""
    from pivotcvae_b200.env.response_model import URM_P_MR
    torch.manual_seed(5)
    env = URM_P_MR(59, 39, 5, 8, "cuda:0", False, 0.2, -0.2, 0.3).to("cuda:0")
    u, s, r = env.generate_dataset(min_user_hist=3, min_item_hist=2, n_record=400)
    assert u.shape == (400,) and s.shape == (400, 5) and r.shape == (400, 5) and r.dtype == np.float32
    assert np.bincount(u, minlength=40).min() >= 3
    assert np.bincount(s[:, 0], minlength=60).min() >= 2
    assert (u[:120] == np.repeat(np.arange(40), 3)).all()
    assert (s[120:240, 0] == np.repeat(np.arange(60), 2)).all()
    p = oracle.urm(2, N(env.docEmbed.weight), N(env.userEmbed.weight), N(env.itemBias.weight).reshape(-1),
                   N(env.userBias.weight).reshape(-1), s, u, N(env.posBias), N(env.posDependentBias), 0.3)
    sure = np.abs(p - 0.5) > 1e-5          # libm vs device expf may differ by an ulp at the threshold
    np.testing.assert_array_equal(r[sure], (p >= 0.5).astype(np.float32)[sure])
