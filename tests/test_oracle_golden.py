"""Pins oracle/ (the CPU restatement) against the reference's own outputs
(tests/golden/*.npz, produced by tests/golden/make_golden.py from /root/reference)
and against the Random123 Philox known-answer vectors.  CPU only."""
import numpy as np
import pytest

import oracle

PIVOT_FIXTURES = ["pivot_small", "pivot_small_nouser", "pivot_c1"]
LIST_FIXTURES = ["list_small", "list_small_user"]
TOL = dict(rtol=2e-5, atol=2e-6)


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 with 10 rounds
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
         [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for ctr, key, want in kat:
        got = oracle.philox4x32_10(ctr, key)
        assert [int(v) for v in got] == want


def test_expf_accuracy():
    x = np.linspace(-80, 80, 200001).astype(np.float32)
    got = oracle.expf(x).astype(np.float64)
    ref = np.exp(x.astype(np.float64))
    assert np.max(np.abs(got - ref) / ref) < 2.5e-7


def test_bitmask_pack_roundtrip():
    rng = np.random.default_rng(0)
    m = rng.random((5, 77)) < 0.3
    b = oracle.pack_bitmask(m)
    un = ((b[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(5, -1)[:, :77].astype(bool)
    assert np.array_equal(un, m)


def test_bernoulli_bitmask_rate():
    bits = oracle.bernoulli_bitmask(7, 0, 16, 4096, 0.25)
    rate = np.unpackbits(bits.view(np.uint8)).mean()
    assert abs(rate - 0.25) < 0.01


@pytest.mark.parametrize("D", [4, 8, 16, 32, 64, 128])
def test_score_select_matches_torch_mm_max(golden, D):
    """cvae.py:97-101 / pivotcvae.py:191: indices bit-exact incl. exact ties; logits bitwise at D=8."""
    fx = golden("dims")
    W, Q = fx["d%d/W" % D], fx["d%d/Q" % D]
    idx, val = oracle.score_select(W, Q)
    assert np.array_equal(idx, fx["d%d/idx" % D])
    assert idx[0] == 13  # duplicated rows 13 / 700 / 1499: lowest index wins
    p = oracle.score_logits(W, Q[:4])
    if D == 8:
        assert np.array_equal(p, fx["d%d/p_rows" % D])  # sequential-k FMA chain
        assert np.array_equal(val, fx["d%d/val" % D])
    else:
        np.testing.assert_allclose(p, fx["d%d/p_rows" % D], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", PIVOT_FIXTURES)
def test_pivot_recommend_greedy(golden, name):
    fx = golden(name)
    sd, cfg = fx.sub("sd/"), fx.cfg
    for k in (1, cfg["L"]):
        tag = "rec_pi_k%d/" % k
        out = oracle.pivot_recommend(sd, fx[tag + "ctx"], fx["in/users"], fx[tag + "eps"], cfg["no_user"], "max")
        np.testing.assert_allclose(out["z_mu"], fx[tag + "z_mu"], **TOL)
        np.testing.assert_allclose(out["rx"], fx[tag + "rx"], rtol=1e-4, atol=1e-5)
        assert np.array_equal(out["items"], fx[tag + "items"])  # slates bit-exact


@pytest.mark.parametrize("name", PIVOT_FIXTURES)
def test_pivot_recommend_sampled(golden, name):
    fx = golden(name)
    sd, cfg = fx.sub("sd/"), fx.cfg
    tag = "rec_spi_k2/"
    out = oracle.pivot_recommend(sd, fx[tag + "ctx"], fx["in/users"], fx[tag + "eps"], cfg["no_user"], "sample",
                                 noise=fx[tag + "noise"])
    assert np.array_equal(out["items"], fx[tag + "items"])
    np.testing.assert_allclose(out["rx"], fx[tag + "rx"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", PIVOT_FIXTURES)
@pytest.mark.parametrize("train", ["gt", "pt", "spt", "sgt"])
def test_pivot_forward_and_loss(golden, name, train):
    fx = golden(name)
    sd, cfg = fx.sub("sd/"), fx.cfg
    how = {"gt": "gt", "pt": "max", "spt": "sample", "sgt": "sample_gt"}[train]
    tag = "fwd_%s/" % train
    noise = fx[tag + "noise"] if (tag + "noise") in fx else None
    f = oracle.pivot_forward(sd, fx["in/slates"], fx["in/resp"], fx["in/users"], fx[tag + "eps"], cfg["no_user"], how, noise)
    np.testing.assert_allclose(f["z_mu"], fx[tag + "z_mu"], **TOL)
    np.testing.assert_allclose(f["z_logvar"], fx[tag + "z_logvar"], **TOL)
    np.testing.assert_allclose(f["z"], fx[tag + "z"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(f["rx"], fx[tag + "rx"], rtol=1e-4, atol=1e-5)
    assert np.array_equal(f["emb"], fx[tag + "emb"])
    W = sd["docEmbed.weight"]
    q = f["rx"].reshape(-1, cfg["D"])
    n = fx[tag + "p_rows"].shape[0]
    np.testing.assert_allclose(oracle.score_logits(W, q[:n]), fx[tag + "p_rows"], rtol=1e-4, atol=1e-5)
    for ntag in ("mask", "full"):
        t2 = "loss_%s_%s/" % (train, ntag)
        noise = fx[t2 + "noise"] if (t2 + "noise") in fx else None
        loss, rec, kld = oracle.gen_loss(sd, fx["in/slates"], fx["in/resp"], fx["in/users"], fx[t2 + "eps"],
                                         cfg["no_user"], 0.01, fx[t2 + "bitmask"], "pivot", how, noise)
        want = fx[t2 + "loss"]
        np.testing.assert_allclose([loss, rec, kld], want, rtol=1e-4)


@pytest.mark.parametrize("name", LIST_FIXTURES)
def test_list_paths(golden, name):
    fx = golden(name)
    sd, cfg = fx.sub("sd/"), fx.cfg
    for k in (1, 3):
        tag = "rec_k%d/" % k
        out = oracle.list_recommend(sd, fx[tag + "ctx"], fx["in/users"], fx[tag + "eps"], cfg["no_user"])
        np.testing.assert_allclose(out["z_mu"], fx[tag + "z_mu"], **TOL)
        np.testing.assert_allclose(out["rx"], fx[tag + "rx"], rtol=1e-4, atol=1e-5)
        assert np.array_equal(out["items"], fx[tag + "items"])
    f = oracle.list_forward(sd, fx["in/slates"], fx["in/resp"], fx["in/users"], fx["fwd/eps"], cfg["no_user"])
    np.testing.assert_allclose(f["rx"], fx["fwd/rx"], rtol=1e-4, atol=1e-5)
    for ntag in ("mask", "full"):
        t2 = "loss_%s/" % ntag
        got = oracle.gen_loss(sd, fx["in/slates"], fx["in/resp"], fx["in/users"], fx[t2 + "eps"], cfg["no_user"],
                              0.01, fx[t2 + "bitmask"], "list")
        np.testing.assert_allclose(got, fx[t2 + "loss"], rtol=1e-4)


def test_ce_gradient_matches_autograd(golden):
    """dq from the oracle == the gradient torch autograd pushed into scm's last bias
    (a direct readout of sum_rows dLoss/dq for slots 1..L-1)."""
    fx = golden("pivot_small")
    sd, cfg = fx.sub("sd/"), fx.cfg
    t2 = "loss_gt_mask/"
    f = oracle.pivot_forward(sd, fx["in/slates"], fx["in/resp"], fx["in/users"], fx[t2 + "eps"], cfg["no_user"], "gt")
    W = sd["docEmbed.weight"]
    L, D, B = cfg["L"], cfg["D"], cfg["B"]
    _, _, dq = oracle.ce(W, f["rx"].reshape(-1, D), fx["in/slates"].reshape(-1), fx[t2 + "bitmask"])
    dq = dq.reshape(B, L, D) / (B * L)
    n_scm = oracle._count_layers(sd, "scm")
    want = fx[t2 + "grad/scm_%d.bias" % n_scm]
    got = dq[:, 1:, :].reshape(B, -1).sum(0)
    np.testing.assert_allclose(got, want, rtol=2e-4, atol=1e-7)
    assert (t2 + "nograd/psm_1.weight") in fx  # PSM never receives a gradient


def test_response_models(golden):
    fx = golden("env_small")
    slates, users = fx["in/slates"], fx["in/users"]
    for tag, nu in (("mlp_user/", False), ("mlp_nouser/", True)):
        got = oracle.resp_mlp(fx.sub(tag + "sd/"), slates, users, nu)
        np.testing.assert_allclose(got, fx[tag + "out"], rtol=1e-4, atol=1e-5)
    for v, name in enumerate(["urm", "urm_p", "urm_p_mr"]):
        sd = fx.sub(name + "/sd/")
        kw = {}
        if v >= 1:
            kw = dict(pos_bias=fx[name + "/posBias"], pos_dep=fx[name + "/posDependentBias"])
        if v == 2:
            kw["mr_factor"] = float(fx[name + "/mrFactor"])
        got = oracle.urm(v, sd["docEmbed.weight"], sd["userEmbed.weight"], sd["itemBias.weight"], sd["userBias.weight"],
                         slates, users, **kw)
        np.testing.assert_allclose(got, fx[name + "/out"], rtol=1e-5, atol=1e-6)


def test_env_scorer_on_recommended_slates(golden):
    fx = golden("pivot_small")
    cfg = fx.cfg
    items = fx["rec_pi_k1/items"].reshape(cfg["B"], -1)
    got = oracle.resp_mlp(fx.sub("env_sd/"), items, fx["in/users"], cfg["no_user"])
    np.testing.assert_allclose(got, fx["env/resp"], rtol=1e-4, atol=1e-5)


def test_threads_do_not_change_results(golden):
    fx = golden("dims")
    W, Q = fx["d8/W"], fx["d8/Q"]
    a = oracle.score_select(W, Q)
    oracle.set_threads(4)
    try:
        b = oracle.score_select(W, Q)
    finally:
        oracle.set_threads(1)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("kind", ["pivot", "list"])
def test_candidate_mode_loss(golden, kind):
    """Sampled soft-max training (the reference's default, train_generative.py:52-56): candidates drawn
    by the reference's own Dataset sampler (data_loader.py:49-58)."""
    fx = golden("cand_small")
    sd = fx.sub(kind + "/sd/")
    got = oracle.gen_loss_candidates(sd, fx["in/slates"], fx["in/resp"], fx["in/users"], fx[kind + "/eps"], False, 0.01,
                                     fx["in/candidates"], fx["in/targets"], kind)
    np.testing.assert_allclose(got, fx[kind + "/loss"], rtol=1e-4)
    f = (oracle.pivot_forward(sd, fx["in/slates"], fx["in/resp"], fx["in/users"], fx[kind + "/eps"], False, "gt")
         if kind == "pivot" else oracle.list_forward(sd, fx["in/slates"], fx["in/resp"], fx["in/users"], fx[kind + "/eps"], False))
    W = sd["docEmbed.weight"]
    _, _, p = oracle.cand_ce(W, f["rx"].reshape(-1, 8), fx["in/candidates"], fx["in/targets"])
    np.testing.assert_allclose(p, fx[kind + "/p"], rtol=1e-4, atol=1e-5)


def test_slate_metrics(golden):
    """analysis.py:5-30: coverage and intra-list similarity."""
    fx = golden("metrics")
    np.testing.assert_allclose(oracle.ils(fx["table"], fx["slates"]), fx["ils"], rtol=1e-5, atol=1e-6)
    assert abs(oracle.ils(fx["table"], fx["slates"])[3] - 1.0) < 1e-6
    assert oracle.coverage(fx["slates"], 900) == float(fx["coverage"])


def test_oracle_topk_and_no_repeat_restatement():
    """score_topk == torch.topk semantics on distinct scores and first-index order on ties; no-repeat = sequential
    arg-max without replacement (an extension: the reference itself never masks)."""
    import torch
    rng = np.random.default_rng(3)
    W = rng.standard_normal((400, 8)).astype(np.float32)
    W /= np.linalg.norm(W, axis=1, keepdims=True)
    Q = rng.standard_normal((30, 8)).astype(np.float32)
    idx, val = oracle.score_topk(W, Q, 6)
    p = torch.from_numpy(oracle.score_logits(W, Q))
    tv, ti = torch.topk(p, 6, dim=1)
    assert np.array_equal(idx, ti.numpy()) and np.array_equal(val, tv.numpy())
    W[10] = W[3]
    W[200] = W[3]
    Q[0] = W[3]
    idx, _ = oracle.score_topk(W, Q, 3)
    assert list(idx[0]) == [3, 10, 200]
    Q[1], Q[2] = Q[0], Q[0]
    items, _ = oracle.slate_no_repeat(W, Q, 5)
    assert list(items[:3]) == [3, 10, 200]
    top1 = oracle.score_select(W, Q[5:10])[0]
    if len(set(top1)) == 5:        # a slate without duplicates keeps the reference's independent picks
        assert np.array_equal(items[5:10], top1)


@pytest.mark.parametrize("tag,no_user", [("mlp_user", False), ("mlp_nouser", True)])
def test_oracle_response_pretrain_step_vs_reference(golden, tag, no_user):
    """the oracle's restatement of one pretrain_env.py training step == the reference's autograd."""
    fx = golden("n4")
    loss, pred, grads = oracle.resp_train_step(fx.sub(tag + "/sd/"), fx[tag + "/slates"], fx[tag + "/users"], fx[tag + "/resp"], no_user)
    assert abs(loss - float(fx[tag + "/loss"])) <= 1e-6
    np.testing.assert_allclose(pred, fx[tag + "/pred"], rtol=1e-5, atol=1e-6)
    ref = fx.sub(tag + "/grad/")
    assert set(ref) == set(grads)
    for k, v in ref.items():
        np.testing.assert_allclose(grads[k], v, rtol=1e-4, atol=1e-8, err_msg=k)


def test_oracle_mf_scores_vs_reference(golden):
    fx = golden("n4")
    p = oracle.mf_scores(fx["mf/doc"], fx["mf/usr"], fx["mf/doc_bias"], fx["mf/user_bias"], fx["mf/users"])
    np.testing.assert_allclose(p, fx["mf/p_all"], rtol=1e-5, atol=1e-6)
    top = np.argsort(-p, axis=1, kind="stable")[:, :fx["mf/items"].shape[1]]
    assert np.array_equal(top, fx["mf/items"])
