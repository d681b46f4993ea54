"""Golden fixture for SURVEY §8f N4, from the UNMODIFIED reference:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_n4.py

* the training step of pretrain_env.py:76-88 — UserResponseModel_MLP.forward (env/response_model.py:76-87) ->
  BCELoss(sigmoid(pred), responses) -> backward: loss, logits and EVERY parameter gradient incl. the two embedding
  tables (with and without user);
* models/deterministic.py MF: forward() on slates and recommend() = torch.topk of the biased scores.
-> tests/golden/n4.npz
"""
import contextlib
import io
import os

import numpy as np
import torch

import make_golden as mg  # noqa: F401  (sets sys.path to the reference, stubs matplotlib, single thread)

with contextlib.redirect_stdout(io.StringIO()):
    from env.response_model import UserResponseModel_MLP
    from models.deterministic import MF

HERE = os.path.dirname(os.path.abspath(__file__))
out = {}
g = torch.Generator().manual_seed(77)
n_items, n_users, Ls, D, B = 300, 40, 5, 8, 96
for tag, no_user in (("mlp_user", False), ("mlp_nouser", True)):
    torch.manual_seed(5 if no_user else 4)
    with contextlib.redirect_stdout(io.StringIO()):
        m = UserResponseModel_MLP(n_items - 1, n_users - 1, D, Ls, [(Ls if no_user else Ls + 1) * D, 64, 32, Ls], "cpu", no_user)
    slates = torch.randint(0, n_items, (B, Ls), generator=g)
    users = torch.randint(0, n_users, (B,), generator=g)
    resp = (torch.rand(B, Ls, generator=g) < 0.4).float()
    pred = m.forward(slates, users)
    loss = torch.nn.BCELoss()(torch.sigmoid(pred.reshape(-1)), resp.reshape(-1))
    loss.backward()
    for k, v in m.state_dict().items():
        out["%s/sd/%s" % (tag, k)] = v.detach().numpy().copy()
    for k, p in m.named_parameters():
        out["%s/grad/%s" % (tag, k)] = p.grad.detach().numpy().copy()
    out[tag + "/slates"], out[tag + "/users"], out[tag + "/resp"] = slates.numpy(), users.numpy(), resp.numpy()
    out[tag + "/pred"], out[tag + "/loss"] = pred.detach().numpy(), np.float32(loss.item())

# biased MF: forward + top-k recommendation
torch.manual_seed(9)
class _E:  # noqa: E302
    def __init__(self, w):
        self.weight = w
docw, usrw = torch.randn(n_items, D, generator=g), torch.randn(n_users, D, generator=g)
with contextlib.redirect_stdout(io.StringIO()):
    mf = MF(_E(docw), _E(usrw), Ls, D, "cpu")
with torch.no_grad():
    mf.userBias.weight.copy_(0.1 * torch.randn(n_users, 1, generator=g))
    mf.docBias.weight.copy_(0.1 * torch.randn(n_items, 1, generator=g))
    users = torch.randint(0, n_users, (32,), generator=g)
    slates = torch.randint(0, n_items, (32, Ls), generator=g)
    pred = mf.forward(slates, None, u=users)
    items, _ = mf.recommend(None, u=users, return_item=True)
    p_all = torch.stack([mf.point_forward(users[i], torch.arange(n_items)) for i in range(32)])
out["mf/doc"], out["mf/usr"] = docw.numpy(), usrw.numpy()
out["mf/user_bias"], out["mf/doc_bias"] = mf.userBias.weight.detach().numpy(), mf.docBias.weight.detach().numpy()
out["mf/users"], out["mf/slates"] = users.numpy(), slates.numpy()
out["mf/pred"], out["mf/items"], out["mf/p_all"] = pred.numpy(), items.numpy(), p_all.numpy()
out["cfg"] = np.array([n_items, n_users, Ls, D, B])
np.savez_compressed(os.path.join(HERE, "n4.npz"), **out)
print("n4.npz", os.path.getsize(os.path.join(HERE, "n4.npz")), "bytes; loss", out["mlp_user/loss"], out["mlp_nouser/loss"])
