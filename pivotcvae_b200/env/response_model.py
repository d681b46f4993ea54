"""User-response simulators — drop-in for the reference's env/response_model.py.

sample_users (:10-13), Environment (:15-38), UserResponseModel_MLP (:47-87),
URM (:97-154), URM_P (:264-302), URM_P_MR (:305-323).  Forward passes run in
libpcv_b200 (fused gather+normalise+MLP, gather-plus-dot); construction and the
parameter containers stay torch so reference checkpoints load.
"""
import math

import torch
from torch import nn

from .. import _lib as L
from .. import ops
from ..models.cvae import with_mlp_engine


def sample_users(environment, batch_size):
    """Uniform user ids (response_model.py:10-13).  The reference draws on the CPU with
    torch.multinomial(ones) and copies; here the draw happens on the device."""
    dev = torch.device(environment.device)
    return torch.randint(0, environment.maxUserId + 1, (batch_size,), device=dev, dtype=torch.int64)


class Environment(nn.Module):
    def __init__(self, maxIID, maxUID, f_size, s_size, device, no_user):
        super().__init__()
        self.maxItemId, self.maxUserId = maxIID, maxUID
        self.featureSize, self.slateSize = f_size, s_size
        self.device = device
        self.noUser = no_user
        a = math.sqrt(2.0 / f_size)
        self.docEmbed = nn.Embedding(maxIID + 1, f_size)
        self.docEmbed.weight.data.uniform_(-a, a)
        if not no_user:
            self.userEmbed = nn.Embedding(maxUID + 1, f_size)
            self.userEmbed.weight.data.uniform_(-a, a)

    def _ids(self, slates, users):
        dev = self.docEmbed.weight.device
        if dev.type != "cuda":
            raise L.PcvError("response models run on a B200 only: move the module with .to('cuda:0')")
        slates = slates.to(dev, torch.int64)
        users = users.to(dev, torch.int64).reshape(-1) if users is not None else None
        return slates, users


class UserResponseModel_MLP(Environment):
    def __init__(self, maxIID, maxUID, f_size, s_size, struct, device, no_user):
        super().__init__(maxIID, maxUID, f_size, s_size, device, no_user)
        assert struct[0] == (s_size if no_user else s_size + 1) * f_size
        assert struct[-1] == s_size
        self.mlp = []
        for i in range(len(struct) - 1):
            m = nn.Linear(struct[i], struct[i + 1])
            nn.init.kaiming_uniform_(m.weight)
            self.mlp.append(m)
            self.add_module("mlp_%d" % (i + 1), m)

    mlp_engine = None   # None = ops.MLP_ENGINE; "tc" = the tcgen05 engine of the fused block (models/cvae.py)

    @with_mlp_engine
    def forward(self, slates, users):
        """gather L rows -> L2-normalise the flattened slate vector -> [+ normalised user row]
        -> Linear/ReLU chain -> (B, L) logits (response_model.py:76-87); one fused kernel."""
        slates, users = self._ids(slates, users)
        B = slates.shape[0]
        if getattr(self, "differentiable", False) and torch.is_grad_enabled():
            return self._forward_train(slates.reshape(B, -1), users)
        segs = [ops.Gather(self.docEmbed.weight.detach(), slates.reshape(B, -1), normalize=True)]
        if not self.noUser:
            segs.append(ops.Gather(self.userEmbed.weight.detach(), users.reshape(B, 1), normalize=True))
        n = len(self.mlp)
        layers = [(m.weight.detach(), m.bias.detach(), L.ACT_RELU if i < n - 1 else L.ACT_NONE)
                  for i, m in enumerate(self.mlp)]
        return ops.mlp_forward(segs, layers, B)["out"]


    # pretrain_env.train_response_model sets `differentiable = True`: forward() then builds the autograd graph (embedding
    # tables included).  Off by default: the generative path only ever scores slates with a frozen simulator, and the
    # fused inference block is the one that is bit-identical to the oracle.
    differentiable = False

    def _forward_train(self, slates, users):
        """Differentiable path (pretrain_env.py:76-88 trains the embeddings too): gather + normalise with a backward,
        then the fused MLP block with its tensor-core backward."""
        from ..autograd import FusedMLPFn, GatherNormFn, MlpSpec
        x0 = GatherNormFn.apply(self.docEmbed.weight, None if self.noUser else self.userEmbed.weight, slates, users)
        n = len(self.mlp)
        spec = MlpSpec([("dense", 0)], [L.ACT_RELU if i < n - 1 else L.ACT_NONE for i in range(n)], save=True)
        flat = []
        for m in self.mlp:
            flat += [m.weight, m.bias]
        return FusedMLPFn.apply(spec, slates.shape[0], 1, x0, *flat)


class URM(Environment):
    variant = L.URM

    def __init__(self, maxIID, maxUID, slate_size, latent_size, device, no_user):
        super().__init__(maxIID, maxUID, latent_size, slate_size, device, no_user)
        assert not no_user
        self.maxIID = torch.tensor(maxIID)
        self.maxUID = torch.tensor(maxUID)
        self.itemBias = nn.Embedding(maxIID + 1, 1)
        self.itemBias.weight.data.zero_()
        self.userBias = nn.Embedding(maxUID + 1, 1)
        self.userBias.weight.data.zero_()
        self.m = nn.Sigmoid()

    def to(self, *args, **kwargs):
        out = super().to(*args, **kwargs)
        out.device = args[0] if args else kwargs.get("device", out.device)
        return out

    def _extra(self):
        return {}

    def forward(self, slates, users):
        """(B, L) click scores: sigmoid(<norm(d), u> + b_i + b_u) [+ positional / relation terms]."""
        slates, users = self._ids(slates, users)
        return ops.urm_forward(self.variant, self.docEmbed.weight.detach(), self.userEmbed.weight.detach(),
                               self.itemBias.weight.detach(), self.userBias.weight.detach(), slates, users,
                               **self._extra())

    def core_forward(self, slates, users):
        """5-tuple of the reference (response_model.py:129-150); the auxiliary tensors are plain gathers."""
        slates_d, users_d = self._ids(slates, users)
        p = self.forward(slates, users)
        B = slates_d.shape[0]
        dEmb = torch.nn.functional.normalize(self.docEmbed.weight[slates_d], p=2, dim=-1).view(B, self.slateSize, -1)
        dBias = self.itemBias.weight[slates_d].view(B, self.slateSize)
        return p, dEmb, dBias, self.userEmbed.weight[users_d], self.userBias.weight[users_d]

    def sample_response(self, slate_p):
        """threshold at 0.5 (response_model.py:156-169)."""
        return (slate_p >= 0.5).to(slate_p.dtype)

    def generate_response_for_dataset(self, sampledU, sampledSlates):
        """One launch over all records (the reference loops 128 at a time, :170-185)."""
        with torch.no_grad():
            return self.sample_response(self.forward(sampledSlates, sampledU.reshape(-1)))

    def generate_dataset(self, min_user_hist=20, min_item_hist=20, n_record=1000000):
        """Simulated (user, slate, responses) records (response_model.py:188-262): every user gets
        >= min_user_hist slates, every item leads >= min_item_hist slates, the rest are uniform.
        The reference draws ids with torch.multinomial(ones, replacement=True) on the CPU, i.e.
        uniformly; here they are drawn uniformly on the device and every record is scored by one
        pcv_urm_fwd launch.  Returns numpy (users (R,), slates (R, L), responses (R, L))."""
        nU, nI, Ls = int(self.maxUID) + 1, int(self.maxIID) + 1, self.slateSize
        assert min_user_hist * (nU - 1) + min_item_hist * (nI - 1) < n_record
        n_record = max(n_record, nU * min_user_hist, nI * min_item_hist)
        dev = self.docEmbed.weight.device
        if dev.type != "cuda":
            raise L.PcvError("response models run on a B200 only: move the module with .to('cuda:0')")
        users, slates = [], []
        if min_user_hist > 0:
            users.append(torch.arange(nU, device=dev).repeat_interleave(min_user_hist))
            slates.append(torch.randint(0, nI, (nU * min_user_hist, Ls), device=dev))
        if min_item_hist > 0:
            users.append(torch.randint(0, nU, (nI * min_item_hist,), device=dev))
            sl = torch.randint(0, nI, (nI, min_item_hist, Ls), device=dev)
            sl[:, :, 0] = torch.arange(nI, device=dev).view(-1, 1)
            slates.append(sl.reshape(-1, Ls))
        done = sum(len(u) for u in users)
        if done < n_record:
            users.append(torch.randint(0, nU, (n_record - done,), device=dev))
            slates.append(torch.randint(0, nI, (n_record - done, Ls), device=dev))
        # the reference sizes its buffers to n_record and lets the guaranteed blocks overflow it
        # only when the assert above already failed, so concatenation yields the same layout
        genU = torch.cat(users)[:n_record]
        genS = torch.cat(slates)[:n_record]
        genR = self.generate_response_for_dataset(genU, genS).to(torch.float32)
        return genU.cpu().numpy(), genS.cpu().numpy(), genR.cpu().numpy()


class URM_P(URM):
    variant = L.URM_P

    def __init__(self, maxIID, maxUID, slate_size, latent_size, device, no_user, p_bias_max, p_bias_min):
        super().__init__(maxIID, maxUID, slate_size, latent_size, device, no_user)
        self.p_bias_max, self.p_bias_min = p_bias_max, p_bias_min
        self.posBias = torch.tensor([p_bias_max - i * (p_bias_max - p_bias_min) / slate_size
                                     for i in range(slate_size)])
        a = math.sqrt(0.5 / latent_size)
        self.posDependentBias = torch.FloatTensor(slate_size * latent_size).uniform_(-a, a).reshape(slate_size, latent_size)

    def to(self, *args, **kwargs):
        out = super().to(*args, **kwargs)
        out.posBias = out.posBias.to(args[0])
        out.posDependentBias = out.posDependentBias.to(args[0])
        return out

    def _extra(self):
        dev = self.docEmbed.weight.device
        return dict(pos_bias=self.posBias.to(dev), pos_dep=self.posDependentBias.to(dev))


class URM_P_MR(URM_P):
    variant = L.URM_P_MR

    def __init__(self, maxIID, maxUID, slate_size, latent_size, device, no_user, p_bias_max, p_bias_min, mr_factor):
        super().__init__(maxIID, maxUID, slate_size, latent_size, device, no_user, p_bias_max, p_bias_min)
        self.mrFactor = mr_factor

    def _extra(self):
        e = super()._extra()
        e["mr_factor"] = float(self.mrFactor)
        return e
