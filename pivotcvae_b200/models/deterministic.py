"""Ranking baselines — the inference half of the reference's models/deterministic.py.

RankModel (:13-72) and the biased MF (:74-124).  MF.recommend in the reference scores every user against every
item with a Python loop over users and calls torch.topk (:114-124); here the biased score
<u, w_j> + b_u + b_j is one exact streaming top-k over an augmented table [w_j | b_j] (pcv_score_topk, the same
fp32 sequential-k FMA chain as every other exact path; b_u is constant per row and does not change the ranking),
and forward() is the fused gather-dot kernel.  Training these baselines is out of scope.
"""
import torch
from torch import nn

from .. import _lib as L
from .. import ops
from .cvae import _require_cuda


class RankModel(nn.Module):
    def __init__(self, embeddings, u_embeddings, slate_size, feature_size, device, fine_tune=True):
        super().__init__()
        assert embeddings.weight.shape[1] == feature_size
        assert u_embeddings.weight.shape[1] == feature_size
        self.candidateFlag = False
        self.slate_size, self.feature_size, self.device = slate_size, feature_size, device
        dev = _require_cuda(device)
        with torch.no_grad():
            src = embeddings.weight.detach().to(dev, torch.float32)
            self.docEmbed = nn.Embedding(src.shape[0], src.shape[1], device=dev)
            self.docEmbed.weight.data.copy_(ops.normalize_rows(src))
            self.docEmbed.weight.requires_grad = fine_tune
            usrc = u_embeddings.weight.detach().to(dev, torch.float32)
            self.userEmbed = nn.Embedding(usrc.shape[0], usrc.shape[1], device=dev)
            self.userEmbed.weight.data.copy_(ops.normalize_rows(usrc))
            self.userEmbed.weight.requires_grad = fine_tune
        self.m = nn.Sigmoid()
        self._table = None

    def get_recommended_item(self, embeddings):
        """arg-max item per row (deterministic.py:64-68)."""
        t = self._plain_table()
        return ops.score_select(t, embeddings.reshape(-1, self.feature_size).detach(), "greedy", want_val=False)[0]

    def _plain_table(self):
        w = self.docEmbed.weight
        t = getattr(self, "_plain", None)
        if t is None or t.weight.data_ptr() != w.data_ptr() or t.version != w._version:
            t = ops.Table(w.detach())
            t.version = w._version
            self._plain = t
        return t

    def __getstate__(self):
        st = self.__dict__.copy()
        st["_table"] = None
        st.pop("_plain", None)
        return st

    def log(self, logger):
        logger.log("\tfeature size: " + str(self.feature_size))
        logger.log("\tslate size: " + str(self.slate_size))
        logger.log("\tdevice: " + str(self.device))


class MF(RankModel):
    """Biased MF (deterministic.py:74-124)."""

    def __init__(self, embeddings, u_embeddings, slate_size, feature_size, device, fine_tune=True):
        super().__init__(embeddings, u_embeddings, slate_size, feature_size, device, fine_tune=fine_tune)
        dev = self.docEmbed.weight.device
        self.userBias = nn.Embedding(self.userEmbed.weight.shape[0], 1, device=dev)
        self.userBias.weight.data.fill_(0.001)
        self.docBias = nn.Embedding(self.docEmbed.weight.shape[0], 1, device=dev)
        self.docBias.weight.data.fill_(0.001)

    def _aug(self):
        """[w_j | b_j | 0-pad] table handle (rebuilt when the embeddings or the biases change) and its width."""
        w, b = self.docEmbed.weight, self.docBias.weight
        key = (w.data_ptr(), w._version, b.data_ptr(), b._version)
        hit = self._table
        if hit is None or hit[0] != key:
            D = self.feature_size
            Da = next(d for d in (4, 8, 16, 32, 64, 128) if d >= D + 1)       # the widths the exact engines are built for
            aug = torch.zeros(w.shape[0], Da, dtype=torch.float32, device=w.device)
            aug[:, :D] = w.detach()
            aug[:, D] = b.detach().reshape(-1)
            hit = (key, ops.Table(aug), Da)
            self._table = hit
        return hit[1], hit[2]

    def _queries(self, users):
        users = users.to(self.docEmbed.weight.device, torch.int64).reshape(-1)
        table, Da = self._aug()
        q = torch.zeros(users.shape[0], Da, dtype=torch.float32, device=users.device)
        q[:, :self.feature_size] = self.userEmbed.weight.detach()[users]
        q[:, self.feature_size] = 1.0
        return table, q, users

    def point_forward(self, users, items):
        """<u, w_i> + b_u + b_i for paired (user, item) ids (deterministic.py:97-112)."""
        table, q, users = self._queries(users)
        items = items.to(q.device, torch.int64).reshape(-1, 1)
        tp = torch.zeros(items.shape[0], dtype=torch.int64, device=q.device)
        _, _, _, p = ops.cand_ce_fwd_bwd(table, q, items, tp, want_dq=False, want_logits=True)
        return p.reshape(-1) + self.userBias.weight.detach()[users].reshape(-1)

    def forward(self, s, r, candidates=None, u=None):
        """pred[:, i] = point_forward(u, s[:, i]) (deterministic.py:56-60) as one gather-dot launch."""
        table, q, users = self._queries(u)
        s = s.to(q.device, torch.int64)
        tp = torch.zeros(s.shape[0], dtype=torch.int64, device=q.device)
        _, _, _, p = ops.cand_ce_fwd_bwd(table, q, s, tp, want_dq=False, want_logits=True)
        return p + self.userBias.weight.detach()[users].reshape(-1, 1)

    def recommend(self, r, u=None, return_item=False):
        """top-`slate_size` items per user (deterministic.py:114-124)."""
        table, q, _ = self._queries(u)
        if self.slate_size > 16:
            raise L.PcvError("MF.recommend: slate_size > 16 is not supported by pcv_score_topk")
        recItems, _ = ops.score_topk(table, q, self.slate_size)
        if return_item:
            return recItems, None
        rx = self.docEmbed.weight.detach()[recItems].reshape(-1, self.slate_size * self.feature_size)
        return rx, None
