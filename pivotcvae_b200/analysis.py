"""Slate metrics — drop-in for the reference's analysis.py:5-30 (coverage, intra-list similarity).

Both run in one fused kernel pass over the recommended slates: one thread
per slate gathers the L embedding rows, normalises them and reduces |sum_i e_i|^2; the same
pass ORs the item ids into a coverage bitmap."""

import torch

from . import _lib as L
from .ops import _f32, _i64, _ptr, _stream


def _metrics(slates, weight, want_ils, want_cov):
    weight = _f32(weight, "embedding table")
    slates = _i64(slates, "slates")
    B, Ls = slates.shape
    n_items, D = weight.shape
    ils = torch.empty(B, dtype=torch.float32, device=weight.device) if want_ils else None
    bitmap = torch.zeros((n_items + 31) // 32, dtype=torch.int32, device=weight.device) if want_cov else None
    with torch.cuda.device(weight.device):
        L.check(L.load().pcv_slate_metrics(_ptr(weight), D, _ptr(slates), B, Ls, _ptr(ils), _ptr(bitmap), _stream()),
                "pcv_slate_metrics")
        count = None
        if want_cov:
            count = torch.zeros(1, dtype=torch.int64, device=weight.device)
            L.check(L.load().pcv_popcount(_ptr(bitmap), bitmap.numel(), _ptr(count), _stream()), "pcv_popcount")
    return ils, count


def get_coverage(slates, N):
    """|unique(slates)| / N (analysis.py:5-11)."""
    slates = _i64(slates, "slates")
    dummy = torch.zeros(int(N), 4, device=slates.device)     # only the ids matter for the bitmap
    _, count = _metrics(slates.reshape(slates.shape[0], -1), dummy, False, True)
    return int(count.item()) * 1.0 / N


def get_ILS(slates, embeds, normalize=False):
    """Intra-list similarity per slate (analysis.py:13-30); diversity = 1 - ILS."""
    assert slates.shape[1] == 5          # the reference's own restriction (analysis.py:20)
    ils, _ = _metrics(slates, embeds.weight.detach(), True, False)
    return ils
