"""Deterministic synthetic weights and noise shared by make_golden_big.py (which feeds them to the
UNMODIFIED reference) and by the tests (which feed them to the CUDA path), so that the fixtures of the
BASELINE-sized configs only have to store seeds, checksums and the reference's outputs.

Weights come from bench.make_weights (torch CPU generators: the reference's own initialisers); the big
noise tensors (the Exp(1) draws of a sampled pivot over the whole catalog, the Bernoulli mask of
train_generative.py:39) come from numpy's PCG64 streams.  Every regenerated tensor is pinned by a checksum
stored in the fixture: a platform that regenerates different bytes fails loudly instead of comparing garbage.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def checksum(a):
    """Order-sensitive 64-bit checksum of an array's bytes (FNV-style over 8-byte words, vectorised)."""
    b = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
    pad = (-b.size) % 8
    if pad:
        b = np.concatenate([b, np.zeros(pad, np.uint8)])
    w = b.view(np.uint64)
    idx = np.arange(1, w.size + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return int(np.bitwise_xor.reduce(w * (idx * np.uint64(0x9E3779B97F4A7C15) + np.uint64(0xD2511F53))))


def weights(workload, kind):
    """(w, sd, env_sd) for a bench.py workload name; kind 'pivot' | 'list'."""
    import bench
    w = bench.WORKLOADS[workload]
    sd, env_sd = bench.make_weights(w, kind)
    return w, sd, env_sd


def weights_checksums(sd, env_sd):
    out = {}
    for k, v in sd.items():
        out["wsum/sd/" + k] = np.array(checksum(v), dtype=np.uint64)
    for k, v in env_sd.items():
        out["wsum/env/" + k] = np.array(checksum(v), dtype=np.uint64)
    return out


def race_noise(seed, B, N):
    """Exp(1) draws [B, N] fp32 for the sampled pivot (Categorical(sigmoid(.)).sample() == argmax(p / Exp(1)))."""
    return np.random.default_rng(seed).standard_exponential((B, N), dtype=np.float32)


def bernoulli_mask(seed, M, N, keep):
    """Bernoulli(keep) draws [M, N] as a bool array, generated row-block by row-block."""
    rng = np.random.default_rng(seed)
    out = np.empty((M, N), dtype=bool)
    step = max(1, (1 << 24) // N)
    for r0 in range(0, M, step):
        r1 = min(M, r0 + step)
        out[r0:r1] = rng.random((r1 - r0, N), dtype=np.float32) < np.float32(keep)
    return out


def pack_bitmask(mask):
    """bool [M, N] -> uint32 [M, ceil(N/32)], bit j%32 of word j/32 (pcv_ce_mask.bitmask layout)."""
    M, N = mask.shape
    words = (N + 31) // 32
    if words * 32 != N:
        pad = np.zeros((M, words * 32), dtype=bool)
        pad[:, :N] = mask
        mask = pad
    return np.packbits(mask, axis=1, bitorder="little").view(np.uint32).reshape(M, words)


def contexts(B, L, k):
    c = np.zeros((B, L), dtype=np.float32)
    c[:, :k] = 1
    return c
