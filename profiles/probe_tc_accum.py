"""How does tcgen05.mma kind::tf32 round its fp32 accumulator?  (Decides the MMA order of csrc/mlp_tc.cu.)

Operands are pre-rounded to tf32, so every product a*b is exact in fp32 and the only error of the single-pass GEMM
(library's gemm_tn, no residual operands) is the accumulation inside the tensor core: K/8 accumulating MMAs per output.
Printed per K: mean and rms of (got - exact) / |exact| in units of 2^-24, for all-positive operands (partial sums grow
monotonically: a truncating accumulator shows a bias of about -0.5 ulp per MMA, a round-to-nearest one none) and for
signed operands, next to a sequential fp32 FMA chain (torch fp32 matmul on the CUDA cores is not used: plain python
float32 loop on a few rows).
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pivotcvae_b200 import ops  # noqa: E402

DEV = "cuda:0"


def tf32(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)


def run(K, positive, M=128, N=128):
    g = torch.Generator(device=DEV).manual_seed(K + positive)
    A = torch.randn(M, K, generator=g, device=DEV)
    B = torch.randn(N, K, generator=g, device=DEV)
    if positive:
        A, B = A.abs() + 0.5, B.abs() + 0.5
    A, B = tf32(A).contiguous(), tf32(B).contiguous()
    C = torch.empty(M, N, device=DEV)
    ops.gemm_tn(A, B, M, N, K, C=C)
    torch.cuda.synchronize()
    exact = A.double() @ B.double().t()
    den = exact.abs() if positive else exact.pow(2).mean().sqrt()     # signed: relative to the rms output (no blow-up at zeros)
    rel = ((C.double() - exact) / den).cpu().numpy() / 2.0 ** -24
    if not positive:
        rel = rel * np.sign(exact.cpu().numpy())                     # > 0: away from zero, < 0: toward zero
    # sequential fp32 chain on the first 4 rows (numpy float32 arithmetic, products exact)
    a, b = A[:4].cpu().numpy(), B.cpu().numpy()
    chain = np.zeros((4, N), np.float32)
    for k in range(K):
        chain = (chain + (a[:, k:k + 1] * b[None, :, k]).astype(np.float32)).astype(np.float32)
    e4 = exact[:4].cpu().numpy()
    rc = (chain.astype(np.float64) - e4) / (np.abs(e4) if positive else float(den)) / 2.0 ** -24
    if not positive:
        rc = rc * np.sign(e4)
    print("K=%4d (%3d MMAs) %s: tensor core mean %+8.2f rms %7.2f ulp(2^-24) | fp32 chain mean %+7.2f rms %6.2f" % (
        K, K // 8, "positive" if positive else "signed  ", rel.mean(), np.sqrt((rel ** 2).mean()), rc.mean(), np.sqrt((rc ** 2).mean())))


if __name__ == "__main__":
    ops.device_ok(0)
    for positive in (1, 0):
        for K in (8, 16, 32, 64, 128, 256, 512):
            run(K, positive)
