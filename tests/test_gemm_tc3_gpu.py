"""GPU: the tcgen05 3 x TF32 GEMM tiles (csrc/gemm_tc3.cuh) through the C ABI (srlx_sgemm_tc3) against a float64 product: every operand
layout the networks use (row- / column-major A and B), ragged sizes around the 128 x {32, 64, 128} tiles and the 32-column k slice,
split-K, ReLU and accumulate epilogues.  Bar: fp32 accuracy -- |C - C64| <= 2e-6 * (|A| . |B|) elementwise (an fp32 FMA chain of length
K is allowed about K * 6e-8 of that bound; 3 x TF32 drops 2^-22 per product)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(M, N, K, a_t, b_t, relu=0, acc=0, ws=True, seed=0):
    from simple_distributed_rl_b200 import _lib

    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn((K, M) if a_t else (M, K), device="cuda", generator=g)
    B = torch.randn((N, K) if b_t else (K, N), device="cuda", generator=g)
    ldc = N + 3
    C0 = torch.randn((M, ldc), device="cuda", generator=g)
    C = C0.clone()
    wsb = torch.empty(1 << 22, device="cuda") if ws else None
    sa = (1, M) if a_t else (K, 1)
    sb = (1, K) if b_t else (N, 1)
    _lib.check(lib.srlx_sgemm_tc3(A.data_ptr(), sa[0], sa[1], B.data_ptr(), sb[0], sb[1], C.data_ptr(), ldc, M, N, K, relu, acc,
                                  wsb.data_ptr() if ws else None, wsb.numel() if ws else 0, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    A64 = (A.t() if a_t else A).double()
    B64 = (B.t() if b_t else B).double()
    want = A64 @ B64
    bound = A64.abs() @ B64.abs()
    if acc:
        want = want + C0[:, :N].double()
    if relu:
        want = want.clamp_min(0)
    err = (C[:, :N].double() - want).abs()
    assert torch.all(err <= 2e-6 * bound + 1e-30), (M, N, K, a_t, b_t, float((err / bound).max()))
    assert torch.equal(C[:, N:], C0[:, N:])  # nothing written beyond N


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 64, 64), (256, 32, 96), (1, 1, 1), (130, 70, 33), (14112, 32, 257), (3872, 64, 513),
                                   (32, 512, 7744), (512, 7745, 32), (64, 577, 3872), (32, 6, 513), (6, 513, 32), (32, 7744, 512), (300, 200, 1000)])
def test_sgemm_tc3_shapes(M, N, K):
    for a_t in (0, 1):
        for b_t in (0, 1):
            _run(M, N, K, a_t, b_t, seed=M + N + K)


def test_sgemm_tc3_epilogues_and_no_workspace():
    _run(200, 90, 700, 0, 1, relu=1)
    _run(200, 90, 700, 0, 1, acc=1)
    _run(200, 90, 700, 1, 0, relu=1, acc=1)
    _run(64, 48, 4096, 0, 0, ws=False)  # few tiles, long k, no split allowed
    _run(64, 48, 4096, 0, 0, relu=1, acc=1)  # the same through split-K: the epilogue runs in the reduction
