"""GPU tests of the tcgen05 tensor-core path (bf16 operands, fp32 accumulate)."""
import ctypes

import numpy as np
import pytest
import torch

from flowchain_iccv2023_b200 import _lib

pytestmark = pytest.mark.gpu


def bf16_round(a):
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return t.to(torch.bfloat16).to(torch.float32).numpy()


@pytest.mark.parametrize("N,K", [(16, 16), (64, 64), (128, 34), (160, 40), (80, 80), (48, 80), (256, 256), (32, 1)])
def test_tcgen05_gemm_selftest(N, K):
    """D[128 x N] = A[128 x K] B[N x K]^T through the UMMA descriptors / TMEM / mbarrier helpers of the fused
    kernel; exact vs a float64 product of the bf16-rounded operands up to fp32 accumulation order."""
    lib = _lib.load()
    rng = np.random.default_rng(N * 1000 + K)
    A = rng.standard_normal((128, K)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    D = np.zeros((128, N), dtype=np.float32)
    rc = lib.fc_tc_gemm_selftest(A.ctypes.data, B.ctypes.data, D.ctypes.data, N, K, 0)
    assert rc == 0, _lib.last_error(None)
    ref = bf16_round(A).astype(np.float64) @ bf16_round(B).astype(np.float64).T
    np.testing.assert_allclose(D, ref, atol=1e-4 * np.sqrt(K), rtol=1e-5)
