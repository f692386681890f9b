"""The FP64 GEMM building blocks (sclmd_dgemm_nt: alpha A . B^T on the DMMA pipe) against numpy: the persistent TMA / stream-K
kernel (dgemm_tma.cuh) over ragged extents, tiles shared by several CTAs, and more tiles than SMs; the cp.async kernel for
skinny products.  Both must give the same numbers to rounding (different summation orders)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [
    (128, 128, 16), (128, 128, 3000),      # one tile: every CTA holds a slice of its K range (stream-K fix-up over many CTAs)
    (1024, 600, 3000),                     # the gather product of the eigenbasis mode: 40 tiles on 148 SMs
    (1024, 3000, 600),                     # the scatter product: 192 tiles
    (1000, 301, 77), (97, 33, 5), (130, 129, 1),       # ragged M, N, K; K not a multiple of the 16-wide slab, odd leading dims
    (2500, 2500, 40),                      # 400 tiles of 3 slabs: ranges span several whole tiles
    (64, 300, 500), (3, 40, 33),           # skinny: the cp.async kernel
])
def test_dgemm_nt_vs_numpy(M, N, K):
    from sclmd_b200 import _lib
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A, B = rng.standard_normal((M, K)), rng.standard_normal((N, K))
    C = _lib.dgemm_nt(A, B, 0.75)
    want = 0.75 * (A @ B.T)
    assert np.max(np.abs(C - want)) <= 1e-13 * np.sqrt(K) * np.max(np.abs(want))
    again = _lib.dgemm_nt(A, B, 0.75)
    assert np.array_equal(C, again)        # fixed summation order: run-to-run identical
