"""CPU: the shared-memory plane layouts of the experimental tcgen05 kernels (csrc/engine_tc5.cuh: t5_a_off,
t5_b_off) against the canonical UMMA SWIZZLE_NONE layouts in 16-byte units (CUTLASS cute/atom/mma_traits_sm100.hpp:
K-major ((8,m),(T,2)):((1T,SBO),(1,LBO)), MN-major ((T,1,m),(8,k)):((1,T,SBO),(1T,LBO)), T = 4 tf32 per 16 bytes):
one plane written as (row, feature) is the K-major A operand of gX with (LBO, SBO) = (1024, 128) AND the MN-major
operand of dW with (128, 1024); the weight chunks are K-major B operands with LBO = (N / 8) * 128.
The index arithmetic is restated here; the kernels themselves have not run on a GPU yet."""
import numpy as np

T = 4


def a_off(r, k):
    return (k >> 2) * 256 + (r >> 3) * 32 + (r & 7) * 4 + (k & 3)


def b_off(n, k, N):
    return (k >> 2) * (N >> 3) * 32 + (n >> 3) * 32 + (n & 7) * 4 + (k & 3)


def k_major(plane, m, k, lbo, sbo):
    return plane[(m % 8) * T + (m // 8) * (sbo // 4) + (k % T) + (k // T) * (lbo // 4)]


def mn_major(plane, m, k, lbo, sbo):
    return plane[(m % T) + (m // T) * (sbo // 4) + (k % 8) * T + (k // 8) * (lbo // 4)]


def test_engine_tc5_header_uses_these_formulas():
    import os
    from conftest import ROOT
    src = open(os.path.join(ROOT, "hgnn-2_b200", "csrc", "engine_tc5.cuh")).read()
    assert "return (k >> 2) * 256 + (r >> 3) * 32 + (r & 7) * 4 + (k & 3);" in src
    assert "return (k >> 2) * (Fout >> 3) * 32 + (n >> 3) * 32 + (n & 7) * 4 + (k & 3);" in src
    assert "t5_desc(t5_smem(Ahi) + ks * 2048, 1024, 128)" in src       # gX: K-major, two core matrices per k-step
    assert "t5_desc(t5_smem(Ahi) + ks * 128, 128, 1024)" in src        # dW: MN-major, one 8-row group per k-step


def test_one_plane_is_k_major_for_gx_and_mn_major_for_dw():
    rng = np.random.default_rng(0)
    A = rng.standard_normal((64, 72)).astype(np.float32)               # 72: input rows + ones column of the dW operand
    plane = np.zeros(64 * 72, np.float32)
    for r in range(64):
        for f in range(72):
            plane[a_off(r, f)] = A[r, f]
    assert max(a_off(r, f) for r in range(64) for f in range(72)) < 64 * 72
    for m in range(64):
        for k in range(64):
            assert k_major(plane, m, k, 1024, 128) == A[m, k]
    for f in range(72):
        for r in range(64):
            assert mn_major(plane, f, r, 128, 1024) == A[r, f]
    # descriptor advance per instruction: K = 8 tf32 = two core matrices (K-major) / one group of 8 rows (MN-major)
    for ks in range(8):
        base_k, base_mn = ks * 2 * 1024 // 4, ks * 128 // 4
        for m in range(0, 64, 7):
            for kk in range(8):
                assert plane[base_k + (m % 8) * T + (m // 8) * 32 + (kk % T) + (kk // T) * 256] == A[m, 8 * ks + kk]
                assert plane[base_mn + (m % T) + (m // T) * 256 + kk * T] == A[8 * ks + kk, m]


def test_weight_chunks_are_k_major_b_operands():
    rng = np.random.default_rng(1)
    for N in (32, 64):
        W = rng.standard_normal((N, 64)).astype(np.float32)            # W[n][k]: output n, chunk column k
        plane = np.zeros(N * 64, np.float32)
        for n in range(N):
            for k in range(64):
                plane[b_off(n, k, N)] = W[n, k]
        for n in range(N):
            for k in range(64):
                assert k_major(plane, n, k, (N // 8) * 128, 128) == W[n, k]
