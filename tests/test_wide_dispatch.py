"""CPU: the width rules that send a layer side to the tensor-core tile kernels (csrc/engine_wide.cuh), through
the host-only C-ABI query hgnn_lg_wide_eligible, and a numpy emulation of the kernels' 3xTF32 arithmetic
(truncation split, large terms summed in round-to-nearest fp32 outside the tensor core) against fp64: the
error model DESIGN.md section 4 states, checked without a GPU."""
import numpy as np


def _lib():
    import hgnn_b200  # noqa: F401
    from hgnn_b200 import _lib
    return _lib.lib


def test_wide_eligibility_rules():
    q = _lib().hgnn_lg_wide_eligible
    # middle layers of GNN_lg at h = 32 / 16 (K = J + 2 operators, every width 2h), forward and backward
    for h in (16, 32):
        for K in (3, 4):
            assert q(K, 2 * h, 2 * h, 2 * h, 0) == 1
            assert q(K, 2 * h, 2 * h, 2 * h, 1) == 1
        assert q(3, 2 * h, 0, 2 * h, 0) == 1 and q(3, 2 * h, 0, 2 * h, 1) == 1      # GNN_simple: no cross part
    # the script default h = 2 and h = 8 stay on their own kernels (threshold: 32 outputs)
    assert q(3, 4, 4, 4, 0) == 0 and q(3, 16, 16, 16, 0) == 0 and q(3, 16, 16, 16, 1) == 0
    # layer 0 (5 node features, 1 line-graph feature) and the readout (2 outputs) do not qualify
    assert q(3, 5, 1, 64, 0) == 0 and q(3, 1, 5, 64, 0) == 0
    assert q(3, 64, 64, 2, 0) == 0 and q(3, 64, 64, 2, 1) == 0
    # shared memory: the 640 x 128 weight block of h = 64 does not fit
    assert q(3, 128, 128, 128, 0) == 0 and q(3, 128, 128, 128, 1) == 0
    # backward: dW blocks per warp are bounded (7 operators x 64 = 28 row blocks > 4 x 4), input width must be
    # 16 / 32 / 64 / 128
    assert q(7, 64, 64, 64, 1) == 0
    assert q(3, 48, 48, 48, 0) == 1 and q(3, 48, 48, 48, 1) == 0
    # nonsense arguments
    assert q(0, 64, 64, 64, 0) == 0 and q(3, 0, 0, 64, 0) == 0 and q(3, 64, -8, 64, 1) == 0


def _trunc_tf32(x):
    """what the tensor core reads of an fp32 operand: the low 13 mantissa bits are ignored"""
    return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def _dot_3xtf32(a, b, chain=2):
    """rows of a (m, k) times b (k,), as engine_wide.cuh does it: hi = trunc(x), lo = x - hi (exact in fp32, read
    truncated); k-steps of 8; the cross terms chain in one accumulator; hi*hi of `chain` k-steps is summed alone
    and added to the running fp32 sum by a round-to-nearest add."""
    a = a.astype(np.float32)
    b = b.astype(np.float32)
    ah, bh = _trunc_tf32(a), _trunc_tf32(b)
    al, bl = (a - ah).astype(np.float32), (b - bh).astype(np.float32)
    assert np.array_equal((ah + al).astype(np.float32), a)              # the split is exact
    alt, blt = _trunc_tf32(al), _trunc_tf32(bl)
    acc = np.zeros(a.shape[0], np.float32)
    small = np.zeros(a.shape[0], np.float64)
    step = 8 * chain
    for k0 in range(0, a.shape[1], step):
        s = slice(k0, k0 + step)
        big = (ah[:, s].astype(np.float64) * bh[s].astype(np.float64)).sum(1)     # products are exact in the TC
        acc = (acc + big.astype(np.float32)).astype(np.float32)                   # FADD, round to nearest
        small += (alt[:, s].astype(np.float64) * bh[s] + ah[:, s].astype(np.float64) * blt[s]).sum(1)
    return (acc + small.astype(np.float32)).astype(np.float32)


def test_3xtf32_split_meets_the_fp32_bound():
    rng = np.random.default_rng(0)
    for k in (128, 192, 320):
        a = rng.standard_normal((4096, k)) * np.exp(rng.standard_normal((4096, k)))
        b = rng.standard_normal(k) * 0.1
        ref = a.astype(np.float32).astype(np.float64) @ b.astype(np.float32).astype(np.float64)
        got = _dot_3xtf32(a, b).astype(np.float64)
        scale = np.abs(a.astype(np.float32).astype(np.float64)) @ np.abs(b.astype(np.float32).astype(np.float64))
        err = np.abs(got - ref) / scale
        assert err.max() < 2e-6, (k, err.max())                # per-output error against sum |a||b|
        assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-5
        # no systematic bias: the signed error summed over many outputs stays far below the 1e-4 parity bound
        # (a truncating accumulator fails exactly here: DESIGN.md section 4)
        assert abs((got - ref).sum()) / np.abs(ref).sum() < 1e-5
    # one TF32 product alone would not do: 2^-11 relative per operand
    single = (_trunc_tf32(a).astype(np.float64) @ _trunc_tf32(b).astype(np.float64))
    assert (np.abs(single - ref) / scale).max() > 1e-5
