"""
CPU check of the arithmetic claim behind the tf32 input layer of the tensor-core path (csrc/naf_policy_tc.cu,
csrc/naf_trunk_tc.cu MODE 2): an observation x is fed to tcgen05.mma.kind::tf32 as hi = tf32(x) and lo = tf32(x - hi),
so the product sees x to 2^-22 relative, and only the weights are rounded to tf32's 10-bit mantissa.  tf32 rounding
(cvt.rna.tf32.f32: nearest, ties away from zero) is emulated on the fp32 bit patterns.
"""
import numpy as np


def tf32_rna(x):
    b = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    b = (b + 0x1000) & 0xFFFFE000                      # magnitude bits: round half away from zero, keep 10 mantissa bits
    return b.astype(np.uint32).view(np.float32)


def test_hi_lo_split_keeps_22_bits():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-3.2, 3.2, 200000), rng.normal(0, 1e-3, 50000), rng.uniform(-100, 100, 50000)]).astype(np.float32)
    hi = tf32_rna(x)
    lo = tf32_rna(x - hi)
    assert np.all(np.abs(hi.astype(np.float64) - x) <= np.abs(x) * 2.0 ** -11 + 1e-45)
    err = np.abs(hi.astype(np.float64) + lo.astype(np.float64) - x.astype(np.float64))
    assert np.all(err <= np.abs(x.astype(np.float64)) * 2.0 ** -22)
    assert np.all((hi.view(np.uint32) & 0x1FFF) == 0) and np.all((lo.view(np.uint32) & 0x1FFF) == 0)


def test_layer1_error_is_the_weight_rounding_only():
    """z1 = s W1^T on KUKA-like observations: split inputs + tf32 weights against fp64.  The error is the weights'
    5e-4 relative rounding (random sign, averaged over K = 21), far below the bf16 rounding (4e-3) the next layer applies."""
    rng = np.random.default_rng(1)
    B, S, H = 2048, 21, 256
    s = np.concatenate([rng.uniform(-2.9, 2.9, (B, 6)), rng.uniform(-1, 1, (B, 6)), rng.uniform(-1, 1.3, (B, 3)),
                        np.tile([0.4, 0.85, 0.71], (B, 1)), np.tile([0.45, 0.55, 0.55], (B, 1))], axis=1).astype(np.float32)
    W = (rng.uniform(-1, 1, (H, S)) / np.sqrt(S)).astype(np.float32)            # nn.Linear default init range
    exact = s.astype(np.float64) @ W.astype(np.float64).T
    hi = tf32_rna(s)
    lo = tf32_rna(s - hi)
    Wt = tf32_rna(W).astype(np.float64)
    split = hi.astype(np.float64) @ Wt.T + lo.astype(np.float64) @ Wt.T
    plain = tf32_rna(s).astype(np.float64) @ Wt.T                               # what a single tf32 MMA would give
    scale = np.abs(exact).max()
    e_split, e_plain = np.abs(split - exact).max() / scale, np.abs(plain - exact).max() / scale
    assert e_split <= 3e-4 and e_split < e_plain
    only_w = np.abs(s.astype(np.float64) @ Wt.T - exact).max() / scale          # weight rounding alone
    assert abs(e_split - only_w) <= 1e-6
