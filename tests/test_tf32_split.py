"""The arithmetic of the tf32 pointwise / 7x7 kernels (csrc/lce_b200_pw.cuh), restated in numpy:
the split hi = rna_tf32(x), lo = rna_tf32(x - hi) done with two integer operations, and the
three-pass product lo*hi + hi*lo + hi*hi accumulated in fp32. No GPU needed; the GPU tests
(tests/test_gpu_builtins.py) hold the kernels to the same bound against an fp64 product."""
import numpy as np


def rna_tf32(x):
    """to_tf32() of the kernels: round to nearest, ties away from zero, to 10 explicit mantissa
    bits, on the integer image of the float."""
    b = np.asarray(x, np.float32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def test_split_is_exact_and_tf32_representable():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(200000).astype(np.float32) * np.float32(3.0),
                        (rng.standard_normal(2000) * 1e-20).astype(np.float32),
                        (rng.standard_normal(2000) * 1e20).astype(np.float32),
                        np.array([0.0, -0.0, 1.0, -1.0, 0.1, 1.0 + 2.0**-11, 1.0 + 2.0**-10], np.float32)])
    hi = rna_tf32(x)
    # hi keeps 11 significant bits and is within half a tf32 ulp (ties away) of x
    assert np.all((hi.view(np.uint32) & np.uint32(0x1FFF)) == 0)
    ulp = np.abs(np.spacing(hi.astype(np.float64).astype(np.float32))).astype(np.float64) * 2.0**13
    assert np.all(np.abs(x.astype(np.float64) - hi.astype(np.float64)) <= 0.5 * ulp + 0.0)
    # x - hi is exact in fp32 (what the kernel computes with one FADD)
    d32 = (x - hi).astype(np.float32)
    assert np.array_equal(d32.astype(np.float64), x.astype(np.float64) - hi.astype(np.float64))
    lo = rna_tf32(d32)
    assert np.all((lo.view(np.uint32) & np.uint32(0x1FFF)) == 0)
    # what the split drops: at most 2^-22 of |x| (normal range)
    normal = np.abs(x) > 1e-30
    resid = np.abs(x.astype(np.float64) - hi.astype(np.float64) - lo.astype(np.float64))
    assert np.all(resid[normal] <= np.abs(x[normal]).astype(np.float64) * 2.0**-22)


def test_three_pass_product_error_bound():
    """lo*hi + hi*lo + hi*hi with fp32 accumulation against the fp64 product: the bound the GPU
    tests use (2e-6 * sum|a||w|) holds with a wide margin, for K up to 512."""
    rng = np.random.default_rng(1)
    for K in (16, 64, 256, 512):
        A = (rng.standard_normal((64, K)) * 1.5).astype(np.float32)
        W = (rng.standard_normal((48, K)) * 0.3).astype(np.float32)
        ah, wh = rna_tf32(A), rna_tf32(W)
        al, wl = rna_tf32(A - ah), rna_tf32(W - wh)
        acc = np.zeros((64, 48), np.float32)
        for k8 in range(0, K, 8):          # K = 8 per MMA, fp32 accumulate, cross terms first
            for a, w in ((al, wh), (ah, wl)):
                acc = (acc + (a[:, k8:k8 + 8].astype(np.float64) @ w[:, k8:k8 + 8].astype(np.float64).T).astype(np.float32)).astype(np.float32)
        for k8 in range(0, K, 8):
            acc = (acc + (ah[:, k8:k8 + 8].astype(np.float64) @ wh[:, k8:k8 + 8].astype(np.float64).T).astype(np.float32)).astype(np.float32)
        ref = A.astype(np.float64) @ W.astype(np.float64).T
        mag = np.abs(A).astype(np.float64) @ np.abs(W).astype(np.float64).T
        rel = (np.abs(acc.astype(np.float64) - ref) / mag).max()
        assert rel < 2e-6, (K, rel)
        # a single pass on truncated operands (what the hardware does with raw fp32) is ~1000x worse
        tr = lambda v: (v.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
        one = (tr(A).astype(np.float64) @ tr(W).astype(np.float64).T)
        assert (np.abs(one - ref) / mag).max() > 50 * rel
