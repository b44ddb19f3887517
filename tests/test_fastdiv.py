"""The launch-constant division used by the gather / epilogue index math
(compute_engine_b200/csrc/lce_b200_kernels.cuh: make_fastdiv / fdiv), restated in numpy and
checked against integer division: mul = floor(2^(31+L)/d) + 1, L = ceil(log2 d),
q = (n * mul) >> (31 + L), exact for 0 <= n < 2^31."""
import numpy as np


def make_fastdiv(d):
    L = 0
    while (1 << L) < d:
        L += 1
    mul = ((1 << (31 + L)) // d) + 1
    assert mul < (1 << 32), (d, mul)          # must fit the kernel's uint32 field
    return mul, 31 + L


def test_fastdiv_matches_integer_division():
    rng = np.random.default_rng(0)
    divisors = list(range(1, 300)) + [3136, 12544, 50176, 802816, 2**16, 2**16 + 1, 2**20 - 1,
                                      2**30, 2**31 - 1] + [int(x) for x in
                                                           rng.integers(1, 2**31 - 1, 200)]
    edge = np.array([0, 1, 2, 2**31 - 1, 2**31 - 2, 2**30, 2**30 - 1], dtype=object)
    for d in divisors:
        mul, shift = make_fastdiv(d)
        n = np.concatenate([edge, rng.integers(0, 2**31, 2000).astype(object),
                            np.array([k * d + e for k in (0, 1, 2, 7, (2**31 - 1) // d)
                                      for e in (-1, 0, 1) if 0 <= k * d + e < 2**31],
                                     dtype=object)])
        q = (n * mul) >> shift
        assert np.array_equal(q, n // d), d
