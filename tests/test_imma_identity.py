"""The arithmetic identity behind the int8 tensor-pipe kernel
(compute_engine_b200/csrc/lce_b200_imma.cuh), checked in numpy:

    popc(a ^ w) = popc(w) + sum_k a_k * w'_k,     a_k in {0,1},  w'_k = +1 if w_k == 0 else -1

and the byte encoding the kernel feeds to mma.sync (bit i of a nibble -> u8 value 2^i on the
activation side, s8 value +-(8 >> i) on the weight side): every product is +-8, so the accumulators
hold exactly 8 * popc(a ^ w) when they start at 8 * popc(w)."""
import numpy as np


def popc(x):
    return np.array([bin(int(v) & 0xFFFFFFFF).count("1") for v in np.ravel(x)]).reshape(np.shape(x))


def act_bytes(word, tig):
    """expand01: byte `tig` of the word -> (lo, hi) registers of four u8, bit i -> value 2^i."""
    byte = (int(word) >> (8 * tig)) & 0xFF
    lo = [(byte & 0xF) & (1 << i) for i in range(4)]
    hi = [(byte >> 4) & (1 << i) for i in range(4)]
    return lo, hi


def weight_bytes(nibble):
    """imma_weight_bytes: byte i = +(8 >> i) for bit 0, -(8 >> i) for bit 1."""
    return [-(8 >> i) if (nibble >> i) & 1 else (8 >> i) for i in range(4)]


def test_xor_popcount_equals_popw_plus_signed_dot():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 2**32, 500, dtype=np.uint64)
    w = rng.integers(0, 2**32, 500, dtype=np.uint64)
    a[:4] = [0, 0xFFFFFFFF, 0, 0xFFFFFFFF]          # padding word, all -1, ...
    w[:4] = [0, 0, 0xFFFFFFFF, 0xFFFFFFFF]
    for x, y in zip(a, w):
        bits_a = [(int(x) >> k) & 1 for k in range(32)]
        signed_w = [-1 if (int(y) >> k) & 1 else 1 for k in range(32)]
        assert popc(int(x) ^ int(y)) == popc(int(y)) + sum(p * q for p, q in zip(bits_a, signed_w))


def test_scaled_byte_encoding_gives_eight_times_the_popcount():
    rng = np.random.default_rng(1)
    for _ in range(300):
        x, y = int(rng.integers(0, 2**32)), int(rng.integers(0, 2**32))
        dot = 0
        for tig in range(4):                         # the four lanes of a quad cover the word
            alo, ahi = act_bytes(x, tig)
            wbyte = (y >> (8 * tig)) & 0xFF
            wlo, whi = weight_bytes(wbyte & 0xF), weight_bytes(wbyte >> 4)
            for i in range(4):
                assert abs(alo[i] * wlo[i]) in (0, 8) and abs(ahi[i] * whi[i]) in (0, 8)
                assert -128 <= wlo[i] <= 127 and 0 <= alo[i] <= 255
            dot += sum(p * q for p, q in zip(alo, wlo)) + sum(p * q for p, q in zip(ahi, whi))
        acc8 = 8 * popc(y) + dot
        assert acc8 == 8 * popc(x ^ y)
        assert (acc8 >> 2) == (popc(x ^ y) << 1)      # OutputTransform's acc << 1
