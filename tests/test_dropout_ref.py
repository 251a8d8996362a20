"""The numpy restatement of the dropout mask generator (oracle/dropout_ref.py): self-checks that run without a GPU.
The GPU kernels are held to it bit for bit in tests/test_gpu_ops.py::test_dropout_mask_matches_numpy_oracle."""
import hashlib

import numpy as np

from oracle import dropout_ref as R


def _digest(mask):
    return hashlib.sha256(np.packbits(mask).tobytes()).hexdigest()[:16]


def test_known_answers_pin_the_generator():
    # digests recorded when the CUDA kernels were verified bit-exact against this file on a B200
    assert _digest(R.keep_mask(20240607, 3, 0.123, 0, 1 << 20)) == "3ffdc0468ea4e38b"
    assert _digest(R.keep_mask((7 << 40) + 12345, 17, 0.1, 5, 4000)) == "64d906105a0b9ed7"
    w = R.philox4x32_7(1234567, np.array([0, 1, 2 ** 33 + 5], dtype=np.uint64), 16)
    assert w.tolist() == [[3481946403, 4133682139, 546616713, 952978656],
                          [1625711112, 2211441789, 2048713179, 871458901],
                          [357070081, 2183093374, 777356037, 2774738466]]


def test_keep_rate_is_exact_to_2_pow_minus_16():
    n = 1 << 22
    for p in (0.1, 0.123, 0.5, 1 / 256, 0.9):
        rate = R.keep_mask(99, 4, p, 0, n).mean()
        assert abs(rate - (1 - R.thr16(p) / 65536)) < 5 * np.sqrt(p * (1 - p) / n) + 2e-5, (p, rate)
    assert R.keep_mask(1, 2, 0.0, 0, 100).all()


def test_mask_is_a_pure_function_of_the_element_index():
    a = R.keep_mask(5, 6, 0.3, 0, 1000)
    b = R.keep_mask(5, 6, 0.3, 137, 500)
    assert np.array_equal(a[137:637], b)
    assert not np.array_equal(a, R.keep_mask(5, 7, 0.3, 0, 1000))   # another site
    assert not np.array_equal(a, R.keep_mask(6, 6, 0.3, 0, 1000))   # another seed
