"""Bit-exact numpy restatement of the dropout mask generator of libtvit_b200 -- TEST INFRASTRUCTURE ONLY.

The reference uses ``nn.Dropout`` (model.py:102,104,138,140,224,250), i.e. ATen's Philox stream, which cannot be
reproduced bit for bit outside ATen.  The CUDA path therefore defines its own counter-based
generator (include/tvit.h, ``tvit_dropout``; csrc/common.cuh) and this file is its oracle: integer arithmetic,
so the parity bar is bit-exactness (tests/test_gpu_ops.py compares the kernels' masks with ``keep_mask``).

    keep(e) <=> byte(e) >= T_g,  g = e >> 4
    byte(e)  = byte (e & 3) of word ((e >> 2) & 3) of Philox4x32-7(key = seed, counter = (g_lo, g_hi, site, 0x5eed5eed))
    T_g      = (thr16 >> 8) + [((uint32(g) * 0x9E3779B1 + uint32(seed)) >> 24) < (thr16 & 255)]
    thr16    = min(round(p * 65536), 65535);  kept elements are scaled by 1 / (1 - thr16 / 65536)
"""
from __future__ import annotations

import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_7(seed: int, group: np.ndarray, site: int) -> np.ndarray:
    """Philox4x32 with 7 rounds; ``group`` is a uint64 array of counters.  Returns uint32 [len(group), 4]."""
    g = np.asarray(group, dtype=np.uint64)
    c0, c1 = g & _MASK32, g >> np.uint64(32)
    c2 = np.full_like(c0, site & 0xFFFFFFFF)
    c3 = np.full_like(c0, 0x5EED5EED)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for _ in range(7):
        p0, p1 = _M0 * c0, _M1 * c2          # 32 x 32 -> 64 bit products (operands < 2^32, no overflow)
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK32
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=1).astype(np.uint32)


def thr16(p: float) -> int:
    return min(int(np.float32(p) * np.float32(65536.0) + np.float32(0.5)), 65535) if p > 0 else 0


def inv_keep(p: float) -> float:
    t = thr16(p)
    return float(np.float32(1.0) / (np.float32(1.0) - np.float32(t) * np.float32(1.0 / 65536.0))) if t else 1.0


def keep_mask(seed: int, site: int, p: float, first: int, count: int) -> np.ndarray:
    """Boolean keep flags of the elements [first, first + count)."""
    t = thr16(p)
    if t == 0:
        return np.ones(count, dtype=bool)
    e = np.arange(first, first + count, dtype=np.uint64)
    g = e >> np.uint64(4)
    ug, inv = np.unique(g, return_inverse=True)
    words = philox4x32_7(seed, ug, site)[inv]                       # [count, 4]
    j = (e & np.uint64(15)).astype(np.int64)
    w = words[np.arange(count), j >> 2]
    byte = (w >> ((j & 3) * 8).astype(np.uint32)) & np.uint32(0xFF)
    weyl = ((g & _MASK32) * np.uint64(0x9E3779B1) + np.uint64(seed & 0xFFFFFFFF)) & _MASK32
    tg = np.uint32(t >> 8) + ((weyl >> np.uint64(24)) < np.uint64(t & 255)).astype(np.uint32)
    return byte >= tg
