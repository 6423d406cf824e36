"""Oracle (test infrastructure): the in-kernel dropout mask generator restated in numpy.

Restates drop_stream / drop_mask4 of csrc/common.cuh (MmrecDropout in include/mmrec_b200.h), the
generator behind nn.Dropout of smore.py:331-333 when the product runs it inside the preference-module
kernels: element `e` (flat index into [planes, n, d]) keeps with probability 1 - floor(p * 65536) / 65536
and is scaled by 1 / (1 - p); the 16 random bits of element e are bits 16 * (e % 4) .. of
mix64(stream + (e // 4) * 0x2545F4914F6CDD1D), stream = mix64(seed ^ counter * 0xD1342543DE82EF95).
The reference draws these masks from torch's global generator; the two streams are different
generators of the same distribution (what is compared bit for bit here is the CUDA kernels against
this restatement, and the fused module against the explicit-mask module on these masks).
"""
from __future__ import annotations

import numpy as np

from .sampler import mix64


def dropout_multipliers(planes, n, d, p, seed, counter=0):
    """float32 [planes, n, d] multipliers of mmrec_dropout_mask_f32 (d % 4 == 0)."""
    assert d % 4 == 0
    total4 = planes * n * (d // 4)
    if p <= 0:
        return np.ones((planes, n, d), dtype=np.float32)
    with np.errstate(over="ignore"):
        stream = mix64(np.uint64(seed & 0xFFFFFFFFFFFFFFFF) ^ (np.uint64(int(counter)) * np.uint64(0xD1342543DE82EF95)))
        r = mix64(stream + np.arange(total4, dtype=np.uint64) * np.uint64(0x2545F4914F6CDD1D))
    thr = np.uint64(int(np.float32(p) * np.float32(65536.0)))
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    bits = np.stack([(r >> np.uint64(16 * j)) & np.uint64(0xFFFF) for j in range(4)], axis=1)   # [total4, 4]
    out = np.where(bits >= thr, scale, np.float32(0.0)).astype(np.float32)
    return out.reshape(planes, n, d)
