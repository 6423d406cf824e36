"""Oracle (test infrastructure): the counter-based negative sampler restated in numpy.

Restates mmrec_neg_sample_counter (csrc/sampler.cu): draw `a` of position `b` in step `s` is
`all_items[mix(seed, s, b, a) mod n_items]`, the first draw outside the user's training history
wins -- the rejection rule of TrainDataLoader._sample_neg_ids (utils/dataloader.py:267-275,
307-309) on a splitmix64 counter stream instead of CPython's Mersenne Twister.
"""
from __future__ import annotations

import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def mix64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def neg_sample_counter(users, all_items, n_items, hist_rowptr, hist_cols, seed, step, max_draws=64):
    users = np.asarray(users, dtype=np.int64)
    out = np.full(len(users), -1, dtype=np.int64)
    with np.errstate(over="ignore"):
        s = mix64(np.uint64(seed) ^ (np.uint64(step) * np.uint64(0xD1342543DE82EF95)))
        base = mix64(s + np.arange(len(users), dtype=np.uint64))
        todo = np.ones(len(users), dtype=bool)
        for a in range(max_draws):
            r = mix64(base + np.uint64(a) * np.uint64(0x2545F4914F6CDD1D))
            idx = ((r >> np.uint64(11)) % np.uint64(n_items)).astype(np.int64)
            cand = idx if all_items is None else np.asarray(all_items, dtype=np.int64)[idx]
            for b in np.flatnonzero(todo):
                u = users[b]
                h = hist_cols[hist_rowptr[u]: hist_rowptr[u + 1]]
                p = np.searchsorted(h, cand[b])
                if not (p < len(h) and h[p] == cand[b]):
                    out[b] = cand[b]
                    todo[b] = False
            if not todo.any():
                break
    return out
