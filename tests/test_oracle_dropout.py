"""CPU checks of oracle/dropout.py (the numpy restatement of the in-kernel dropout generator of
csrc/common.cuh): known answers against pure-Python integers, keep rate, scale, stream separation."""
import numpy as np

from oracle import dropout as odrop


def _mix(z):
    M = (1 << 64) - 1
    z = (z + 0x9E3779B97F4A7C15) & M
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
    return z ^ (z >> 31)


def _python_mask(planes, n, d, p, seed, counter):
    M = (1 << 64) - 1
    stream = _mix((seed ^ (counter * 0xD1342543DE82EF95)) & M)
    thr = int(np.float32(p) * np.float32(65536.0))
    scale = float(np.float32(1.0) / (np.float32(1.0) - np.float32(p)))
    out = np.empty(planes * n * d, dtype=np.float32)
    for e4 in range(planes * n * d // 4):
        r = _mix((stream + e4 * 0x2545F4914F6CDD1D) & M)
        for j in range(4):
            out[4 * e4 + j] = scale if ((r >> (16 * j)) & 0xFFFF) >= thr else 0.0
    return out.reshape(planes, n, d)


def test_matches_python_integers():
    for (planes, n, d, p, seed, counter) in [(3, 5, 8, 0.5, 999, 0), (1, 3, 4, 0.1, (999 << 32) ^ 0x9E3779B97F4A7C15, 7),
                                             (3, 2, 64, 0.25, 2 ** 64 - 1, 123456)]:
        got = odrop.dropout_multipliers(planes, n, d, p, seed, counter)
        assert np.array_equal(got, _python_mask(planes, n, d, p, seed, counter))


def test_distribution_and_streams():
    p = 0.3
    a = odrop.dropout_multipliers(3, 4000, 64, p, seed=5, counter=10)
    assert set(np.unique(a)) == {np.float32(0.0), np.float32(1.0) / (np.float32(1.0) - np.float32(p))}
    keep = (a > 0).mean()
    assert abs(keep - (1 - p)) < 3e-3                       # 768k draws: sigma = 5e-4
    assert abs(float(a.mean()) - 1.0) < 5e-3                # E[multiplier] = 1
    # per plane and per column group the rate holds too (no lane / bit-field bias)
    assert np.abs((a > 0).mean(axis=(1, 2)) - (1 - p)).max() < 5e-3
    assert np.abs((a > 0).reshape(-1, 4).mean(axis=0) - (1 - p)).max() < 5e-3
    # a different counter or seed is a different stream; the same pair repeats exactly
    b = odrop.dropout_multipliers(3, 4000, 64, p, seed=5, counter=11)
    c = odrop.dropout_multipliers(3, 4000, 64, p, seed=6, counter=10)
    assert abs(((a > 0) == (b > 0)).mean() - ((1 - p) ** 2 + p ** 2)) < 5e-3
    assert abs(((a > 0) == (c > 0)).mean() - ((1 - p) ** 2 + p ** 2)) < 5e-3
    assert np.array_equal(a, odrop.dropout_multipliers(3, 4000, 64, p, seed=5, counter=10))
    assert np.array_equal(odrop.dropout_multipliers(2, 3, 8, 0.0, 1, 1), np.ones((2, 3, 8), np.float32))
