"""CPU checks of the counter-based sampler's oracle restatement (oracle/sampler.py): the numpy
splitmix64 mixer against pure-Python integers, and the sampling rule of dataloader.py:267-275."""
import numpy as np

from oracle import sampler as osampler


def _mix(z):
    M = (1 << 64) - 1
    z = (z + 0x9E3779B97F4A7C15) & M
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
    return z ^ (z >> 31)


def test_mixer_matches_python_integers():
    xs = [0, 1, 2, 0x9E3779B97F4A7C15, 12345678901234567890, 2 ** 64 - 1]
    got = osampler.mix64(np.array(xs, dtype=np.uint64))
    assert [int(g) for g in got] == [_mix(x) for x in xs]
    assert int(osampler.mix64(np.uint64(0))) == 0xE220A8397B1DCDAF          # splitmix64's first output for seed 0


def test_rule_and_known_answer():
    rng = np.random.default_rng(3)
    U, I = 40, 25
    hist = [np.sort(rng.choice(I, size=rng.integers(1, 20), replace=False)) for _ in range(U)]
    rowptr = np.concatenate(([0], np.cumsum([len(h) for h in hist]))).astype(np.int64)
    cols = np.concatenate(hist).astype(np.int32)
    users = rng.integers(0, U, size=500)
    a = osampler.neg_sample_counter(users, None, I, rowptr, cols, seed=11, step=5)
    b = osampler.neg_sample_counter(users, None, I, rowptr, cols, seed=11, step=5)
    c = osampler.neg_sample_counter(users, None, I, rowptr, cols, seed=11, step=6)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert all(n not in set(hist[u].tolist()) for u, n in zip(users, a)) and (a >= 0).all()
    # first draw of position 0, written out with Python integers
    s = _mix(11 ^ ((5 * 0xD1342543DE82EF95) & ((1 << 64) - 1)))
    base = _mix((s + 0) & ((1 << 64) - 1))
    draws = [(_mix((base + k * 0x2545F4914F6CDD1D) & ((1 << 64) - 1)) >> 11) % I for k in range(64)]
    first_ok = next(d for d in draws if d not in set(hist[users[0]].tolist()))
    assert a[0] == first_ok
    # all_items remaps the draw
    perm = rng.permutation(I).astype(np.int64)
    d = osampler.neg_sample_counter(users[:50], perm, I, rowptr, cols, seed=11, step=5)
    assert all(n not in set(hist[u].tolist()) for u, n in zip(users[:50], d))
