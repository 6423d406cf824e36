"""CPU oracle: a restatement of the reference's hot-path arithmetic. TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this package, and only as the checker (or as the timed CPU baseline) -- never as
part of the product path. The product (`recommendar-systems_b200/`) must not import it and
fails loudly when its CUDA library is missing.

The arithmetic of the reference lives in third-party libraries that are not under /root/reference:
torch (requirements.txt pins 1.11.0; installed 2.11.0), scipy (pinned 1.7.3; installed 1.18.1) and
numpy. The oracle restates each reference function on those same libraries' CPU kernels (plain
numpy for integer/byte work, torch CPU float32/float64 for the floating-point chains, autograd
for gradients) and cites the reference file:line it follows.

Parity is PINNED: `tests/test_oracle_golden.py` checks every function here against golden vectors
produced by executing the unmodified reference in-process (`tests/golden/make_golden.py`, which
imports /root/reference/src; the committed `tests/golden/*.npz` travel to the GPU box).
"""
