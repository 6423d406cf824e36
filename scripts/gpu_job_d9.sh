#!/bin/bash
# per-step loss read through a pinned slot + event (truly one step behind): trajectories, then the bench line
O=gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "trainer or epoch or trajectory or train" 2>&1 | tail -4 | tee $O/d9_tests.txt
python bench.py --steps 20 --warmup 5 > $O/d9_bench_1gpu.json 2> $O/d9_bench_1gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/d9_bench_1gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','train_epoch_s','train_epoch_s_e2e','eval_users_per_s')}, d['e2e']['value'])
PY
