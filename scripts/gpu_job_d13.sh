#!/bin/bash
# final code: whole GPU suite, smoke, bench line, step profile, other configs
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/r02g_gpu_tests.txt; tail -1 $O/r02g_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > $O/r02g_bench_1gpu.json 2> $O/r02g_bench_1gpu.err || exit 1
MMREC_OVERLAP=0 python scripts/profile_step.py SMORE 5 baby > $O/r02g_step_profile_smore.txt 2>/dev/null
python scripts/profile_step.py SMORE 5 baby > $O/r02g_step_profile_smore_overlap.txt 2>/dev/null
python scripts/configs_bench.py > $O/r02g_other_configs.jsonl 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02g_bench_1gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','train_epoch_s','train_epoch_s_e2e','eval_users_per_s','gpu_launches')}, d['e2e']['value'], d['clocks'])
PY
cat $O/r02g_other_configs.jsonl
