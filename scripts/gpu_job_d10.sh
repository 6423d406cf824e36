#!/bin/bash
# final code on a 2-GPU box: smoke(), the NCCL tests, the bench line under torchrun (replicas + sharded config 5 / 4 blocks)
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee $O/d10_smoke.txt
timeout 600 python -m pytest tests -x -q -m gpu -k "nccl or sharded or two_gpu or multi" 2>&1 | tail -4 | tee $O/d10_nccl_tests.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02f_bench_2gpu.json 2> $O/r02f_bench_2gpu.err
tail -3 $O/r02f_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02f_bench_2gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','clocks')}, d['e2e']['value'])
for k in ('sharded_config5','sharded_config4'):
    print(k, json.dumps(d['config'].get(k))[:700])
PY
