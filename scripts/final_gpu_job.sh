#!/bin/bash
# The round's measurement job on ONE B200 (run through gpurun from the repo root); outputs under gpurun_out/.
set -x
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/r02_gpu_tests.txt
python bench.py --steps 20 --warmup 5 > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu.err
python bench.py --impl reference --steps 6 --warmup 3 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err
MMREC_OVERLAP=0 python scripts/profile_step.py SMORE 5 baby > $O/r02_step_profile_smore.txt 2>/dev/null
python scripts/profile_step.py SMORE 5 baby > $O/r02_step_profile_smore_overlap.txt 2>/dev/null
python scripts/configs_bench.py > $O/r02_other_configs.jsonl 2>/dev/null
# ncu AFTER the plain runs exited 0: launch list of the bench command (timed steps + timed eval), then --set full of the hot kernels
ncu --nvtx --nvtx-include "timed_steps/" --nvtx-include "timed_eval/" --metrics gpu__time_duration.sum --clock-control none \
    --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 8 --no-sharded-blocks > $O/r02_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"spmm_csr|gemm_tc05|side_fwd|side_bwd|infonce_tc|score_topk_tc|mgcn_fuse|smore_combine|adam_kernel" -c 80 \
    -o $O/r02_kernels python scripts/ncu_kernels.py > $O/r02_ncu_kernels.log 2>&1
ncu -i $O/r02_kernels.ncu-rep --page raw --csv > $O/r02_kernels_raw.csv 2>/dev/null
ls -la $O | tail -20
