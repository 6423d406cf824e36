#!/bin/bash
# Final measurement job of round 2 (second session) on ONE B200: tests, bench (both arms), step profiles, the other
# configs; then -- only after the plain runs exited 0 -- the ncu launch list of the bench command and --set full of one
# launch of every hot kernel.
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/r02f_gpu_tests.txt
tail -2 $O/r02f_gpu_tests.txt
python bench.py --steps 20 --warmup 5 > $O/r02f_bench_1gpu.json 2> $O/r02f_bench_1gpu.err || exit 1
python bench.py --impl reference --steps 6 --warmup 3 > $O/r02f_bench_reference.json 2> $O/r02f_bench_reference.err
MMREC_OVERLAP=0 python scripts/profile_step.py SMORE 5 baby > $O/r02f_step_profile_smore.txt 2>/dev/null
python scripts/profile_step.py SMORE 5 baby > $O/r02f_step_profile_smore_overlap.txt 2>/dev/null
python scripts/configs_bench.py > $O/r02f_other_configs.jsonl 2>/dev/null
head -c 600 $O/r02f_bench_1gpu.json; echo; tail -2 $O/r02f_bench_1gpu.err; cat $O/r02f_other_configs.jsonl
ncu --nvtx --nvtx-include "timed_steps/" --nvtx-include "timed_eval/" --metrics gpu__time_duration.sum --clock-control none \
    --csv --log-file $O/r02f_launches.csv python bench.py --steps 2 --warmup 8 --no-sharded-blocks > $O/r02f_ncu_bench.log 2>&1
gzip -f $O/r02f_launches.csv
ncu --set full --clock-control none --import-source on \
    -k regex:"spmm_csr|gemm_tc05|side_fwd|side_bwd|side_partial|infonce_tc|score_topk_tc|mgcn_fuse|smore_combine|adam_kernel" -c 80 \
    -o /tmp/r02f_kernels python scripts/ncu_kernels.py > $O/r02f_ncu_kernels.log 2>&1
ncu -i /tmp/r02f_kernels.ncu-rep --page raw --csv > $O/r02f_kernels_raw.csv 2>/dev/null
ncu -i /tmp/r02f_kernels.ncu-rep --page details > $O/r02f_kernels_details.txt 2>/dev/null
gzip -f $O/r02f_kernels_details.txt
du -sh $O
