#!/bin/bash
# Final measurement job of round 2 (last session) on ONE B200: tests, bench (both arms), step profile, the other
# configs; then -- only after the plain runs exited 0 -- the ncu launch list of the bench command and --set full of
# the batch-row kernels added in this session.
O=gpurun_out
SECONDS=0
python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/r02h_gpu_tests.txt
tail -2 $O/r02h_gpu_tests.txt; echo "t=$SECONDS"
python bench.py --steps 20 --warmup 5 > $O/r02h_bench_1gpu.json 2> $O/r02h_bench_1gpu.err || { tail -20 $O/r02h_bench_1gpu.err; exit 1; }
echo "bench t=$SECONDS"
python bench.py --impl reference --steps 6 --warmup 3 > $O/r02h_bench_reference.json 2> $O/r02h_bench_reference.err
echo "ref t=$SECONDS"
MMREC_OVERLAP=0 python scripts/profile_step.py SMORE 5 baby > $O/r02h_step_profile_smore.txt 2>/dev/null
python scripts/profile_step.py SMORE 5 baby > $O/r02h_step_profile_smore_overlap.txt 2>/dev/null
python scripts/profile_step.py MGCN 5 sports > $O/r02h_step_profile_mgcn.txt 2>/dev/null
python scripts/configs_bench.py > $O/r02h_other_configs.jsonl 2>/dev/null
echo "profiles t=$SECONDS"
head -c 700 $O/r02h_bench_1gpu.json; echo; tail -2 $O/r02h_bench_1gpu.err; cat $O/r02h_other_configs.jsonl
python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -3 > $O/r02h_smoke.txt; tail -1 $O/r02h_smoke.txt
ncu --nvtx --nvtx-include "timed_steps/" --nvtx-include "timed_eval/" --metrics gpu__time_duration.sum --clock-control none \
    --csv --log-file $O/r02h_launches.csv python bench.py --steps 2 --warmup 8 --no-sharded-blocks > $O/r02h_ncu_bench.log 2>&1
gzip -f $O/r02h_launches.csv
echo "launch list t=$SECONDS"
ncu --set full --clock-control none --import-source on \
    -k regex:"batch_rows|batch_views|side_fwd_tc|side_bwd|side_partial" -c 24 \
    -o /tmp/r02h_kernels python scripts/ncu_batch_rows.py > $O/r02h_ncu_kernels.log 2>&1
ncu -i /tmp/r02h_kernels.ncu-rep --page raw --csv > $O/r02h_kernels_raw.csv 2>/dev/null
echo "ncu t=$SECONDS"
du -sh $O
