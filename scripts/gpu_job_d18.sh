#!/bin/bash
# side partial reduce with 256-thread CTAs for the <= 128 slabs of a batch-row backward: tests + A/B on one box
O=gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "side or smore or SMORE or batch_rows or trainer or baseline" 2>&1 | tail -3 | tee $O/d18_tests.txt
MMREC_SIDE_REDUCE_WIDE=1 timeout 300 python scripts/configs_bench.py SMORE:baby 2>/dev/null | sed "s/^/WIDE=1 /" | tee $O/d18_step.txt
timeout 300 python scripts/configs_bench.py SMORE:baby SMORE:sports 2>/dev/null | sed "s/^/WIDE=0 /" | tee -a $O/d18_step.txt
MMREC_SIDE_REDUCE_WIDE=1 timeout 300 python scripts/configs_bench.py SMORE:baby 2>/dev/null | sed "s/^/WIDE=1 /" | tee -a $O/d18_step.txt
MMREC_OVERLAP=0 timeout 300 python scripts/profile_step.py SMORE 5 baby 2>/dev/null > $O/d18_step_profile_smore.txt
grep -E "^# SMORE|side_partial|side_bwd" $O/d18_step_profile_smore.txt | cut -c1-130
