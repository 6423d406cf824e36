#!/bin/bash
# preference module on the batch rows only (training): equivalence test, SMORE parity tests, step A/B, profile
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2b.py -x -q -m gpu -k "batch_rows" 2>&1 | tail -12 | tee $O/d14_tests_a.txt
timeout 900 python -m pytest tests -x -q -m gpu -k "smore or SMORE or trainer or baseline" 2>&1 | tail -6 | tee $O/d14_tests_b.txt
MMREC_BATCH_ROWS=0 timeout 300 python scripts/configs_bench.py SMORE:baby 2>/dev/null | sed "s/^/ROWS=0 /" | tee $O/d14_step.txt
MMREC_BATCH_ROWS=1 timeout 300 python scripts/configs_bench.py SMORE:baby SMORE:sports SMORE:clothing 2>/dev/null | sed "s/^/ROWS=1 /" | tee -a $O/d14_step.txt
MMREC_BATCH_ROWS=0 timeout 300 python scripts/configs_bench.py SMORE:sports SMORE:clothing 2>/dev/null | sed "s/^/ROWS=0 /" | tee -a $O/d14_step.txt
MMREC_OVERLAP=0 timeout 300 python scripts/profile_step.py SMORE 5 baby 2>/dev/null > $O/d14_step_profile_smore.txt
head -24 $O/d14_step_profile_smore.txt | cut -c1-130
