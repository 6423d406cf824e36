#!/bin/bash
# batch views: user rows R x' of the modality views for the batch users only
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2b.py -x -q -m gpu -k "batch_rows" 2>&1 | tail -12 | tee $O/d15_tests_a.txt
timeout 900 python -m pytest tests -x -q -m gpu -k "smore or SMORE or trainer or baseline" 2>&1 | tail -6 | tee $O/d15_tests_b.txt
MMREC_BATCH_VIEWS=0 timeout 300 python scripts/configs_bench.py SMORE:baby 2>/dev/null | sed "s/^/VIEWS=0 /" | tee $O/d15_step.txt
MMREC_BATCH_VIEWS=1 timeout 300 python scripts/configs_bench.py SMORE:baby SMORE:sports SMORE:clothing 2>/dev/null | sed "s/^/VIEWS=1 /" | tee -a $O/d15_step.txt
MMREC_BATCH_VIEWS=0 timeout 300 python scripts/configs_bench.py SMORE:baby 2>/dev/null | sed "s/^/VIEWS=0 /" | tee -a $O/d15_step.txt
MMREC_OVERLAP=0 timeout 300 python scripts/profile_step.py SMORE 5 baby 2>/dev/null > $O/d15_step_profile_smore.txt
head -30 $O/d15_step_profile_smore.txt | cut -c1-130
