#!/bin/bash
# batch views with the CTA-level chunk task list; MGCN on the batch rows
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2b.py -x -q -m gpu -k "batch_rows" 2>&1 | tail -6 | tee $O/d16_tests_a.txt
timeout 900 python -m pytest tests -x -q -m gpu -k "smore or SMORE or mgcn or MGCN or trainer or baseline" 2>&1 | tail -4 | tee $O/d16_tests_b.txt
MMREC_BATCH_VIEWS=0 timeout 300 python scripts/configs_bench.py SMORE:baby 2>/dev/null | sed "s/^/VIEWS=0 /" | tee $O/d16_step.txt
MMREC_BATCH_VIEWS=1 timeout 300 python scripts/configs_bench.py SMORE:baby SMORE:sports SMORE:clothing MGCN:sports 2>/dev/null | sed "s/^/VIEWS=1 /" | tee -a $O/d16_step.txt
MMREC_BATCH_ROWS=0 timeout 300 python scripts/configs_bench.py MGCN:sports 2>/dev/null | sed "s/^/ROWS=0 /" | tee -a $O/d16_step.txt
MMREC_OVERLAP=0 timeout 300 python scripts/profile_step.py SMORE 5 baby 2>/dev/null > $O/d16_step_profile_smore.txt
head -12 $O/d16_step_profile_smore.txt | cut -c1-130; grep -E "batch_views|gather_batch|scatter_batch|spmm_csr_multi" $O/d16_step_profile_smore.txt | cut -c1-130
