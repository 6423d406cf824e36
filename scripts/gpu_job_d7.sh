#!/bin/bash
# flat (row tile, column tile) partition of the low-rank table Adam / norm kernels
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "lowrank or table" 2>&1 | tail -4 | tee $O/d7_tests.txt
MMREC_TA_FLAT=0 timeout 200 python scripts/table_adam_micro.py 2>&1 | tail -8 | tee $O/d7_micro_rect.txt
MMREC_TA_FLAT=1 timeout 200 python scripts/table_adam_micro.py 2>&1 | tail -8 | tee $O/d7_micro_flat.txt
MMREC_TA_FLAT=0 timeout 300 python scripts/configs_bench.py SMORE:baby 2>/dev/null | tee $O/d7_step.txt
MMREC_TA_FLAT=1 timeout 300 python scripts/configs_bench.py SMORE:baby SMORE:clothing 2>/dev/null | tee -a $O/d7_step.txt
MMREC_TA_FLAT=0 timeout 300 python scripts/configs_bench.py SMORE:baby SMORE:clothing 2>/dev/null | tee -a $O/d7_step.txt
MMREC_TA_FLAT=1 timeout 300 python scripts/configs_bench.py SMORE:baby 2>/dev/null | tee -a $O/d7_step.txt
