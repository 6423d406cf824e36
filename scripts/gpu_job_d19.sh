#!/bin/bash
# free A/B on the final code: which preference-module forward wins on the 6 144 batch rows (tcgen05, 48 CTAs of one
# 128-row tile, against mma.sync, 96 CTAs of one 64-row tile)
O=gpurun_out
MMREC_SIDE_TC=0 timeout 200 python scripts/configs_bench.py SMORE:baby SMORE:sports 2>/dev/null | sed "s/^/SIDE_TC=0 /" | tee $O/d19_step.txt
timeout 200 python scripts/configs_bench.py SMORE:baby SMORE:sports 2>/dev/null | sed "s/^/SIDE_TC=1 /" | tee -a $O/d19_step.txt
MMREC_SIDE_TC=0 timeout 200 python scripts/configs_bench.py SMORE:baby 2>/dev/null | sed "s/^/SIDE_TC=0 /" | tee -a $O/d19_step.txt
