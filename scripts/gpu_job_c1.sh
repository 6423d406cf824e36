#!/bin/bash
# round 2, session 2, call 1: tests for in-kernel dropout + packed SpMM tasks, pack micro-benchmark, step A/B
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/c1_tests.txt
for W in 32 48 16; do
  MMREC_SPMM_PACK_WINDOW=$W timeout 300 python scripts/spmm_pack_micro.py baby sports knn > $O/c1_pack_W$W.txt 2>&1
done
timeout 300 python scripts/spmm_pack_micro.py big > $O/c1_pack_big.txt 2>&1
for cfg in "1 1" "0 1" "1 0" "0 0"; do
  set -- $cfg
  MMREC_SPMM_PACK=$1 MMREC_FUSED_DROPOUT=$2 timeout 300 python scripts/configs_bench.py SMORE:baby 2>/dev/null | sed "s/^/pack=$1 fused_dropout=$2 /" >> $O/c1_step_ab.txt
done
MMREC_SPMM_PACK=1 timeout 600 python scripts/configs_bench.py 2>/dev/null | sed "s/^/pack=1 /" >> $O/c1_step_ab.txt
MMREC_OVERLAP=0 timeout 300 python scripts/profile_step.py SMORE 5 baby > $O/c1_step_profile_smore.txt 2>/dev/null
cat $O/c1_tests.txt | tail -3; cat $O/c1_step_ab.txt
