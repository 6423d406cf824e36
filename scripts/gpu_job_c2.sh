#!/bin/bash
# round 2, session 2, call 2: 256-user scoring kernel, split-K reduced in the GEMM launch
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > $O/c2_tests.txt
tail -3 $O/c2_tests.txt
for v in 1 0; do MMREC_TOPK_256=$v timeout 200 python scripts/topk_time.py 2>/dev/null >> $O/c2_topk.txt; done
cat $O/c2_topk.txt
for cfg in "1 1" "0 1" "1 0"; do
  set -- $cfg
  MMREC_GEMM_FUSED_SPLITK=$1 MMREC_TOPK_256=$2 timeout 300 python scripts/configs_bench.py SMORE:baby 2>/dev/null | sed "s/^/fused_splitk=$1 topk256=$2 /" >> $O/c2_step_ab.txt
done
cat $O/c2_step_ab.txt
