#!/bin/bash
# first GPU run of the tcgen05 preference-module forward: its own tests in their own process (a trap poisons
# the context), timing A/B, then the whole GPU suite and the step A/B
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2b.py -x -q -m gpu -k "tcgen05 or fresh_masks" 2>&1 | tail -25 > $O/d1_tc_tests.txt
tail -3 $O/d1_tc_tests.txt
if grep -q " passed" $O/d1_tc_tests.txt && ! grep -q "failed" $O/d1_tc_tests.txt; then TC=1; else TC=0; fi
echo "TC=$TC"
timeout 200 python scripts/side_time.py 2>&1 | tail -14 | tee $O/d1_side_time.txt
MMREC_SIDE_TC=$TC timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > $O/d1_tests.txt
tail -3 $O/d1_tests.txt
MMREC_SIDE_TC=0 timeout 300 python scripts/configs_bench.py SMORE:baby 2>/dev/null | tee $O/d1_step.txt
MMREC_SIDE_TC=$TC timeout 300 python scripts/configs_bench.py SMORE:baby SMORE:sports 2>/dev/null | tee -a $O/d1_step.txt
