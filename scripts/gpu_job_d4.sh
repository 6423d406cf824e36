#!/bin/bash
# in-step A/B of the tcgen05 preference forward at Baby size (the eager op timing carries ~10 us of host work)
O=gpurun_out
timeout 200 python -m pytest tests/test_gpu_round2b.py -x -q -m gpu 2>&1 | tail -3 | tee $O/d4_tests.txt
MMREC_SIDE_TC=0 timeout 300 python scripts/configs_bench.py SMORE:baby 2>/dev/null | tee $O/d4_step.txt
MMREC_SIDE_TC=1 timeout 300 python scripts/configs_bench.py SMORE:baby 2>/dev/null | tee -a $O/d4_step.txt
MMREC_SIDE_TC=0 timeout 300 python scripts/configs_bench.py SMORE:baby 2>/dev/null | tee -a $O/d4_step.txt
MMREC_SIDE_TC=1 timeout 300 python scripts/configs_bench.py SMORE:baby 2>/dev/null | tee -a $O/d4_step.txt
MMREC_SIDE_TC=1 MMREC_OVERLAP=0 timeout 300 python scripts/profile_step.py SMORE 5 baby 2>/dev/null > $O/d4_step_profile_smore_tc.txt
grep -E "side_|timeline|total device" $O/d4_step_profile_smore_tc.txt | cut -c1-120
