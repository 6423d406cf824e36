#!/bin/bash
# tcgen05 preference forward v3: tile loads hoisted above the previous write-out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2b.py -x -q -m gpu -k "tcgen05 or fresh_masks" 2>&1 | tail -4 > $O/d6_tc_tests.txt
tail -2 $O/d6_tc_tests.txt
timeout 300 python scripts/configs_bench.py SMORE:baby SMORE:sports 2>/dev/null | tee $O/d6_step.txt
MMREC_OVERLAP=0 timeout 300 python scripts/profile_step.py SMORE 5 baby 2>/dev/null > $O/d6_step_profile_smore.txt
grep -E "side_|timeline|total device" $O/d6_step_profile_smore.txt | cut -c1-120
MMREC_OVERLAP=0 timeout 300 python scripts/profile_step.py SMORE 5 sports 2>/dev/null > $O/d6_step_profile_smore_sports.txt
grep -E "side_|timeline|total device" $O/d6_step_profile_smore_sports.txt | cut -c1-120
