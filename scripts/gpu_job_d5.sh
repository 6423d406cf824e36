#!/bin/bash
# tcgen05 preference forward as the default for d = 64 (biases in shared memory, no re-staging of F / C)
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2b.py -x -q -m gpu -k "tcgen05 or fresh_masks" 2>&1 | tail -4 > $O/d5_tc_tests.txt
tail -2 $O/d5_tc_tests.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > $O/d5_tests.txt
tail -3 $O/d5_tests.txt
MMREC_SIDE_TC=0 timeout 300 python scripts/configs_bench.py SMORE:baby 2>/dev/null | tee $O/d5_step.txt
timeout 300 python scripts/configs_bench.py SMORE:baby SMORE:sports 2>/dev/null | tee -a $O/d5_step.txt
MMREC_OVERLAP=0 timeout 300 python scripts/profile_step.py SMORE 5 baby 2>/dev/null > $O/d5_step_profile_smore.txt
grep -E "side_|timeline|total device" $O/d5_step_profile_smore.txt | cut -c1-120
