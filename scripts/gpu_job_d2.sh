#!/bin/bash
# tcgen05 side forward (test fix + cost-model dispatch), conflict-free dW fragment loads, PDL on the SpMM chain
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2b.py -x -q -m gpu -k "tcgen05 or fresh_masks" 2>&1 | tail -8 > $O/d2_tc_tests.txt
tail -2 $O/d2_tc_tests.txt
timeout 200 python scripts/pdl_micro.py baby 2>&1 | tail -6 | tee $O/d2_pdl_micro.txt
timeout 200 python scripts/pdl_micro.py sports 2>&1 | tail -6 | tee -a $O/d2_pdl_micro.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > $O/d2_tests.txt
tail -3 $O/d2_tests.txt
MMREC_PDL=0 timeout 300 python scripts/configs_bench.py SMORE:baby 2>/dev/null | tee $O/d2_step.txt
MMREC_PDL=1 timeout 300 python scripts/configs_bench.py SMORE:baby SMORE:sports 2>/dev/null | tee -a $O/d2_step.txt
MMREC_PDL=0 timeout 300 python scripts/configs_bench.py SMORE:sports LayerGCN:baby 2>/dev/null | tee -a $O/d2_step.txt
MMREC_PDL=1 timeout 300 python scripts/configs_bench.py LayerGCN:baby 2>/dev/null | tee -a $O/d2_step.txt
MMREC_OVERLAP=0 timeout 300 python scripts/profile_step.py SMORE 5 baby 2>/dev/null > $O/d2_step_profile_smore.txt
grep -E "side_|spmm|dense_bwd|timeline|total device" $O/d2_step_profile_smore.txt | cut -c1-120
