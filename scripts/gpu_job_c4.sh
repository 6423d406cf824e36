#!/bin/bash
# fast activations + pre-split weights in side_fwd; tightened trajectory tests (FREEDOM, LayerGCN-drop added)
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > $O/c4_tests.txt
tail -4 $O/c4_tests.txt
timeout 300 python scripts/configs_bench.py SMORE:baby 2>/dev/null | tee $O/c4_step.txt
timeout 300 python scripts/configs_bench.py SMORE:clothing MGCN:sports 2>/dev/null | tee -a $O/c4_step.txt
MMREC_OVERLAP=0 timeout 300 python scripts/profile_step.py SMORE 5 baby 2>/dev/null > $O/c4_step_profile_smore.txt
grep -E "side_|dense_fwd|timeline|total device" $O/c4_step_profile_smore.txt | cut -c1-120
