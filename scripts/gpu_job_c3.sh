#!/bin/bash
O=gpurun_out
timeout 300 python scripts/trajectory_gap.py 2>/dev/null > $O/c3_trajectory_gap.txt
cat $O/c3_trajectory_gap.txt
MMREC_OVERLAP=0 timeout 300 python scripts/profile_step.py SMORE 5 baby 2>/dev/null | tail -3
timeout 300 python scripts/profile_step.py SMORE 5 baby 2>/dev/null | tail -3
timeout 300 python scripts/profile_step.py LayerGCN 5 baby 2>/dev/null | tail -2
timeout 300 python scripts/profile_step.py MGCN 5 sports 2>/dev/null | tail -2
