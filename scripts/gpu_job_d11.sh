#!/bin/bash
# split-K projections: several row tiles per CTA (one wave instead of two)
O=gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "gemm or project or linear or table" 2>&1 | tail -4 | tee $O/d11_tests.txt
for mt in 0 1 0 1; do
  MMREC_GEMM_MT=$mt timeout 300 python scripts/configs_bench.py SMORE:baby 2>/dev/null | sed "s/^/MT=$mt /" | tee -a $O/d11_step.txt
done
MMREC_GEMM_MT=0 timeout 300 python scripts/configs_bench.py SMORE:clothing MGCN:sports 2>/dev/null | sed "s/^/MT=0 /" | tee -a $O/d11_step.txt
MMREC_GEMM_MT=1 timeout 300 python scripts/configs_bench.py SMORE:clothing MGCN:sports 2>/dev/null | sed "s/^/MT=1 /" | tee -a $O/d11_step.txt
MMREC_OVERLAP=0 timeout 300 python scripts/profile_step.py SMORE 5 baby 2>/dev/null > $O/d11_step_profile_smore.txt
grep -E "gemm_tc05|splitk|total device" $O/d11_step_profile_smore.txt | cut -c1-130
