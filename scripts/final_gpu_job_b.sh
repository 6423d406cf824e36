#!/bin/bash
# Part B: ncu, only after the plain runs of part A exited 0. Launch list of the bench command (timed steps + timed
# evaluation, NVTX-filtered), then --set full of one launch of every hot kernel. The .ncu-rep stays on the box
# (gpurun_out/ is capped at 64 MiB); the raw CSV and the details page come back.
O=gpurun_out
ncu --nvtx --nvtx-include "timed_steps/" --nvtx-include "timed_eval/" --metrics gpu__time_duration.sum --clock-control none \
    --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 8 --no-sharded-blocks > $O/r02_ncu_bench.log 2>&1
gzip -f $O/r02_launches.csv
ncu --set full --clock-control none --import-source on \
    -k regex:"spmm_csr|gemm_tc05|side_fwd|side_bwd|side_partial|infonce_tc|score_topk_tc|mgcn_fuse|smore_combine|adam_kernel" -c 76 \
    -o /tmp/r02_kernels python scripts/ncu_kernels.py > $O/r02_ncu_kernels.log 2>&1
ncu -i /tmp/r02_kernels.ncu-rep --page raw --csv > $O/r02_kernels_raw.csv 2>/dev/null
ncu -i /tmp/r02_kernels.ncu-rep --page details > $O/r02_kernels_details.txt 2>/dev/null
gzip -f $O/r02_kernels_details.txt
du -sh $O
