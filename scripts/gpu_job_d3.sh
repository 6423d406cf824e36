#!/bin/bash
# where do side_fwd_tc_kernel (Baby size, 2 rounds of tiles) and side_bwd_kernel stall? --set full with source
# counters of one launch each (only after the plain run exited 0)
O=gpurun_out
timeout 200 python scripts/side_time.py 26495 2>&1 | tail -6 | tee $O/d3_side_time.txt || exit 1
MMREC_SIDE_TC=1 ncu --set full --clock-control none --import-source on -k regex:"side_fwd_tc_kernel|side_bwd_kernel" \
    --launch-skip 8 -c 1 -o /tmp/d3_tc python scripts/side_time.py 26495 > $O/d3_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"side_bwd_kernel" \
    --launch-skip 2 -c 1 -o /tmp/d3_bwd python scripts/side_time.py 26495 >> $O/d3_ncu.log 2>&1
for k in tc bwd; do
  ncu -i /tmp/d3_$k.ncu-rep --page raw --csv > $O/d3_${k}_raw.csv 2>/dev/null
  ncu -i /tmp/d3_$k.ncu-rep --page details > $O/d3_${k}_details.txt 2>/dev/null
  ncu -i /tmp/d3_$k.ncu-rep --page source --csv > $O/d3_${k}_source.csv 2>/dev/null
  gzip -f $O/d3_${k}_source.csv
done
ls -la $O/d3_*
