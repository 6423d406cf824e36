#!/bin/bash
O=gpurun_out
timeout 300 python scripts/host_timeline.py 2>/dev/null | tee $O/d8_host_timeline.txt
timeout 300 python scripts/host_profile.py 2>/dev/null | head -50 > $O/d8_host_profile.txt
head -45 $O/d8_host_profile.txt | cut -c1-150
