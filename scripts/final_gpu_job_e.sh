#!/bin/bash
# Last measurement job of round 2: the full GPU suite and both bench arms on the final code (side partial reduce change).
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/r02i_gpu_tests.txt
tail -1 $O/r02i_gpu_tests.txt
python bench.py --steps 20 --warmup 5 > $O/r02i_bench_1gpu.json 2> $O/r02i_bench_1gpu.err || { tail -20 $O/r02i_bench_1gpu.err; exit 1; }
python bench.py --impl reference --steps 6 --warmup 3 > $O/r02i_bench_reference.json 2> $O/r02i_bench_reference.err
head -c 420 $O/r02i_bench_1gpu.json; echo; head -c 300 $O/r02i_bench_reference.json; echo
python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
