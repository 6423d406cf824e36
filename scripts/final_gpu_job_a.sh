#!/bin/bash
# Part A of the round's measurement job on ONE B200: tests, bench (both arms), step profiles, the other configs.
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/r02_gpu_tests.txt
tail -2 $O/r02_gpu_tests.txt
python bench.py --steps 20 --warmup 5 > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu.err
python bench.py --impl reference --steps 6 --warmup 3 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err
MMREC_OVERLAP=0 python scripts/profile_step.py SMORE 5 baby > $O/r02_step_profile_smore.txt 2>/dev/null
python scripts/profile_step.py SMORE 5 baby > $O/r02_step_profile_smore_overlap.txt 2>/dev/null
python scripts/configs_bench.py > $O/r02_other_configs.jsonl 2>/dev/null
python scripts/topk_time.py > $O/r02_topk_time.txt 2>/dev/null
head -c 1500 $O/r02_bench_1gpu.json; echo; tail -3 $O/r02_bench_1gpu.err; cat $O/r02_other_configs.jsonl
du -sh $O
