#!/bin/bash
# 2-GPU check of the final code: the NCCL parity test, then the bench command as the driver launches it at N = 2.
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "nccl" 2>&1 | tail -2 | tee $O/r02i_nccl_tests.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02i_bench_2gpu.json 2> $O/r02i_bench_2gpu.err
echo "rc=$?"; head -c 300 $O/r02i_bench_2gpu.json; echo; tail -3 $O/r02i_bench_2gpu.err | cut -c1-300
