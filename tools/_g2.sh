#!/bin/bash
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r4q_bench_2gpu.log 2>&1; echo "rc=$?"
tail -c 400 gpurun_out/r4q_bench_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r4q_ref_2gpu.log 2>&1; echo "rc=$?"
tail -c 300 gpurun_out/r4q_ref_2gpu.log
