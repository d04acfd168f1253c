#!/bin/bash
# 2-GPU validation of the final build: multi-rank / multi-context tests, the bench line under torchrun
mkdir -p gpurun_out
{
echo "=== multi tests"; timeout 900 python -m pytest tests -x -q -m gpu -k "multi" 2>&1 | tail -4
} > gpurun_out/final_n2.log 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > gpurun_out/bench_r2b_n2.json 2> gpurun_out/bench_r2b_n2.err
echo "bench rc=$?" >> gpurun_out/final_n2.log
cat gpurun_out/final_n2.log; tail -c 300 gpurun_out/bench_r2b_n2.err
