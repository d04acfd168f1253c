#!/bin/bash
mkdir -p gpurun_out
{
echo "=== stream + tc5 tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stream or tc5" 2>&1 | tail -3
} > gpurun_out/final_n1b.log 2>&1
timeout 1500 python bench.py > gpurun_out/bench_r2b_n1.json 2> gpurun_out/bench_r2b_n1.err
echo "bench rc=$?" >> gpurun_out/final_n1b.log
N=1000000 D=16 K=8 STEPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:assign_stream_kernel -s 2 -c 1 -f -o gpurun_out/ncu_r2b_c2_assign_stream_final python bench/step_probe.py > /dev/null 2>&1
cat gpurun_out/final_n1b.log
