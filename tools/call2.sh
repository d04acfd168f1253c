#!/bin/bash
# GPU call 2: ncu --set full + source of the 3xFP16 kernel (4M rows, third step) and of the C2 streaming kernel
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
N=4000000 D=32 K=4096 DTYPE=f32 STEPS=3 timeout 600 $NCU -k regex:assign_tc5h_kernel -s 3 -c 1 -f -o gpurun_out/ncu_r2b_c5_assign_tc5h python bench/step_probe.py > gpurun_out/call2.log 2>&1
N=1000000 D=16 K=8 STEPS=2 timeout 600 $NCU -k regex:assign_stream_kernel -s 2 -c 1 -f -o gpurun_out/ncu_r2b_c2_assign_stream python bench/step_probe.py >> gpurun_out/call2.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -5 gpurun_out/call2.log
