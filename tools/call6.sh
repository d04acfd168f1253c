#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
N=4000000 D=32 K=4096 DTYPE=f32 STEPS=3 timeout 600 $NCU -k regex:assign_tc5h_kernel -s 3 -c 1 -f -o gpurun_out/ncu_r2b_c5_assign_tc5h_v2 python bench/step_probe.py > gpurun_out/call6.log 2>&1
{
echo "=== c5 probe 50M"; N=50000000 timeout 600 python bench/c5_probe.py 2>&1 | tail -8
} >> gpurun_out/call6.log 2>&1
tail -12 gpurun_out/call6.log
