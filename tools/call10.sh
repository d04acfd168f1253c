#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
N=4000000 timeout 900 $NCU -k regex:assign_tc5h_kernel -s 8 -c 1 -f -o gpurun_out/ncu_r2b_c5_tc5h_v3_kpp python bench/c5_probe.py > gpurun_out/call10.log 2>&1
tail -5 gpurun_out/call10.log; ls -la gpurun_out/*.ncu-rep
