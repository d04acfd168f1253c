#!/bin/bash
mkdir -p gpurun_out
{
echo "=== EW16 c5 probe 10M"; SCKM_TC5H_EW16=1 N=10000000 timeout 300 python bench/c5_probe.py 2>&1 | grep -E "mean step"
echo "=== EW8 c5 probe 10M"; N=10000000 timeout 300 python bench/c5_probe.py 2>&1 | grep -E "mean step"
} > gpurun_out/call13.log 2>&1
tail -40 gpurun_out/call13.log
