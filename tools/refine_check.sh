#!/bin/bash
mkdir -p gpurun_out
{
echo "=== full gpu suite"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "=== c5 probe 10M"; N=10000000 timeout 300 python bench/c5_probe.py 2>&1 | grep -E "mean step|same sizes"
} > gpurun_out/refine_check.log 2>&1
bash tools/launch_c5.sh >> gpurun_out/refine_check.log 2>&1
cat gpurun_out/refine_check.log
