#!/bin/bash
mkdir -p gpurun_out
{
echo "=== full gpu suite"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "=== TF32 kernel probe 10M (d=32 forced)"; SCKM_TC5_TF32=1 N=10000000 timeout 300 python bench/c5_probe.py 2>&1 | grep -E "mean step"
echo "=== d=64 f32 step probe"; N=5000000 D=64 K=2048 DTYPE=f32 STEPS=6 timeout 300 python bench/step_probe.py 2>&1 | tail -1
} > gpurun_out/call14.log 2>&1
tail -20 gpurun_out/call14.log
