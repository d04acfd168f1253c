#!/bin/bash
mkdir -p gpurun_out
{
for n in 2000000 4000000 10000000 20000000; do
  N=$n D=32 K=4096 DTYPE=f32 STEPS=6 timeout 300 python bench/step_probe.py 2>&1 | tail -1
done
for n in 4000000 10000000; do
  echo "c5_probe N=$n"; N=$n timeout 300 python bench/c5_probe.py 2>&1 | grep -E "^primed"
done
} > gpurun_out/call8.log 2>&1
tail -30 gpurun_out/call8.log
