#!/bin/bash
mkdir -p gpurun_out
{
echo "=== tc5 tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tc5" 2>&1 | tail -4
echo "=== c5 probe EW8 grouped tracking"; N=10000000 timeout 300 python bench/c5_probe.py 2>&1 | tail -8
echo "=== EW16"; SCKM_TC5H_EW16=1 N=10000000 timeout 300 python bench/c5_probe.py 2>&1 | grep -E "mean step"
for n in 4000000 10000000; do
  N=$n D=32 K=4096 DTYPE=f32 STEPS=6 timeout 300 python bench/step_probe.py 2>&1 | tail -1
done
} > gpurun_out/call9.log 2>&1
tail -30 gpurun_out/call9.log
