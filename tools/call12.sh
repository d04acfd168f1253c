#!/bin/bash
mkdir -p gpurun_out
{
echo "=== tc5 tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tc5" 2>&1 | tail -4
echo "=== c5 probe 10M"; N=10000000 timeout 300 python bench/c5_probe.py 2>&1 | tail -8
echo "=== trace"; N=10000000 timeout 600 python bench/c5_trace_probe.py 2>&1 | tail -10
echo "=== c5 probe 50M"; N=50000000 timeout 600 python bench/c5_probe.py 2>&1 | grep -E "mean step"
} > gpurun_out/call12.log 2>&1
tail -40 gpurun_out/call12.log
