#!/bin/bash
mkdir -p gpurun_out
{
echo "=== tc5 tests EW16"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tc5" 2>&1 | tail -4
echo "=== c5 probe EW16"; N=10000000 timeout 300 python bench/c5_probe.py 2>&1 | tail -8
echo "=== c5 probe EW8"; SCKM_TC5H_EW8=1 N=10000000 timeout 300 python bench/c5_probe.py 2>&1 | grep -E "mean step"
echo "=== c5 probe 50M EW16"; N=50000000 timeout 600 python bench/c5_probe.py 2>&1 | grep -E "mean step"
} > gpurun_out/call7.log 2>&1
tail -30 gpurun_out/call7.log
