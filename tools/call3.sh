#!/bin/bash
mkdir -p gpurun_out
{
echo "=== tc5 + stream tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tc5 or stream or far_from_origin or centring" 2>&1 | tail -8
echo "=== c5 probe FOLD v2"; N=10000000 timeout 300 python bench/c5_probe.py 2>&1 | tail -8
echo "=== c2 probe"; timeout 300 python bench/c2_probe.py 2>&1 | tail -12
echo "=== c2 trace"; timeout 300 python bench/c2_trace_probe.py 2>&1 | tail -10
} > gpurun_out/call3.log 2>&1
tail -60 gpurun_out/call3.log
