#!/bin/bash
mkdir -p gpurun_out
{
echo "=== stream tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stream" 2>&1 | tail -3
echo "=== c2 probe LDG.256"; timeout 300 python bench/c2_probe.py 2>&1 | grep -E "default  |lloyd_fit batch=default"
echo "=== c2 probe LDG.128"; SCKM_STREAM_NO256=1 timeout 300 python bench/c2_probe.py 2>&1 | grep -E "default  |lloyd_fit batch=default"
echo "=== 10M LDG.256"; N=10000000 timeout 300 python bench/c2_probe.py 2>&1 | grep -E "default  "
echo "=== 10M LDG.128"; SCKM_STREAM_NO256=1 N=10000000 timeout 300 python bench/c2_probe.py 2>&1 | grep -E "default  "
} > gpurun_out/call5.log 2>&1
tail -30 gpurun_out/call5.log
