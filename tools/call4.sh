#!/bin/bash
mkdir -p gpurun_out
{
echo "=== stream tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stream or far_from_origin or centring or fit" 2>&1 | tail -5
for v in "" ur1 ur2 ur8; do
  echo "=== c2 probe variant '$v'"; SCKM_LIB_VARIANT=$v timeout 300 python bench/c2_probe.py 2>&1 | grep -E "default  |lloyd_fit batch=default"
  echo "=== 10M rows variant '$v'"; N=10000000 SCKM_LIB_VARIANT=$v timeout 300 python bench/c2_probe.py 2>&1 | grep -E "default  "
done
echo "=== c2 trace"; timeout 300 python bench/c2_trace_probe.py 2>&1 | tail -9
} > gpurun_out/call4.log 2>&1
tail -60 gpurun_out/call4.log
