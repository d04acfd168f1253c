#!/bin/bash
# GPU call 1 of the 3xFP16 work: parity of the new kernel (fold / no fold), timing against the TF32 kernel, C2 timeline
mkdir -p gpurun_out
{
echo "=== tc5 tests, FOLD"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tc5" 2>&1 | tail -15
echo "=== tc5 tests, NOFOLD"; SCKM_TC5H_NOFOLD=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tc5" 2>&1 | tail -15
echo "=== c5 probe FOLD"; N=10000000 timeout 300 python bench/c5_probe.py 2>&1 | tail -8
echo "=== c5 probe NOFOLD"; SCKM_TC5H_NOFOLD=1 N=10000000 timeout 300 python bench/c5_probe.py 2>&1 | tail -8
echo "=== c5 probe TF32"; SCKM_TC5_TF32=1 N=10000000 timeout 300 python bench/c5_probe.py 2>&1 | tail -8
echo "=== c2 trace"; timeout 300 python bench/c2_trace_probe.py 2>&1 | tail -40
} > gpurun_out/call1.log 2>&1
tail -80 gpurun_out/call1.log
