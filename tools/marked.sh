#!/bin/bash
mkdir -p gpurun_out
SCKM_TRACE_MARKED=1 N=10000000 timeout 300 python bench/c5_probe.py > gpurun_out/marked.log 2>&1
grep -E "marked" gpurun_out/marked.log | sort | uniq -c | sort -rn | head -12; grep -E "mean step" gpurun_out/marked.log
