#!/bin/bash
mkdir -p gpurun_out
{ N=10000000 timeout 600 python bench/c5_trace_probe.py 2>&1 | tail -12; } > gpurun_out/call11.log 2>&1
tail -14 gpurun_out/call11.log
