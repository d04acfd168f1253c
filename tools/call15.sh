#!/bin/bash
mkdir -p gpurun_out
{
echo "=== memcheck"; timeout 1500 compute-sanitizer --tool memcheck python tests/sanitizer_smoke.py 2>&1 | tail -25
echo "=== racecheck"; timeout 1500 compute-sanitizer --tool racecheck python tests/sanitizer_smoke.py 2>&1 | tail -12
} > gpurun_out/sanitizer_r2b.txt 2>&1
tail -45 gpurun_out/sanitizer_r2b.txt
