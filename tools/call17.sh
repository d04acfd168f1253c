#!/bin/bash
mkdir -p gpurun_out
{
echo "=== memcheck"; timeout 1500 compute-sanitizer --tool memcheck python tests/sanitizer_smoke.py 2>&1 | tail -30
echo "=== racecheck"; timeout 1500 compute-sanitizer --tool racecheck python tests/sanitizer_smoke.py 2>&1 | tail -30
} > gpurun_out/sanitizer_r2b.txt 2>&1
grep -E "SUMMARY|sanitizer smoke done" gpurun_out/sanitizer_r2b.txt
