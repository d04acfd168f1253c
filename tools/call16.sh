#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis python tests/sanitizer_smoke.py > gpurun_out/racecheck_full.txt 2>&1
grep -E "hazard|Race reported|ERROR|WARN" gpurun_out/racecheck_full.txt | sed 's/0x[0-9a-f]*//g' | sort | uniq -c | sort -rn | head -30
grep -E "at .*sckm|in .*kernel|========= .*(Write|Read) Thread" gpurun_out/racecheck_full.txt | sed 's/0x[0-9a-f]*//g' | sort | uniq -c | sort -rn | head -30
