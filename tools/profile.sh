#!/bin/bash
# Run under gpurun (one GPU): ncu --set full captures of the dominant kernel of each BASELINE shape, at sizes small
# enough for ~40 replays per launch, plus the launch list of one bench.py run.  Reports land in gpurun_out/ and are
# summarised on the CPU box by tools/ncu_summary.py into profiles/.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
N=2000000 D=128 K=1024 STEPS=2 $NCU -k regex:assign_dmma_kernel -s 2 -c 1 -f -o gpurun_out/ncu_r2_c4_assign_dmma_streamed python bench/step_probe.py
N=10000000 D=64 K=256 STEPS=2 $NCU -k regex:assign_dmma_resident_kernel -s 2 -c 1 -f -o gpurun_out/ncu_r2_c3_assign_dmma_resident python bench/step_probe.py
N=1000000 D=16 K=8 STEPS=2 $NCU -k regex:assign_stream_kernel -s 2 -c 1 -f -o gpurun_out/ncu_r2_c2_assign_stream python bench/step_probe.py
N=4000000 D=32 K=4096 DTYPE=f32 STEPS=3 $NCU -k regex:assign_tc5_kernel -s 3 -c 1 -f -o gpurun_out/ncu_r2_c5_assign_tc5 python bench/step_probe.py
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2_bench.csv python bench.py --steps 3 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
ls -la gpurun_out/*.ncu-rep
