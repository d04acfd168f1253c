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
# launch list of the headline command (kmeans++ of 256 seeds comes first: ~1500 launches before the Lloyd steps)
ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches_r2_bench.csv python bench.py --steps 5 --warmup 1 --no-cpu --no-e2e --no-configs > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
# ... and of one C2 fit (the small-launch regime)
N=1000000 D=16 K=8 STEPS=6 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_c2.csv python bench/step_probe.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
