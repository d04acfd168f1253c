#!/bin/bash
# Final 1-GPU validation of the build: GPU tests, smoke, bench line, ncu captures of the changed kernels, launch list
mkdir -p gpurun_out
{
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
} > gpurun_out/final_n1.log 2>&1
timeout 1500 python bench.py > gpurun_out/bench_r2b_n1.json 2> gpurun_out/bench_r2b_n1.err
echo "bench rc=$?" >> gpurun_out/final_n1.log
NCU="ncu --set full --clock-control none --import-source on"
N=4000000 timeout 900 $NCU -k regex:assign_tc5h_kernel -s 8 -c 1 -f -o gpurun_out/ncu_r2b_c5_assign_tc5h_final python bench/c5_probe.py > /dev/null 2>&1
N=1000000 D=16 K=8 STEPS=2 timeout 600 $NCU -k regex:assign_stream_kernel -s 2 -c 1 -f -o gpurun_out/ncu_r2b_c2_assign_stream_final python bench/step_probe.py > /dev/null 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches_r2b_bench.csv python bench.py --steps 5 --warmup 1 --no-cpu --no-e2e --no-configs > gpurun_out/bench_under_ncu.json 2> /dev/null
N=4000000 D=32 K=4096 DTYPE=f32 STEPS=4 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2b_c5.csv python bench/step_probe.py > /dev/null 2>&1
tail -12 gpurun_out/final_n1.log; head -c 600 gpurun_out/bench_r2b_n1.json; ls -la gpurun_out | tail -12
