#!/bin/bash
# round-end measurement: default bench line, ncu launch list of the same command, one full capture of the dominant kernel
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c3_sparse_f32.json 2> gpurun_out/bench_c3_sparse_f32.err
tail -c 400 gpurun_out/bench_c3_sparse_f32.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r1b.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:propagate_group --launch-skip 12 --launch-count 1 \
   -f -o gpurun_out/prop_group_r1b python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
