#!/bin/bash
# one ncu --set full capture of a steady-state propagate launch per configuration
mkdir -p gpurun_out
for cfg in ${CFGS:-CRWR_GROUP=16 CRWR_GROUP=8}; do
  tag=$(echo $cfg | tr '=' '_')
  env $cfg timeout 600 ncu --set full --import-source on --clock-control none -k regex:propagate_group --launch-skip 10 --launch-count 1 \
     -f -o gpurun_out/prop_$tag python tools/perf_probe.py --workload ${WORKLOAD:-c3_sparse} --reps 1 > gpurun_out/ncu_$tag.log 2>&1
  tail -3 gpurun_out/ncu_$tag.log
done
