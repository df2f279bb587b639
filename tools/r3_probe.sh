#!/bin/bash
mkdir -p gpurun_out
for w in ${WORKLOADS:-c3_sparse c3_dense}; do
  for cfg in ${CFGS:-CRWR_PROP=owner CRWR_PROP=group}; do
    echo "== $w $cfg"
    env $cfg timeout 60 python tools/perf_probe.py --workload $w --reps 3 2>&1 | grep -E "^rep 2|^owner shape|^step 16|^step  5|Error|error"
  done
done > gpurun_out/ab_probe.log 2>&1
cat gpurun_out/ab_probe.log
