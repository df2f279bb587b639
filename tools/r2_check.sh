#!/bin/bash
# round-2 probe: parity suite + A/B of the propagate kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_parity.log
cat gpurun_out/pytest_parity.log
for w in ${WORKLOADS:-c3_sparse c3_dense}; do
  for cfg in ${CFGS:-CRWR_PROP=warp CRWR_GROUP=8 CRWR_GROUP=16 CRWR_GROUP=32}; do
    echo "== $w $cfg"
    env $cfg timeout 300 python tools/perf_probe.py --workload $w --reps 3 2>&1 | grep -E "^rep 2|^group shape|Error|error" 
  done
done > gpurun_out/ab_probe.log 2>&1
cat gpurun_out/ab_probe.log
