#!/bin/bash
# probe: GPU test suite (per-test and overall timeouts) + A/B of launch / kernel variants
mkdir -p gpurun_out
timeout 240 python -m pytest tests -x -q -m gpu --timeout 60 2>&1 | grep -v Warning | tail -${TAILN:-8} > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
for w in ${WORKLOADS:-c3_sparse c3_dense}; do
  for cfg in ${CFGS:-CRWR_PDL=1 CRWR_PDL=0}; do
    echo "== $w $cfg"
    env $cfg timeout 60 python tools/perf_probe.py --workload $w --reps 3 2>&1 | grep -E "^rep 2|^group shape|^step 16|Error|error"
  done
done > gpurun_out/ab_probe.log 2>&1
cat gpurun_out/ab_probe.log
