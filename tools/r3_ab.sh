#!/bin/bash
# A/B of env settings with the real bench (no profiling events): lines "workload tag value us/step e2e prop_us"
# usage: CFG_A="X=1" CFG_B="CRWR_PDL=0" WORKLOADS="c3_sparse" tools/r3_ab.sh
mkdir -p gpurun_out
run() { # tag env -- workload
  tag=$1; envs=$2; w=$3
  out=$(env $envs timeout 120 python bench.py --steps 6 --warmup 3 --workload $w --no-cpu-baseline 2>/dev/null | tail -1)
  echo "$w $tag $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('value %.0f us/step %.1f e2e %.0f e2e_ms %.2f prop_us %.1f' % (d['value'], d['us_per_walk_step'], d['e2e']['value'], d['imputation_wall_ms_e2e'], r['propagate_kernel_alone']['us_per_launch']))")"
}
{
for w in ${WORKLOADS:-c3_sparse c3_dense}; do
  for rep in 1 2; do
    run A "${CFG_A:-X=1}" $w
    run B "${CFG_B:-CRWR_PDL=0}" $w
  done
done
} > gpurun_out/ab_bench.log 2>&1
cat gpurun_out/ab_bench.log
