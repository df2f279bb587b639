#!/bin/bash
# A/B library builds with the real bench: tools/ab_bench.sh workload libA.so libB.so ...
w=$1; shift
for rep in 1 2; do for v in "$@"; do
  out=$(CRWR_LIB=$PWD/$v python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline 2>/dev/null | tail -1)
  echo "$v $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('value %.0f us/step %.1f e2e %.0f prop_us %.1f' % (d['value'], d['us_per_walk_step'], d['e2e']['value'], r['propagate_kernel_alone']['us_per_launch']))")"
done; done
