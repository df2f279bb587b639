#!/bin/bash
# A/B env settings on N GPUs with the real bench: tools/ab_4gpu.sh N "ENV_A" "ENV_B"
n=$1; shift
mkdir -p gpurun_out
for rep in 1 2; do for cfg in "$@"; do
  out=$(env $cfg timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$((RANDOM % 90 + 10)) bench.py --gpus $n --steps 6 --warmup 3 2>/dev/null | tail -1)
  echo "$cfg $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.0f us/step %.1f e2e %.0f prop_us %.1f' % (d['value'], d['us_per_walk_step'], d['e2e']['value'], d['roofline']['us_per_launch']))")"
done; done 2>&1 | tee gpurun_out/ab_4gpu.log
