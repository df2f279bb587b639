#!/bin/bash
# A/B two library builds on a 2-GPU box: tools/ab_2gpu.sh libA.so libB.so
for rep in 1 2; do
for v in "$@"; do
  out=$(CRWR_LIB=$PWD/$v python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 2>/dev/null | tail -1)
  echo "$v $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.0f us/step %.1f e2e %.0f' % (d['value'], d['us_per_walk_step'], d['e2e']['value']))")"
done; done
