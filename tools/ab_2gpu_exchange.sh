#!/bin/bash
# peer-memory exchange vs NCCL collectives on a 2-GPU box
for rep in 1 2; do for x in peer nccl; do
  out=$(CRWR_SHARDED_EXCHANGE=$x python -m torch.distributed.run --nnodes=1 --nproc-per-node ${1:-2} --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus ${1:-2} --steps 8 --warmup 3 2>/dev/null | tail -1)
  echo "$x $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.0f us/step %.1f e2e %.0f launches %d | %s' % (d['value'], d['us_per_walk_step'], d['e2e']['value'], d['gpu_launches'], d['config']['parallelism'][-40:]))")"
done; done
