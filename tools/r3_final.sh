#!/bin/bash
# round-end bench table: every workload x dtype (1 GPU), default line with the CPU baseline, reference arm
mkdir -p gpurun_out/bench_r1b
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1b/bench_c3_sparse_f32.json 2>/dev/null
for w in c2_sparse c3_sparse c3_dense c4_sparse; do
  for d in f32 f64; do
    [ "$w$d" = "c3_sparsef32" ] && continue
    timeout 300 python bench.py --steps 10 --warmup 3 --workload $w --dtype $d --no-cpu-baseline > gpurun_out/bench_r1b/bench_${w}_${d}.json 2>/dev/null
  done
done
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1b/ref_c3_sparse.json 2>/dev/null
for f in gpurun_out/bench_r1b/*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
if d.get("impl")=="reference": print(sys.argv[1], "reference value", d["value"]); sys.exit()
r=d["roofline"]; print(sys.argv[1].split("/")[-1], "value %.0f e2e %.0f e2e_ms %.2f us/step %.1f prop_us %.1f frac %.4f stepfrac %.4f launches %d" % (d["value"], d["e2e"]["value"], d["imputation_wall_ms_e2e"], d["us_per_walk_step"], r["us_per_launch"], r["frac"], r["whole_step"]["frac"], d["gpu_launches"]))
PY
done
