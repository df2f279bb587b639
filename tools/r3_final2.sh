#!/bin/bash
# last check of the round: GPU suite, default bench line, the dense workloads re-timed after the overflow-list change
mkdir -p gpurun_out/bench_r1b
timeout 400 python -m pytest tests -x -q -m gpu --timeout 120 2>&1 | tail -1
timeout 200 python bench.py --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_r1b/bench_c3_sparse_f32.json
for d in f32 f64; do timeout 200 python bench.py --steps 10 --warmup 3 --workload c3_dense --dtype $d --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_r1b/bench_c3_dense_$d.json; done
timeout 300 python bench.py --steps 6 --warmup 3 --workload c4_dense --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_r1b/bench_c4_dense_f32.json
for f in bench_c3_sparse_f32 bench_c3_dense_f32 bench_c3_dense_f64 bench_c4_dense_f32; do python - gpurun_out/bench_r1b/$f.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read()); r=d["roofline"]
print(sys.argv[1].split("/")[-1], "value %.0f e2e %.0f e2e_ms %.2f us/step %.1f prop_us %.1f frac %.4f stepfrac %.4f B_prop %.1f MB steps %d" % (d["value"], d["e2e"]["value"], d["imputation_wall_ms_e2e"], d["us_per_walk_step"], r["us_per_launch"], r["frac"], r["whole_step"]["frac"], r["bytes_per_launch"]/1e6, d["walk_steps_per_imputation"]))
PY
done
