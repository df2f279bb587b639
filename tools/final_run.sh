mkdir -p gpurun_out/r2z
timeout 1300 python -m pytest tests -m gpu -q > gpurun_out/r2z/pytest_gpu_1gpu.log 2>&1; tail -2 gpurun_out/r2z/pytest_gpu_1gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z/smoke.log 2>&1; tail -1 gpurun_out/r2z/smoke.log
for w in c2_sparse c3_sparse c3_dense c4_sparse c4_dense; do timeout 400 python bench.py --steps 10 --warmup 3 --workload $w --dtype f32 > gpurun_out/r2z/bench_${w}_f32.json 2> gpurun_out/r2z/bench_${w}_f32.err; done
timeout 400 python bench.py --steps 10 --warmup 3 --workload c3_sparse --dtype f64 > gpurun_out/r2z/bench_c3_sparse_f64.json 2>/dev/null
CRWR_STRUCT=2 timeout 400 python bench.py --steps 10 --warmup 3 --workload c3_sparse --dtype f32 > gpurun_out/r2z/bench_c3_sparse_f32_structures_forced.json 2>/dev/null
timeout 400 python bench.py > gpurun_out/r2z/bench_default.json 2>/dev/null
python tools/bench_line.py gpurun_out/r2z/bench_*.json
