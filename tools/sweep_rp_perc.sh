for w in c3_sparse c3_dense; do
for cfg in "0.5 50" "0.5 70" "0.5 90" "0.5 95" "0.3 80" "0.7 80" "0.7 95" "0.3 50"; do set -- $cfg
  out=$(timeout 200 python tools/perf_probe.py --workload $w --rp $1 --perc $2 --reps 3 2>&1)
  fb=$(echo "$out" | grep "^step" | awk '{s+=$13} END {print s}')
  echo "$w rp=$1 perc=$2 fb_rows=$fb $(echo "$out" | grep "^rep 2" | cut -c1-160) $(echo "$out" | grep "retries")"
done; done
