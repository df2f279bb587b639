#!/bin/bash
# Developer aid: wall time of a probed walk under a few environment settings.  usage: sweep_env.sh WORKLOAD "VAR=a VAR=b ..."
w=$1; shift
for setting in "$@"; do
  echo "== $setting"
  env $setting python tools/perf_probe.py --workload $w --reps 3 2>/dev/null | grep "^rep" | tail -2 | cut -c1-60
done
