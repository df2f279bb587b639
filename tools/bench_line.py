#!/usr/bin/env python
"""Developer aid: the few numbers of a bench.py JSON line one wants to see at a glance."""
import json, sys
for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
        print("%-28s N=%d value %9.1f  %6.1f us/step  e2e %8.1f (%.2f ms)  hash %s  %s" % (
            path.split("/")[-1], d["n_gpus"], d["value"], d["us_per_walk_step"], d["e2e"]["value"], d.get("imputation_wall_ms_e2e", 0.0),
            d.get("result_sha256", "")[:10], d["config"]["workload"][:12]))
    except Exception as e:
        print(path, "unreadable:", e)
