#!/usr/bin/env bash
# CTA-budget sweep of the component kernel (PGW_CTA_BUDGET: tuning knob read by pgw_create).
#   gpurun -- bash tools/budget_sweep.sh "hs c2 c3" "1480 2368 2960"
for b in ${2:-1480 2368 2960}; do
  for w in ${1:-hs c2 c3}; do
    PGW_CTA_BUDGET=$b python bench.py --no-cpu-baseline --workload $w --steps 100 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('budget $b', '$w', 'step us', round(d['ms_per_step']*1e3,2), 'component kernel us', round(d.get('roofline_components',d['roofline'])['avg_launch_ms']*1e3,2))"
  done
done
