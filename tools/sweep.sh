#!/usr/bin/env bash
# Size / kernel sweep of bench.py on one GPU; one JSON line per configuration.
set -u
OUT=${1:-gpurun_out/sweep.jsonl}
: > "$OUT"
for E in 4096 32768 262144; do
  for PF in fp64 tc; do
    python bench.py --steps 60 --warmup 10 --no-cpu-baseline --envs $E --pf-kernel $PF >> "$OUT" 2>> gpurun_out/sweep.err
  done
done
python - "$OUT" <<'PY'
import json, sys
for line in open(sys.argv[1]):
    d = json.loads(line)
    print(f"E={d['config']['envs_per_gpu']:7d} pf={d['config']['pf_kernel'][:8]:8s} "
          f"value={d['value']/1e6:9.1f} M/s  step={d['ms_per_step']*1e3:8.1f} us  "
          f"comp={d['roofline']['avg_launch_ms']*1e3:8.1f} us ({d['roofline']['frac']*100:5.1f}% hbm)  "
          f"pf={d['roofline_pf']['avg_launch_ms']*1e3:8.1f} us ({d['roofline_pf']['achieved']:7.2f} TF/s) "
          f"iters={d['roofline_pf']['mean_iterations']:.1f} warm={d['warm_l2']['ms_per_step']*1e3:8.1f} us "
          f"e2e={d['e2e']['value']/1e6:7.1f} M/s")
PY
