#!/usr/bin/env bash
# Round-2 evidence on one GPU box: launch list of the bench command, one `ncu --set full` capture per
# workload / batch size (exported as raw CSV for bench.py's `roofline.traffic`), SASS-order stall walks.
# Usage: tools/r2_profile.sh <tag>   (files land in gpurun_out/<tag>/; copy what is judged into profiles/)
T=${1:-r2}; O=gpurun_out/$T; mkdir -p $O
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra > $O/bench_c1_steps30.json 2> $O/err.log || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_c1_steps30.csv \
    python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra > $O/ncu_launch.log 2>&1
for spec in c1:4096 c1:262144 c3:16384 c2:65536; do
  w=${spec%%:*}; e=${spec##*:}
  python bench.py --steps 6 --warmup 5 --no-cpu-baseline --no-extra --workload $w --envs $e > $O/plain_${w}_$e.json 2>> $O/err.log || continue
  ncu --set full --clock-control none --import-source on -k regex:'step_fused_kernel|component_kernel|pf_tc2_kernel' -s 20 -c 4 \
      -o $O/full_${w}_$e -f python bench.py --steps 6 --warmup 5 --no-cpu-baseline --no-extra --workload $w --envs $e > $O/ncu_full_${w}_$e.log 2>&1
  ncu -i $O/full_${w}_$e.ncu-rep --page raw --csv > $O/ncu_full_raw_${w}_$e.csv 2>> $O/err.log
  ncu -i $O/full_${w}_$e.ncu-rep --page source --csv > $O/src_${w}_$e.csv 2>> $O/err.log
  python tools/ncu_walk.py $O/src_${w}_$e.csv 0 15 > $O/ncu_stall_walk_${w}_$e.txt 2>> $O/err.log
  python tools/ncu_walk.py $O/src_${w}_$e.csv 2 15 >> $O/ncu_stall_walk_${w}_$e.txt 2>> $O/err.log   # (the source page lists every launch twice: 0 = first kernel, 2 = second)
  rm -f $O/src_${w}_$e.csv $O/full_${w}_$e.ncu-rep
done
tail -5 $O/err.log
