#!/usr/bin/env bash
# One GPU box: tests, every bench line, launch list + ncu --set full exports, phase stamps.
O=gpurun_out/final; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/gpu_tests.log 2>&1; tail -2 $O/gpu_tests.log
python bench.py > $O/bench_c1.json 2> $O/err.log
python bench.py --impl reference --steps 100 --warmup 3 > $O/bench_reference.json 2>> $O/err.log
for e in 32768 262144; do python bench.py --no-cpu-baseline --envs $e >> $O/sweep_c1_sizes.jsonl 2>> $O/err.log; done
python bench.py --workload c2 > $O/bench_c2.json 2>> $O/err.log
python bench.py --no-cpu-baseline --workload c3 --steps 100 > $O/bench_c3.json 2>> $O/err.log
python bench.py --no-cpu-baseline --workload c3 > $O/bench_c3_steps200.json 2>> $O/err.log
python bench.py --no-cpu-baseline --workload c4 > $O/bench_c4.json 2>> $O/err.log
python bench.py --workload hs > $O/bench_hs.json 2>> $O/err.log
python tools/phase_probe.py c1 > $O/phase_c1.txt 2>> $O/err.log
python tools/phase_probe.py c3 > $O/phase_c3.txt 2>> $O/err.log
# launch list of the bench command (already exited 0 above)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_c1_steps30.csv \
    python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/ncu_launch.log 2>&1
# full captures: two launches of each kernel after the warm-up
for w in c1 c3; do
  ncu --set full --clock-control none --import-source on -k regex:'component_kernel|pf_tc2_kernel' -s 20 -c 4 \
      -o $O/full_$w -f python bench.py --steps 6 --warmup 5 --no-cpu-baseline --workload $w > $O/ncu_full_$w.log 2>&1
  ncu -i $O/full_$w.ncu-rep --page raw --csv > $O/ncu_full_raw_$w.csv 2>> $O/err.log
  ncu -i $O/full_$w.ncu-rep --page source --csv > $O/src_$w.csv 2>> $O/err.log
  python tools/ncu_walk.py $O/src_$w.csv 0 15 > $O/ncu_stall_walk_$w.txt 2>> $O/err.log
  python tools/ncu_walk.py $O/src_$w.csv 1 15 >> $O/ncu_stall_walk_$w.txt 2>> $O/err.log
  rm -f $O/src_$w.csv $O/full_$w.ncu-rep
done
tail -5 $O/err.log
