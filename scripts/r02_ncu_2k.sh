#!/bin/bash
# launch list of the 2k-cell MGN rollout (BASELINE configs[0]) - per-kernel durations of a launch-bound step
O=gpurun_out/r02_ncu_2k; mkdir -p $O
timeout 200 python bench.py --workload mgn_rollout_2k --steps 50 --warmup 5 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; python scripts/print_bench.py $O/bench.json | head -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv python bench.py --workload mgn_rollout_2k --steps 3 --warmup 2 --no-cpu-baseline > $O/ncu.log 2>&1; tail -2 $O/ncu.log | cut -c1-300
