#!/bin/bash
# launch lists of the 200k-cell Flux / Conservative rollouts (BASELINE configs[2])
O=gpurun_out/r02_ncu_flux; mkdir -p $O
for w in flux_rollout_200k cons_rollout_200k; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/$w.csv python bench.py --workload $w --steps 3 --warmup 2 --no-cpu-baseline > $O/$w.log 2>&1; tail -1 $O/$w.log | cut -c1-200
done
