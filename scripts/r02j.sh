#!/bin/bash
O=gpurun_out/r02j; mkdir -p $O
for w in vertpot_train_8x20k; do
timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > $O/bench_$w.json 2> $O/$w.err; tail -4 $O/$w.err; python scripts/print_bench.py $O/bench_$w.json
done
