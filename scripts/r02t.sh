#!/bin/bash
# A/B on one box: training forward's gathered k-blocks via TMA gather4 (1) vs register staging (0)
O=gpurun_out/r02t; mkdir -p $O
for i in 1 2; do for v in 0 1; do
GNNFD_TRAIN_TMA_GATHER=$v timeout 600 python bench.py --steps 10 --warmup 3 --strong-4m off --no-cpu-baseline > $O/bench_${v}_$i.json 2> $O/bench_${v}_$i.err
echo "gather=$v run $i: $(python scripts/print_bench.py $O/bench_${v}_$i.json 2>/dev/null | head -1)"
done; done
