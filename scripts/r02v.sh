#!/bin/bash
# PDL A/B on one box (after reverting the training-epilogue split shadow)
O=gpurun_out/r02v; mkdir -p $O
for i in 1 2; do for v in 0 1; do
GNNFD_PDL=$v timeout 600 python bench.py --steps 10 --warmup 3 --strong-4m off --no-cpu-baseline > $O/bench_${v}_$i.json 2> $O/bench_${v}_$i.err
echo "pdl=$v run $i: $(python scripts/print_bench.py $O/bench_${v}_$i.json 2>/dev/null | head -2 | tr '\n' ' ')"
done; done
for i in 1 2; do for v in 0 1; do
for w in mgn_rollout_2k flux_rollout_200k cons_rollout_200k fvgn_fwd_8x20k; do
GNNFD_PDL=$v timeout 600 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline > $O/${w}_${v}_$i.json 2> $O/${w}_${v}_$i.err
echo "pdl=$v $w run $i: $(python scripts/print_bench.py $O/${w}_${v}_$i.json 2>/dev/null | head -1)"
done; done; done
