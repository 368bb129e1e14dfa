#!/bin/bash
# PDL A/B on one box
O=gpurun_out/r02u; mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 > $O/pytest.log; cat $O/pytest.log
for i in 1 2; do for v in 0 1; do
GNNFD_PDL=$v timeout 600 python bench.py --steps 10 --warmup 3 --strong-4m off --no-cpu-baseline > $O/bench_${v}_$i.json 2> $O/bench_${v}_$i.err
echo "pdl=$v run $i: $(python scripts/print_bench.py $O/bench_${v}_$i.json 2>/dev/null | head -1)"
done; done
for v in 0 1; do
GNNFD_PDL=$v timeout 600 python bench.py --workload mgn_rollout_2k --steps 50 --warmup 5 --no-cpu-baseline > $O/roll2k_$v.json 2> $O/roll2k_$v.err
echo "pdl=$v 2k rollout: $(python scripts/print_bench.py $O/roll2k_$v.json 2>/dev/null | head -1)"
GNNFD_PDL=$v timeout 600 python bench.py --workload flux_rollout_200k --steps 20 --warmup 3 --no-cpu-baseline > $O/flux_$v.json 2> $O/flux_$v.err
echo "pdl=$v flux 200k rollout: $(python scripts/print_bench.py $O/flux_$v.json 2>/dev/null | head -1)"
done
