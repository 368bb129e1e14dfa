#!/bin/bash
# A/B of the LEAN instantiations of mlp_tc_kernel (GNNFD_LEAN=0/1): full GPU suite with them on, then the bench lines per arm
O=gpurun_out/r02_lean; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu -x > $O/pytest.log 2>&1; tail -3 $O/pytest.log
for r in 1 2; do for v in 0 1; do
  GNNFD_LEAN=$v timeout 300 python bench.py --steps 20 --warmup 5 --strong-4m off --no-cpu-baseline > $O/train_$v.json 2> $O/train_$v.err
  echo "lean=$v train: $(python scripts/print_bench.py $O/train_$v.json 2>/dev/null | head -4 | tr '\n' ' ')"
done; done
for v in 0 1; do
  GNNFD_LEAN=$v timeout 200 python bench.py --workload fvgn_fwd_8x20k --steps 20 --warmup 5 --no-cpu-baseline > $O/fwd_$v.json 2>/dev/null; echo "lean=$v fwd: $(python scripts/print_bench.py $O/fwd_$v.json 2>/dev/null | head -1)"
  GNNFD_LEAN=$v timeout 200 python bench.py --workload flux_rollout_200k --steps 20 --warmup 5 --no-cpu-baseline > $O/flux_$v.json 2>/dev/null; echo "lean=$v flux: $(python scripts/print_bench.py $O/flux_$v.json 2>/dev/null | head -1)"
  GNNFD_LEAN=$v timeout 200 python bench.py --workload mgn_rollout_2k --steps 200 --warmup 20 --no-cpu-baseline > $O/mgn2k_$v.json 2>/dev/null; echo "lean=$v 2k: $(python scripts/print_bench.py $O/mgn2k_$v.json 2>/dev/null | head -1)"
done
