#!/bin/bash
# A/B of the L2 eviction-priority hints (GNNFD_L2_HINTS bit mask, common.cuh) on one B200: parity subset with every hint
# on, then the default training bench, the inference forward and the 200k-cell Flux rollout per mask.
O=gpurun_out/r02_l2hints; mkdir -p $O
GNNFD_L2_HINTS=31 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_training.py -q -m gpu -x > $O/pytest_mask31.log 2>&1; tail -2 $O/pytest_mask31.log
for m in 0 1 3 11 15 31 0 3; do
  GNNFD_L2_HINTS=$m timeout 300 python bench.py --steps 20 --warmup 5 --strong-4m off --no-cpu-baseline > $O/train_m$m.json 2> $O/train_m$m.err
  echo "mask $m train: $(python scripts/print_bench.py $O/train_m$m.json | head -3 | tr '\n' ' ')"
done
for m in 0 2 10 14 30 0; do
  GNNFD_L2_HINTS=$m timeout 200 python bench.py --workload fvgn_fwd_8x20k --steps 20 --warmup 5 --no-cpu-baseline > $O/fwd_m$m.json 2> $O/fwd_m$m.err
  echo "mask $m fwd: $(python scripts/print_bench.py $O/fwd_m$m.json | head -1)"
  GNNFD_L2_HINTS=$m timeout 200 python bench.py --workload flux_rollout_200k --steps 20 --warmup 5 --no-cpu-baseline > $O/flux_m$m.json 2> $O/flux_m$m.err
  echo "mask $m flux: $(python scripts/print_bench.py $O/flux_m$m.json | head -1)"
done
