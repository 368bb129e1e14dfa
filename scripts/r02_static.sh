#!/bin/bash
O=gpurun_out/r02_static; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_launch_overlap.py -q -m gpu -x > $O/pytest.log 2>&1; tail -3 $O/pytest.log
for r in 1 2; do for v in 0 1; do
  GNNFD_STATIC_OPERANDS=$v timeout 200 python bench.py --workload mgn_rollout_2k --steps 200 --warmup 20 --no-cpu-baseline > $O/mgn2k_$v.json 2> $O/mgn2k_$v.err; echo "static=$v 2k: $(python scripts/print_bench.py $O/mgn2k_$v.json | head -1)"
done; done
for v in 0 1; do
  GNNFD_STATIC_OPERANDS=$v timeout 200 python bench.py --workload flux_rollout_200k --steps 20 --warmup 5 --no-cpu-baseline > $O/flux_$v.json 2> $O/flux_$v.err; echo "static=$v flux: $(python scripts/print_bench.py $O/flux_$v.json | head -1)"
done
