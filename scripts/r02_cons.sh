#!/bin/bash
O=gpurun_out/r02_cons; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rollout.py tests/test_gpu_dist.py tests/test_gpu_training.py -q -m gpu -x > $O/pytest.log 2>&1; tail -3 $O/pytest.log
for i in 1 2; do timeout 200 python bench.py --workload cons_rollout_200k --steps 30 --warmup 5 --no-cpu-baseline > $O/cons.json 2> $O/cons.err; python scripts/print_bench.py $O/cons.json 2>/dev/null | head -1; done
