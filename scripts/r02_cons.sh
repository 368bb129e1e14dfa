#!/bin/bash
O=gpurun_out/r02_cons; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu -x > $O/pytest.log 2>&1; tail -8 $O/pytest.log
timeout 200 python bench.py --workload cons_rollout_200k --steps 20 --warmup 5 --no-cpu-baseline > $O/cons.json 2> $O/cons.err; python scripts/print_bench.py $O/cons.json | head -1
