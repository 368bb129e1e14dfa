#!/bin/bash
O=gpurun_out/r02l; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rollout.py -q -m gpu -x 2>&1 | tail -6 > $O/pytest.log; cat $O/pytest.log
timeout 200 python bench.py --workload cons_rollout_200k --steps 10 --warmup 3 > $O/bench_cons_rollout_200k.json 2> $O/cons.err; tail -3 $O/cons.err; python scripts/print_bench.py $O/bench_cons_rollout_200k.json
