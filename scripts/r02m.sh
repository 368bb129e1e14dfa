#!/bin/bash
O=gpurun_out/r02m; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -3 > $O/pytest.log; cat $O/pytest.log
timeout 200 python scripts/bench_kernels.py 2>&1 | head -7 > $O/kernel_microbench.log; cat $O/kernel_microbench.log
timeout 600 python bench.py --steps 10 --warmup 3 --strong-4m off --no-cpu-baseline > $O/bench_train.json 2> $O/bench_train.err; tail -3 $O/bench_train.err; python scripts/print_bench.py $O/bench_train.json
timeout 200 python bench.py --workload cons_rollout_200k --steps 10 --warmup 3 > $O/bench_cons_rollout_200k.json 2> $O/cons.err; python scripts/print_bench.py $O/bench_cons_rollout_200k.json
