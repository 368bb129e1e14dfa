#!/bin/bash
O=gpurun_out/r02f; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -4 > $O/pytest_train.log; cat $O/pytest_train.log
timeout 200 python scripts/bench_kernels.py > $O/kernel_microbench.log 2>&1; cat $O/kernel_microbench.log
timeout 600 python bench.py --steps 10 --warmup 3 --strong-4m off --no-cpu-baseline > $O/bench_train.json 2> $O/bench_train.err; tail -3 $O/bench_train.err; python scripts/print_bench.py $O/bench_train.json 2>/dev/null || head -c 600 $O/bench_train.json
