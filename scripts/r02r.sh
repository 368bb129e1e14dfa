#!/bin/bash
O=gpurun_out/r02r; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -3 > $O/pytest.log; cat $O/pytest.log
for rep in 1 2; do
echo "== LN in chain"; timeout 200 python scripts/bench_kernels.py 2>&1 | sed -n 5,6p
echo "== separate LN kernel"; GNNFD_LN_IN_CHAIN=0 timeout 200 python scripts/bench_kernels.py 2>&1 | sed -n 5,6p
done > $O/ab.log 2>&1; cat $O/ab.log
timeout 600 python bench.py --steps 10 --warmup 3 --strong-4m off --no-cpu-baseline > $O/bench_train.json 2> $O/bench_train.err; tail -3 $O/bench_train.err; python scripts/print_bench.py $O/bench_train.json | head -2
GNNFD_LN_IN_CHAIN=0 timeout 600 python bench.py --steps 10 --warmup 3 --strong-4m off --no-cpu-baseline > $O/bench_train_sep.json 2> $O/bench_train_sep.err; python scripts/print_bench.py $O/bench_train_sep.json | head -2
