#!/bin/bash
O=gpurun_out/r02r; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_training.py -q -m gpu -x 2>&1 | tail -3 > $O/pytest.log; cat $O/pytest.log
for rep in 1 2; do
echo "== new"; timeout 200 python scripts/bench_kernels.py 2>&1 | sed -n 5,6p
echo "== previous"; GNNFD_LIB=$PWD/gnn_fluid_dynamics_b200/lib_abl/libgnnfd_prev.so timeout 200 python scripts/bench_kernels.py 2>&1 | sed -n 5,6p
done > $O/ab.log 2>&1; cat $O/ab.log
timeout 600 python bench.py --steps 10 --warmup 3 --strong-4m off --no-cpu-baseline > $O/bench_train.json 2> $O/bench_train.err; tail -3 $O/bench_train.err; python scripts/print_bench.py $O/bench_train.json
