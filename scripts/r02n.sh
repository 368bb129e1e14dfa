#!/bin/bash
O=gpurun_out/r02n; mkdir -p $O
for rep in 1 2; do
echo "== prefetch"; timeout 200 python scripts/bench_kernels.py 2>&1 | sed -n 2,6p
echo "== no prefetch"; GNNFD_LIB=$PWD/gnn_fluid_dynamics_b200/lib_abl/libgnnfd_nopf.so timeout 200 python scripts/bench_kernels.py 2>&1 | sed -n 2,6p
done > $O/ab.log 2>&1; cat $O/ab.log
